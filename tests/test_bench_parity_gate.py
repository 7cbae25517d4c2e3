"""bench.py's in-bench parity gate (host logic only): how differences between the GPU records and the reference
run's records are classified -- Mapper carry (explained by a fresh-Mapper re-map), tie order (explained by the
exact-ties kernel), or unexplained (fails the run)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

FIELDS = [("mapped", "i4"), ("fwd", "i4"), ("rid", "i4"), ("n_events", "u4"), ("events_used", "u4"), ("matches", "u4"),
          ("rd_len", "u8"), ("rd_st", "u8"), ("rd_en", "u8"), ("rf_st", "u8"), ("rf_en", "u8"), ("rf_len", "u8")]


def _recs(n):
    a = np.zeros(n, dtype=FIELDS)
    a["mapped"] = 1
    a["fwd"] = 1
    a["n_events"] = 780
    a["events_used"] = np.arange(n) + 100
    a["matches"] = 40
    a["rd_len"] = 450
    a["rf_st"] = 1000 + np.arange(n)
    a["rf_en"] = 1400 + np.arange(n)
    a["rf_len"] = 4700000
    return a


def _keys(a):
    import uncalled_b200 as U
    return [U.paf_key(a[i]) for i in range(len(a))]


def test_gate_classifies_carry_tie_and_unexplained(monkeypatch):
    n = 8
    truth = _recs(n)                       # what a fresh Mapper of the reference gives
    fresh = dict(enumerate(_keys(truth)))
    ref_run = _recs(n)
    ref_run["matches"][2] = 41             # read 2: the timed run's Mapper carried flags from its previous read
    gpu_exact = _recs(n)                   # the exact-ties kernel = the reference on every read
    gpu_default = _recs(n)
    gpu_default["rf_en"][5] += 1           # read 5: a tie the stable order decides differently
    monkeypatch.setattr(bench, "cpu_fresh_mapper_keys", lambda prefix, sig, ids: {i: fresh[i] for i in ids})
    rep = bench.parity_report("p", None, _keys(ref_run), gpu_default, gpu_exact)
    assert rep["ok"]
    assert rep["default_kernel"]["identical"] == n - 2
    assert rep["default_kernel"]["explained_by_mapper_carry"] == 1
    assert rep["default_kernel"]["explained_by_tie_order"] == 1
    assert rep["exact_ties_kernel"]["identical"] == n - 1 and rep["exact_ties_kernel"]["unexplained"] == []
    # a real mismatch: neither carry nor tie
    gpu_default["rd_st"][6] += 3
    gpu_exact["rd_st"][6] += 3
    rep = bench.parity_report("p", None, _keys(ref_run), gpu_default, gpu_exact)
    assert not rep["ok"]
    assert rep["default_kernel"]["unexplained"] == [6] and rep["exact_ties_kernel"]["unexplained"] == [6]


def test_host_cpus_respects_affinity_and_is_positive():
    c = bench.host_cpus()
    assert 1 <= c["usable"] <= c["affinity"]
    assert bench.cpu_sample_size(16, 10000) == 768 and bench.cpu_sample_size(128, 10000) == 4096 and bench.cpu_sample_size(2, 10000) == 256
