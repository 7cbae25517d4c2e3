"""CPU tier: the warp-parallel event detector (uncalled_b200/csrc/unc_k1.cuh) run under the warp
emulator against the oracle and the reference's golden vectors, and the exactly-rounded
constant-divisor sequences it relies on against IEEE division (exhaustive in float)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(example_prefix):
    import emulib
    return emulib.Emu(example_prefix)


def _check(emu, sigs, dtype=0, cal=(1.0, 0.0, 1.0), pas=None):
    import emulib
    import orclib
    O = orclib.Oracle()
    recs, ev, nm, mel = emu.map_batch(sigs, run_k2=False, dtype=dtype, cal=cal)
    for i, s in enumerate(sigs):
        x = pas[i] if pas is not None else np.ascontiguousarray(s, dtype=np.float32)
        m, _, _, omel = O.detect(x)
        assert len(m) == len(ev[i]) and np.array_equal(m, ev[i]), i
        assert omel == mel[i] or (np.isnan(omel) and np.isnan(mel[i])), i
        if len(m):
            assert np.array_equal(O.normalize(m), nm[i], equal_nan=True), i
    return emulib.k1_stats()


def test_constant_division_sequences_exhaustive(tmp_path):
    exe = str(tmp_path / "k1_arith")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-o", exe,
                    os.path.join(ROOT, "tests", "arith", "k1_arith_check.c"), "-lm"], check=True)
    r = subprocess.run([exe, "20000000"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr


def test_golden_read_and_windows(emu, golden_read):
    raw = golden_read["raw"]
    sigs = [raw] + [raw[o:o + 4000] for o in (0, 4000, 8000, 20000)]
    recs, ev, nm, mel = emu.map_batch(sigs, run_k2=False)
    assert np.array_equal(ev[0], golden_read["ev_mean"]) and np.array_equal(nm[0], golden_read["normed"])
    assert mel[0] == golden_read["mean_event_len"]
    assert [len(e) for e in ev[1:]] == list(golden_read["win_counts"])


def test_edge_lengths_and_tile_boundaries(emu, golden_read):
    raw = golden_read["raw"]
    lens = [0, 1, 5, 6, 7, 11, 12, 13, 40, 1151, 1152, 1153, 1157, 1158, 1159, 2304, 2309, 2310, 3456 + 5]
    sigs = [raw[7 * i:7 * i + L] for i, L in enumerate(lens)]          # odd offsets: unaligned bulk-copy sources
    st = _check(emu, sigs)
    assert st[3] == 0


def test_synthetic_ragged_f32_and_i16(emu):
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 24, 9000, seed=11)
    rng = np.random.default_rng(5)
    lens = [int(x) for x in rng.integers(100, 9000, 24)]
    sigs = [sig[i, :L] for i, L in enumerate(lens)]
    st = _check(emu, sigs)
    assert st[3] == 0 and st[1] < st[0]          # no serial redo; FSM speculation rarely re-runs
    for cal in [(1467.61, 10.0, 8192.0), (1534.14, 3.0, 8000.0)]:   # exact-reciprocal and true-division calibration
        i16s, pas = [], []
        for s in sigs[:10]:
            raw = np.clip(np.round(s.astype(np.float64) * cal[2] / cal[0] - cal[1]), -50, 32000).astype(np.int16)
            raw[50:53] = -7                                            # negative DAC values wrap to ~65k
            i16s.append(raw)
            pas.append(((np.float32(cal[0]) * (raw.astype(np.uint16).astype(np.float32) + np.float32(cal[1]))) /
                        np.float32(cal[2])).astype(np.float32))
        _check(emu, i16s, dtype=1, cal=cal, pas=pas)


def test_inexact_sums_fall_back_to_the_serial_routine(emu):
    """samples whose prefix sums round (tiny next to huge values) must take the serial path and
    still match; flat and all-zero signals exercise the variance clamp and the never-firing FSM."""
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 2, 5000, seed=2)
    s = sig[0].copy(); s[1000] = 1e-20; s[2000] = 3e7
    t = sig[1].copy(); t[10] = np.float32(1e-41)                       # subnormal
    z = np.zeros(3000, np.float32)
    c = np.full(3000, 87.25, np.float32); c[1500:] = 90.5
    st = _check(emu, [s, z, c, t, sig[1]])
    assert st[3] == 2


def test_non_finite_samples_take_the_serial_path(emu):
    """NaN / Inf samples (a corrupt read) must neither hang nor diverge: the exactness check sends the read to the
    serial routine, whose arithmetic is the reference's."""
    import orclib
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 2, 3000, seed=8)
    a = sig[0].copy(); a[700] = np.nan
    b = sig[1].copy(); b[100] = np.inf; b[2000] = -np.inf
    recs, ev, nm, mel = emu.map_batch([a, b], run_k2=False)
    import emulib
    assert emulib.k1_stats()[3] == 2
    O = orclib.Oracle()
    for i, s in enumerate((a, b)):
        m, _, _, omel = O.detect(s)
        assert len(m) == len(ev[i]) and np.array_equal(m, ev[i], equal_nan=True), i
