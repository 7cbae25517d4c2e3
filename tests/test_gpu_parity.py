"""GPU parity tests: the CUDA path, called through the C-ABI, against the golden vectors of
the real reference and against the oracle on the same seeded inputs.  Bit-exact throughout
(integer/index work; the float pipeline reproduces the reference's rounding exactly, so the
1e-4 tolerance of the north star is tightened to equality here)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def U():
    import uncalled_b200
    uncalled_b200._native.lib()
    return uncalled_b200


@pytest.fixture(scope="module")
def ex(U, example_prefix):
    idx = U.Index(example_prefix, device=0)
    bm = U.BatchMapper(idx, max_reads=64, max_samples=64 * 32000)
    return idx, bm


def test_events_and_normalisation_match_reference_golden(U, ex, golden_read):
    idx, bm = ex
    raw = golden_read["raw"]
    offs = [0, 4000, 8000, 20000]
    sigs = [raw] + [raw[o:o + 4000] for o in offs]
    ev, nm, ne, mel = bm.events(np.concatenate(sigs), U.make_descs([len(s) for s in sigs]))
    assert ne[0] == len(golden_read["ev_mean"]) == 6171
    assert np.array_equal(ev[0, :ne[0]], golden_read["ev_mean"])
    assert np.array_equal(nm[0, :ne[0]], golden_read["normed"])     # bit-exact (tolerance 0 <= 1e-4)
    assert mel[0] == golden_read["mean_event_len"]
    assert list(ne[1:]) == list(golden_read["win_counts"])


def test_events_i16_calibration(U, ex, golden_read):
    """raw DAC input calibrated on the device as read_buffer.cpp:239-242 (u16 reinterpretation)."""
    import orclib
    idx, bm = ex
    rng = np.random.default_rng(3)
    i16 = rng.integers(200, 1200, 6000).astype(np.int16)
    i16[100:110] = -5  # negative DAC values wrap to ~65k and are rejected by max_mean
    cal = (1467.61, 10.0, 8192.0)
    pa = (np.float32(cal[0]) * (i16.astype(np.uint16).astype(np.float32) + np.float32(cal[1]))) / np.float32(cal[2])
    ev, nm, ne, mel = bm.events(i16, U.make_descs([len(i16)], dtype=1, cal=cal))
    O = orclib.Oracle()
    m, s, l, omel = O.detect(pa.astype(np.float32))
    assert ne[0] == len(m) and np.array_equal(ev[0, :ne[0]], m) and mel[0] == omel


def _events_vs_oracle(U, bm, sigs, dtype=0, cal=(1.0, 0.0, 1.0), pas=None, offsets=None):
    import orclib
    O = orclib.Oracle()
    lens = [len(s) for s in sigs]
    flat = np.concatenate(sigs) if offsets is None else offsets[1]
    d = U.make_descs(lens, dtype=dtype, cal=cal, offsets=None if offsets is None else offsets[0])
    ev, nm, ne, mel = bm.events(flat, d)
    for i, s in enumerate(sigs):
        x = pas[i] if pas is not None else np.ascontiguousarray(s, np.float32)
        m, _, _, omel = O.detect(x)
        assert ne[i] == len(m) and np.array_equal(ev[i, :ne[i]], m), i
        assert mel[i] == omel or (np.isnan(mel[i]) and np.isnan(omel)), i
        if len(m):
            assert np.array_equal(nm[i, :ne[i]], O.normalize(m), equal_nan=True), i
    return bm.k1_stats()


def test_events_edge_lengths_unaligned_and_tile_boundaries(U, ex, golden_read):
    """warp-parallel event detector: empty/tiny reads, lengths around the 1152-position tile
    boundary, reads starting at odd sample offsets (unaligned bulk-copy sources, buffer tail)."""
    idx, bm = ex
    raw = golden_read["raw"]
    lens = [0, 1, 5, 6, 7, 11, 12, 13, 40, 1151, 1152, 1153, 1157, 1158, 1159, 2304, 2309, 2310, 3461]
    sigs = [raw[7 * i:7 * i + L] for i, L in enumerate(lens)]
    st = _events_vs_oracle(U, bm, sigs)
    assert st[3] == 0


def test_events_inexact_sums_take_the_serial_path(U, ex):
    import synth
    import synthdata
    idx, bm = ex
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 3, 5000, seed=2)
    s = sig[0].copy(); s[1000] = 1e-20; s[2000] = 3e7
    t = sig[1].copy(); t[10] = np.float32(1e-41)
    z = np.zeros(3000, np.float32)
    c = np.full(3000, 87.25, np.float32); c[1500:] = 90.5
    st = _events_vs_oracle(U, bm, [s, z, c, t, sig[2]])
    assert st[3] == 2


def test_events_many_synthetic_reads_f32_and_i16(U, ex):
    import synth
    import synthdata
    idx, bm = ex
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 60, 9000, seed=11)
    rng = np.random.default_rng(5)
    lens = [int(x) for x in rng.integers(100, 9000, 60)]
    sigs = [sig[i, :L] for i, L in enumerate(lens)]
    st = _events_vs_oracle(U, bm, sigs)
    assert st[3] == 0
    for cal in [(1467.61, 10.0, 8192.0), (1534.14, 3.0, 8000.0)]:
        i16s, pas = [], []
        for s in sigs[:24]:
            raw = np.clip(np.round(s.astype(np.float64) * cal[2] / cal[0] - cal[1]), -50, 32000).astype(np.int16)
            raw[50:53] = -7
            i16s.append(raw)
            pas.append(((np.float32(cal[0]) * (raw.astype(np.uint16).astype(np.float32) + np.float32(cal[1]))) /
                        np.float32(cal[2])).astype(np.float32))
        _events_vs_oracle(U, bm, i16s, dtype=1, cal=cal, pas=pas)


def test_pore_model_matches_reference_golden(U, ex):
    idx, _ = ex
    g = np.load(os.path.join(ROOT, "tests", "golden", "example_model.npz"))
    for e, want in zip(g["events"], g["probs"]):
        assert np.array_equal(idx.match_probs(e), want)


def test_fm_index_matches_reference_golden(U, ex, example_prefix):
    import orclib
    idx, _ = ex
    g = np.load(os.path.join(ROOT, "tests", "golden", "example_index.npz"))
    st, en = idx.kmer_ranges()
    assert np.array_equal(en - st + 1, g["kmer_count"])
    n = int(g["size"])
    assert idx.n_rows == n
    assert np.array_equal(idx.sa(np.arange(1, n + 1)), g["sa_1_to_n"])
    O = orclib.Oracle(example_prefix)
    rng = np.random.default_rng(11)
    s = rng.integers(1, n, 4000).astype(np.uint64)
    e = np.minimum(s + rng.integers(0, 40, 4000).astype(np.uint64), n)
    b = rng.integers(0, 4, 4000).astype(np.uint8)
    os_, oe = idx.neighbors(s, e, b)
    import ctypes as C
    a, c = C.c_uint64(), C.c_uint64()
    for i in range(4000):
        O.lib.orc_get_neighbor(O.idx, int(s[i]), int(e[i]), int(b[i]), C.byref(a), C.byref(c))
        assert (a.value, c.value) == (int(os_[i]), int(oe[i])), i


def _golden_fields(key):
    f = json.load(open(os.path.join(ROOT, "tests", "golden", "example_paf.json")))[key]["fields"]
    return f


def test_map_example_read_matches_reference_paf(U, ex, golden_read):
    """config 1: the example read's PAF line, field for field, as `uncalled map` prints it."""
    idx, bm = ex
    raw = golden_read["raw"]
    out = bm.map(np.concatenate([raw, raw[:4000]]), U.make_descs([len(raw), 4000]))
    for r, key in ((out[0], "default"), (out[1], "max_chunks_1")):
        f = _golden_fields(key)
        assert r["mapped"] == 1 and r["status"] == 0
        got = [str(int(r["rd_len"])), str(int(r["rd_st"])), str(int(r["rd_en"])), "+" if r["fwd"] else "-",
               idx.seqs[r["rid"]][0], str(int(r["rf_len"])), str(int(r["rf_st"])), str(int(r["rf_en"])),
               str(int(r["matches"])), str(int(r["rf_en"] - r["rf_st"] + 1)), "255"]
        assert got == f[1:], (got, f)


def test_map_example_max_events(U, example_prefix, golden_read):
    idx = U.Index(example_prefix, device=0)
    p = U.default_params()
    p.max_events = 100
    bm = U.BatchMapper(idx, params=p, max_reads=4, max_samples=40000)
    raw = golden_read["raw"]
    out = bm.map(raw, U.make_descs([len(raw)]))
    f = _golden_fields("max_events_100")
    assert out[0]["mapped"] == 0 and str(int(out[0]["rd_len"])) == f[1] and out[0]["events_used"] == 100


def _compare_with_oracle(U, name, n_reads, n_samples, seed, threads=8, params_mod=None):
    import orclib
    import synth
    import synthdata
    prefix, g = synthdata.get_index(name)
    sig, truth = synth.reads(g, n_reads, n_samples, seed=seed)
    idx = U.Index(prefix, device=0)
    p = U.default_params()
    O = orclib.Oracle(prefix)
    if params_mod:
        params_mod(p)
        params_mod(O.params)
    bm = U.BatchMapper(idx, params=p, max_reads=n_reads, max_samples=n_reads * n_samples)
    out = bm.map(sig.ravel(), U.make_descs([n_samples] * n_reads))
    offs = np.arange(n_reads, dtype=np.uint64) * n_samples
    want = O.map_batch(sig.ravel(), offs, np.full(n_reads, n_samples, np.uint32), threads)
    bad = []
    for i in range(n_reads):
        a, b = orclib.paf_tuple(want[i]), U.paf_key(out[i])
        ca = (want[i].n_children, want[i].n_sources, want[i].n_seeds, want[i].n_clusters)
        cb = (int(out[i]["n_children"]), int(out[i]["n_sources"]), int(out[i]["n_seeds"]), int(out[i]["n_clusters"]))
        if a != b or ca != cb or out[i]["status"] != 0:
            bad.append((i, a, b, ca, cb, int(out[i]["status"])))
    assert not bad, bad[:5]
    return out, want


def test_map_synthetic_200k_matches_oracle(U):
    out, want = _compare_with_oracle(U, "g200k", 256, 4000, seed=7)
    assert sum(r.mapped for r in want) > 150


def test_map_synthetic_4m7_matches_oracle(U):
    """E. coli-sized index (4.7 Mb): includes reads that hit the max_paths cap."""
    out, want = _compare_with_oracle(U, "g4m7", 48, 4000, seed=21)
    assert max(r.max_paths_seen for r in want) == 10000


def test_map_long_reads_on_20mb_index(U):
    """config-4-like at reduced scale: a 20 Mb genome (40 M FM rows: 26-bit keys -> 4 radix passes,
    160 MB expanded SA) and 30 000-sample reads (~5 800 events, 26 K1 tiles) incl. never-mapping
    reads that sit at the max_paths cap for thousands of events."""
    import orclib
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g20m")
    n, L = 12, 30000
    sig, truth = synth.reads(g, n, L, seed=41, frac_random=0.25)
    idx = U.Index(prefix, device=0)
    bm = U.BatchMapper(idx, max_reads=n, max_samples=n * L)
    out = bm.map(sig.ravel(), U.make_descs([L] * n))
    O = orclib.Oracle(prefix)
    want = O.map_batch(sig.ravel(), np.arange(n, dtype=np.uint64) * L, np.full(n, L, np.uint32), 8)
    for i in range(n):
        assert orclib.paf_tuple(want[i]) == U.paf_key(out[i]) and out[i]["status"] == 0, i
        assert (want[i].n_children, want[i].n_sources, want[i].n_seeds) == \
               (int(out[i]["n_children"]), int(out[i]["n_sources"]), int(out[i]["n_seeds"])), i
    assert any(w.mapped == 0 and w.events_used > 5000 for w in want) and sum(w.mapped for w in want) >= 6


def test_map_ragged_and_tiny_reads(U):
    """ragged batch: empty-event reads, very short reads, long reads in one call."""
    import orclib
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 6, 12000, seed=5)
    lens = [12000, 0 + 5, 40, 300, 4000, 7777]
    sigs = [sig[i, :L] for i, L in enumerate(lens)]
    idx = U.Index(prefix, device=0)
    bm = U.BatchMapper(idx, max_reads=8, max_samples=sum(lens) + 16)
    out = bm.map(np.concatenate(sigs), U.make_descs(lens))
    O = orclib.Oracle(prefix)
    for i, s in enumerate(sigs):
        assert orclib.paf_tuple(O.map_read(s)) == U.paf_key(out[i]), i


def test_small_max_paths_overflow_semantics(U):
    """max_paths tiny: exercises the full-buffer break, stale sources_added_ flags and the
    source caps on every event."""
    def mod(p):
        p.max_paths = 300
    _compare_with_oracle(U, "g200k", 64, 3000, seed=13, params_mod=mod)


def _stream_vs_oracle(U, prefix, sigs, n_channels, chunk_len, max_chunks=1000000, params_mod=None, tie_order=0):
    import orclib
    idx = U.Index(prefix, device=0)
    p = U.default_params()
    O = orclib.Oracle(prefix)
    if params_mod:
        params_mod(p)
        params_mod(O.params)
    sm = U.StreamMapper(idx, n_channels, chunk_len, max_chunks=max_chunks, params=p)
    if tie_order:
        sm.set_tie_order(tie_order)
    res = sm.map_reads(sigs)
    sm.close()
    states = []
    for c in range(n_channels):
        idxs = list(range(c, len(sigs), n_channels))
        want = O.stream_channel([sigs[i] for i in idxs], chunk_len, max_chunks)
        for i, (rec, nu, en) in zip(idxs, want):
            r = res[i]
            if nu == 0:
                assert r is None, i
                continue
            assert (U.paf_key(r[3]), r[2], r[1]) == (orclib.paf_tuple(rec), nu, en), i
            assert (r[3].n_children, r[3].n_sources, r[3].n_seeds) == (rec.n_children, rec.n_sources, rec.n_seeds), i
            assert r[3].status == 0 and r[0] == (2 if rec.mapped else 3)
            states.append((r[0], r[1]))
    return states


def test_stream_chunks_match_streaming_oracle(U):
    """config-5-like at test scale: 450-sample chunks, 8 channels, 3 reads per channel one after the other (the
    channel's normaliser statistics and sources_added_ persist), against the oracle's streaming restatement."""
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 24, 8000, seed=15, frac_random=0.3)
    sigs = [sig[i][:8000 - 53 * i] for i in range(24)] + [sig[0][:200]]
    st = _stream_vs_oracle(U, prefix, sigs, 8, 450)
    assert (2, 0) in st and (3, 1) in st


def test_stream_limits(U):
    """max_chunks, max_events in the middle of a chunk, a tiny path buffer, one-second chunks."""
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 10, 6000, seed=9, frac_random=0.5)
    sigs = [sig[i] for i in range(10)]
    _stream_vs_oracle(U, prefix, sigs, 5, 450, max_chunks=4)
    _stream_vs_oracle(U, prefix, sigs, 4, 4000)

    def mod(p):
        p.max_events = 150
        p.max_paths = 300
    st = _stream_vs_oracle(U, prefix, sigs, 3, 450, params_mod=mod)
    assert (3, 1) in st


def test_map_matches_records_computed_by_the_reference_itself(U):
    """No oracle in between: tests/golden/synth_paf_golden.json holds what the reference's own code computes for 320 seeded
    reads (tools/make_synth_paf_golden.py) -- with its child sort made stable, which is the order this path defines for
    equal children, and as it is (pdqsort): identical to the former on every read, to the latter on all but the two reads
    the fixture lists (tie order)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synth_paf_golden as M
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "synth_paf_golden.json")))
    for name, n, ns, seed, frac in M.SETS:
        prefix, sig = M.signals(name, n, ns, seed, frac)
        idx = U.Index(prefix, device=0)
        bm = U.BatchMapper(idx, max_reads=n, max_samples=n * ns)
        out = bm.map(np.ascontiguousarray(sig.reshape(-1)), U.make_descs([ns] * n))
        got = [[int(v) for v in U.paf_key(r)] for r in out]
        assert got == gold["reference_stable_sort"][name], name
        assert [i for i in range(n) if got[i] != gold["reference"][name][i]] == gold["differ"][name]
        bm.close()

