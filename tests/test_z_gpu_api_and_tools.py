"""GPU tests that sort LAST on purpose: the Python surface (MapPool, CLI) and the paths added after this round's GPU
minutes were spent (fast5 -> map end to end, `uncalled index` on the GPU, submit/wait on two pools).  `pytest -x` reaches
them only after the parity suite of tests/test_gpu_parity.py has run."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "example_paf.json")))


@pytest.fixture(scope="module")
def U():
    import uncalled_b200
    uncalled_b200._native.lib()
    return uncalled_b200


def test_submit_wait_on_two_pools_matches_the_synchronous_call(U):
    """unc_map_batch_submit / _wait: batches in flight on two pools at once give the records of unc_map_batch."""
    import synth
    import synthdata
    prefix, g = synthdata.get_index("g200k")
    idx = U.Index(prefix, device=0)
    sig, _ = synth.reads(g, 96, 3000, seed=13, frac_random=0.3)
    halves = [np.ascontiguousarray(sig[:48].reshape(-1)), np.ascontiguousarray(sig[48:].reshape(-1))]
    d = U.make_descs([3000] * 48)
    pools = [U.BatchMapper(idx, max_reads=48, max_samples=48 * 3000) for _ in range(2)]
    want = [pools[0].map(halves[0], d).copy(), pools[0].map(halves[1], d).copy()]
    for rep in range(3):
        pools[0].submit(halves[0], d)
        pools[1].submit(halves[1], d)
        with pytest.raises(U.UncError):
            pools[0].submit(halves[1], d)                      # a pool holds one batch at a time
        got1 = pools[1].wait()
        got0 = pools[0].wait()
        assert np.array_equal(got0, want[0]) and np.array_equal(got1, want[1]), rep
    pools[0].record(0)                                             # device-side timing across the two pools
    pools[0].submit(halves[0], d)
    pools[1].submit(halves[1], d)
    pools[0].record(1)
    pools[1].record(1)
    assert np.array_equal(pools[0].wait(), want[0]) and np.array_equal(pools[1].wait(), want[1])
    span = max(pools[0].elapsed_ms(0, pools[0], 1), pools[0].elapsed_ms(0, pools[1], 1))
    assert 0.0 < span < 60000.0
    with pytest.raises(U.UncError):
        pools[0].wait()                                            # nothing submitted
    with pytest.raises(U.UncError):
        U._native.check(pools[0].L.unc_map_batch_wait(pools[0].h, want[0].ctypes.data))   # the C entry point says so too
    for p in pools:
        p.close()


def test_self_align_matches_oracle_and_reference_digest(U, example_prefix, tmp_path):
    """`uncalled index`: unc_self_align against the oracle, the reference's digests and, end to end, the
    reference-made .uncl files."""
    import hashlib
    import orclib
    import synthdata
    from uncalled_b200 import index as UI, index_params as IP
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "self_align_golden.json")))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a, "<u8").tobytes()).hexdigest()   # noqa: E731
    for row in gold:
        prefix = example_prefix if row["index"] == "example" else synthdata.get_index(row["index"])[0]
        off, val = UI.self_align_csr(prefix, row["sample_dist"])
        assert (len(off) - 1, len(val)) == (row["n_paths"], row["n_values"])
        assert sha(off) == row["offsets_sha256"] and sha(val) == row["values_sha256"], row["index"]
    for name, sd in (("g4m7", 94), ("g200k", 1)):
        prefix = synthdata.get_index(name)[0]
        a, b = UI.self_align_csr(prefix, sd), orclib.self_align(prefix, sd)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), name
    # repeats, tiny and ambiguous sequences: paths beyond the staging depth are re-walked in pass 2
    import test_selfalign_emul as tse
    prefix = tse.repeat_index(str(tmp_path))
    for sd in (1, 3):
        a, b = UI.self_align_csr(prefix, sd), orclib.self_align(prefix, sd)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), sd
    assert UI.self_align(prefix, 3) == [[int(v) for v in b[1][int(b[0][i]):int(b[0][i + 1])]] for i in range(len(b[0]) - 1)]


def test_index_cmd_reproduces_reference_uncl_files(U, tmp_path):
    """FASTA -> .bwt/.sa/... -> self_align on the GPU -> .uncl, against what the real `uncalled index` wrote."""
    import shutil
    import synth
    import orclib
    from uncalled_b200 import index as UI
    synth_uncl = json.load(open(os.path.join(ROOT, "tests", "golden", "synth_uncl.json")))
    fa = str(tmp_path / "g.fa")
    synth.write_fasta(fa, synth.genome(synth_uncl["g200k"]["size"], synth_uncl["g200k"]["seed"]))
    prefix = UI.index_cmd(fa, str(tmp_path / "g200k"))
    assert open(prefix + ".uncl").read() == synth_uncl["g200k"]["uncl"]
    ex = orclib.materialise_example_index(str(tmp_path))
    shipped = open(ex + ".uncl").read()
    os.remove(ex + ".uncl")
    UI.index_cmd(ex + ".fa", ex)                                   # reuses the shipped BWA files
    assert open(ex + ".uncl").read() == shipped
    presets = json.load(open(os.path.join(ROOT, "tests", "golden", "uncl_presets.json")))[0]
    assert UI.write_uncl(ex, probs=presets["probs"], speeds=presets["speeds"]) == presets["uncl"]


def test_map_pool_prints_the_reference_paf_lines(example_prefix, golden_read):
    """`uncalled map`, `-c 1` and `-e 100` on the example read (config 1), line for line up to the mt tag."""
    from uncalled_b200.api import Conf, MapPool
    raw = golden_read["raw"]
    rid, ch, st = str(golden_read["read_id"]), int(golden_read["channel"]), int(golden_read["start"])
    for key, mod in (("default", {}), ("max_chunks_1", {"max_chunks": 1}), ("max_events_100", {"max_events": 100})):
        conf = Conf()
        conf.bwa_prefix = example_prefix
        for k, v in mod.items():
            setattr(conf, k, v)
        pool = MapPool(conf)
        assert not pool.running()
        pool.add_read(rid, raw, channel=ch, number=0, start_sample=st)
        lines = []
        while pool.running():
            lines += [p.line() for p in pool.update()]
        pool.stop()
        assert len(lines) == 1
        body, mt = lines[0].rsplit("\t", 1)
        assert body == GOLD[key]["line"] and mt.startswith("mt:f:"), (key, lines[0])


def test_map_pool_i16_reads_and_filters(example_prefix, golden_read, tmp_path):
    from uncalled_b200.api import Conf, MapPool
    raw = golden_read["raw"]
    cal = (1534.14, 10.0, 8192.0)       # the example read's calibration attributes
    dac = np.round(raw.astype(np.float64) * cal[2] / cal[0] - cal[1]).astype(np.int64)
    dac16 = dac.astype(np.uint16).astype(np.int16)       # the >32767 spike wraps like the fast5 payload does
    pa = (np.float32(cal[0]) * (dac16.astype(np.uint16).astype(np.float32) + np.float32(cal[1]))) / np.float32(cal[2])
    assert np.array_equal(pa.astype(np.float32), raw)
    rl = tmp_path / "reads.txt"
    rl.write_text("keep_me\n")
    conf = Conf()
    conf.bwa_prefix, conf.read_list = example_prefix, str(rl)
    pool = MapPool(conf)
    assert pool.add_read("keep_me", dac16, channel=486, start_sample=257117, calibration=cal)
    assert not pool.add_read("drop_me", dac16, calibration=cal)
    out = pool.update()
    pool.stop()
    assert len(out) == 1 and out[0].fields()[1:] == GOLD["default"]["fields"][1:]


def test_uncalled_map_on_the_example_fast5(example_prefix, tmp_path):
    """config 1 end to end from the FILE: `uncalled map -t 1 <index> <fast5>`, `-c 1` and `-e 100` (SURVEY 8c golden
    lines), the fast5 decoded by the library's own reader, calibrated on the GPU; and a fast5 list + read filter."""
    from uncalled_b200.api import Conf, MapPool
    f5 = os.path.join(ROOT, "tests", "golden", "fast5", "example_single.fast5")
    for key, mod in (("default", {}), ("max_chunks_1", {"max_chunks": 1}), ("max_events_100", {"max_events": 100})):
        conf = Conf()
        conf.bwa_prefix = example_prefix
        for k, v in mod.items():
            setattr(conf, k, v)
        pool = MapPool(conf)
        pool.add_fast5(f5)
        assert pool.running()
        lines = []
        while pool.running():
            lines += [p.line() for p in pool.update()]
        pool.stop()
        assert len(lines) == 1 and lines[0].rsplit("\t", 1)[0] == GOLD[key]["line"], (key, lines)
    # fast5_list + read_list + max_reads over multi-read files (none of these reads map to the example reference)
    fl, rl = tmp_path / "files.txt", tmp_path / "reads.txt"
    multi = os.path.join(ROOT, "tests", "golden", "fast5", "multi_gzip.fast5")
    fl.write_text(multi + "\n" + f5 + "\n")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fast5", "golden.json")))
    ids = sorted(set(r["id"] for r in gold if r["file"] == "multi_gzip.fast5"))
    rl.write_text("\n".join(ids[:5] + ["f41a60f7-de4a-4b17-9f54-387e52d60b65"]) + "\n")
    conf = Conf()
    conf.bwa_prefix, conf.fast5_list, conf.read_list, conf.batch_reads = example_prefix, str(fl), str(rl), 4
    pool = MapPool(conf)
    out = []
    while pool.running():
        out += pool.update()
    pool.stop()
    assert sorted(p.fields()[0] for p in out) == sorted(ids[:5] + ["f41a60f7-de4a-4b17-9f54-387e52d60b65"])
    assert [p.is_mapped() for p in out if p.fields()[0].startswith("f41a60f7")] == [True]


def test_cli_index_then_map_end_to_end(tmp_path, capsys):
    """`uncalled index example_ref.fa` from the FASTA alone, then `uncalled map` of the example fast5 against it: the
    reference's golden PAF line (a rebuilt index is byte-identical to the shipped one)."""
    import orclib
    from uncalled_b200 import cli
    os.makedirs(tmp_path / "src")
    src = orclib.materialise_example_index(str(tmp_path / "src"))
    fa = str(tmp_path / "example_ref.fa")
    open(fa, "wb").write(open(src + ".fa", "rb").read())
    assert cli.main(["index", fa]) == 0
    for ext in (".bwt", ".sa", ".ann", ".amb", ".pac", ".uncl"):
        assert open(fa + ext, "rb").read() == open(src + ext, "rb").read(), ext
    capsys.readouterr()
    assert cli.main(["map", fa, os.path.join(ROOT, "tests", "golden", "fast5", "example_single.fast5")]) == 0
    cap = capsys.readouterr()
    lines = cap.out.strip().split("\n")
    assert len(lines) == 1 and lines[0].rsplit("\t", 1)[0] == GOLD["default"]["line"].replace("\n", "")
    assert cap.err.count("Mapping\n") == 1 and cap.err.count("Finishing\n") == 1


def test_multi_contig_index_built_and_mapped_on_the_gpu(U, tmp_path):
    """Three contigs with N runs: `uncalled index` entirely on this library (GPU self-alignments), then reads of every
    contig through unc_map_batch against the oracle -- rid / offsets come from bns_pos2rid on the device."""
    import orclib
    import test_multi_contig as M
    prefix, gens = M.build_multi_contig_index(str(tmp_path), None)           # None: unc_self_align on the GPU
    sigs = M.contig_reads(gens)
    idx = U.Index(prefix, device=0)
    bm = U.BatchMapper(idx, max_reads=len(sigs), max_samples=sum(len(s) for s in sigs))
    out = bm.map(np.concatenate(sigs), U.make_descs([len(s) for s in sigs]))
    O = orclib.Oracle(prefix)
    for i, s in enumerate(sigs):
        assert orclib.paf_tuple(O.map_read(s)) == U.paf_key(out[i]), i
    assert len(set(int(r["rid"]) for r in out if r["mapped"])) == 3
    bm.close()


def test_ordered_mode_equals_one_long_lived_mapper(U):
    """unc_map_batch_ordered (`uncalled map -t 1`): the whole batch in one launch plus re-mapping of the reads whose
    predecessor left flags set == the oracle's ONE Mapper mapping the reads one after the other (pinned to the
    reference's own long-lived Mapper by tests/test_oracle_pinned.py).  A small path buffer makes most reads depend
    on their predecessor; the 4.7 Mb case is the bench workload's shape."""
    import orclib
    import synth
    import synthdata
    cnt = ("n_children", "n_sources", "n_seeds", "n_clusters")
    for name, max_paths, n, L, seed, fr in (("g200k", 300, 96, 2000, 21, 0.4), ("g4m7", 10000, 48, 4000, 7, 0.15)):
        prefix, g = synthdata.get_index(name)
        idx, O = U.Index(prefix, device=0), orclib.Oracle(prefix)
        p = U._native.default_params()
        p.max_paths = O.params.max_paths = max_paths
        sig, _ = synth.reads(g, n, L, seed=seed, frac_random=fr)
        flat = np.ascontiguousarray(sig.reshape(-1), np.float32)
        lens = np.full(n, L, np.uint32)
        offs = (np.arange(n, dtype=np.uint64) * L).astype(np.uint64)
        want = O.map_reads_one_mapper(flat, offs, lens)
        d = U.make_descs([L] * n)
        bm = U.BatchMapper(idx, params=p, max_reads=n, max_samples=n * L)
        recs, carry, n_remapped, n_rounds = bm.map_ordered(flat, d)
        for i in range(n):
            assert U.paf_key(recs[i]) == orclib.paf_tuple(want[i]), (name, i)
            assert tuple(int(recs[i][k]) for k in cnt) == tuple(int(getattr(want[i], k)) for k in cnt), (name, i)
        assert n_rounds <= n and (n_remapped >= 1) == (n_rounds >= 1), (name, n_remapped, n_rounds)
        if name == "g200k":
            assert n_remapped >= 1
        t = bm.timing()
        assert t["kernel_launches"] == 5 + 4 * n_rounds and t["k2_ms"] > 0      # round 0 adds the candidate-mask kernel
        # the flags after the last read, and two batches linked by them
        prev = np.zeros(32, np.uint32)
        for i in range(n):
            _, prev = O.map_read_flags(sig[i], prev)
        assert np.array_equal(carry, prev), name
        k = n // 3
        a, c1, _, _ = bm.map_ordered(flat[:k * L], d[:k])
        b, c2, _, _ = bm.map_ordered(flat[k * L:], U.make_descs([L] * (n - k)), carry=c1)
        assert np.array_equal(np.concatenate([a, b]), recs) and np.array_equal(c2, carry), name
        # plain batch mapping is unaffected: every read from a new Mapper
        plain = bm.map(flat, d)
        fresh = O.map_batch(flat, offs, lens, threads=8)
        assert [U.paf_key(r) for r in plain] == [orclib.paf_tuple(r) for r in fresh], name
        if name == "g200k":
            assert any(U.paf_key(plain[i]) != U.paf_key(recs[i]) or int(plain[i]["n_sources"]) != int(recs[i]["n_sources"])
                       for i in range(n))
        bm.close()
        idx.close()


def test_exact_ties_kernel_equals_the_unmodified_reference_on_every_golden_read(U):
    """unc_pool_set_tie_order(1) -> k2_map_exact: the 320 reads of tests/golden/synth_paf_golden.json against the records
    the UNMODIFIED reference computed for them (its pdqsort as it is), no oracle in between; the default kernel still
    gives the stable-sort records.  Then both exact modes together against the oracle's one-Mapper chain with pdqsort."""
    import sys
    import orclib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synth_paf_golden as M
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "synth_paf_golden.json")))
    for name, n, ns, seed, frac in M.SETS:
        prefix, sig = M.signals(name, n, ns, seed, frac)
        idx = U.Index(prefix, device=0)
        bm = U.BatchMapper(idx, max_reads=n, max_samples=n * ns)
        flat, d = np.ascontiguousarray(sig.reshape(-1)), U.make_descs([ns] * n)
        bm.set_tie_order(1)
        exact = bm.map(flat, d)
        assert int((exact["status"] != 0).sum()) == 0
        assert [[int(v) for v in U.paf_key(r)] for r in exact] == gold["reference"][name], name
        bm.set_tie_order(0)
        plain = bm.map(flat, d)
        assert [[int(v) for v in U.paf_key(r)] for r in plain] == gold["reference_stable_sort"][name], name
        if name == "g4m7":        # exact ties + ordered = the unmodified reference with one long-lived Mapper (`-t 1`)
            k = 40
            O = orclib.Oracle(prefix)
            O.lib.orc_set_child_sort(1)
            try:
                want = O.map_reads_one_mapper(flat[:k * ns], (np.arange(k, dtype=np.uint64) * ns).astype(np.uint64),
                                              np.full(k, ns, np.uint32))
            finally:
                O.lib.orc_set_child_sort(0)
            bm.set_tie_order(1)
            recs, _, _, _ = bm.map_ordered(flat[:k * ns], d[:k])
            assert [U.paf_key(r) for r in recs] == [orclib.paf_tuple(w) for w in want]
        bm.close()
        idx.close()


def test_stream_with_exact_ties(U):
    """unc_stream_set_tie_order(1) -> k2_map_stream_exact against the streaming oracle in pdqsort mode."""
    import orclib
    import synth
    import synthdata
    import test_gpu_parity as TG
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 12, 6000, seed=15, frac_random=0.3)
    sigs = [sig[i][:6000 - 53 * i] for i in range(12)]
    lib = orclib.orc()
    lib.orc_set_child_sort(1)
    try:
        st = TG._stream_vs_oracle(U, prefix, sigs, 4, 450, tie_order=1)
    finally:
        lib.orc_set_child_sort(0)
    assert (2, 0) in st


@pytest.mark.gpu
def test_stream_chunk_timeout_fails_the_reads_still_mapping():
    """Mapper::PRMS.chunk_timeout (reference src/mapper.cpp:40,384-390): a read whose chunk stayed with the mapper longer
    than the limit fails and is marked ended.  With an impossible limit every read that does not map within its first
    chunk ends FAILURE + ended after that step; without a limit the same reads go on."""
    import synth
    import synthdata
    import uncalled_b200 as U
    from uncalled_b200 import stream as S
    prefix, g = synthdata.get_index("g200k")
    sig, _ = synth.reads(g, 6, 4000, seed=3, frac_random=1.0)          # random reads: nothing maps in one chunk
    idx = U.Index(prefix, device=0)
    sm = U.StreamMapper(idx, 6, 450)
    free = sm.map_reads([sig[i] for i in range(6)])
    assert all(r is not None and r[2] > 1 for r in free)               # several chunks consumed
    assert sm.last_step_ms() > 0
    sm.set_chunk_timeout(1e-6)
    cut = sm.map_reads([sig[i] for i in range(6)])
    assert all(r is not None and r[0] == S.FAILURE and r[1] == 1 and r[2] == 1 for r in cut)
    sm.close()
