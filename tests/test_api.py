"""The reference-shaped Python surface (uncalled_b200/api.py: Conf, Paf, MapPool) against the
golden PAF lines of the real `uncalled map` (tests/golden/example_paf.json)."""
import io
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "example_paf.json")))


def test_paf_formats_like_the_reference():
    from uncalled_b200.api import Paf
    f = GOLD["default"]["fields"]
    p = Paf(f[0], 486, 257117)
    p.set_read_len(106)
    p.set_mapped(73, 106, f[5], 6938, 6976, 10000, False, 38)
    assert p.line() == GOLD["default"]["line"]
    p.set_float(Paf.Tag.MAP_TIME, 12.5)
    assert p.line() == GOLD["default"]["line"] + "\tmt:f:12.500000"
    u = Paf("r1", 3, 9)
    u.set_read_len(3562)
    assert u.line() == "r1\t3562" + "\t*" * 9 + "\t255\tch:i:3\tst:i:9"
    assert not u.is_mapped() and p.is_mapped() and not p.is_ended()
    assert Paf.MAP_TIME == Paf.Tag.MAP_TIME and Paf.KEEP == 11      # export_values()
    buf = io.StringIO()
    u.print_paf(buf)
    assert buf.getvalue().endswith("st:i:9\n")


def test_conf_has_real_properties_with_docs():
    from uncalled_b200.api import Conf
    for name in ("read_list", "max_reads", "fast5_list", "bwa_prefix", "max_events", "max_chunks", "threads"):
        assert isinstance(getattr(Conf, name), property) and getattr(Conf, name).__doc__
    c = Conf()
    c.max_events = "100"                       # args.load_conf assigns parsed values (uncalled/args.py:296-302)
    assert c.max_events == 100 and c.chunk_time == 1.0 and c.idx_preset == "default"


def test_realtime_pool_surface_and_constants():
    from uncalled_b200.api import Chunk, Conf, RealtimePool
    assert (RealtimePool.DEPLETE, RealtimePool.ENRICH) == (0, 1) and (RealtimePool.FULL, RealtimePool.EVEN, RealtimePool.ODD) == (0, 1, 2)
    for name in ("host", "port", "duration"):                 # uncalled/args.py:223-260 reads these docstrings
        assert getattr(Conf, name).__doc__
    c = Chunk("r", 3, 7, 100, np.arange(10, dtype=np.float32), 2, 100)   # clipped like Chunk::Chunk (chunk.cpp:74-83)
    assert (c.channel, c.number, c.size(), c.empty()) == (3, 7, 8, False)
    assert np.array_equal(c.pop(), np.arange(2, 10, dtype=np.float32)) and c.empty()


def test_realtime_pool_chunk_protocol_with_emulated_device():
    """RealtimePool driven like the reference's simulator: try_add_chunk per channel with the next chunk (an empty
    one when the read has no more signal), update() for the finished reads -- on the EMULATED device code, against
    the oracle's streaming restatement (reads following each other on a channel share its Mapper)."""
    import emulib
    import orclib
    import synth
    import synthdata
    from uncalled_b200.api import Chunk, Conf, RealtimePool
    prefix, g = synthdata.get_index("g200k")
    E, O = emulib.Emu(prefix), orclib.Oracle(prefix)
    conf = Conf()
    conf.num_channels, conf.chunk_time = 2, 0.1125
    backend = emulib.EmuStream(E, 2, 450)

    class _Idx:
        seqs = [(O.lib.orc_seq_name(O.idx, i).decode(), int(O.lib.orc_seq_len(O.idx, i))) for i in range(O.lib.orc_n_seqs(O.idx))]
    pool = RealtimePool(conf, backend=backend, index=_Idx)
    sig, _ = synth.reads(g, 4, 4000, seed=21, frac_random=0.25)
    reads = {0: [0, 2], 1: [1, 3]}                       # channel -> reads, one after the other
    pos = {0: [0, 0], 1: [0, 0]}                         # [index into reads[ch], next chunk]
    results = {}
    for _ in range(200):
        for ch in (0, 1):
            if pos[ch][0] >= len(reads[ch]):
                continue
            i, k = reads[ch][pos[ch][0]], pos[ch][1]
            nfull = len(sig[i]) // 450
            c = Chunk("read%d" % i, ch + 1, i + 1, k * 450, sig[i], k * 450, 450) if k < nfull else Chunk("read%d" % i, ch + 1, i + 1, 0, [])
            if pool.try_add_chunk(c) or c.empty():
                pos[ch][1] += 1
        for channel, number, paf in pool.update():
            results[number - 1] = paf
            ch = channel - 1
            pos[ch] = [pos[ch][0] + 1, 0]
        if len(results) == 4:
            break
    assert pool.all_finished() and len(results) == 4
    pool.stop_all()
    for ch in (0, 1):
        want = O.stream_channel([sig[i] for i in reads[ch]], 450)
        for i, (rec, nu, en) in zip(reads[ch], want):
            p = results[i]
            assert p.is_mapped() == bool(rec.mapped) and p.is_ended() == bool(en), i
            assert p.rd_len == rec.rd_len, i
            if rec.mapped:
                assert (p.rd_st, p.rd_en, p.rf_st, p.rf_en, p.matches, p.fwd) == \
                       (rec.rd_st, rec.rd_en, rec.rf_st, rec.rf_en, rec.matches, bool(rec.fwd)), i
                assert p.fields()[0] == "read%d" % i and p.int_tags[0] == (4, ch + 1)


def test_cli_parser_and_fast5_discovery(tmp_path, capsys):
    """`uncalled map` / `uncalled index` option names and defaults (uncalled/args.py:87-161,218-286) and the fast5
    path expansion of scripts/uncalled:80-118."""
    from uncalled_b200 import cli
    f5dir = os.path.join(ROOT, "tests", "golden", "fast5")
    _, conf, args = cli.load_conf(["map", "-t", "3", "-c", "1", "-e", "100", "-n", "7", "-l", "reads.txt", "-p", "fast",
                                   "idx/ecoli", f5dir, "-r"])
    assert (conf.bwa_prefix, conf.threads, conf.max_chunks, conf.max_events, conf.max_reads, conf.read_list, conf.idx_preset) == \
        ("idx/ecoli", 3, 1, 100, 7, "reads.txt", "fast")
    assert args.recursive and conf.chunk_time == 1.0
    found = sorted(os.path.basename(p) for p in cli.load_fast5s([f5dir], False) if p)
    assert found == sorted(f for f in os.listdir(f5dir) if f.endswith(".fast5")) and len(found) == 8
    lst = tmp_path / "list.txt"
    lst.write_text(os.path.join(f5dir, "multi_gzip.fast5") + "\n#comment.fast5\nnot_a_fast5.txt\n" + str(tmp_path / "gone.fast5") + "\n")
    assert [os.path.basename(p) for p in cli.load_fast5s([str(lst)], False) if p] == ["multi_gzip.fast5"]
    assert "is not a fast5 file" in capsys.readouterr().err
    _, _, a = cli.load_conf(["index", "ref.fa", "-o", "out/ref", "--probs", "0.5,0.2", "-s", "50", "-1", "0.55"])
    assert (a.fasta_filename, a.bwa_prefix, a.probs, a.speeds, a.max_sample_dist, a.matchpr1, a.matchpr2, a.min_samples) == \
        ("ref.fa", "out/ref", "0.5,0.2", None, 50, 0.55, 0.9838, 50000)
    with pytest.raises(SystemExit):
        list(cli.load_fast5s([str(tmp_path / "absent_dir")], False))


class _EmuBackend:
    """The emulated device code standing in for the GPU BatchMapper (test vehicle): maps each read with the same
    device source compiled for the CPU (tests/emulib.py)."""

    def __init__(self, E):
        self.E, self.calls = E, []

    def map(self, flat, descs):
        self.calls.append(len(descs))
        out = []
        for d in descs:
            s = flat[int(d["offset"]):int(d["offset"]) + int(d["n_samples"])]
            cal = (float(d["cal_range"]), float(d["cal_offset"]), float(d["cal_digit"]))
            out.append(self.E.map_batch([s], dtype=int(d["dtype"]), cal=cal)[0][0])
        return out


class _EmuBackendOrdered(_EmuBackend):
    """... with unc_map_batch_ordered's counterpart (the product's ordered-mode host logic over the emulated kernels)."""

    def map_ordered(self, flat, descs, carry=None, on_device=False):
        self.calls.append(len(descs))
        assert all(int(d["dtype"]) == int(descs[0]["dtype"]) for d in descs)
        sigs = []
        for d in descs:
            s = flat[int(d["offset"]):int(d["offset"]) + int(d["n_samples"])]
            if int(d["dtype"]) == 1:        # the emulator entry takes pA: calibrate as src/read_buffer.cpp:239-242
                rng, off, dig = (np.float32(d[k]) for k in ("cal_range", "cal_offset", "cal_digit"))
                s = (rng * (s.view(np.uint16).astype(np.float32) + off) / dig).astype(np.float32)
            sigs.append(np.ascontiguousarray(s, np.float32))
        return self.E.map_ordered(sigs, carry=carry)


class _EmuBackendTwoCall(_EmuBackend):
    """The same with the submit()/wait() form of the GPU mapper, so that MapPool pipelines its batches."""

    def __init__(self, E):
        _EmuBackend.__init__(self, E)
        self.fifo, self.max_in_flight = [], 0

    def submit(self, flat, descs):
        self.fifo.append(self.map(flat, descs))
        self.max_in_flight = max(self.max_in_flight, len(self.fifo))

    def wait(self):
        return self.fifo.pop(0)


def test_uncalled_map_from_fast5_files_on_the_emulated_device(tmp_path):
    """`uncalled map` end to end WITHOUT a GPU: fast5 files -> the library's reader -> MapPool (queueing, read
    filter, max_reads, max_chunks truncation, decode prefetch) -> the device source under the CPU emulator -> PAF.
    The example fast5 must give the reference's golden lines for `-c 1` and `-e 100`."""
    import emulib
    import orclib
    from uncalled_b200.api import Conf, MapPool
    prefix = orclib.materialise_example_index(str(tmp_path))
    O = orclib.Oracle(prefix)

    class _Idx:
        seqs = [(O.lib.orc_seq_name(O.idx, i).decode(), int(O.lib.orc_seq_len(O.idx, i))) for i in range(O.lib.orc_n_seqs(O.idx))]
    f5dir = os.path.join(ROOT, "tests", "golden", "fast5")
    ex, multi = os.path.join(f5dir, "example_single.fast5"), os.path.join(f5dir, "multi_many_reads.fast5")
    for key, mod in (("max_chunks_1", {"max_chunks": 1}), ("max_events_100", {"max_events": 100})):
        E = emulib.Emu(prefix)
        conf = Conf()
        for k, v in mod.items():
            setattr(conf, k, v)
            if hasattr(E.params, k):
                setattr(E.params, k, v)
        pool = MapPool(conf, backend=_EmuBackend(E), index=_Idx)
        pool.add_fast5(ex)
        assert pool.running()
        lines = []
        while pool.running():
            lines += [p.line() for p in pool.update()]
        pool.stop()
        assert len(lines) == 1 and lines[0].rsplit("\t", 1)[0] == GOLD[key]["line"], (key, lines)
    # batches of 16 reads over two files (70 tiny reads + the example), decode of batch k+1 overlapping batch k;
    # a read list and max_reads applied the way Fast5Reader does
    gold = json.load(open(os.path.join(f5dir, "golden.json")))
    ids = sorted(set(r["id"] for r in gold if r["file"] == "multi_many_reads.fast5"))
    E = emulib.Emu(prefix)
    conf = Conf()
    conf.batch_reads, conf.max_chunks, conf.threads = 16, 1, 2
    for be in (_EmuBackend(E), _EmuBackendTwoCall(E)):
        pool = MapPool(conf, backend=be, index=_Idx)
        pool.add_fast5(multi)
        pool.add_fast5(ex)
        out, per_call = [], []
        while pool.running():
            got = pool.update()
            per_call.append(len(got))
            out += got
        pool.stop()
        assert sorted(p.fields()[0] for p in out) == sorted(ids + ["f41a60f7-de4a-4b17-9f54-387e52d60b65"])
        assert be.calls == [16, 16, 16, 16, 7] and not pool.running()
        assert [p.line().rsplit("\t", 1)[0] for p in out if p.is_mapped()] == [GOLD["max_chunks_1"]["line"]]
        if isinstance(be, _EmuBackendTwoCall):      # batch k+1 was submitted before batch k was collected
            assert be.max_in_flight == 2 and per_call == [0, 16, 16, 16, 16 + 7] and not be.fifo
    rl = tmp_path / "reads.txt"
    rl.write_text("\n".join(ids[10:40]) + "\n")
    conf = Conf()
    conf.batch_reads, conf.read_list, conf.max_reads = 8, str(rl), 12
    be = _EmuBackend(E)
    pool = MapPool(conf, backend=be, index=_Idx)
    pool.add_fast5(multi)
    out = []
    while pool.running():
        out += pool.update()
    pool.stop()
    assert len(out) == 12 and set(p.fields()[0] for p in out) <= set(ids[10:40]) and sum(be.calls) == 12
    # an unreadable file surfaces as an error from update(), like the reference's reader throwing in fill_buffer
    bad = tmp_path / "bad.fast5"
    bad.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 100)
    pool = MapPool(Conf(), backend=_EmuBackend(E), index=_Idx)
    pool.add_fast5(str(bad))
    with pytest.raises(RuntimeError):
        pool.update()
    pool.stop()


def test_ordered_conf_keeps_input_order_and_threads_the_carry():
    """conf.ordered: batches are not re-sorted, go through map_ordered one after the other, and each batch starts
    from the flags the previous one ended with (the kernels behind it: tests/test_ordered_emul.py)."""
    import uncalled_b200._native as N
    from uncalled_b200.api import Conf, MapPool

    class _Idx:
        seqs = [("chr", 1000)]

    class _Stub:
        def __init__(self):
            self.calls = []

        def map_ordered(self, flat, descs, carry=None, on_device=False):
            self.calls.append(([int(x) for x in descs["n_samples"]], None if carry is None else carry.copy()))
            out = np.zeros(len(descs), dtype=N.PAF_DTYPE)
            nxt = np.full(32, len(self.calls), np.uint32)
            return out, nxt, 0, 0

        def map(self, flat, descs):
            raise AssertionError("ordered mode must not use the plain batch call")

    conf = Conf()
    conf.ordered, conf.batch_reads = 1, 3
    be = _Stub()
    pool = MapPool(conf, backend=be, index=_Idx)
    lens = [50, 400, 30, 200, 10]
    for i, n in enumerate(lens):
        pool.add_read("r%d" % i, np.zeros(n, np.float32))
    out = []
    while pool.running():
        out += pool.update()
    pool.stop()
    assert [p.fields()[0] for p in out] == ["r%d" % i for i in range(5)]
    assert [c[0] for c in be.calls] == [[50, 400, 30], [200, 10]]
    assert be.calls[0][1] is None and np.array_equal(be.calls[1][1], np.full(32, 1, np.uint32))
    from uncalled_b200 import cli
    _, c2, _ = cli.load_conf(["map", "--ordered", "idx", "reads.fast5"])
    _, c3, _ = cli.load_conf(["map", "idx", "reads.fast5"])
    assert c2.ordered == 1 and c3.ordered == 0
    _, c4, _ = cli.load_conf(["map", "--exact-ties", "--ordered", "idx", "reads.fast5"])
    assert (c4.exact_ties, c4.ordered, c3.exact_ties) == (1, 1, 0)


def test_uncalled_map_ordered_from_fast5_files_on_the_emulated_device(tmp_path):
    """`uncalled map --ordered -c 1` end to end without a GPU: reads come out in file order, batch after batch through the
    ordered-mode entry point with the flag carry threaded, and the example read gives the reference's golden `-c 1` line
    (the reference's own `-t 1` run is ordered by construction)."""
    import emulib
    import orclib
    from uncalled_b200.api import Conf, MapPool
    prefix = orclib.materialise_example_index(str(tmp_path))
    O = orclib.Oracle(prefix)

    class _Idx:
        seqs = [(O.lib.orc_seq_name(O.idx, i).decode(), int(O.lib.orc_seq_len(O.idx, i))) for i in range(O.lib.orc_n_seqs(O.idx))]
    f5dir = os.path.join(ROOT, "tests", "golden", "fast5")
    ex, multi = os.path.join(f5dir, "example_single.fast5"), os.path.join(f5dir, "multi_many_reads.fast5")
    gold = json.load(open(os.path.join(f5dir, "golden.json")))
    file_order = list(dict.fromkeys(r["id"] for r in gold if r["file"] == "multi_many_reads.fast5"))   # two rows per read
    E = emulib.Emu(prefix)
    conf = Conf()
    conf.batch_reads, conf.max_chunks, conf.ordered = 32, 1, 1
    be = _EmuBackendOrdered(E)
    pool = MapPool(conf, backend=be, index=_Idx)
    pool.add_fast5(multi)
    pool.add_fast5(ex)
    out = []
    while pool.running():
        out += pool.update()
    pool.stop()
    ids = [p.fields()[0] for p in out]
    assert ids[-1] == "f41a60f7-de4a-4b17-9f54-387e52d60b65" and len(ids) == len(file_order) + 1
    assert sorted(ids[:-1]) == sorted(file_order)
    assert ids[:-1] == file_order                          # input order, not longest-first
    assert [p.line().rsplit("\t", 1)[0] for p in out if p.is_mapped()] == [GOLD["max_chunks_1"]["line"]]
    assert sum(be.calls) == len(ids) and max(be.calls) <= 32
