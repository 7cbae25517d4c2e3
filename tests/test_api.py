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


@pytest.mark.gpu
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


@pytest.mark.gpu
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
