"""The library's own fast5 (HDF5 subset) reader against what the REFERENCE's Fast5Reader + ReadBuffer, built with
its vendored libhdf5, deliver for the same files (tests/golden/fast5/golden.json, made by
tools/fast5_fixtures/build.sh): the file the reference ships in example/, and files written with the real libhdf5
that cover vlen / fixed strings, big-endian and 32-bit attributes, gzip / shuffle / fletcher32 / unfiltered chunks,
contiguous and compact layouts, multi-level group and chunk B-trees, header continuation blocks, superblock v2 with
v2 object headers, empty signals, start_time beyond 32 bits, negative DAC values."""
import json
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F5DIR = os.path.join(ROOT, "tests", "golden", "fast5")
GOLD = json.load(open(os.path.join(F5DIR, "golden.json")))
FILES = sorted(set(r["file"] for r in GOLD))


def _rows(fname, max_chunks):
    return {r["id"]: r for r in GOLD if r["file"] == fname and r["max_chunks"] == max_chunks}


@pytest.mark.parametrize("fname", FILES)
@pytest.mark.parametrize("max_chunks", [1000000, 2])
def test_reads_match_reference_reader(fname, max_chunks):
    from uncalled_b200.fast5 import Fast5File
    want = _rows(fname, max_chunks)
    limit = 0 if max_chunks == 1000000 else max_chunks * 4000       # ReadBuffer::PRMS: chunk_time 1.0 s x 4000 Hz
    with Fast5File(os.path.join(F5DIR, fname)) as f:
        assert f.single_read_format == ("single" in fname)
        got = f.load(max_samples_per_read=limit, threads=3)
    assert sorted(r.read_id for r in got) == sorted(want)
    for r in got:
        w = want[r.read_id]
        pa = r.pa()
        assert r.number == w["number"] and r.channel == w["channel"] and len(pa) == w["n"], r.read_id
        assert r.start_sample & 0xFFFFFFFF == w["start"] & 0xFFFFFFFF        # atoi: the low 32 bits, sign-extended there
        assert zlib.crc32(pa.tobytes()) & 0xFFFFFFFF == w["crc32_f32"], r.read_id
        assert [float(np.float32(x)) for x in w["head"]] == [float(x) for x in pa[:4]]


def test_example_fast5_known_answers():
    """31 668 samples, range 1534.14 / offset 10 / digitisation 8192, first pA values, raw f32 SHA-1."""
    import hashlib
    from uncalled_b200.fast5 import Fast5File
    with Fast5File(os.path.join(F5DIR, "example_single.fast5")) as f:
        assert len(f) == 1
        i = f.info(0)
        assert (i.read_id, i.number, i.start_sample, i.channel) == ("f41a60f7-de4a-4b17-9f54-387e52d60b65", 101, 257117, 486)
        r = f.load()[0]
    assert r.calibration == (np.float32(1534.14), 10.0, 8192.0) and len(r.signal) == 31668
    pa = r.pa()
    assert np.allclose(pa[:4], [140.45471191, 94.01102448, 93.63647461, 94.19829559], rtol=0, atol=1e-6)
    assert hashlib.sha1(pa.tobytes()).hexdigest() == "ebf1855f152585c6b95990aff540d5f274b6c36f"
    gold = np.load(os.path.join(ROOT, "tests", "golden", "example_read.npz"))["raw"]
    assert np.array_equal(pa, gold)


def test_ranges_threads_and_small_buffers():
    from uncalled_b200.fast5 import Fast5File
    with Fast5File(os.path.join(F5DIR, "multi_gzip.fast5")) as f:
        full = f.load(threads=1)
        part = f.load(3, 5, threads=8)
        assert [r.read_id for r in part] == [r.read_id for r in full[3:8]]
        assert all(np.array_equal(a.signal, b.signal) for a, b in zip(part, full[3:8]))
        cut = f.load(0, 4, max_samples_per_read=1234)
        assert all(np.array_equal(c.signal, r.signal[:1234]) for c, r in zip(cut, full))
        assert f.load(2, 0) == []


FUZZ = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
from uncalled_b200.fast5 import Fast5File, Fast5Error
src, tmp = sys.argv[1], sys.argv[2]
data = np.fromfile(src, np.uint8)
rng = np.random.default_rng(len(data))
ok = err = 0
for it in range(int(sys.argv[3])):
    d = data.copy()
    if it %% 3 == 0:
        d = d[:rng.integers(0, len(d))]                              # truncated
    else:
        for _ in range(int(rng.integers(1, 6))):                     # a few corrupted bytes, mostly in the metadata
            pos = int(rng.integers(0, min(len(d), 60000)))
            d[pos] = rng.integers(0, 256)
    d.tofile(tmp)
    try:
        with Fast5File(tmp) as f:
            f.load(max_samples_per_read=20000, threads=2)
        ok += 1
    except Fast5Error:
        err += 1
print("FUZZ-DONE", ok, err)
"""


@pytest.mark.parametrize("fname", ["multi_gzip.fast5", "multi_latest.fast5", "multi_contig.fast5"])
def test_corrupt_files_give_errors_not_crashes(fname, tmp_path):
    """Truncated and bit-flipped files: every outcome is a result or a Fast5Error -- no crash, no hang."""
    r = subprocess.run([sys.executable, "-c", FUZZ % ROOT, os.path.join(F5DIR, fname), str(tmp_path / "x.fast5"), "150"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FUZZ-DONE" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_not_hdf5_and_missing_file(tmp_path):
    from uncalled_b200.fast5 import Fast5File, Fast5Error
    p = tmp_path / "x.fast5"
    p.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(Fast5Error):
        Fast5File(str(p))
    with pytest.raises(Fast5Error):
        Fast5File(str(tmp_path / "absent.fast5"))
