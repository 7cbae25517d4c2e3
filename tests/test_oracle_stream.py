"""The oracle's restatement of the STREAMING path (EventProfiler, streaming Normalizer, chunked
process_chunk / map_chunk) against (a) committed results of the reference's own streaming path
(tests/golden/stream_golden.json, made by tools/make_stream_golden.py from oracle/_ref) and
(b) oracle/_ref itself when it is present."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_stream_golden as msg  # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "stream_golden.json")))


def _chunk_len(chunk_time):
    return int(np.float32(chunk_time) * np.float32(4000.0))      # ReadBuffer::PRMS.chunk_len(): u16(chunk_time * sample_rate)


@pytest.mark.parametrize("which", ["example", "g200k"])
def test_stream_port_matches_reference_golden(which):
    import orclib
    prefix, sigs = msg.signals(which)
    O = orclib.Oracle(prefix)
    for row in GOLD[which]:
        rec, nu, en = O.stream_read(sigs[row["read"]], _chunk_len(row["chunk_time"]), row["max_chunks"])
        assert list(orclib.paf_tuple(rec)) == row["paf"] and nu == row["chunks"] and en == row["ended"], row


def test_example_read_streams_like_uncalled_map_ord():
    """`uncalled_map_ord` on the example read prints ... 67 41 67 - ... 10000 6948 6977 29 30 255."""
    row = [r for r in GOLD["example"] if r["read"] == 0 and r["chunk_time"] == 1.0][0]
    p = row["paf"]
    assert p[0] == 1 and p[1] == 0 and p[6:11] == [67, 41, 67, 6948, 6977] and p[5] == 29


LIVE = r"""
import ctypes as C, sys
sys.path[:0] = [%r, %r]
import numpy as np, orclib, synth, synthdata
prefix, g = synthdata.get_index("g200k")
R = orclib.ref()
assert R.ref_load(prefix.encode(), b"default") == 0
O = orclib.Oracle(prefix)
sig, _ = synth.reads(g, 10, 9000, seed=77, frac_random=0.3)
for i in range(10):
    s = np.ascontiguousarray(sig[i], np.float32)
    for ct, mc in ((0.1125, 1000000), (0.1125, 7), (0.25, 3)):
        out, nu, en = orclib.RefPaf(), C.c_uint32(), C.c_int32()
        R.ref_stream_read(orclib.fp(s), len(s), ct, mc, C.byref(out), C.byref(nu), C.byref(en))
        rec, nu2, en2 = O.stream_read(s, int(np.float32(ct) * np.float32(4000.0)), mc)
        assert (orclib.paf_tuple(out), nu.value, en.value) == (orclib.paf_tuple(rec), nu2, en2), (i, ct, mc)
print("LIVE-OK")
"""


def test_stream_port_matches_ref_library_live():
    """bit-for-bit against oracle/_ref on fresh reads; in its own process because the reference keeps
    its FM index in process-global statics (one index per process)."""
    import subprocess
    import orclib
    if not orclib.ref_available():
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, "-c", LIVE % (os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"))],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "LIVE-OK" in r.stdout, r.stdout + r.stderr
