#!/usr/bin/env python
"""Generates tests/golden/stream_golden.json: results of the reference's OWN streaming path
(oracle/_ref: Mapper::new_read(Chunk&) / process_chunk / map_chunk / add_chunk driven chunk by chunk,
see oracle/ref_build/ref_shim.cpp ref_stream_read) on the example read and on seeded synthetic reads.
Needs oracle/_ref (built from /root/reference by `make -C oracle ref`).  Run twice: once per index,
because the reference keeps its FM index in process-global statics.
    python tools/make_stream_golden.py example ; python tools/make_stream_golden.py g200k"""
import ctypes as C
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402
import orclib  # noqa: E402
import synth  # noqa: E402
import synthdata  # noqa: E402

CONFIGS = [(1.0, 1000000), (0.1125, 1000000), (0.1125, 5), (0.05, 12)]   # (chunk_time s, max_chunks)


def signals(which):
    if which == "example":
        raw = np.load(os.path.join(ROOT, "tests", "golden", "example_read.npz"))["raw"]
        return orclib.materialise_example_index(tempfile.mkdtemp()), [raw, raw[:4000], raw[4000:20000]]
    prefix, g = synthdata.get_index(which)
    sig, _ = synth.reads(g, 24, 12000, seed=5, frac_random=0.3)
    return prefix, [sig[i] for i in range(len(sig))]


def main():
    which = sys.argv[1]
    prefix, sigs = signals(which)
    R = orclib.ref()
    R.ref_load(prefix.encode(), b"default")
    rows = []
    for i, s in enumerate(sigs):
        s = np.ascontiguousarray(s, np.float32)
        for ct, mc in CONFIGS:
            out, nu, en = orclib.RefPaf(), C.c_uint32(), C.c_int32()
            R.ref_stream_read(orclib.fp(s), len(s), ct, mc, C.byref(out), C.byref(nu), C.byref(en))
            rows.append({"read": i, "chunk_time": ct, "max_chunks": mc, "paf": list(orclib.paf_tuple(out)),
                         "chunks": nu.value, "ended": en.value})
    path = os.path.join(ROOT, "tests", "golden", "stream_golden.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[which] = rows
    json.dump(d, open(path, "w"), indent=0)
    print(which, len(rows), "rows")


if __name__ == "__main__":
    main()
