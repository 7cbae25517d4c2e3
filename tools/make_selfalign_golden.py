#!/usr/bin/env python
"""Generates tests/golden/self_align_golden.json: digests of what the reference's OWN `self_align`
(src/self_align_ref.cpp, through oracle/_ref) returns for the shipped example index and for seeded
synthetic indexes -- the input of `uncalled index`'s parameter search.  Needs oracle/_ref.
    python tools/make_selfalign_golden.py"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402
import orclib  # noqa: E402
import synthdata  # noqa: E402

CASES = [("example", 1), ("example", 3), ("g200k", 4), ("g200k", 37), ("g1m", 20)]   # (index, sample_dist)


def prefix_of(which):
    if which == "example":
        return orclib.materialise_example_index(tempfile.mkdtemp())
    return synthdata.get_index(which)[0]


def digest(off, val):
    return {"n_paths": int(len(off) - 1), "n_values": int(len(val)),
            "offsets_sha256": hashlib.sha256(np.ascontiguousarray(off, "<u8").tobytes()).hexdigest(),
            "values_sha256": hashlib.sha256(np.ascontiguousarray(val, "<u8").tobytes()).hexdigest(),
            "head": [[int(v) for v in val[int(off[i]):int(off[i + 1])]] for i in range(3)]}


def main():
    rows = []
    for which, sd in CASES:
        off, val = orclib.ref_self_align(prefix_of(which), sd)
        rows.append(dict(index=which, sample_dist=sd, **digest(off, val)))
        print(which, sd, rows[-1]["n_paths"], rows[-1]["n_values"])
    json.dump(rows, open(os.path.join(ROOT, "tests", "golden", "self_align_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
