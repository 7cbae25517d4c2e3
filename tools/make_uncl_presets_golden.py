#!/usr/bin/env python
"""Generates tests/golden/uncl_presets.json by running the reference's OWN Python parameter search
(/root/reference/uncalled/index.py, IndexParameterizer, imported from where it lies) the way
`uncalled index` drives it (scripts/uncalled:38-78), with `unc.self_align` served by oracle/_ref (the
reference's own C++ self_align).  The compiled `_uncalled` module is not needed for anything else here, so
a stand-in module named `uncalled` offers just that one function.
    python tools/make_uncl_presets_golden.py"""
import argparse
import importlib.util
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import orclib  # noqa: E402
import synthdata  # noqa: E402

# (index, probs, speeds, overrides of the `uncalled index` defaults)
CASES = [
    ("example", "0.5,0.2,0.9", "60,200,115", {}),
    ("g200k", "0.4,0.05", "80,150.5", {}),
    ("g200k", "0.3", "100", {"matchpr1": 0.55, "matchpr2": 0.99, "pathlen_percentile": 0.1, "max_replen": 60}),
    ("g1m", "0.25", "30,400", {"max_sample_dist": 50, "min_samples": 20000}),
]
DEFAULTS = dict(max_sample_dist=100, min_samples=50000, max_samples=1000000, kmer_len=5, matchpr1=0.6334,
                matchpr2=0.9838, pathlen_percentile=0.05, max_replen=100)


def load_reference_index_module():
    stub = types.ModuleType("uncalled")

    def self_align(prefix, sample_dist):
        off, val = orclib.ref_self_align(prefix, sample_dist)
        return [[int(v) for v in val[int(off[i]):int(off[i + 1])]] for i in range(len(off) - 1)]
    stub.self_align = self_align
    sys.modules["uncalled"] = stub
    spec = importlib.util.spec_from_file_location("uncalled_reference_index", "/root/reference/uncalled/index.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref_index = load_reference_index_module()
    rows = []
    for which, probs, speeds, over in CASES:
        if which == "example":
            prefix = orclib.materialise_example_index(tempfile.mkdtemp())
        else:
            src = synthdata.get_index(which)[0]
            d = tempfile.mkdtemp()
            prefix = os.path.join(d, which)
            for e in (".bwt", ".sa", ".ann", ".amb", ".pac"):
                os.symlink(src + e, prefix + e)
        args = argparse.Namespace(bwa_prefix=prefix, **dict(DEFAULTS, **over))
        p = ref_index.IndexParameterizer(args)
        p.add_preset("default", tgt_speed=115)                    # scripts/uncalled:57-76
        for t in probs.split(","):
            try:
                p.add_preset("prob_%s" % t, tgt_prob=float(t))
            except Exception:
                pass
        for t in speeds.split(","):
            try:
                p.add_preset("speed_%s" % t, tgt_speed=float(t))
            except Exception:
                pass
        p.write()
        rows.append({"index": which, "probs": probs, "speeds": speeds, "opts": over,
                     "uncl": open(prefix + ".uncl").read()})
        print(which, over, "\n" + rows[-1]["uncl"])
    json.dump(rows, open(os.path.join(ROOT, "tests", "golden", "uncl_presets.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
