#!/usr/bin/env python
"""Generates tests/golden/synth_paf_golden.json: PAF records that the REFERENCE's own code (oracle/_ref, built from
/root/reference) computes for seeded synthetic reads, fresh Mapper per read -- once as it is (pdqsort) and once with its
child sort made stable (oracle/ref_build/stubs_stable).  The GPU parity tests compare the CUDA path with the stable-sort
records directly (no oracle in between) and list the reads where the unmodified reference differs.
One build per process (the reference keeps its index in process-global statics):
    python tools/make_synth_paf_golden.py            # runs itself twice"""
import ctypes as C
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), ROOT]
SETS = [("g200k", 120, 3000, 31, 0.25), ("g4m7", 200, 4000, 123, 0.15)]      # index, reads, samples, seed, frac_random
OUT = os.path.join(ROOT, "tests", "golden", "synth_paf_golden.json")


def signals(name, n, n_samples, seed, frac):
    import synth
    import synthdata
    prefix, g = synthdata.get_index(name)
    total = 2400 if name == "g4m7" else n        # the g4m7 set is the head of the 2400-read set DESIGN.md quotes
    sig, _ = synth.reads(g, total, n_samples, seed=seed, frac_random=frac)
    return prefix, sig[:n]


def one_build(stable):
    import numpy as np
    import orclib
    rows = {}
    for name, n, ns, seed, frac in SETS:
        code = ("import sys, json, ctypes as C; sys.path[:0]=%r\n"
                "from concurrent.futures import ThreadPoolExecutor\n"
                "import numpy as np, orclib, make_synth_paf_golden as M\n"
                "prefix, sig = M.signals(%r, %d, %d, %d, %r)\n"
                "R = orclib.ref(stable_sort=%r); assert R.ref_load(prefix.encode(), b'default') == 0\n"
                "def one(i):\n"
                "    s = np.ascontiguousarray(sig[i], np.float32); out = orclib.RefPaf()\n"
                "    R.ref_map_read(orclib.fp(s), len(s), C.byref(out)); return [int(v) for v in orclib.paf_tuple(out)]\n"
                "with ThreadPoolExecutor(8) as ex: print(json.dumps(list(ex.map(one, range(len(sig))))))\n"
                ) % (sys.path[:3], name, n, ns, seed, frac, stable)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
        rows[name] = json.loads(r.stdout.strip().splitlines()[-1])
    return rows


def main():
    ref, stable = one_build(False), one_build(True)
    out = {"sets": [dict(zip(("index", "reads", "samples", "seed", "frac_random"), s)) for s in SETS],
           "fields": "orclib.paf_tuple: mapped, fwd, rid, rd_len, rd_st, rd_en, ... (see tests/orclib.py)",
           "reference_stable_sort": stable, "reference": ref,
           "differ": {k: [i for i in range(len(ref[k])) if ref[k][i] != stable[k][i]] for k in ref}}
    json.dump(out, open(OUT, "w"))
    print({k: (len(v), out["differ"][k]) for k, v in ref.items()})


if __name__ == "__main__":
    main()
