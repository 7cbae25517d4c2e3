#!/usr/bin/env python
"""Reproduces the numbers of DESIGN.md section 2 (needs oracle/_ref, i.e. /root/reference at build time; minutes of CPU):
    python tools/measure_divergences.py pdqsort [n]   reference as is        vs the oracle, fresh Mapper per read
    python tools/measure_divergences.py stable  [n]   reference, stable sort vs the oracle, fresh Mapper per read
    python tools/measure_divergences.py pdq-oracle [n] reference as is     vs the oracle with pdqsort restated (orc_set_child_sort(1))
    python tools/measure_divergences.py carry   [n]   reference, ONE long-lived Mapper vs the oracle's fresh / carried modes
One build of the reference per process (it keeps its index in process-global statics)."""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), ROOT]
import numpy as np  # noqa: E402
import orclib  # noqa: E402
import synth  # noqa: E402
import synthdata  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "pdqsort"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (600 if mode == "carry" else 2400)
prefix, g = synthdata.get_index("g4m7")
R = orclib.ref(stable_sort=(mode == "stable"))
assert R.ref_load(prefix.encode(), b"default") == 0
O = orclib.Oracle(prefix)
if mode == "pdq-oracle":
    O.lib.orc_set_child_sort(1)
sig, _ = synth.reads(g, n, 4000, seed=7 if mode == "carry" else 123, frac_random=0.15)
flat = np.ascontiguousarray(sig.reshape(-1))
offs, lens = (np.arange(n) * 4000).astype(np.uint64), np.full(n, 4000, np.uint32)
t = time.time()
if mode == "carry":
    ref = (orclib.RefPaf * n)()
    R.ref_map_batch_mt(orclib.fp(flat), offs.ctypes.data_as(orclib.u64p), lens.ctypes.data_as(orclib.u32p), n, 1, ref)
    ref = [orclib.paf_tuple(r) for r in ref]
    carried = [orclib.paf_tuple(r) for r in O.map_reads_one_mapper(flat, offs, lens)]
    print("reference long-lived Mapper vs oracle with carried flags:", [i for i in range(n) if ref[i] != carried[i]])
else:
    def one(i):
        s = np.ascontiguousarray(sig[i], np.float32)
        out = orclib.RefPaf()
        R.ref_map_read(orclib.fp(s), len(s), C.byref(out))
        return orclib.paf_tuple(out)
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        ref = list(ex.map(one, range(n)))
fresh = [orclib.paf_tuple(r) for r in O.map_batch(flat, offs, lens, threads=os.cpu_count() or 1)]
d = [i for i in range(n) if ref[i] != fresh[i]]
print("%s: %d reads, %d mapped, %d differ from the oracle's fresh-Mapper result: %s  (%.0f s)" %
      (mode, n, sum(1 for r in fresh if r[0]), len(d), d, time.time() - t))
for i in d[:8]:
    print(" ", i, ref[i], fresh[i])
