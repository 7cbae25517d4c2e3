"""One-off CPU check at full scale (a few minutes; not part of the test suite): ordered mapping
(unc_ordered_logic.hpp + the shipped mapper kernel under the warp emulator) against the oracle's one-Mapper chain on
the 4.7 Mb index with max_paths 10 000, over bench-workload reads around one that never maps and leaves 65 flags set
(read 99 of the seed-7 set).   python tools/emul_ordered_fullscale.py [n_set lo hi]
`600 575 595` covers the read of DESIGN.md section 2 whose result depends on its predecessor."""
import sys, time
sys.path[:0] = ['.', 'tests', 'tools']
import numpy as np, emulib, orclib, synth, synthdata

prefix, g = synthdata.get_index("g4m7")
E, O = emulib.Emu(prefix), orclib.Oracle(prefix)
n_set, lo, hi = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (120, 97, 103)
sig, _ = synth.reads(g, n_set, 4000, seed=7, frac_random=0.15)
sigs = [np.ascontiguousarray(sig[i], np.float32) for i in range(lo, hi)]
flat = np.concatenate(sigs)
lens = np.array([len(s) for s in sigs], np.uint32)
offs = (np.arange(len(sigs), dtype=np.uint64) * 4000).astype(np.uint64)
want = O.map_reads_one_mapper(flat, offs, lens)
t = time.time()
recs, carry, n_re, n_ro = E.map_ordered(sigs)
print("emulated ordered batch: %.1f s, %d read(s) mapped again in %d extra round(s)" % (time.time() - t, n_re, n_ro), flush=True)
cnt = lambda r: (r.n_children, r.n_sources, r.n_seeds, r.n_clusters)
prev = np.zeros(32, np.uint32)
for i, s in enumerate(sigs):
    assert (emulib.paf_tuple(recs[i]), cnt(recs[i])) == (orclib.paf_tuple(want[i]), cnt(want[i])), i
    _, prev = O.map_read_flags(s, prev)
    print("read", lo + i, "mapped" if want[i].mapped else "unmapped", "children", want[i].n_children,
          "flags set after it:", int(sum(bin(int(x)).count("1") for x in prev)), flush=True)
assert np.array_equal(carry, prev)
fresh = [O.map_read(s) for s in sigs]
print("reads whose chained result differs from a new Mapper's:", [lo + i for i in range(len(sigs)) if (orclib.paf_tuple(want[i]), cnt(want[i])) != (orclib.paf_tuple(fresh[i]), cnt(fresh[i]))])
print("ORDERED-BIG-OK")
