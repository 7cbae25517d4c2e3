"""One-off CPU check at full scale (minutes; not part of the test suite): the fully combined prototype build of the
mapper kernel under the warp emulator against the oracle on the 4.7 Mb index with max_paths 10 000 -- reads that never map
run all ~780 events with ~5 200 children each.   python tools/emul_fullscale_check.py"""
import sys, time
sys.path[:0]=['.','tests','tools']
import numpy as np, emulib, orclib, synth, synthdata
import test_emul_kernel as T
prefix, g = synthdata.get_index("g4m7")
flags=("-DK2_TRK_INLINE","-DK2_LEAN_B","-DK2_PAR_E","-DK2_SCAN2","-DK2_PF2","-DK2_DFUSE","-DK2_BMATCH")
E = emulib.Emu(prefix, extra_flags=flags, tag="_everything")
O = orclib.Oracle(prefix)
sig, truth = synth.reads(g, 6, 4000, seed=11, frac_random=0.34)
for i in range(6):
    t=time.time()
    T._check(E, O, [sig[i]])
    rec=O.map_read(sig[i])
    print("read", i, "mapped" if rec.mapped else "unmapped", "children", rec.n_children, "%.1fs"%(time.time()-t), flush=True)
print("BIG-OK")
