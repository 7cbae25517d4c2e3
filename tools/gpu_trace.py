"""Debug: timeline of one read's events in the mapper kernel -- every warp's clock at every phase mark
(-DUNC_PHASE_TIMING build; the read and events are UNC_PT_TRACE_READ / UNC_PT_TRACE_E0 of unc_device.cuh).
    python tools/gpu_trace.py g4m7 2368 uncalled_b200/variants_pt/x0.so"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import uncalled_b200._native as N
N.LIB_PATH = sys.argv[3]
import uncalled_b200 as U
import synth, synthdata
name, n_reads = sys.argv[1], int(sys.argv[2])
prefix, g = synthdata.get_index(name)
sig, truth = synth.reads(g, n_reads, 4000, seed=7)
idx = U.Index(prefix, device=0)
bm = U.BatchMapper(idx, max_reads=n_reads, max_samples=n_reads * 4000)
d = U.make_descs([4000] * n_reads)
out = bm.map(sig.ravel(), d)
tr = np.zeros((8, 16, 32), np.uint64)
L = N.lib()
L.unc_pool_debug_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
N.check(L.unc_pool_debug_trace(bm.h, n_reads, tr.ctypes.data))
order = [9, 7, 0, 16, 1, 17, 10, 26, 12, 2, 18, 3, 19, 11, 27, 5, 30, 6, 8]
label = {9: "top", 7: "load", 0: "A", 16: "A|", 1: "B", 17: "B|", 10: "B2", 26: "B2|", 12: "cut", 2: "C1", 18: "C1|", 3: "C2", 19: "C2|", 11: "D0",
         27: "D0|", 5: "D1E", 30: "X|", 6: "X|f", 8: "end"}
print("trace read: events_used", out["events_used"][40], "mapped", out["mapped"][40])
for e in range(8):
    t = tr[e].astype(np.int64)
    if not t.any():
        continue
    t0 = t[1:][t[1:] > 0].min()
    print("event %d (clock %d)" % (e, t0))
    print("warp  " + " ".join("%7s" % label[m] for m in order))
    for w in range(16):
        if not t[w].any():
            continue
        if w == 0:
            print("trk   arrive %d  released %d  n_rows read %d  verdict out %d   seeds %d -> %.0f cycles/seed" % (tuple(int(x - t0) if x else -1 for x in t[0, :4]) + (int(t[0, 4]), (t[0, 3] - t[0, 2]) / max(int(t[0, 4]), 1))))
            continue
        print("%4d  " % w + " ".join("%7d" % (t[w, m] - t0) if t[w, m] else "      -" for m in order))
