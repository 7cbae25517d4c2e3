"""Quick GPU timing probe (not the bench): maps N synthetic reads against a synthetic index."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import synth, synthdata
import uncalled_b200._native as _N
if os.environ.get('UNC_LIB'): _N.LIB_PATH = os.path.join(ROOT, os.environ['UNC_LIB'])
import uncalled_b200 as U
name = sys.argv[1] if len(sys.argv) > 1 else "g4m7"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
n_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
t = time.time(); prefix, g = synthdata.get_index(name); print("index %.1fs" % (time.time() - t), flush=True)
t = time.time(); sig, truth = synth.reads(g, n_reads, n_samples, seed=7); print("reads %.1fs" % (time.time() - t), flush=True)
idx = U.Index(prefix, device=0)
bm = U.BatchMapper(idx, max_reads=n_reads, max_samples=n_reads * n_samples)
d = U.make_descs([n_samples] * n_reads)
for it in range(int(os.environ.get('PROBE_ITERS', '3'))):
    t = time.time(); out = bm.map(sig.ravel(), d); wall = time.time() - t
    tm = bm.timing()
    print("iter", it, "wall %.3fs" % wall, {k: round(v, 3) if isinstance(v, float) else v for k, v in tm.items()},
          "reads/s %.1f" % (n_reads / wall), "mapped", int(out["mapped"].sum()), "status!=0", int((out["status"] != 0).sum()), flush=True)
ev = out["events_used"].astype(np.float64)
print("events used mean %.1f children/event %.1f" % (ev.mean(), out["n_children"].sum() / max(ev.sum(), 1)),
      "occ blocks/read %.0f sa steps/read %.0f clusters max %d" % (out["n_occ_blocks"].mean(), out["n_sa_steps"].mean(), out["n_clusters"].max()))
