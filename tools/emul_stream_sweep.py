"""Randomised CPU sweep of the emulated STREAMING path (minutes; not part of the test suite): random chunk lengths,
channel counts, reads per channel, max_chunks / max_events / max_paths, CTA shapes, with and without exact ties -- against
the streaming oracle (per-channel persistent Mapper; pdqsort mode for exact ties).
    python tools/emul_stream_sweep.py [seed [configs]]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), ROOT]
import numpy as np, emulib, orclib, synth, synthdata
import test_stream_emul as TS

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
n_cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 12
prefix, g = synthdata.get_index("g200k")
bad = 0
t0 = time.time()
for rep in range(n_cfg):
    E, O = emulib.Emu(prefix), orclib.Oracle(prefix)
    mp = int(rng.choice([10000, 300, 77])); me = int(rng.choice([30000, 30000, 150]))
    E.params.max_paths = O.params.max_paths = mp
    E.params.max_events = O.params.max_events = me
    chunk = int(rng.choice([200, 450, 450, 1000, 4000])); nch = int(rng.integers(1, 4)); per = int(rng.integers(1, 4))
    maxc = int(rng.choice([1000000, 1000000, 3, 6])); nw = int(rng.choice([2, 3, 5, 8]))
    L = int(rng.integers(600, 7000))
    sig, _ = synth.reads(g, nch * per, L, seed=int(rng.integers(1, 1 << 30)), frac_random=0.35)
    sigs = [sig[i][:int(rng.integers(max(1, L // 3), L + 1))] for i in range(nch * per)]
    exact = int(rng.integers(0, 2))
    O.lib.orc_set_child_sort(exact)
    E.set_tie_order(exact)
    try:
        TS._check(E, O, sigs, nch, chunk, max_chunks=maxc, n_warps=nw)
    except AssertionError as e:
        bad += 1
        print("MISMATCH", dict(mp=mp, me=me, chunk=chunk, nch=nch, per=per, maxc=maxc, nw=nw, L=L, exact=exact), e, flush=True)
    finally:
        O.lib.orc_set_child_sort(0)
        E.set_tie_order(0)
print("STREAM-SWEEP configs", n_cfg, "bad", bad, "%.0fs" % (time.time() - t0))
