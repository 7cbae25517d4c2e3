"""Randomised CPU sweep of the emulated event-detection kernel (k1_events + serial redo + normaliser statistics) against
the oracle: synthetic reads, noise at several scales, step functions, quantised values, spikes / tiny / huge values that
force the serial path, float and int16 input with random calibrations, lengths 0..9000, 1..4 warps per CTA.
    python tools/emul_k1_sweep.py [seed [batches]]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), ROOT]
import numpy as np, emulib, orclib, synth, synthdata

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_batches = int(sys.argv[2]) if len(sys.argv) > 2 else 40
prefix, g = synthdata.get_index("g200k")
E, O = emulib.Emu(prefix), orclib.Oracle()
base, _ = synth.reads(g, 8, 9000, seed=4)


def make(kind, n):
    if kind == 0:
        s = base[rng.integers(0, 8)][:n].copy()
    elif kind == 1:
        s = rng.normal(90, float(rng.choice([0.01, 1, 10, 40])), n)
    elif kind == 2:                                    # steps with little noise
        s = np.repeat(rng.uniform(60, 130, n // 7 + 1), 7)[:n] + rng.normal(0, 0.3, n)
    elif kind == 3:                                    # quantised (exact sums)
        s = np.round(rng.normal(90, 12, n) * 4) / 4
    else:
        s = np.full(n, float(rng.uniform(50, 120)))
        if n:
            s[rng.integers(0, n, max(1, n // 50))] = rng.uniform(0, 500, max(1, n // 50))
    s = np.asarray(s, np.float32)
    if n and rng.random() < 0.25:                      # values that break the exact-sum condition
        idx = rng.integers(0, n, 2)
        s[idx[0]] = np.float32(rng.choice([1e-20, 3e7, 1e-41, 5e5]))
    return s


bad = 0
total = 0
t0 = time.time()
for b in range(n_batches):
    k = int(rng.integers(1, 9))
    os.environ["UNC_EMU_K1_WARPS"] = str(int(rng.integers(1, 5)))
    lens = [int(rng.choice([0, 1, 5, 12, 13, 40, 1151, 1152, 1153, int(rng.integers(0, 9000))])) for _ in range(k)]
    sigs = [make(int(rng.integers(0, 5)), L) for L in lens]
    if rng.random() < 0.35:                            # int16 DAC input, one calibration per batch
        cal = (float(rng.uniform(1200, 1600)), float(rng.integers(-5, 30)), float(rng.choice([8192.0, 8000.0, 2048.0])))
        raws = [np.clip(np.round(s.astype(np.float64) * cal[2] / cal[0] - cal[1]), -60, 32000).astype(np.int16) for s in sigs]
        pas = [((np.float32(cal[0]) * (r.view(np.uint16).astype(np.float32) + np.float32(cal[1]))) / np.float32(cal[2])).astype(np.float32)
               for r in raws]
        if sum(lens) == 0:
            continue
        _, ev, nm, mel = E.map_batch(raws, run_k2=False, dtype=1, cal=cal)
    else:
        pas = sigs
        if sum(lens) == 0:
            continue
        _, ev, nm, mel = E.map_batch(sigs, run_k2=False)
    for i, x in enumerate(pas):
        m, _, _, omel = O.detect(np.ascontiguousarray(x, np.float32))
        ok = len(ev[i]) == len(m) and np.array_equal(ev[i], m) and (mel[i] == omel or (np.isnan(mel[i]) and np.isnan(omel)))
        if ok and len(m):
            ok = np.array_equal(nm[i], O.normalize(m), equal_nan=True)
        total += 1
        if not ok:
            bad += 1
            print("MISMATCH batch", b, "read", i, "len", len(x), flush=True)
print("K1-SWEEP reads", total, "bad", bad, "%.0fs" % (time.time() - t0))
