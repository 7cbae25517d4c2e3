"""DTW: unc_dtw_batch on the GPU next to the reference's own DTWr94p (oracle/_ref, one thread per problem
over the usable host CPUs) on the same problems -- event means of synthetic reads against the k-mers they come from.
    python tools/bench_dtw.py [--problems 128] [--kmers 2000] [--events 3000]"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=128)
    ap.add_argument("--kmers", type=int, default=2000)
    ap.add_argument("--events", type=int, default=3000)
    ap.add_argument("--cpu-problems", type=int, default=32)
    args = ap.parse_args()
    import torch
    import bench
    import orclib
    from uncalled_b200 import dtw as D
    tab = D.model_table()
    rng = np.random.default_rng(5)
    probs = []
    for _ in range(args.problems):
        km = rng.integers(0, 1024, args.kmers).astype(np.uint16)
        idx = np.clip((np.arange(args.events) * args.kmers) // args.events, 0, args.kmers - 1)
        probs.append(((tab[2 * km[idx].astype(np.int64)] + rng.normal(0, 2.5, args.events)).astype(np.float32), km))
    D.dtw_batch(probs[:4], D.DTW_EVENT_GLOB)                     # warm-up (context, module load)
    times, kms = [], []
    import uncalled_b200._native as N
    N.lib().unc_dtw_last_kernel_ms.restype = C.c_float
    for _ in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        got = D.dtw_batch(probs, D.DTW_EVENT_GLOB)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t)
        kms.append(float(N.lib().unc_dtw_last_kernel_ms()))
    cells = float(args.problems) * args.kmers * args.events
    line = {"what": "unc_dtw_batch, DTWr94p, DTW_EVENT_GLOB: %d problems of %d k-mers x %d events, host buffers in and out (Python packing included), device "
                    "workspace kept between calls" % (args.problems, args.kmers, args.events),
            "gpu_s": min(times), "gpu_kernel_ms": min(kms), "kernel_cells_per_s": cells / (min(kms) / 1e3),
            "kernel_breadcrumb_gb_per_s": cells / (min(kms) / 1e3) / 1e9, "gpu_cells_per_s": cells / min(times), "gpu_problems_per_s": args.problems / min(times)}
    if orclib.ref_available():
        R = orclib.ref()
        u64p, u16p, f32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint16), C.POINTER(C.c_float)
        R.ref_dtw.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, f32p, C.c_uint32, u16p, C.c_uint32, u64p, u64p, f32p, f32p]
        ncpu = min(args.cpu_problems, args.problems)
        cpus = bench.host_cpus()["usable"]

        def one(i):
            m, k = probs[i]
            p = np.zeros(2 * (len(m) + len(k)), np.uint64)
            n, s, ms = C.c_uint64(), C.c_float(), C.c_float()
            R.ref_dtw(0, 0, 2, 1, 100, m.ctypes.data_as(f32p), len(m), k.ctypes.data_as(u16p), len(k), p.ctypes.data_as(u64p), C.byref(n),
                      C.byref(s), C.byref(ms))
            return p[:2 * n.value].reshape(-1, 2), s.value
        t = time.perf_counter()
        with ThreadPoolExecutor(cpus) as ex:
            want = list(ex.map(one, range(ncpu)))
        dt = time.perf_counter() - t
        same = sum(1 for i in range(ncpu) if got[i][1] == want[i][1] and np.array_equal(got[i][0], want[i][0]))
        line.update({"cpu_reference_s": dt, "cpu_problems": ncpu, "cpu_threads": cpus, "cpu_cells_per_s": ncpu * args.kmers * args.events / dt,
                     "identical_to_reference": same, "speedup": (cells / min(times)) / (ncpu * args.kmers * args.events / dt)})
    print(json.dumps(line))


if __name__ == "__main__":
    main()
