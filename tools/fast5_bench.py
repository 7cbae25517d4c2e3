#!/usr/bin/env python
"""Host-side fast5 decode rate of unc_fast5_load (reads/s and MB/s of int16 signal) for 1..N threads.
    python tools/fast5_bench.py [file.fast5] [max_samples_per_read]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uncalled_b200.fast5 import Fast5File  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "fast5", "multi_gzip.fast5")
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with Fast5File(path) as f:
    for threads in (1, 2, 4, 8, 16, 32):
        if threads > (os.cpu_count() or 1):
            break
        reps, n, samples = 0, 0, 0
        t0 = time.time()
        while time.time() - t0 < 1.0:
            rs = f.load(max_samples_per_read=limit, threads=threads)
            reps += 1
            n += len(rs)
            samples += sum(len(r.signal) for r in rs)
        dt = time.time() - t0
        print("%2d threads: %8.0f reads/s  %7.1f MB/s of int16 signal  (%d reads of %s per pass)" %
              (threads, n / dt, samples * 2 / dt / 1e6, len(rs), os.path.basename(path)))
