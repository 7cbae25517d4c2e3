"""bench_data/bench_reads_4000x4000.fast5: 4000 synthetic r9.4 reads x 4000 samples of the bench genome (g4m7), quantised to
int16 DAC values with the calibration SURVEY.md 8(d) names (digitisation 8192, range 1467.61, offset 10), written as ONE
multi-read fast5 by the reference's vendored libhdf5 (tools/fast5_fixtures/make_bench_fast5.c; needs /root/reference, so it runs
in the build container -- the file travels to the GPU box with the snapshot, it is not committed).
    python tools/make_bench_fast5.py [n_reads [n_samples]]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import synth, synthdata  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
H5 = os.environ.get("H5", "/tmp/h5")
if not os.path.exists(os.path.join(H5, "inst", "lib", "libhdf5.a")):
    subprocess.run(["bash", "-c", "mkdir -p %s && cp -r /root/reference/submods/hdf5 %s/src && chmod -R u+w %s/src && cd %s/src && "
                    "./configure --disable-hl --prefix=%s/inst --enable-shared=no --with-pic=yes >/dev/null && make -j16 >/dev/null && "
                    "make install >/dev/null" % (H5, H5, H5, H5, H5)], check=True)
tool = os.path.join(H5, "bin", "make_bench_fast5")
os.makedirs(os.path.dirname(tool), exist_ok=True)
subprocess.run(["gcc", "-O1", "-I" + os.path.join(H5, "inst", "include"), os.path.join(ROOT, "tools", "fast5_fixtures", "make_bench_fast5.c"),
                os.path.join(H5, "inst", "lib", "libhdf5.a"), "-lz", "-ldl", "-lm", "-o", tool], check=True)
prefix, g = synthdata.get_index("g4m7")
sig, truth = synth.reads(g, n, L, seed=7, noise_mult=1.5)
# pA = range * (raw + offset) / digitisation  ->  raw = pA * digitisation / range - offset
raw = np.clip(np.rint(sig.astype(np.float64) * 8192.0 / 1467.61 - 10.0), 0, 32767).astype(np.int16)
tmp = "/tmp/bench_raw_i16.bin"
raw.tofile(tmp)
out = os.path.join(ROOT, "bench_data", "bench_reads_%dx%d.fast5" % (n, L))
subprocess.run([tool, out, tmp, str(n), str(L)], check=True)
os.remove(tmp)
print(out, os.path.getsize(out))
