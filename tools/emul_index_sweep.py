"""Randomised CPU sweep of `uncalled index`'s pieces: random multi-FASTA inputs (repeats, N runs, tiny and low-complexity
sequences, 1..6 contigs, 20 bp .. 30 kb) -> the product's FM-index builder vs the reference's own bwa build (oracle/_ref,
byte-identical files) -> the emulated self_align kernels vs the oracle (and the oracle vs the reference's self_align).
    python tools/emul_index_sweep.py [seed [genomes]]"""
import filecmp
import os
import sys
import tempfile
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), ROOT]
import numpy as np, emulib, orclib
from uncalled_b200 import _native as N

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
n_genomes = int(sys.argv[2]) if len(sys.argv) > 2 else 20
have_ref = orclib.ref_available()
R = orclib.ref() if have_ref else None


def rand_seq(n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def make_fasta(path):
    seqs = []
    for c in range(int(rng.integers(1, 7))):
        kind = int(rng.integers(0, 5))
        n = int(rng.choice([3, 20, 200, 3000, int(rng.integers(50, 30000))]))
        if kind == 0:
            s = rand_seq(n)
        elif kind == 1:
            unit = rand_seq(max(2, n // 6)); s = unit * int(rng.integers(2, 5)) + rand_seq(n // 4)
        elif kind == 2:
            s = rand_seq(n // 2) + "N" * int(rng.integers(1, 60)) + rand_seq(n - n // 2) + ("N" if rng.random() < 0.5 else "")
        elif kind == 3:
            s = "A" * (n // 3) + "AC" * (n // 6) + rand_seq(n // 3 + 1)
        else:
            s = rand_seq(n).lower() if rng.random() < 0.5 else rand_seq(n)
        seqs.append(("c%d some description" % c, s))
    w = int(rng.choice([60, 70, 1000000]))
    with open(path, "w") as f:
        for name, s in seqs:
            f.write(">%s\n" % name)
            for i in range(0, len(s), w):
                f.write(s[i:i + w] + "\n")
    return sum(len(s) for _, s in seqs)


bad = 0
t0 = time.time()
for gi in range(n_genomes):
    d = tempfile.mkdtemp()
    fa = os.path.join(d, "g.fa")
    total = make_fasta(fa)
    prefix = os.path.join(d, "ours")
    rc = N.lib().unc_index_build(fa.encode(), prefix.encode())
    if rc != 0:
        print("genome", gi, "unc_index_build rc", rc, N.lib().unc_last_error().decode()[:120], flush=True)
        continue
    if have_ref:
        rp = os.path.join(d, "ref")
        R.ref_index_build(fa.encode(), rp.encode())
        for ext in (".bwt", ".sa", ".pac", ".ann", ".amb"):
            if not filecmp.cmp(prefix + ext, rp + ext, shallow=False):
                bad += 1
                print("MISMATCH genome", gi, "file", ext, "total bp", total, flush=True)
    for sd in (1, int(rng.integers(2, 9))):
        a, b = emulib.self_align(prefix, sd), orclib.self_align(prefix, sd)
        if not (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])):
            bad += 1
            print("MISMATCH genome", gi, "self_align sample_dist", sd, flush=True)
print("INDEX-SWEEP genomes", n_genomes, "bad", bad, "reference build compared:", have_ref, "%.0fs" % (time.time() - t0))
