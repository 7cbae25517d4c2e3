"""Deterministic synthetic genomes and r9.4-like raw reads.

genome(n, seed)            uniform random ACGT of length n (numpy default_rng(seed))
write_fasta(path, ...)     60-column FASTA
reads(genome, n_reads...)  per read: start ~ U[0, len-span), strand ~ Bernoulli(0.5),
                           k-mer levels from the r9.4 TEMPLATE table, dwell per k-mer
                           1 + Geometric(p=1/7.9) samples, noise N(0, noise_mult*level_stdv),
                           truncated to n_samples; a fraction `frac_random` of reads is drawn
                           from an unrelated random sequence (never-maps path).
Returns pA float32 signals (what Fast5Reader hands the mapper after calibration).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MODEL_TABLE = os.path.join(_HERE, "..", "uncalled_b200", "data", "r94_5mer_template.f32")


def genome(n, seed=1234):
    return np.random.default_rng(seed).integers(0, 4, n, dtype=np.uint8)


def write_fasta(path, seq, name="synthetic_chr"):
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[seq].tobytes().decode()
    with open(path, "w") as f:
        f.write(">" + name + "\n")
        for i in range(0, len(s), 60):
            f.write(s[i:i + 60] + "\n")


def _kmers(seq):
    """5-mer codes (first base most significant) for every position of seq."""
    n = len(seq) - 4
    k = np.zeros(n, dtype=np.int64)
    for i in range(5):
        k = (k << 2) | seq[i:i + n]
    return k


def reads(gen, n_reads, n_samples=4000, seed=7, frac_random=0.15, noise_mult=1.0, dwell_mean=7.9):
    """Returns (signals[n_reads, n_samples] f32, truth dict)."""
    tab = np.fromfile(MODEL_TABLE, dtype=np.float32).reshape(1024, 2)
    lv_mean, lv_stdv = tab[:, 0].astype(np.float64), tab[:, 1].astype(np.float64)
    rng = np.random.default_rng(seed)
    span = n_samples // 4 + 64  # bases: more than enough k-mers for n_samples
    sig = np.empty((n_reads, n_samples), dtype=np.float32)
    starts = np.zeros(n_reads, dtype=np.int64)
    strands = np.zeros(n_reads, dtype=np.int8)
    is_random = rng.random(n_reads) < frac_random
    comp = np.array([3, 2, 1, 0], dtype=np.uint8)
    for i in range(n_reads):
        if is_random[i]:
            s = rng.integers(0, 4, span, dtype=np.uint8)
            starts[i] = -1
        else:
            st = int(rng.integers(0, len(gen) - span))
            s = gen[st:st + span]
            starts[i] = st
            if rng.random() < 0.5:
                s = comp[s[::-1]]
                strands[i] = 1
        km = _kmers(s)
        dwell = 1 + rng.geometric(1.0 / dwell_mean, size=len(km))
        idx = np.repeat(km, dwell)[:n_samples]
        if len(idx) < n_samples:  # extremely unlikely; pad by repeating the last k-mer
            idx = np.concatenate([idx, np.full(n_samples - len(idx), idx[-1])])
        x = lv_mean[idx] + rng.standard_normal(n_samples) * (noise_mult * lv_stdv[idx])
        sig[i] = x.astype(np.float32)
    return sig, {"start": starts, "strand": strands, "random": is_random}
