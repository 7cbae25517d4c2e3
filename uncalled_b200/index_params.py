"""`uncalled index`'s parameter search: from the FM range lengths of sampled self-alignments
(`self_align`, reference src/self_align_ref.cpp:34-91) to the `.uncl` threshold lines the mapper loads
(reference uncalled/index.py:53-209, scripts/uncalled:38-78).

Host-side Python like the reference's (it runs once per index); the heavy part -- the sampled backward
searches -- comes from `unc_self_align` on the GPU (or, in the CPU tests, from the oracle).  Every
floating-point step uses the same numpy operation on the same operands as the reference, so the produced
`.uncl` text is character-identical (tests/test_index_params.py compares it with lines made by the real
reference).

Data: `data/r94_5mer_threshs.f64` = the r9.4 model's (threshold, match frequency, k-mer count) table
(reference uncalled/conf/r94_5mers_threshs.txt, 4901 rows) as float64.
"""
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
THRESH_TABLE = os.path.join(PKG_DIR, "data", "r94_5mer_threshs.f64")

# argument defaults of `uncalled index` (reference uncalled/args.py:87-137)
DEFAULTS = dict(max_sample_dist=100, min_samples=50000, max_samples=1000000, kmer_len=5, matchpr1=0.6334,
                matchpr2=0.9838, pathlen_percentile=0.05, max_replen=100)


def reference_length(bwa_prefix):
    """l_pac, the first field of the .ann header (uncalled/index.py:71-74)."""
    with open(bwa_prefix + ".ann") as f:
        return int(f.readline().split()[0])


def sample_distance(ref_len, max_sample_dist=100, min_samples=50000, max_samples=1000000):
    """Average distance between sampled start positions (uncalled/index.py:76-82)."""
    approx = ref_len / max_sample_dist
    if approx < min_samples:
        return int(np.ceil(ref_len / min_samples))
    if approx > max_samples:
        return int(np.floor(ref_len / max_samples))
    return max_sample_dist


class PathStatistics:
    """What calc_map_stats derives from the self-alignment paths (uncalled/index.py:84-116): where along a
    path the FM range has which order of magnitude."""

    def __init__(self, offsets, values, kmer_len=5, max_replen=100, pathlen_percentile=0.05):
        offsets = np.asarray(offsets, dtype=np.int64)
        values = np.asarray(values, dtype=np.uint64)
        n = len(offsets) - 1
        lens = offsets[1:] - offsets[:-1]
        # a path enters the statistics from its k-th range on; shorter paths count as one range of length 1
        klen = np.where(lens >= kmer_len, lens - (kmer_len - 1), 1)
        kstart = offsets[:-1] + (kmer_len - 1)
        short = lens < kmer_len
        # fraction of the (not overly repetitive) paths that are still ambiguous after i steps
        kept = klen[klen <= max_replen]
        still = np.zeros(int(kept.max()))
        counts = np.bincount(kept, minlength=len(still) + 1)
        # paths of k-length l contribute to positions 0..l-1
        still += np.cumsum(counts[::-1])[::-1][1:len(still) + 1]
        self.max_pathlen = int(np.flatnonzero(still / len(kept) <= pathlen_percentile)[0])
        first = np.where(short, np.uint64(1), values[np.minimum(kstart, len(values) - 1)])
        max_fmexp = int(np.log2(int(first.max()))) + 1
        mat = np.zeros((max_fmexp, self.max_pathlen))
        for i in range(self.max_pathlen):
            have = (~short) & (klen > i)
            if have.any():
                bins = np.log2(values[kstart[have] + i]).astype(np.int64)     # int(np.log2(range length))
                mat[:, i] += np.bincount(bins, minlength=max_fmexp)[:max_fmexp]
            if i == 0:
                mat[0, 0] += np.count_nonzero(short)                         # the stand-in range of length 1
            else:
                mat[0, i] += np.count_nonzero(klen <= i)                     # the path has ended: unique
        self.fm_path_mat = mat
        steps = np.arange(self.max_pathlen)
        exps = np.arange(max_fmexp)
        # an order of magnitude no path passes through gives 0/0 = nan here, as in the reference; the nan
        # reaches the .uncl line and the mapper's loader accepts it
        with np.errstate(invalid="ignore", divide="ignore"):
            self.fm_locs = np.array([np.sum(steps * (mat[f] / np.sum(mat[f]))) for f in range(max_fmexp)])
            self.loc_fms = np.array([np.sum(exps * (mat[:, p] / np.sum(mat[:, p]))) for p in range(self.max_pathlen)])
        self.speed_denom = np.sum(self.loc_fms)
        self.conf_locs = np.arange(np.round(self.fm_locs[0]))
        self.all_locs = steps


class Parameterizer:
    """The preset search of IndexParameterizer (uncalled/index.py:118-209)."""

    def __init__(self, stats, matchpr1=0.6334, matchpr2=0.9838, thresh_table=THRESH_TABLE):
        self.st, self.pck1, self.pck2 = stats, matchpr1, matchpr2
        tab = np.fromfile(thresh_table, dtype=np.float64).reshape(-1, 3)
        self.model_ekms = np.flip(tab[:, 0], 0)
        self.model_pcks = np.flip(tab[:, 1], 0)
        self.model_counts = np.flip(tab[:, 2], 0)
        self.presets = {}

    def _curve(self, exp, N=100):
        dt = 1.0 / N
        t = np.arange(0, 1 + dt, dt)
        return t * self.st.fm_locs[0], (t ** exp) * (self.pck2 - self.pck1) + self.pck1

    def _speed(self, locs, pcks):
        p = np.interp(self.st.all_locs, locs, pcks)
        counts = np.interp(p, self.model_pcks, self.model_counts)
        return np.dot(counts, self.st.loc_fms) / (self.st.speed_denom)

    def _prob(self, locs, pcks):
        return np.prod(np.interp(self.st.conf_locs, locs, pcks))

    def add_preset(self, name, tgt_prob=None, tgt_speed=None, exp_st=2, init_fac=2, eps=0.00001):
        """Bisection on the exponent of the match-probability curve until the target is met."""
        exp, lo, hi, last = exp_st, None, None, None
        while True:
            locs, pcks = self._curve(exp)
            delta = (self._prob(locs, pcks) - tgt_prob) if tgt_prob is not None else (self._speed(locs, pcks) - tgt_speed)
            if abs(delta) <= eps or delta == last:
                break
            last = delta
            if delta < 0:
                hi = exp
            else:
                lo = exp
            prev = exp
            if hi is None:
                exp *= init_fac
            elif lo is None:
                exp /= init_fac
            else:
                exp = lo + ((hi - lo) / 2.0)
            if exp == prev:
                break
        fm_pcks = np.interp(self.st.fm_locs, locs, pcks)
        ekms = np.interp(fm_pcks, self.model_pcks, self.model_ekms)
        self.presets[name] = (ekms, self._prob(locs, pcks), self._speed(locs, pcks))

    def text(self):
        """The .uncl file: one TAB-separated line per preset (uncalled/index.py:202-209)."""
        return "".join("%s\t%s\t%.5f\t%.3f\n" % (name, ",".join(map(str, ekms)), prob, speed)
                       for name, (ekms, prob, speed) in self.presets.items())


def uncl_text(offsets, values, probs=None, speeds=None, **opts):
    """`uncalled index` after the BWA build (scripts/uncalled:61-78): the "default" preset (speed 115) plus the
    requested prob_/speed_ presets."""
    o = dict(DEFAULTS)
    o.update(opts)
    st = PathStatistics(offsets, values, o["kmer_len"], o["max_replen"], o["pathlen_percentile"])
    pz = Parameterizer(st, o["matchpr1"], o["matchpr2"])
    pz.add_preset("default", tgt_speed=115)
    for kind, targets in (("prob", probs), ("speed", speeds)):
        for t in (targets.split(",") if isinstance(targets, str) else (targets or [])):
            try:                                                    # a target that cannot be parsed or searched is
                pz.add_preset("%s_%s" % (kind, t), **{"tgt_" + kind: float(t)})   # skipped (scripts/uncalled:61-74)
            except Exception:
                pass
    return pz.text()
