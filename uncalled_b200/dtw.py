"""DTW of event means against reference k-mers: the reference's `DTWr94p`, `DTWr94d`, `DTWParams` and the `DTW_*` presets
(src/dtw.hpp:9-28,188-232, bound at src/pybinder.cpp:75-91) over `unc_dtw_batch` (include/unc_b200.h).

    d = DTWr94p(means, kmers, DTW_EVENT_GLOB); d.get_path(); d.score(); d.mean_score()

`dtw_batch` aligns many (means, kmers) pairs in one call (one CTA per problem on the GPU).  There is no CPU path."""
import ctypes as C

import numpy as np

from . import _native as N


class DTWSubSeq:                      # enum class DTWSubSeq {NONE, ROW, COL}, src/dtw.hpp:9
    NONE, ROW, COL = 0, 1, 2


class DTWParams(C.Structure):
    """DTWParams (src/dtw.hpp:10-13): subseq, dw, hw, vw -- the weights of the diagonal, horizontal and vertical moves."""
    _fields_ = [("subseq", C.c_int32), ("dw", C.c_float), ("hw", C.c_float), ("vw", C.c_float)]

    def __init__(self, subseq=DTWSubSeq.NONE, dw=1.0, hw=1.0, vw=1.0):
        super().__init__(int(subseq), float(dw), float(hw), float(vw))


# src/dtw.hpp:15-28
DTW_EVENT_GLOB = DTWParams(DTWSubSeq.NONE, 2, 1, 100)
DTW_EVENT_QSUB = DTWParams(DTWSubSeq.COL, 2, 1, 100)
DTW_EVENT_RSUB = DTWParams(DTWSubSeq.ROW, 2, 1, 100)
DTW_RAW_QSUB = DTWParams(DTWSubSeq.COL, 10, 1, 1000)
DTW_RAW_RSUB = DTWParams(DTWSubSeq.ROW, 10, 1, 1000)
DTW_RAW_GLOB = DTWParams(DTWSubSeq.NONE, 10, 1, 1000)

_model = None


def model_table():
    """The r9.4 template model as (mean, stdv) pairs per 5-mer (src/model_r94.inl)."""
    global _model
    if _model is None:
        _model = np.fromfile(N.MODEL_TABLE, dtype=np.float32)
        assert _model.size == 2048
    return _model


def dtw_batch(problems, prms, cost="r94p", model=None):
    """problems: sequence of (means, kmers).  Returns a list of (path, score): path = uint64 array [n, 2] of (column =
    event index, row = k-mer index) pairs from the end of the alignment back to its start, as `get_path()` gives them."""
    L = N.lib()
    L.unc_dtw_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(DTWParams), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n = len(problems)
    if n == 0:
        return []
    means = [np.ascontiguousarray(m, dtype=np.float32).ravel() for m, _ in problems]
    kmers = [np.ascontiguousarray(k, dtype=np.uint16).ravel() for _, k in problems]
    moff = np.zeros(n + 1, np.uint64)
    koff = np.zeros(n + 1, np.uint64)
    poff = np.zeros(n + 1, np.uint64)
    moff[1:] = np.cumsum([len(m) for m in means])
    koff[1:] = np.cumsum([len(k) for k in kmers])
    poff[1:] = np.cumsum([len(m) + len(k) for m, k in zip(means, kmers)])
    am, ak = np.concatenate(means), np.concatenate(kmers)
    path = np.zeros((int(poff[-1]), 2), np.uint64)
    plen = np.zeros(n, np.uint64)
    score = np.zeros(n, np.float32)
    tab = np.ascontiguousarray(model if model is not None else model_table(), dtype=np.float32)
    kind = {"r94p": 0, "r94d": 1}[cost]
    N.check(L.unc_dtw_batch(tab.ctypes.data, kind, C.byref(prms), n, am.ctypes.data, moff.ctypes.data, ak.ctypes.data, koff.ctypes.data,
                            path.ctypes.data, poff.ctypes.data, plen.ctypes.data, score.ctypes.data))
    return [(path[int(poff[i]):int(poff[i]) + int(plen[i])].copy(), float(score[i])) for i in range(n)]


class _DTW:
    _cost = None

    def __init__(self, means, kmers, prms):
        (self._path, self._score), = dtw_batch([(means, kmers)], prms, self._cost)

    def get_path(self):
        """[(event index, k-mer index), ...] from the end of the alignment back to its start (src/dtw.hpp:124-126)."""
        return [(int(a), int(b)) for a, b in self._path]

    def score(self):
        return self._score

    def mean_score(self):             # score_sum_ / path_.size() in float (src/dtw.hpp:132-134)
        return float(np.float32(self._score) / np.float32(len(self._path)))


class DTWr94p(_DTW):
    """cost = -match_prob(event, k-mer) of the r9.4 template model (src/dtw.hpp:188-209)"""
    _cost = "r94p"


class DTWr94d(_DTW):
    """cost = abs(event - model mean of the k-mer), as the reference compiles it: the difference is truncated to an
    integer first (src/dtw.hpp:212-232)"""
    _cost = "r94d"
