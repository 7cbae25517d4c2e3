"""Dry run of GPU tests without a GPU: the bodies of the ordered-mode and exact-ties tests of
tests/test_z_gpu_api_and_tools.py executed against the EMULATED device (a stand-in for uncalled_b200.BatchMapper built on
tests/emulib.py) at reduced sizes, to catch mistakes in the tests themselves before they meet hardware (a few minutes).
    python tools/gpu_tests_on_emulator.py"""
import sys, inspect, types, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np
import emulib, orclib
import uncalled_b200 as RU
import uncalled_b200._native as N

class Index:
    def __init__(self, prefix, device=0): self.prefix = prefix
    def close(self): pass

class BatchMapper:
    def __init__(self, idx, params=None, max_reads=0, max_samples=0):
        self.E = emulib.Emu(idx.prefix)
        if params is not None:
            for f, _ in emulib.UncParams._fields_:
                setattr(self.E.params, f, getattr(params, f))
        self._t = {}
    def _sigs(self, flat, d):
        return [np.asarray(flat[int(x["offset"]):int(x["offset"]) + int(x["n_samples"])], np.float32) for x in d]
    def _np(self, recs):
        arr = (emulib.UncPaf * len(recs))(*recs)
        return np.frombuffer(bytes(arr), dtype=N.PAF_DTYPE).copy()
    def map_ordered(self, flat, d, carry=None, on_device=False):
        recs, c, nre, nro = self.E.map_ordered(self._sigs(flat, d), carry=carry)
        self._t = {"kernel_launches": 5 + 4 * nro, "k2_ms": 1.0}
        return self._np(recs), c, nre, nro
    def map(self, flat, d):
        return self._np(self.E.map_batch(self._sigs(flat, d))[0])
    def set_tie_order(self, m): self.E.set_tie_order(m)
    def timing(self): return self._t
    def close(self): self.E.set_tie_order(0)

U = types.SimpleNamespace(Index=Index, BatchMapper=BatchMapper, make_descs=RU.make_descs, paf_key=RU.paf_key,
                          _native=N, default_params=N.default_params, UncError=RU.UncError)
import importlib
T = importlib.import_module("test_z_gpu_api_and_tools")
src = inspect.getsource(T.test_ordered_mode_equals_one_long_lived_mapper)
src = src.replace('("g200k", 300, 96, 2000, 21, 0.4)', '("g200k", 300, 24, 2000, 21, 0.4)').replace('("g4m7", 10000, 48, 4000, 7, 0.15)', '("g4m7", 10000, 6, 4000, 7, 0.15)')
ns = dict(T.__dict__)
exec(src, ns)
ns["test_ordered_mode_equals_one_long_lived_mapper"](U)
print("ORDERED-GPU-TEST-BODY-OK")
import make_synth_paf_golden as M
src = inspect.getsource(T.test_exact_ties_kernel_equals_the_unmodified_reference_on_every_golden_read)
src = src.replace('gold["reference"][name]', 'gold["reference"][name][:n]').replace('gold["reference_stable_sort"][name]', 'gold["reference_stable_sort"][name][:n]').replace("k = 40", "k = 4").replace("prefix, sig = M.signals(name, n, ns, seed, frac)", "prefix, sig = M.signals(name, n, ns, seed, frac); n = 8 if name == 'g200k' else 5; sig = sig[:n]")
ns = dict(T.__dict__)
exec(src, ns)
ns["test_exact_ties_kernel_equals_the_unmodified_reference_on_every_golden_read"](U)
print("EXACT-GPU-TEST-BODY-OK")
