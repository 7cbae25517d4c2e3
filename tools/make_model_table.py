#!/usr/bin/env python
"""Write uncalled_b200/data/r94_5mer_template.f32 : the ONT r9.4 5-mer template model as
1024 x (level_mean, level_stdv) float32 pairs in k-mer order AAAAA..TTTTT.

Source of the numbers: the reference's own table (identical values in
/root/reference/src/model_r94.inl:6-1031 and uncalled/conf/r94_5mers.txt); rounding is
decimal text -> double -> float32, the same path the C++ initialiser list takes.
"""
import os, re, sys
import numpy as np
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/model_r94.inl"
rows = re.findall(r"^\s+([0-9.]+),\s+([0-9.]+),?\s*//([ACGT]{5})", open(src).read(), re.M)
assert len(rows) == 1024
order = ["".join(k) for k in __import__("itertools").product("ACGT", repeat=5)]
assert [r[2] for r in rows] == order
tab = np.array([[float(r[0]), float(r[1])] for r in rows], dtype=np.float64).astype(np.float32)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "uncalled_b200", "data")
os.makedirs(out, exist_ok=True)
tab.tofile(os.path.join(out, "r94_5mer_template.f32"))
print("wrote", tab.shape, tab[:2], tab[-1])
