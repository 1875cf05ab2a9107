"""Determinism stress: the E-step is a fixed-order computation, so repeated calls with the same parameters must
return bit-identical statistics.  Any difference is a race.  usage: python scripts/race_check.py [N] [reps] [chunk]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
Z = 3
data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
p = bench.production_params(T.loadDefaultParameters, data, Z)
g, sP = p["genotypeParams"], p["cellPrevParams"]
sol = T.Solver(data, g, Z, 1e15, 1.0)
sol.configure(chunk, 192, 1e-8)
args = (0.45, sP["s_0"], 2.04, g["var_0"], g["piG_0"], sP["piZ_0"], g)
ref = None
bad = 0
for r in range(reps):
    out = sol.estep(*args, posteriors=(r % 2 == 0))
    key = np.concatenate([out["stats"][k].ravel() for k in "abcdeg"] + [out["loglik_chr"]])
    if ref is None:
        ref = key
    elif not np.array_equal(ref, key):
        bad += 1
        d = np.abs(ref - key)
        print(f"rep {r}: {np.count_nonzero(ref != key)} values differ, max abs {d.max():.3e} (rel {(d / np.maximum(np.abs(ref), 1e-300)).max():.3e})")
print(f"N={N} chunk={chunk}: {bad} of {reps - 1} repeats differ from the first")
