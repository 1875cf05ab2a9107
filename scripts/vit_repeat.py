"""Determinism of the Viterbi decode at 36 genotype states (two per lane): repeated decodes must be identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
data, _ = synth.make_data(300000, Z=5, maxCN=8, n_chr=23, seed=9)
p = T.loadDefaultParameters(copyNumber=10, numberClonalClusters=5, data=data, extendedTable=True)
g, sP = p["genotypeParams"], p["cellPrevParams"]
args = (0.4, sP["s_0"], 2.05, g["var_0"], g["piG_0"], sP["piZ_0"], g)
sol = T.Solver(data, g, 5, 1e15, 1.0)
ref = sol.viterbi(*args)
bad = sum(int(not np.array_equal(ref, sol.viterbi(*args))) for _ in range(6))
os.environ["TITAN_VIT_MODE"] = "0"
sol0 = T.Solver(data, g, 5, 1e15, 1.0)
same = bool(np.array_equal(ref, sol0.viterbi(*args)))
print("K", sol.K, "repeats differing", bad, "equals dense scan", same, "segments", 1 + np.count_nonzero(np.diff(ref)))
sol.close(); sol0.close()
