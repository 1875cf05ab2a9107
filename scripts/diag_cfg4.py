"""Cycles per E-step of the 2 M-SNP K = 125 workload (BASELINE configs[3], U = 25 variant): TITAN_TRACE=2 on stderr."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TITAN_TRACE", "2")
import numpy as np
import titancna_b200 as T
from tests import synth
import bench
Z = int(sys.argv[1]) if len(sys.argv) > 1 else 5
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED + 2)
if os.environ.get("MAXCN") == "10":      # BASELINE configs[3] as named: 36 genotype states, K = 180
    p = T.loadDefaultParameters(copyNumber=10, numberClonalClusters=Z, data=data, extendedTable=True)
    K0 = p["genotypeParams"]["rt"].size
    p["genotypeParams"]["alphaKHyper"] = np.full(K0, 10000.0); p["genotypeParams"]["betaKHyper"] = np.full(K0, 25.0)
else:
    p = bench.production_params(T.loadDefaultParameters, data, Z, 2.0)
sol = T.Solver(data, p["genotypeParams"], Z, 1e15, 1.0)
if os.environ.get("CHUNK"):
    sol.configure(int(os.environ["CHUNK"]), int(os.environ.get("WARM", 192)), 1e-8)
cp = T.runEMclonalCN(data, p, maxiter=4, verbose=False, solver=sol, return_timing=True, **bench.EM_KW)
print("estep_ms", cp["timing"]["estep_ms"], "iters", cp["n"].size - 1)
g = p["genotypeParams"]
for i in range(cp["n"].size - 1):
    sol.estep(cp["n"][i], cp["s"][:, i], cp["phi"][i], cp["var"][:, i], cp["piG"][:, i], cp["piZ"][:, i], g)
    tm = sol.timing()
    print(i, "estep %.2f = emission %.2f + round0 %.2f + verification %.2f + final %.2f  cycles %d" % (
        tm["last_estep_ms"], tm["last_emission_ms"], tm["last_fwd_r0_ms"], tm["last_rounds_ms"], tm["last_bwd_r0_ms"], tm["last_fwd_rounds"]))
import time
t0 = time.perf_counter()
path = T.viterbiClonalCN(data, cp, solver=sol)
print("viterbi wall %.1f ms device %.1f ms" % (1e3 * (time.perf_counter() - t0), sol.timing()["last_viterbi_ms"]), "K", sol.K, "segments", 1 + np.count_nonzero(np.diff(path)))
sol.close()
