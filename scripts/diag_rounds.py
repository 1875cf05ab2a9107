"""Diagnostic: phases / verification rounds / chunk re-runs of the E-step along a production EM trajectory.
usage: python scripts/diag_rounds.py   (env DIAG_N, DIAG_Z, DIAG_CFG="chunk,warm,tol;...")"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

N, Z = int(os.environ.get("DIAG_N", 1_000_000)), int(os.environ.get("DIAG_Z", 3))
cfgs = [tuple(float(x) for x in c.split(",")) for c in os.environ.get("DIAG_CFG", "1024,192,1e-8;512,192,1e-8").split(";")]
data, truth = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
p = bench.production_params(T.loadDefaultParameters, data, Z)
g, sP = p["genotypeParams"], p["cellPrevParams"]
sol = T.Solver(data, g, Z, 1e15, 1.0)
cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=6, maxiterUpdate=1500, verbose=False, solver=sol)
print("n", cp["n"], "loglik", cp["loglik"])
for it in range(cp["n"].size):
    args = (cp["n"][it], cp["s"][:, it], cp["phi"][it], cp["var"][:, it], cp["piG"][:, it], cp["piZ"][:, it], g)
    for chunk, warm, tol in cfgs:
        sol.configure(int(chunk), int(warm), tol)
        sol.estep(*args)
        out = sol.estep(*args)
        tm = sol.timing()
        print(f"iter {it} chunk {int(chunk)} warm {int(warm)} tol {tol:g}: estep {tm['last_estep_ms']:.2f} ms = emission {tm['last_emission_ms']:.2f} + round0 {tm['last_fwd_r0_ms']:.2f} "
              f"+ rounds {tm['last_rounds_ms']:.2f} + final {tm['last_bwd_r0_ms']:.2f} + reduce {tm['last_reduce_ms']:.2f}; rounds {tm['last_fwd_rounds']} runs {tm['last_fwd_chunk_runs']} "
              f"of 2x{tm['n_chunks']} chunks  ll {out['loglik']:.6f}", flush=True)
