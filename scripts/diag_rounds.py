"""Diagnostic: verification rounds / chunk re-runs of the chunked scan vs parameters and tolerance."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

N, Z = 1_000_000, 3
data, truth = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
p = bench.production_params(T.loadDefaultParameters, data, Z)
g, sP = p["genotypeParams"], p["cellPrevParams"]
sol = T.Solver(data, g, Z, 1e15, 1.0)
cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=6, maxiterUpdate=1500, verbose=False, solver=sol)
print("s trajectory:\n", cp["s"].T, "\nn", cp["n"], "phi", cp["phi"])
for it in range(cp["n"].size):
    args = (cp["n"][it], cp["s"][:, it], cp["phi"][it], cp["var"][:, it], cp["piG"][:, it], cp["piZ"][:, it], g)
    for chunk, warm, tol in ((512, 160, 1e-10), (512, 160, 1e-6), (512, 512, 1e-10), (2048, 512, 1e-10)):
        sol.configure(chunk, warm, tol)
        sol.estep(*args)
        sol.estep(*args)
        tm = sol.timing()
        print(f"iter {it} chunk {chunk} warm {warm} tol {tol:g}: estep {tm['last_estep_ms']:.2f} ms fwd rounds {tm['last_fwd_rounds']} "
              f"runs {tm['last_fwd_chunk_runs']} / bwd rounds {tm['last_bwd_rounds']} runs {tm['last_bwd_chunk_runs']} of {tm['n_chunks']} chunks; "
              f"r0 fwd {tm['last_fwd_r0_ms']:.2f} bwd {tm['last_bwd_r0_ms']:.2f}")
