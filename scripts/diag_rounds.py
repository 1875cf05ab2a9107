"""Diagnostic: verification rounds / chunk re-runs of the chunked scan vs parameters and tolerance."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

N, Z = int(os.environ.get("DIAG_N", 1_000_000)), int(os.environ.get("DIAG_Z", 3))
data, truth = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
p = bench.production_params(T.loadDefaultParameters, data, Z)
g, sP = p["genotypeParams"], p["cellPrevParams"]
sol = T.Solver(data, g, Z, 1e15, 1.0)
cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=6, maxiterUpdate=1500, verbose=False, solver=sol)
print("s trajectory:\n", cp["s"].T, "\nn", cp["n"], "phi", cp["phi"], "loglik", cp["loglik"])
for it in range(cp["n"].size):
    args = (cp["n"][it], cp["s"][:, it], cp["phi"][it], cp["var"][:, it], cp["piG"][:, it], cp["piZ"][:, it], g)
    for chunk, warm, tol in ((512, 160, 1e-10), (1024, 160, 1e-10), (256, 128, 1e-10)):
        sol.configure(chunk, warm, tol)
        sol.estep(*args)
        out = sol.estep(*args)
        tm = sol.timing()
        print(f"iter {it} chunk {chunk} warm {warm} tol {tol:g}: estep {tm['last_estep_ms']:.2f} ms (fwd {tm['last_fwd_ms']:.2f} bwd {tm['last_bwd_ms']:.2f}) "
              f"fwd rounds {tm['last_fwd_rounds']} runs {tm['last_fwd_chunk_runs']} / bwd rounds {tm['last_bwd_rounds']} runs {tm['last_bwd_chunk_runs']} "
              f"of {tm['n_chunks']} chunks; main pass fwd {tm['last_fwd_r0_ms']:.2f} bwd {tm['last_bwd_r0_ms']:.2f}  ll {out['loglik']:.6f}")
