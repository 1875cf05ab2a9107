"""Run the production EM on the bench workload and dump the parameter trajectory (research input for the
round-criteria prototype).  usage: python scripts/dump_trajectory.py [Z] [out.npz]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

Z = int(sys.argv[1]) if len(sys.argv) > 1 else 3
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/traj_full.npz"
data, truth = synth.make_data(1_000_000, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
p = bench.production_params(T.loadDefaultParameters, data, Z)
sol = T.Solver(data, p["genotypeParams"], Z, 1e15, 1.0)
cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=20, maxiterUpdate=1500, verbose=True, solver=sol)
np.set_printoptions(precision=6, linewidth=200)
for k in ("n", "phi", "s", "loglik"):
    print(k, np.asarray(cp[k]))
np.savez(out, **{k: np.asarray(cp[k]) for k in ("n", "phi", "s", "var", "piG", "piZ", "loglik")})
