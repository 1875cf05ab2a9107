"""Where a solution's wall time goes: solver creation, EM (device vs host), Viterbi; per (ploidy, Z)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--snps", type=int, default=1_000_000)
ap.add_argument("--configs", type=str, nargs="+", default=["2:1", "2:3", "2:4", "3:3", "4:1"])
ap.add_argument("--repeat", type=int, default=2)
a = ap.parse_args()
data, _ = synth.make_data(a.snps, Z=3, maxCN=8, n_chr=23, seed=bench.SEED)
for cfg in a.configs:
    ploidy, Z = (int(x) for x in cfg.split(":"))
    for rep in range(a.repeat):
        p = bench.production_params(T.loadDefaultParameters, data, Z)
        p["ploidyParams"]["phi_0"] = float(ploidy)
        t0 = time.perf_counter()
        sol = T.Solver(data, p["genotypeParams"], Z, 1e15, 1.0)
        t1 = time.perf_counter()
        cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=20, maxiterUpdate=1500, verbose=False,
                             return_timing=True, solver=sol)
        t2 = time.perf_counter()
        tm = sol.timing()
        path = T.viterbiClonalCN(data, cp, solver=sol)
        t3 = time.perf_counter()
        print(f"ploidy {ploidy} Z {Z} rep {rep}: create {t1-t0:.3f} em {t2-t1:.3f} (iters {cp['n'].size-1}, estep device "
              f"{cp['timing']['estep_ms']:.1f} ms, mstep {cp['timing']['mstep_ms']:.1f} ms, last rounds {tm['last_fwd_rounds']}/{tm['last_bwd_rounds']}) "
              f"viterbi {t3-t2:.3f} (device {sol.timing()['last_viterbi_ms']:.1f} ms) total {t3-t0:.3f}", flush=True)
        sol.close()
