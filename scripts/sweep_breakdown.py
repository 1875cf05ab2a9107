"""Where a sweep solution's wall time goes: solver creation / EM / Viterbi per (ploidy, numClusters), one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench
data, _ = synth.make_data(1_000_000, Z=3, maxCN=8, n_chr=23, seed=bench.SEED)
T.lib().titan_set_device(0)
mom = T.dataMoments(data)
for rep in range(2):
    tot = 0.0
    for pl in (2, 3, 4):
        for Z in (1, 2, 3, 4, 5):
            t0 = time.perf_counter()
            p = bench.production_params(T.loadDefaultParameters, data, Z, pl, mom)
            t1 = time.perf_counter()
            sol = T.Solver(data, p["genotypeParams"], Z, 1e15, 1.0)
            t2 = time.perf_counter()
            cp = T.runEMclonalCN(data, p, maxiter=20, verbose=False, solver=sol, return_timing=True, posteriors=False, **bench.EM_KW)
            t3 = time.perf_counter()
            path = T.viterbiClonalCN(data, cp, solver=sol)
            t4 = time.perf_counter()
            vit = sol.timing()["last_viterbi_ms"]
            sol.close()
            t5 = time.perf_counter()
            tot += t5 - t0
            print(f"rep {rep} ploidy {pl} Z {Z}: params {1e3*(t1-t0):6.1f} create {1e3*(t2-t1):6.1f} EM {1e3*(t3-t2):6.1f} (iters {cp['n'].size-1}, device {cp['timing']['estep_ms']:.1f}) "
                  f"viterbi {1e3*(t4-t3):6.1f} (device {vit:.1f}) close {1e3*(t5-t4):5.1f}", flush=True)
    print("rep", rep, "total", round(tot, 3))
