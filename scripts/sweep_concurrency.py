"""15-solution sweep on one GPU with 1, 2, 3 solutions in flight (host threads + separate streams), same process, caches warm."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

data, _ = synth.make_data(1_000_000, Z=3, maxCN=8, n_chr=23, seed=bench.SEED)
T.lib().titan_set_device(0)


def run_one(d, pp, **kw):
    T.lib().titan_set_device(0)
    t1 = time.perf_counter()
    so = T.Solver(d, pp["genotypeParams"], len(pp["cellPrevParams"]["s_0"]), kw["txnExpLen"], kw["txnZstrength"], device=0)
    try:
        c = T.runEMclonalCN(d, pp, verbose=False, solver=so, posteriors=False, **kw)
        pth = T.viterbiClonalCN(d, c, solver=so)
    finally:
        so.close()
    return {"iters": int(c["n"].size - 1), "loglik": float(c["loglik"][-1]), "segments": int(1 + np.count_nonzero(np.diff(pth))),
            "seconds": time.perf_counter() - t1}


ref = None
for conc in (1, 1, 2, 3, 4, 1):
    t0 = time.perf_counter()
    mom = T.dataMoments(data)
    res, info = T.run_sweep(data, 0, 1, device=0, maxiter=20, run_one=run_one, concurrent=conc,
                            make_params=lambda pl, z: bench.production_params(T.loadDefaultParameters, data, z, pl, mom), **bench.EM_KW)
    wall = time.perf_counter() - t0
    key = [(r["ploidy_0"], r["numClusters"], r["iters"], r["loglik"], r["segments"]) for r in res]
    if ref is None:
        ref = key
    print(f"concurrent {conc}: sweep wall {wall:.3f} s, sum of solution seconds {sum(r['seconds'] for r in res):.3f}, identical results {key == ref}", flush=True)
