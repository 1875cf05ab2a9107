"""15-solution sweep (ploidy {2,3,4} x numClusters {1..5}) on the visible GPUs.
   1 GPU:  python scripts/run_sweep_gpu.py [--concurrent k]
   N GPUs: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 scripts/run_sweep_gpu.py"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--snps", type=int, default=1_000_000)
ap.add_argument("--concurrent", type=int, default=1)
ap.add_argument("--maxiter", type=int, default=20)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
T.lib().titan_set_device(local)
data, _ = synth.make_data(a.snps, Z=3, maxCN=8, n_chr=23, seed=bench.SEED)


def make_params(ploidy, Z):
    p = bench.production_params(T.loadDefaultParameters, data, Z)
    p["ploidyParams"]["phi_0"] = float(ploidy)
    return p


def run_one(d, p, **kw):
    T.lib().titan_set_device(local)
    t0 = time.perf_counter()
    sol = T.Solver(d, p["genotypeParams"], len(p["cellPrevParams"]["s_0"]), kw.get("txnExpLen", 1e15), kw.get("txnZstrength", 5e5))
    try:
        cp = T.runEMclonalCN(d, p, verbose=False, solver=sol, **kw)
        path = T.viterbiClonalCN(d, cp, solver=sol)
    finally:
        sol.close()
    return {"n": float(cp["n"][-1]), "phi": float(cp["phi"][-1]), "loglik": float(cp["loglik"][-1]), "iters": int(cp["n"].size - 1),
            "segments": int(1 + np.count_nonzero(np.diff(path))), "seconds": time.perf_counter() - t0}


t0 = time.perf_counter()
out, info = T.run_sweep(data, rank, world, make_params=make_params, run_one=run_one, concurrent=a.concurrent,
                        txnExpLen=1e15, txnZstrength=1.0, maxiter=a.maxiter, maxiterUpdate=1500)
wall = time.perf_counter() - t0
if rank == 0:
    for o in out:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in o.items()})
    print({"world": world, "concurrent": a.concurrent, "solutions": len(out), "sweep_wall_s": round(wall, 3),
           "sum_solution_s": round(sum(o["seconds"] for o in out), 3), "lpt_load": info["load"].tolist()})
if world > 1:
    dist.barrier(); dist.destroy_process_group()
