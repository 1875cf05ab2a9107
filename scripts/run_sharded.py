"""torchrun driver: chromosome-sharded EM over NCCL vs the single-GPU EM on rank 0 (same data, same parameters).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/run_sharded.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import titancna_b200 as T
from tests import synth
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, Z = int(os.environ.get("SHARD_N", 400000)), int(os.environ.get("SHARD_Z", 3))
data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
p = bench.production_params(T.loadDefaultParameters, data, Z)
kw = dict(txnExpLen=1e15, txnZstrength=1.0, maxiter=4, maxiterUpdate=1500)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
cp = T.runEMclonalCN_sharded(data, p, rank, world, device=local, **kw)
dist.barrier(); torch.cuda.synchronize(); t_sh = time.perf_counter() - t0
if rank == 0:
    t0 = time.perf_counter()
    ref = T.runEMclonalCN(data, p, verbose=False, **kw)
    t_1 = time.perf_counter() - t0
    err = {k: float(np.abs(np.asarray(cp[k]) - np.asarray(ref[k])).max()) for k in ("n", "phi", "s", "var", "piG", "piZ")}
    llerr = float(np.abs(cp["loglik"][1:] - ref["loglik"][1:]).max() / np.abs(ref["loglik"][1:]).max())
    comm = T.Comm(T.Comm.unique_id(), 0, 1)
    ver = comm.nccl_version()
    comm.close()
    print({"world": world, "nccl_version": ver, "iters": int(cp["n"].size - 1), "sharded_s": t_sh, "single_gpu_s": t_1, "max_abs_err": err, "loglik_rel_err": llerr,
           "shard_sizes": [int(x) for x in T.shard_chromosomes(data, world)[2]]})
    # the shards and the whole genome are chunked differently (chunk length follows the SNP count), so the two EMs agree
    # to the boundary tolerance of the chunked scan, not bit for bit: contract 1e-6 on parameters, 1e-9 on log-likelihoods
    assert max(err.values()) < 1e-7 and llerr < 1e-10, (err, llerr)
dist.barrier()
dist.destroy_process_group()
