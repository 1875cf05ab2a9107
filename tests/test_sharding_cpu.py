"""CPU (gloo, world_size 2): the multi-rank host logic -- LPT partitions, the one all-reduce of the
M-step statistics and the replicated M-step -- with the oracle's E-step standing in for the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import titancna_b200 as T
from titancna_b200 import sweep
from oracle import titan_oracle as O
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_assignment_balances_the_sweep():
    grid, costs = sweep.solution_grid()
    assert len(grid) == 15 and costs.sum() == 3 * 25 * 15
    where, load = sweep.lpt_assign(costs, 8)
    assert sorted(np.bincount(where, minlength=8).tolist())[0] >= 1
    assert load.max() <= 1.35 * costs.sum() / 8 + 1e-9          # K=125 runs dominate; LPT stays close to the mean
    where1, load1 = sweep.lpt_assign(costs, 1)
    assert load1[0] == costs.sum()


def test_sweep_runs_every_solution_once_with_and_without_host_concurrency():
    data, _ = synth.make_data(600, Z=2, maxCN=5, n_chr=3, seed=2)
    seen = []
    run_one = lambda d, p, **kw: (seen.append((p["ploidyParams"]["phi_0"], len(p["cellPrevParams"]["s_0"]))) or {"ok": True})
    for conc in (1, 3):
        seen.clear()
        out, info = sweep.run_sweep(data, 0, 1, run_one=run_one, gather=False, concurrent=conc)
        assert len(out) == 15 and sorted(seen) == sorted((float(p), z) for p in (2, 3, 4) for z in (1, 2, 3, 4, 5))
        assert [(o["ploidy_0"], o["numClusters"]) for o in out] == [(p, z) for p in (2, 3, 4) for z in (1, 2, 3, 4, 5)]


def test_chromosome_shards_cover_the_genome():
    data, _ = synth.make_data(5000, Z=2, maxCN=5, n_chr=23, seed=2)
    for world in (1, 2, 4, 8):
        shards, off, load = sweep.shard_chromosomes(data, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(23))
        assert load.sum() == 5000 and load.max() <= 1.3 * 5000 / world + 600


def _oracle_estep(sub, n, s, phi, var, piG, piZ, g, txn=1e15, zfac=1.0):
    Z, U = len(s), len(g["rt"])
    if sub["posn"].size == 0:
        return {"loglik": 0.0, "stats_flat": np.zeros(6 * U * Z), "rhoG0": np.zeros((U, 0)), "rhoZ0": np.zeros((Z, 0))}
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi)
    py = O.compute_py_c(sub["ref"], sub["tumDepth"], sub["logR"], muR, muC, var)
    chrsI = O.chromosome_slices(sub["chr"])
    es = O.e_step(sub, chrsI, py, piG, piZ, g["ZS"], g["ct"], Z, txn, zfac)
    st = O.sufficient_statistics(sub["ref"], sub["tumDepth"], sub["logR"], es["rho"])
    c0 = np.array([ix[0] for ix in chrsI])
    return {"loglik": es["loglik"], "stats_flat": np.concatenate([st[k].reshape(-1, order="F") for k in "abcdeg"]),
            "rhoG0": es["rhoG"][:, c0], "rhoZ0": es["rhoZ"][:, c0]}


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data, _ = synth.make_data(900, Z=2, maxCN=5, n_chr=5, seed=4, mean_seg=80)
        p = T.loadDefaultParameters(copyNumber=5, numberClonalClusters=2)
        cp = sweep.runEMclonalCN_sharded(data, p, rank, world, txnExpLen=1e15, txnZstrength=1.0, maxiter=2,
                                         estep_fn=_oracle_estep)
        out, info = sweep.run_sweep(data, rank, world, ploidies=(2, 3), clusters=(1, 2),
                                    run_one=lambda d, pp, **kw: {"phi_0": pp["ploidyParams"]["phi_0"],
                                                                 "K": pp["genotypeParams"]["rt"].size * len(pp["cellPrevParams"]["s_0"])})
        q.put((rank, cp["n"], cp["s"], cp["phi"], cp["loglik"], sorted((o["ploidy_0"], o["numClusters"]) for o in out)))
    finally:
        dist.destroy_process_group()


def test_sharded_em_equals_single_rank_em_world2():
    """2 ranks over gloo: the all-reduced statistics reproduce the single-process EM trajectory."""
    data, _ = synth.make_data(900, Z=2, maxCN=5, n_chr=5, seed=4, mean_seg=80)
    po = O.load_default_parameters(5, 2)
    want = O.run_em_clonal_cn(data, po, txnExpLen=1e15, txnZstrength=1.0, maxiter=2)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, n, s, phi, ll, sols in res:
        assert n.shape == want["n"].shape
        assert np.abs(n - want["n"]).max() < 1e-9 and np.abs(phi - want["phi"]).max() < 1e-9
        assert np.abs(s - want["s"]).max() < 1e-9
        assert np.allclose(ll[1:], want["loglik"][1:], rtol=1e-12)
        assert sols == [(2, 1), (2, 2), (3, 1), (3, 2)]          # every rank sees the gathered sweep
    assert np.array_equal(res[0][1], res[1][1])                   # replicated M-step: identical on both ranks
