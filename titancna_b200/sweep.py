"""Multi-GPU sharding of the path along its two natural shards.

1. Solution grid (ploidy_0 x numClusters; scripts/snakemake/TitanCNA.snakefile:8-9,14): independent
   EM runs on the same data, one process per GPU, longest-processing-time assignment by cost ~ K,
   NO data-path collective -- only a gather of the final parameter sets.
2. Chromosomes within one solution (the reference's foreach %dopar%, R/hmmClonal.R:203): every rank
   holds a subset of the chromosomes; per EM iteration ONE all-reduce(sum, fp64) of
   [a,b,c,d,e,g (6*U*Z) | loglik | rhoG@chrPos_0 (U*nChr) | rhoZ@chrPos_0 (Z*nChr)], then the O(U*Z)
   M-step is replicated on every rank (deterministic, so no broadcast).

torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_double_p, check, lib
from .hmm import Solver, _check_params, _hyper, chromosome_offsets, runEMclonalCN, viterbiClonalCN
from .params import dataMoments, loadDefaultParameters
from .results import correctIntegerCN, outputTitanResults, outputTitanSegments
from .selection import outputModelParameters, selectSolution


def lpt_assign(costs, n_bins):
    """Longest-processing-time-first: -> (bin index per item, load per bin)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(n_bins)
    where = np.zeros(costs.size, dtype=np.int64)
    for i in order:
        b = int(np.argmin(load))
        where[i] = b
        load[b] += costs[i]
    return where, load


def solution_grid(ploidies=(2, 3, 4), clusters=(1, 2, 3, 4, 5)):
    """The sweep of TitanCNA.snakefile:8-9: every (ploidy_0, numClusters) pair, cost ~ K = 25*Z."""
    grid = [(p, z) for p in ploidies for z in clusters]
    return grid, np.array([25.0 * z for _, z in grid])


def shard_chromosomes(data, world):
    """Contiguous chromosome runs -> per-rank lists of chromosome indices (LPT by SNP count)."""
    off = chromosome_offsets(data["chr"])
    sizes = np.diff(off)
    where, load = lpt_assign(sizes, world)
    return [np.flatnonzero(where == r) for r in range(world)], off, load


def subset_data(data, off, chrs):
    idx = np.concatenate([np.arange(off[c], off[c + 1]) for c in chrs]) if len(chrs) else np.zeros(0, dtype=np.int64)
    return {k: np.asarray(v)[idx] for k, v in data.items()}, idx


def _all_reduce_sum(vec, group=None):
    """fp64 sum over ranks; NCCL needs the tensor on the GPU, gloo on the CPU."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return vec
    backend = dist.get_backend(group)
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
    if backend == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def make_comm(rank, world, group=None):
    """A library communicator (NCCL inside libtitancna_b200) over the ranks of a torch.distributed group: rank 0 draws
    the 128-byte id, torch.distributed ships it (plumbing only), every rank joins on its current CUDA device."""
    import torch.distributed as dist
    ids = [_lib.Comm.unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0, group=group)
    return _lib.Comm(ids[0], rank, world)


def host_mstep(params, U, Z, nChr, maxiterUpdate, normalEstimateMethod, estimateS, estimatePloidy, stats_flat, rhoG0,
               rhoZ0, prev):
    """titan_mstep (host C++ coordinate descent + priors) -> dict(n, s, var, phi, piG, piZ, sweeps, prior)."""
    keep = []
    h = _hyper(params, keep)
    P = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    st, g0, z0 = P(stats_flat), P(np.asarray(rhoG0).reshape(-1, order="F")), P(np.asarray(rhoZ0).reshape(-1, order="F"))
    ps, pv, pg, pz = P(prev["s"]), P(prev["var"]), P(prev["piG"]), P(prev["piZ"])
    out_s, out_var, out_piG, out_piZ = np.zeros(Z), np.zeros(U), np.zeros(U), np.zeros(Z)
    n, phi, prior = C.c_double(), C.c_double(), C.c_double()
    sweeps = C.c_int()
    D = lambda a: a.ctypes.data_as(c_double_p)
    check(lib().titan_mstep(C.byref(h), U, Z, nChr, int(maxiterUpdate), int(normalEstimateMethod == "map"),
                            int(bool(estimateS)), int(bool(estimatePloidy)), D(st), D(g0), D(z0), float(prev["n"]), D(ps),
                            D(pv), float(prev["phi"]), D(pg), D(pz), C.byref(n), D(out_s), D(out_var), C.byref(phi),
                            D(out_piG), D(out_piZ), C.byref(sweeps), C.byref(prior)))
    return {"n": n.value, "s": out_s, "var": out_var, "phi": phi.value, "piG": out_piG, "piZ": out_piZ,
            "sweeps": sweeps.value, "prior": prior.value}


def runEMclonalCN_sharded(data, params, rank, world, txnExpLen=1e15, txnZstrength=5e5, maxiter=15, maxiterUpdate=1500,
                          normalEstimateMethod="map", estimateS=True, estimatePloidy=True, likChangeThreshold=0.001,
                          group=None, estep_fn=None, device=None, verbose=False, comm=None, solver=None,
                          return_timing=False):
    """runEMclonalCN with one solution's chromosomes sharded over `world` ranks.

    On GPUs (default) every rank builds a Solver over its own chromosomes, attaches a library communicator
    (titan_solver_set_comm) and runs the ordinary in-library EM loop: the ONE collective of an iteration is an
    ncclAllReduce of the statistics vector on the device, inside titan_estep; the M-step is replicated.
    The N-sized outputs (rhoG, rhoZ) cover this rank's chromosomes (`shard`, `shard_index`).

    `estep_fn(sub_data, n, s, phi, var, piG, piZ, genotypeParams) -> dict(loglik, stats_flat, rhoG0, rhoZ0)` lets the CPU
    tests stand in for the GPU E-step; that path keeps the loop and the all-reduce on the host (torch.distributed)."""
    _check_params(params)
    g, pP, nP, sP = (params[k] for k in ("genotypeParams", "ploidyParams", "normalParams", "cellPrevParams"))
    U, Z = len(g["rt"]), len(sP["s_0"])
    K = U * Z
    shards, off, _ = shard_chromosomes(data, world)
    nChr = off.size - 1
    mine = shards[rank]
    sub, sub_idx = subset_data(data, off, mine)
    if estep_fn is None:
        if len(mine) == 0:
            raise ValueError("more ranks than chromosomes: every rank of a sharded solution needs at least one chromosome")
        own_solver, own_comm = solver is None, comm is None
        if own_solver:
            solver = Solver(sub, g, Z, txnExpLen, txnZstrength, device=device)
        if own_comm:
            comm = make_comm(rank, world, group)
        try:
            solver.set_comm(comm, nChr, mine)
            cp = runEMclonalCN(sub, params, txnExpLen=txnExpLen, txnZstrength=txnZstrength, maxiter=maxiter,
                               maxiterUpdate=maxiterUpdate, normalEstimateMethod=normalEstimateMethod, estimateS=estimateS,
                               estimatePloidy=estimatePloidy, likChangeThreshold=likChangeThreshold, verbose=verbose and rank == 0,
                               solver=solver, return_timing=return_timing)
        finally:
            if own_solver:
                solver.close()
            else:
                solver.set_comm(None, 0, None)
            if own_comm:
                comm.close()
        cp["shard"], cp["shard_index"] = mine, sub_idx
        return cp
    solver = None            # host-side loop below: stand-in E-step, all-reduce through torch.distributed (CPU tests)
    M = int(maxiter) + 1
    cur = {"n": float(nP["n_0"]), "s": np.array(sP["s_0"], float), "var": np.array(g["var_0"], float),
           "phi": float(pP["phi_0"]), "piG": np.array(g["piG_0"], float), "piZ": np.array(sP["piZ_0"], float)}
    hist = {k: [np.copy(v)] for k, v in cur.items()}
    loglik = [-np.inf]
    converged, i = False, 0
    try:
        while not converged and (i + 1) < M:
            i += 1
            o = estep_fn(sub, cur["n"], cur["s"], cur["phi"], cur["var"], cur["piG"], cur["piZ"], g)
            vec = np.zeros(6 * K + 1 + (U + Z) * nChr)
            vec[:6 * K] = o["stats_flat"]
            vec[6 * K] = o["loglik"]
            g0 = np.zeros((U, nChr))
            z0 = np.zeros((Z, nChr))
            g0[:, mine] = o["rhoG0"]
            z0[:, mine] = o["rhoZ0"]
            vec[6 * K + 1:6 * K + 1 + U * nChr] = g0.reshape(-1, order="F")
            vec[6 * K + 1 + U * nChr:] = z0.reshape(-1, order="F")
            vec = _all_reduce_sum(vec, group)                       # the ONE collective of the iteration
            stats = vec[:6 * K]
            ll = float(vec[6 * K])
            g0 = vec[6 * K + 1:6 * K + 1 + U * nChr].reshape((U, nChr), order="F")
            z0 = vec[6 * K + 1 + U * nChr:].reshape((Z, nChr), order="F")
            new = host_mstep(params, U, Z, nChr, maxiterUpdate, normalEstimateMethod, estimateS, estimatePloidy, stats, g0,
                             z0, cur)
            loglik.append(ll + new["prior"])
            for k in cur:
                hist[k].append(np.copy(new[k]))
            if verbose and rank == 0:
                print(f"sharded EM iter {i}: loglik={loglik[i]:.4f} n={new['n']:.4f} phi={new['phi']:.4f}")
            if abs(loglik[i] - loglik[i - 1]) / abs(loglik[i]) <= likChangeThreshold and loglik[i] >= loglik[i - 1]:
                converged = True
                cur = {k: new[k] for k in cur}
            elif loglik[i] < loglik[i - 1]:
                converged = True
                i -= 1
            else:
                cur = {k: new[k] for k in cur}
    finally:
        if solver is not None:
            solver.close()
    m = i + 1
    return {"n": np.array([float(v) for v in hist["n"][:m]]), "phi": np.array([float(v) for v in hist["phi"][:m]]),
            "s": np.stack(hist["s"][:m], axis=1), "var": np.stack(hist["var"][:m], axis=1),
            "piG": np.stack(hist["piG"][:m], axis=1), "piZ": np.stack(hist["piZ"][:m], axis=1),
            "loglik": np.array(loglik[:m]), "txnExpLen": txnExpLen, "txnZstrength": txnZstrength,
            "useOutlierState": False, "genotypeParams": g, "ploidyParams": pP, "normalParams": nP, "cellPrevParams": sP,
            "shard": mine}


def run_sweep(data, rank, world, ploidies=(2, 3, 4), clusters=(1, 2, 3, 4, 5), maxCN=8, device=None, make_params=None,
              run_one=None, gather=True, concurrent=4, keep_path=False, **em_kwargs):
    """The snakemake sweep in one job: rank r runs the solutions LPT assigns to it and the final parameter sets
    are gathered (all_gather_object) -- no collective on the data path.  `run_one(data, params, **em_kwargs)` defaults
    to runEMclonalCN + viterbiClonalCN on this rank's GPU followed by the result table, segments and S_Dbw index;
    `selectSolution(out)` then applies the reference's selection rule to the gathered records."""
    grid, costs = solution_grid(ploidies, clusters)
    where, load = lpt_assign(costs, world)
    mine = [grid[i] for i in range(len(grid)) if where[i] == rank]
    if device is None and run_one is None and concurrent > 1 and len(mine) > 1:
        # worker threads start on CUDA device 0, not on the caller's: pin the solvers to the device this rank is on
        device = _lib.current_device()
    if make_params is None:
        moments = dataMoments(data)          # the data-dependent defaults once, not once per solution

        def make_params(ploidy, Z):
            p = loadDefaultParameters(copyNumber=maxCN, numberClonalClusters=Z, data=data, moments=moments)
            p["ploidyParams"]["phi_0"] = float(ploidy)
            return p
    if run_one is None:
        def run_one(d, p, **kw):
            # one device-resident solver per solution: the SNP columns are uploaded once and the Viterbi
            # transition tables are built in the background while the EM runs
            sol = Solver(d, p["genotypeParams"], len(p["cellPrevParams"]["s_0"]), kw.get("txnExpLen", 1e15),
                         kw.get("txnZstrength", 5e5), device=device)
            try:
                kw.setdefault("posteriors", False)      # nothing below reads the N-sized marginals (posteriorProbs = FALSE)
                cp = runEMclonalCN(d, p, verbose=False, solver=sol, **kw)
                path = viterbiClonalCN(d, cp, solver=sol)
            finally:
                sol.close()
            # what scripts/R_scripts/titanCNA.R:263-285 derives per solution: corrected result table, segments and
            # the S_Dbw validity index that selectSolution.R ranks the cluster counts by
            res = outputTitanResults(d, cp, path, proportionThreshold=0.05, proportionThresholdClonal=0.05,
                                     recomputeLogLik=False, verbose=False)
            cpc = res["convergeParams"]
            segs = outputTitanSegments(res["corrResults"], "sweep", cpc)
            chrs = list(dict.fromkeys(np.asarray(d["chr"]).astype(str).tolist()))
            corr = correctIntegerCN(res["corrResults"], segs, 1 - float(cpc["n"][-1]), float(cpc["phi"][-1]),
                                    maxCNtoCorrect_autosomes=maxCN, minPurityToCorrect=0.2, gender="male", chrs=chrs)
            segs = corr["segs"]
            mp = outputModelParameters(cpc, corr["cn"])
            out = {"n": cp["n"][-1], "phi": cp["phi"][-1], "s": cp["s"][:, -1], "loglik": cp["loglik"][-1],
                   "iters": cp["n"].size - 1, "sdbw": mp["S_Dbw"], "dens_bw": mp["dens_bw"], "scat": mp["scat"],
                   "segs": {"TITAN_call": segs["TITAN_call"], "Length.snp.": segs["Length.snp."]}}
            if keep_path:
                out["path"] = path
            return out
    # `concurrent` > 1 runs that many of this rank's solutions at once (one host thread and one pair of CUDA
    # streams each; ctypes releases the GIL): the verification rounds of one solution are latency-bound and the
    # decode keeps one CTA per chromosome busy, so most SMs are idle for most of a solution and other solutions'
    # launches fill them.  Measured on one B200, 15 solutions at 1 M SNPs: 0.99 s with one solution in flight,
    # 0.82 / 0.70 / 0.65 s with two / three / four, identical results (every solution's arithmetic is its own).
    def one(pz):
        ploidy, Z = pz
        return {"ploidy_0": ploidy, "numClusters": Z, **run_one(data, make_params(ploidy, Z), **em_kwargs)}

    if concurrent > 1 and len(mine) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=int(concurrent)) as ex:
            out = list(ex.map(one, mine))
    else:
        out = [one(pz) for pz in mine]
    if gather and world > 1:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            allres = [None] * world
            dist.all_gather_object(allres, out)
            out = [r for part in allres for r in part]
    return out, {"assignment": where, "load": load}
