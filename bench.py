#!/usr/bin/env python
"""bench.py -- TitanCNA E-step hot path on B200: fwd-bwd SNP*state*iters/s.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's own C on the host cores)

Workload = BASELINE.json configs[1]: synthetic genome-wide WGS, 1,000,000 heterozygous SNPs, 22+X chromosomes,
numClusters=3, ploidy=2, maxCN=8 (U=25, K=75), production transition settings txnExpLen=1e15, txnZstrength=1
(scripts/R_scripts/titanCNA.R:55-64).

A "step" is one E-step over the whole genome: emission -> scaled forward -> backward + posteriors -> the six
sufficient-statistic sums (reference: R/hmmClonal.R:160-171,191-234, src/fwd_backC_clonalCN.c, R/paramEstimation.R:27-44).
The steps walk through the parameter sets of a REAL production EM run on this data (untimed run first; step i uses the
parameters of EM iteration i mod I), so `value` is the mean over the E-steps that EM executes and no two consecutive
steps share parameters.
  value : N*K*steps / device time of the E-step kernels (CUDA events inside the library), data resident
  e2e   : the same through the C-ABI call titan_estep with HOST buffers: per step the parameter block goes host->device,
          the statistics come device->host, host-side means/logs are recomputed, and the call synchronises.  (The SNP
          columns are resident state, uploaded once by titan_solver_create: config.engine.upload_ms and e2e_cold.)
  roofline.frac     : whole-E-step algorithmic bytes (B_alg) / time / measured HBM peak
  roofline.fp64_frac: whole-E-step algorithmic flops (F_alg = 4K + 40) / time / MEASURED DFMA peak
With N>1 GPUs every rank runs the same solution shape on its own tumour sample (same generator, seed + 100 rank): independent
solutions, the path's natural shard, no data-path collective: weak scaling.  Two more records ride on the same line:
  sweep   : the 15-solution ploidy x numClusters sweep strong-scaled over the N ranks (LPT), EM + Viterbi per solution
  sharded : BASELINE configs[3] (2 M SNPs, K = 125) with its chromosomes sharded over the N ranks, the statistics
            all-reduced by NCCL inside the library once per E-step
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402

METRIC = "fwd-bwd SNP*state*iters/s"
UNIT = "SNP*state*iters/s"
SEED = 20261019 + 1
PLOIDIES = (2, 3, 4)
EM_KW = dict(txnExpLen=1e15, txnZstrength=1.0, maxiterUpdate=1500)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--snps", type=int, default=1_000_000)
    ap.add_argument("--clusters", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=None)
    ap.add_argument("--warm", type=int, default=None)
    ap.add_argument("--tol", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-em", action="store_true", help="skip the whole-solution record")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 15-solution sweep record")
    ap.add_argument("--no-sharded", action="store_true", help="skip the chromosome-sharded configs[3] record")
    ap.add_argument("--sweep-concurrent", type=int, default=4, help="solutions in flight per GPU in the sweep record")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="reference arm: CPU seconds for all its steps")
    return ap.parse_args()


def workload_config(a, n_gpus):
    """The same dictionary in both arms (the driver compares them)."""
    K = 25 * a.clusters
    return {"workload": f"synthetic WGS {a.snps} het SNPs, 23 chromosomes, numClusters={a.clusters}, ploidy_0=2, maxCN=8 "
                        f"(U=25, K={K}), txnExpLen=1e15, txnZstrength=1 [BASELINE.json configs[1]]",
            "snps": a.snps, "clusters": a.clusters, "K": K, "seed": SEED,
            "step": "one E-step; step i uses the parameters of iteration i mod I of a production EM run on this data",
            "parallelism": "1 solution per GPU (one tumour sample of this shape per rank, seed + 100 rank)" if n_gpus > 1 else "single solution",
            "l2": "working set (alpha + scaled emissions, K x N fp64 each = %.0f MB, streamed) exceeds the 126 MB L2" % (
                2 * a.snps * K * 8 / 1e6)}


def production_params(mod_load, data, Z, ploidy=2.0, moments=None):
    """scripts/R_scripts/titanCNA.R:235-247"""
    p = mod_load(8, Z, data=data) if moments is None else mod_load(8, Z, data=data, moments=moments)
    K = p["genotypeParams"]["rt"].size
    p["genotypeParams"]["alphaKHyper"] = np.full(K, 10000.0)
    p["genotypeParams"]["betaKHyper"] = np.full(K, 25.0)
    p["ploidyParams"]["phi_0"] = float(ploidy)
    p["normalParams"]["n_0"] = 0.5
    return p


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.1)               # first sample in flight before the timed region opens
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes of one whole E-step (sum over its launches) from the committed ncu capture, else None."""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json")), reverse=True):
        try:
            v = json.load(open(p)).get("estep", {}).get("dram_bytes_per_estep")
            if v:
                return v
        except Exception:
            continue
    return None


# ------------------------------------------------------------------------------------------------
# CPU arms (the ONLY place bench.py touches oracle/): the reference's own C when oracle/_ref holds
# it, else this repo's plain-C port; R-side emission / marginals / sums restated in oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_sample(data, per_chr):
    idx = np.concatenate([np.flatnonzero(data["chr"] == c)[:per_chr] for c in np.unique(data["chr"])])
    return {k: v[idx] for k, v in data.items()}


def cpu_estep(sample, Z, threads):
    from oracle import refshim, titan_oracle as O
    engine = "reference" if refshim.have_reference() else "port"
    p = production_params(O.load_default_parameters, sample, Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    chrsI = O.chromosome_slices(sample["chr"])

    def step():
        muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], 0.5, sP["s_0"], g["ct"], 2.0)
        py = O.compute_py_c(sample["ref"], sample["tumDepth"], sample["logR"], muR, muC, g["var_0"])
        es = O.e_step(sample, chrsI, py, g["piG_0"], sP["piZ_0"], g["ZS"], g["ct"], Z, 1e15, 1.0, engine=engine,
                      threads=threads)
        O.sufficient_statistics(sample["ref"], sample["tumDepth"], sample["logR"], es["rho"])
        return es["loglik"]

    return step, engine


def run_reference_arm(a, rank, world, emit):
    """The reference's own C (one chromosome per worker, its only parallel mode) on the host cores.  One step of the
    full configuration costs ~280 core-seconds (O(N K^2)) and its longest chromosome ~22 s on one thread, so the arm
    runs every step on the LARGEST per-chromosome prefix sample that keeps steps + warm-up inside --ref-budget-s
    (the whole genome when that fits, e.g. --steps 1 --warmup 0), and says which."""
    if rank != 0:
        return
    data, _ = synth.make_data(a.snps, Z=a.clusters, maxCN=8, n_chr=23, seed=SEED)
    cores = os.cpu_count() or 1
    threads = min(cores, 23)
    K = 25 * a.clusters
    sizes = np.bincount(data["chr"])[1:]
    # calibration (untimed): seconds per SNP on one thread for this K
    cal = cpu_sample(data, 150)
    step, engine = cpu_estep(cal, a.clusters, threads=threads)
    t0 = time.perf_counter()
    step()
    waves = int(np.ceil(23 / threads))
    per_snp_s = (time.perf_counter() - t0) / (150 * waves)
    n_calls = a.steps + a.warmup
    per_chr = int(max(50, min(sizes.max(), a.ref_budget_s / max(1, n_calls) / (per_snp_s * waves))))
    sample = cpu_sample(data, per_chr)
    Ns = sample["posn"].size
    step, engine = cpu_estep(sample, a.clusters, threads=threads)
    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    val = Ns * K * a.steps / dt
    whole = Ns == data["posn"].size
    sample_txt = ("the whole genome" if whole else f"first {per_chr} SNPs of each of the 23 chromosomes ({Ns} SNPs, "
                  f"{100.0 * Ns / data['posn'].size:.1f} % of the workload; O(N) algorithm, per-SNP rate)") + \
        f" per step, one chromosome per worker (the reference's foreach %dopar% mode), {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference" if engine == "reference" else "port",
                             "sample": sample_txt, "whole_workload": bool(whole)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
def em_trajectory(T, data, p, sol, maxiter=20):
    """Untimed production EM -> the parameter sets its E-steps ran with (column i-1 feeds E-step i)."""
    cp = T.runEMclonalCN(data, p, maxiter=maxiter, verbose=False, solver=sol, return_timing=True, **EM_KW)
    g = p["genotypeParams"]
    iters = cp["n"].size - 1
    traj = [(cp["n"][i], cp["s"][:, i].copy(), cp["phi"][i], cp["var"][:, i].copy(), cp["piG"][:, i].copy(),
             cp["piZ"][:, i].copy(), g) for i in range(iters)]
    return cp, traj


def main():
    # stdout carries the one JSON line and nothing else: native libraries (NCCL's version banner) write to file
    # descriptor 1 directly, so the descriptor itself is pointed at stderr until the line is printed
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    import titancna_b200 as T

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T.lib().titan_set_device(local)

    Z, K = a.clusters, 25 * a.clusters
    # rank 0 = the named workload; rank r > 0 = another tumour sample of the same shape (same generator, seed + 100 r):
    # independent solutions, fixed work per GPU, no data-path collective
    data, _ = synth.make_data(a.snps, Z=Z, maxCN=8, n_chr=23, seed=SEED + 100 * rank)
    N = data["posn"].size
    ploidy = 2
    p = production_params(T.loadDefaultParameters, data, Z, ploidy)
    g = p["genotypeParams"]

    t0 = time.perf_counter()
    sol = T.Solver(data, g, Z, 1e15, 1.0, device=local)
    upload_ms = (time.perf_counter() - t0) * 1e3
    if a.chunk is not None or a.warm is not None or a.tol is not None:
        sol.configure(1024 if a.chunk is None else a.chunk, 192 if a.warm is None else a.warm, 1e-8 if a.tol is None else a.tol)
    sol.set_stream(torch.cuda.current_stream().cuda_stream)
    fp64_peak, _ = T.measure_fp64_peak()

    cp, traj = em_trajectory(T, data, p, sol)
    I = len(traj)

    # parity guard, outside the timed region: the chunked scan at the library defaults against the plain sequential
    # recursion (one chunk per chromosome) at one of the timed parameter sets
    chk = sol.estep(*traj[min(2, I - 1)])
    seq = T.Solver(data, g, Z, 1e15, 1.0, device=local)
    seq.configure(0, 0, 1e-8)
    want = seq.estep(*traj[min(2, I - 1)])
    seq.close()
    ll_rel = float(np.max(np.abs(chk["loglik_chr"] - want["loglik_chr"]) / np.abs(want["loglik_chr"])))
    # a sum of w_t * posterior_t moves by at most max|d posterior| * sum_t |w_t| (the scale tests/test_gpu_fullsize.py uses)
    x, nx, lr = np.abs(data["ref"]), np.abs(data["tumDepth"] - data["ref"]), np.abs(data["logR"])
    scale = {"a": x.sum(), "b": nx.sum(), "c": lr.sum(), "d": 12.0 * N, "e": float(N), "g": (lr ** 2).sum()}
    st_rel = float(max(np.max(np.abs(chk["stats"][k] - want["stats"][k])) / scale[k] for k in "abcdeg"))
    if not (ll_rel <= 1e-9 and st_rel <= 1e-9):
        raise SystemExit(f"bench.py: chunked scan disagrees with the sequential recursion (loglik {ll_rel:.3g}, sums {st_rel:.3g})")

    for w in range(max(a.warmup, 3)):
        out = sol.estep(*traj[w % I])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = sol.timing()["kernel_launches"]
    phases = {k: [] for k in ("last_estep_ms", "last_emission_ms", "last_fwd_r0_ms", "last_rounds_ms", "last_bwd_r0_ms",
                              "last_reduce_ms", "last_fwd_rounds")}
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(a.steps):
        out = sol.estep(*traj[i % I])              # C-ABI call: host params in, host statistics out
        tm = sol.timing()
        for k in phases:
            phases[k].append(tm[k])
    ev1.record()
    barrier()
    wall_ms = ev0.elapsed_time(ev1)
    dev_ms = float(np.sum(phases["last_estep_ms"]))
    launches = sol.timing()["kernel_launches"] - launches0
    clocks = sampler.stop()
    if not np.isfinite(out["loglik"]):
        raise SystemExit("bench.py: non-finite log-likelihood")

    t = torch.tensor([dev_ms, wall_ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    rank_ms = [dev_ms / a.steps if world == 1 else None]
    if world > 1:      # every rank's own device time per step: shows whether ranks slow each other down or merely differ
        mine = torch.tensor([float(np.sum(phases["last_estep_ms"])) / a.steps], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_ms = [float(x[0]) for x in allr]
    units = float(N) * K * a.steps * world
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (wall_ms * 1e-3)

    # roofline of the whole E-step (rank 0's own times), SURVEY.md 8(d): B_alg = 16 + (96 + 8(U+Z))/K bytes and
    # F_alg = 4K + 40 flop per SNP*state*iteration
    peak, peak_src = measured_peak()
    U = 25
    b_alg = 16.0 + (96.0 + 8.0 * (U + Z)) / K
    f_alg = 4.0 * K + 40.0
    my_ms = float(np.mean(phases["last_estep_ms"]))
    achieved = N * K * b_alg / (my_ms * 1e-3) / 1e9
    mean = {k: float(np.mean(v)) for k, v in phases.items()}
    Kp = K + 1 + (K + 1) % 2
    r0_bytes = N * (2 * 8 * Kp + 8 * Kp + 2 * 128)     # scaled emissions read by both directions, alpha written, records twice
    fin_bytes = N * (2 * 8 * Kp + 128 + 32)            # scaled emissions + alpha read, both records
    em_bytes = N * (8 * Kp + 32)
    roofline = {"bound": "hbm", "kernel": "fb2_round_kernel (round 0 + re-runs: largest share of the E-step; frac is the WHOLE E-step)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                "peak_source": peak_src, "algorithmic_bytes_per_unit": b_alg, "algorithmic_bytes_per_estep": N * K * b_alg,
                "estep_ms": my_ms,
                "fp64_frac": N * K * f_alg / (my_ms * 1e-3) / 1e12 / fp64_peak, "fp64_peak_tflops_measured": fp64_peak,
                "algorithmic_flop_per_unit": f_alg,
                "phases_ms": {"emission": mean["last_emission_ms"], "round0": mean["last_fwd_r0_ms"], "verification": mean["last_rounds_ms"],
                              "final": mean["last_bwd_r0_ms"], "reduce": mean["last_reduce_ms"]},
                "launch_gbs": {"fb2_emission_kernel": em_bytes / (mean["last_emission_ms"] * 1e-3) / 1e9,
                               "fb2_round_kernel (round 0)": r0_bytes / (mean["last_fwd_r0_ms"] * 1e-3) / 1e9,
                               "fb2_final_kernel": fin_bytes / (mean["last_bwd_r0_ms"] * 1e-3) / 1e9}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, world),
            "engine": {"upload_ms": upload_ms, "chunks": sol.timing()["n_chunks"], "em_iterations": I,
                       "cycles_per_step": [int(x) for x in phases["last_fwd_rounds"]], "rank_ms_per_step": rank_ms,
                       "parity_guard": {"loglik_rel": ll_rel, "sums_err_over_scale": st_rel, "against": "sequential recursion, same kernels, one chunk per chromosome"}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (5 * K + 3 * 25) * 8,
                    "d2h_bytes_per_step": (6 * K + 23 + 23 * (25 + Z)) * 8, "ms_per_step": wall_ms / a.steps},
            "gpu_launches": launches, "roofline": roofline}

    if rank == 0:
        # cold call: host columns -> solver (upload + host-exact preprocessing) -> one E-step -> statistics
        t0 = time.perf_counter()
        cold = T.Solver(data, g, Z, 1e15, 1.0, device=local)
        cold.estep(*traj[0])
        cold_s = time.perf_counter() - t0
        cold.close()
        line["e2e_cold"] = {"value": N * K / cold_s, "unit": UNIT, "seconds": cold_s, "h2d_bytes": int(N * (32 + 16)),
                            "note": "titan_solver_create from host columns + first titan_estep"}
        if not a.no_em:
            # whole solution from host columns: solver creation (upload + host-exact preprocessing),
            # runEMclonalCN with the production settings, viterbiClonalCN on the same device-resident data
            t0 = time.perf_counter()
            whole = T.Solver(data, g, Z, 1e15, 1.0, device=local)
            full = T.runEMclonalCN(data, p, maxiter=20, verbose=False, return_timing=True, solver=whole, **EM_KW)
            em_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            path = T.viterbiClonalCN(data, full, solver=whole)
            vit_s = time.perf_counter() - t0
            vit_ms = whole.timing()["last_viterbi_ms"]
            whole.close()
            its = int(full["n"].size - 1)
            line["solution"] = {"em_iterations": its, "em_wall_s": em_s, "estep_device_ms": full["timing"]["estep_ms"],
                                "estep_device_ms_per_iteration": full["timing"]["estep_ms"] / its,
                                "mstep_host_ms": full["timing"]["mstep_ms"], "viterbi_wall_s": vit_s, "viterbi_device_ms": vit_ms,
                                "n": float(full["n"][-1]), "phi": float(full["phi"][-1]), "segments": int(1 + np.count_nonzero(np.diff(path)))}
    sol.close()

    # ---- the two real shards, every N: strong scaling of fixed total work ----
    if world > 1 and rank > 0 and not a.no_sweep:
        data, _ = synth.make_data(a.snps, Z=Z, maxCN=8, n_chr=23, seed=SEED)      # the sweep runs on ONE data set: rank 0's
    if not a.no_sweep:
        barrier()
        t0 = time.perf_counter()
        def run_one(d, pp, **kw):
            # the path of one solution: device-resident solver from host columns, production EM, Viterbi decode
            t1 = time.perf_counter()
            so = T.Solver(d, pp["genotypeParams"], len(pp["cellPrevParams"]["s_0"]), kw["txnExpLen"], kw["txnZstrength"], device=local)
            try:
                # the sweep keeps the parameter sets, the paths and the selection indices; the N-sized posterior
                # marginals stay on the device (T.run_sweep's own run_one does the same)
                c = T.runEMclonalCN(d, pp, verbose=False, solver=so, posteriors=False, **kw)
                pth = T.viterbiClonalCN(d, c, solver=so)
            finally:
                so.close()
            return {"iters": int(c["n"].size - 1), "loglik": float(c["loglik"][-1]), "segments": int(1 + np.count_nonzero(np.diff(pth))),
                    "seconds": time.perf_counter() - t1}

        mom = T.dataMoments(data)                  # inside the timed region: computed once per sweep, not once per solution
        res, info = T.run_sweep(data, rank, world, device=local, maxiter=20, run_one=run_one, concurrent=a.sweep_concurrent,
                                make_params=lambda pl, z: production_params(T.loadDefaultParameters, data, z, pl, mom), **EM_KW)
        barrier()
        sw = time.perf_counter() - t0
        if rank == 0:
            work = float(sum(N * 25.0 * r["numClusters"] * r["iters"] for r in res))
            line["sweep"] = {"solutions": len(res), "wall_s": sw, "value": work / sw, "unit": UNIT, "scaling": "strong",
                             "em_iterations_total": int(sum(r["iters"] for r in res)), "lpt_load_K": [float(x) for x in info["load"]],
                             "solution_seconds": [round(r["seconds"], 4) for r in res], "solutions_in_flight_per_gpu": a.sweep_concurrent,
                             "note": "ploidy_0 {2,3,4} x numClusters {1..5}: solver creation from host columns, production EM and Viterbi "
                                     "decode per solution (wall clock, barrier to barrier); CUDA context and input generation outside"}
    if not a.no_sharded:
        Z4 = 5
        data4, _ = synth.make_data(2 * a.snps, Z=Z4, maxCN=8, n_chr=23, seed=SEED + 2)
        p4 = production_params(T.loadDefaultParameters, data4, Z4, 2.0)
        comm = T.make_comm(rank, world)
        barrier()
        t0 = time.perf_counter()
        cp4 = T.runEMclonalCN_sharded(data4, p4, rank, world, maxiter=4, device=local, comm=comm, return_timing=True, **EM_KW)
        barrier()
        sh = time.perf_counter() - t0
        ver = comm.nccl_version()
        comm.close()
        its4 = int(cp4["n"].size - 1)
        est = torch.tensor([cp4["timing"]["estep_ms"], cp4["timing"]["estep_ms"] - cp4["timing"]["reduce_ms"]], dtype=torch.float64,
                           device="cuda")
        if world > 1:
            dist.all_reduce(est, op=dist.ReduceOp.MAX)
        if rank == 0:
            N4 = data4["posn"].size
            line["sharded"] = {"workload": f"{N4} SNPs, numClusters=5 (K=125), chromosomes LPT-sharded over {world} rank(s), "
                                           f"{its4} EM iterations [BASELINE.json configs[3], U=25 variant]",
                               "wall_s": sh, "estep_device_ms_max_rank": float(est[0]), "estep_compute_ms_max_rank": float(est[1]), "value": N4 * 125.0 * its4 / (float(est[0]) * 1e-3),
                               "unit": UNIT, "scaling": "strong", "collective": "ncclAllReduce(sum, f64) of %d doubles per E-step, inside titan_estep" % (
                                   6 * 125 + 23 * (1 + 25 + Z4)), "nccl_version": ver, "n": float(cp4["n"][-1]), "loglik": float(cp4["loglik"][-1])}

    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            per_chr = max(50, int(3000 * (75 / K) ** 2))
            sample = cpu_sample(data, per_chr)
            step, engine = cpu_estep(sample, Z, threads=1)
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample["posn"].size * K / dt, "unit": UNIT, "cores": 1,
                                    "kind": "reference" if engine == "reference" else "port",
                                    "sample": f"one E-step over the first {per_chr} SNPs of each chromosome "
                                              f"({sample['posn'].size} SNPs, K={K}), {dt:.1f} s on one host core"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
