#!/usr/bin/env python
"""bench.py -- TitanCNA E-step hot path on B200: fwd-bwd SNP*state*iters/s.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's own C on the host cores)

Workload = BASELINE.json configs[1]: synthetic genome-wide WGS, 1,000,000 heterozygous SNPs,
22+X chromosomes, numClusters=3, ploidy=2, maxCN=8 (U=25, K=75), production transition settings
txnExpLen=1e15, txnZstrength=1 (scripts/R_scripts/titanCNA.R:55-64).

A "step" is one E-step over the whole genome: emission -> scaled forward -> backward + posteriors ->
the six sufficient-statistic sums (reference: R/hmmClonal.R:160-171,191-234, src/fwd_backC_clonalCN.c,
R/paramEstimation.R:27-44), at the parameter values reached after two EM iterations.
  value : N*K*steps / device time of the E-step kernels (CUDA events inside the library), data resident
  e2e   : the same through the C-ABI call titan_estep with HOST buffers: per step the parameter block
          goes host->device, the statistics come device->host, host-side means/logs are recomputed, and
          the call synchronises.  (The SNP columns are resident state, uploaded once by
          titan_solver_create; that one-off cost is reported as config.upload_ms and as e2e_cold.)
With N>1 GPUs every rank runs an independent solution of the snakemake sweep on its own copy of the
data (no data-path collective): weak scaling.
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-em", action="store_true")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {"workload": f"synthetic WGS {a.snps} het SNPs, 23 chromosomes, numClusters={a.clusters}, ploidy_0=2, maxCN=8 "
                        f"(U=25, K={25 * a.clusters}), txnExpLen=1e15, txnZstrength=1 [BASELINE.json configs[1]]",
            "snps": a.snps, "clusters": a.clusters, "K": 25 * a.clusters, "seed": SEED,
            "parallelism": "1 solution per GPU (solution-grid shard)" if n_gpus > 1 else "single solution",
            "l2": "working set (alpha K x N fp64 = %.0f MB, streamed once per direction) exceeds the 126 MB L2" % (
                a.snps * 25 * a.clusters * 8 / 1e6)}


def production_params(mod_load, data, Z):
    """scripts/R_scripts/titanCNA.R:235-247"""
    p = mod_load(8, Z, data=data)
    K = p["genotypeParams"]["rt"].size
    p["genotypeParams"]["alphaKHyper"] = np.full(K, 10000.0)
    p["genotypeParams"]["betaKHyper"] = np.full(K, 25.0)
    p["ploidyParams"]["phi_0"] = 2.0
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
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
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
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("bwd_chunk_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            return None
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


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    data, _ = synth.make_data(a.snps, Z=a.clusters, maxCN=8, n_chr=23, seed=SEED)
    cores = os.cpu_count() or 1
    per_chr = max(50, int(1200 * (75 / (25 * a.clusters)) ** 2))
    sample = cpu_sample(data, per_chr)
    Ns = sample["posn"].size
    K = 25 * a.clusters
    step, engine = cpu_estep(sample, a.clusters, threads=min(cores, 23))
    for _ in range(min(a.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    val = Ns * K * a.steps / dt
    sample_txt = (f"first {per_chr} SNPs of each of the 23 chromosomes ({Ns} SNPs) per step, one chromosome per worker "
                  f"(the reference's foreach %dopar% mode), {min(cores, 23)} threads")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": min(a.warmup, 1), "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": min(cores, 23), "kind": "reference" if engine == "reference" else "port",
                             "sample": sample_txt},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
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
    # every rank = one independent solution on its own copy of the data: the workload `config` names, on every
    # rank (weak scaling: per-GPU work fixed as N grows)
    data, _ = synth.make_data(a.snps, Z=Z, maxCN=8, n_chr=23, seed=SEED)
    N = data["posn"].size
    p = production_params(T.loadDefaultParameters, data, Z)
    p["ploidyParams"]["phi_0"] = 2.0
    g, sP = p["genotypeParams"], p["cellPrevParams"]

    t0 = time.perf_counter()
    sol = T.Solver(data, g, Z, 1e15, 1.0, device=local)
    upload_ms = (time.perf_counter() - t0) * 1e3
    if a.chunk is not None or a.warm is not None:
        sol.configure(1024 if a.chunk is None else a.chunk, 192 if a.warm is None else a.warm, 1e-8)
    sol.set_stream(torch.cuda.current_stream().cuda_stream)

    # representative parameters: two EM iterations from the production initial values (untimed)
    cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=2, maxiterUpdate=1500, verbose=False,
                         solver=sol)
    i = cp["n"].size - 1
    args = (cp["n"][i], cp["s"][:, i], cp["phi"][i], cp["var"][:, i], cp["piG"][:, i], cp["piZ"][:, i], g)

    for _ in range(max(a.warmup, 3)):
        out = sol.estep(*args)
    if not np.isfinite(out["loglik"]):
        raise SystemExit("bench.py: non-finite log-likelihood")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = sol.timing()["kernel_launches"]
    dev_ms, fwd0, bwd0, frounds, brounds = 0.0, [], [], [], []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        out = sol.estep(*args)              # C-ABI call: host params in, host statistics out
        tm = sol.timing()
        dev_ms += tm["last_estep_ms"]
        fwd0.append(tm["last_fwd_r0_ms"]); bwd0.append(tm["last_bwd_r0_ms"])
        frounds.append(tm["last_fwd_rounds"]); brounds.append(tm["last_bwd_rounds"])
    ev1.record()
    barrier()
    wall_ms = ev0.elapsed_time(ev1)
    launches = sol.timing()["kernel_launches"] - launches0
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, wall_ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    units = float(N) * K * a.steps * world
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (wall_ms * 1e-3)

    # roofline of the dominant kernel (bwd_chunk_kernel, round-0 launch: every chunk runs)
    peak, peak_src = measured_peak()
    bwd_ms = float(np.mean(bwd0))
    fwd_ms = float(np.mean(fwd0))
    bwd_bytes = N * (K * 8 + 96)            # alpha read + SnpRec (32 B) + TxnRec (64 B) per SNP
    estep_bytes = N * K * (16 + (96 + 96) / K)
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": f"bwd_chunk_kernel<{Z},false,false> (round 0)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bwd_bytes, "launch_ms": bwd_ms,
                "fwd_kernel": {"achieved": N * (K * 8 + 96) / (fwd_ms * 1e-3) / 1e9, "launch_ms": fwd_ms},
                "estep": {"achieved": estep_bytes * a.steps / (dev_ms * 1e-3) / 1e9 if world == 1 else None,
                          "algorithmic_bytes_per_unit": estep_bytes / (N * K)}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(workload_config(a, world), upload_ms=upload_ms,
                                                                 chunks=sol.timing()["n_chunks"],
                                                                 fwd_rounds=int(np.max(frounds)), bwd_rounds=int(np.max(brounds))),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (5 * K + 3 * 25) * 8,
                    "d2h_bytes_per_step": (6 * K + 23 + 23 * (25 + Z)) * 8, "ms_per_step": wall_ms / a.steps},
            "gpu_launches": launches, "roofline": roofline}

    if rank == 0:
        # cold call: host columns -> solver (upload + host-exact preprocessing) -> one E-step -> statistics
        t0 = time.perf_counter()
        cold = T.Solver(data, g, Z, 1e15, 1.0, device=local)
        cold.estep(*args)
        cold_s = time.perf_counter() - t0
        cold.close()
        line["e2e_cold"] = {"value": N * K / cold_s, "unit": UNIT, "seconds": cold_s, "h2d_bytes": int(N * (32 + 16)),
                            "note": "titan_solver_create from host columns + first titan_estep"}
        if not a.no_em:
            # whole solution from host columns: solver creation (upload + host-exact preprocessing),
            # runEMclonalCN with the production settings, viterbiClonalCN on the same device-resident data
            t0 = time.perf_counter()
            whole = T.Solver(data, g, Z, 1e15, 1.0, device=local)
            full = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=20, maxiterUpdate=1500, verbose=False,
                                   return_timing=True, solver=whole)
            em_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            path = T.viterbiClonalCN(data, full, solver=whole)
            vit_s = time.perf_counter() - t0
            vit_ms = whole.timing()["last_viterbi_ms"]
            whole.close()
            line["solution"] = {"em_iterations": int(full["n"].size - 1), "em_wall_s": em_s, "estep_device_ms": full["timing"]["estep_ms"],
                                "mstep_host_ms": full["timing"]["mstep_ms"], "viterbi_wall_s": vit_s, "viterbi_device_ms": vit_ms,
                                "n": float(full["n"][-1]), "phi": float(full["phi"][-1]), "segments": int(1 + np.count_nonzero(np.diff(path)))}
        if world == 1 and not a.no_cpu_baseline:
            per_chr = max(50, int(600 * (75 / K) ** 2))
            sample = cpu_sample(data, per_chr)
            step, engine = cpu_estep(sample, Z, threads=1)
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample["posn"].size * K / dt, "unit": UNIT, "cores": 1,
                                    "kind": "reference" if engine == "reference" else "port",
                                    "sample": f"one E-step over the first {per_chr} SNPs of each chromosome "
                                              f"({sample['posn'].size} SNPs, K={K}), {dt:.1f} s on one host core"}
        print(json.dumps(line))
    sol.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
