"""BASELINE config 5b: 50k-SNP exome case (small-N / latency regime). E-step, EM and Viterbi wall and device times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

for mode, N in (("exome", 50000), ("haplotype", 200000)):
    kw = {"exome": True} if mode == "exome" else {"haplotype_bins": True, "bin_size": 100000}
    data, _ = synth.make_data(N, Z=3, maxCN=8, n_chr=23, seed=bench.SEED + 5, **kw)
    p = bench.production_params(T.loadDefaultParameters, data, 3)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    t0 = time.perf_counter()
    sol = T.Solver(data, g, 3, 1e15, 1.0)
    t_create = time.perf_counter() - t0
    args = (0.45, sP["s_0"], 2.04, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    for _ in range(3):
        sol.estep(*args)
    t0 = time.perf_counter()
    for _ in range(10):
        out = sol.estep(*args)
    wall = (time.perf_counter() - t0) / 10
    tm = sol.timing()
    print(f"{mode} N={N}: create {t_create*1e3:.1f} ms; E-step wall {wall*1e3:.3f} ms, device {tm['last_estep_ms']:.3f} ms "
          f"(fwd r0 {tm['last_fwd_r0_ms']:.3f}, bwd r0 {tm['last_bwd_r0_ms']:.3f}, rounds {tm['last_fwd_rounds']}/{tm['last_bwd_rounds']}, "
          f"chunks {tm['n_chunks']}, launches/step {0})", flush=True)
    for chunk in (512, 256, 128):      # small-N regime: shorter chunks shorten every latency-bound pass
        sol.configure(chunk, min(192, chunk), 1e-8)
        for _ in range(3):
            sol.estep(*args)
        tm = sol.timing()
        print(f"   chunk {chunk}: E-step device {tm['last_estep_ms']:.3f} ms, cycles {tm['last_fwd_rounds']}, chunks {tm['n_chunks']}", flush=True)
    sol.configure(1024, 192, 1e-8)
    t0 = time.perf_counter()
    cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=20, verbose=False, solver=sol)
    t_em = time.perf_counter() - t0
    t0 = time.perf_counter()
    path = T.viterbiClonalCN(data, cp, solver=sol)
    t_v = time.perf_counter() - t0
    print(f"   EM {cp['n'].size - 1} iterations {t_em*1e3:.1f} ms; Viterbi wall {t_v*1e3:.1f} ms (device {sol.timing()['last_viterbi_ms']:.2f} ms); "
          f"segments {1 + np.count_nonzero(np.diff(path))}", flush=True)
    sol.close()
