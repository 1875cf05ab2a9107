"""Viterbi kernel comparison: device time per decode for each forward-kernel mode, paths cross-checked."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import titancna_b200 as T
from tests import synth
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--snps", type=int, default=1000000)
ap.add_argument("--clusters", type=int, nargs="+", default=[3])
ap.add_argument("--modes", type=str, nargs="+", default=["0", "2", "2:8"])
a = ap.parse_args()
for Z in a.clusters:
    data, _ = synth.make_data(a.snps, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
    p = bench.production_params(T.loadDefaultParameters, data, Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    args = (0.45, sP["s_0"], 2.04, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    ref = None
    for m in a.modes:
        mode, _, P = m.partition(":")
        os.environ["TITAN_VIT_MODE"] = mode
        if P:
            os.environ["TITAN_VIT_P"] = P
        else:
            os.environ.pop("TITAN_VIT_P", None)
        sol = T.Solver(data, g, Z, 1e15, 1.0)
        sol.estep(*args)
        ts = []
        for _ in range(3):
            path = sol.viterbi(*args)
            ts.append(sol.timing()["last_viterbi_ms"])
        ll = sol.estep(*args)["loglik"]
        same = "ref" if ref is None else str(bool(np.array_equal(ref, path)))
        if ref is None:
            ref = path
        print(f"Z={Z} K={25*Z} mode={m}: viterbi ms {min(ts):.2f} (runs {['%.1f' % t for t in ts]}) same_path={same} "
              f"segments={1 + np.count_nonzero(np.diff(path))} loglik_after={ll:.6f}", flush=True)
        sol.close()
