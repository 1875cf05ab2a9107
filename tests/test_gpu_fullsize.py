"""Full-size parity: the benchmarked configuration itself (BASELINE.json configs[1], 1 M SNPs, K = 75, library
defaults -- no configure() call) against the reference's own C (oracle/_ref, src/fwd_backC_clonalCN.c:40-211,
src/viterbiC_clonalCN.c:35-134) run on all host threads, one chromosome per worker.  ~20-40 s of host time.

Tolerances are north_star's: per-chromosome log-likelihood 1e-9 relative, posteriors 1e-6, Viterbi bit-exact.
"""
import math
import os

import numpy as np
import pytest

import bench
import titancna_b200 as T
from oracle import titan_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu

ENGINE = "reference" if os.path.exists(os.path.join(os.path.dirname(O.HERE), "oracle", "_ref", "libtitanref.so")) else "port"


def _reference_estep(data, g, Z, n, s, phi, var, piG, piZ, txn, zfac, idx=None):
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi)
    py = O.compute_py_c(data["ref"], data["tumDepth"], data["logR"], muR, muC, var)
    chrsI = O.chromosome_slices(data["chr"])
    es = O.e_step(data, chrsI, py, piG, piZ, g["ZS"], g["ct"], Z, txn, zfac, engine=ENGINE, threads=os.cpu_count() or 8)
    es["stats"] = O.sufficient_statistics(data["ref"], data["tumDepth"], data["logR"], es["rho"])
    lp = np.vectorize(math.log)(np.outer(piG, piZ).reshape(-1, order="F"))
    posn = np.asarray(data["posn"], float)
    paths = O._pmap(lambda ix: O.viterbi(lp, py[:, ix], g["ZS"], Z, posn[ix], zfac * txn, txn, 0, engine=ENGINE, ct=g["ct"]),
                    chrsI, os.cpu_count() or 8)
    path = np.empty(posn.size, dtype=np.int32)
    for ix, p in zip(chrsI, paths):
        path[ix] = p
    es["path"], es["chrsI"] = path, chrsI
    return es


def _stat_scales(data):
    x, nx, lr = np.abs(data["ref"]), np.abs(data["tumDepth"] - data["ref"]), np.abs(data["logR"])
    N = x.size
    return {"a": x.sum(), "b": nx.sum(), "c": lr.sum(), "d": 12.0 * N, "e": float(N), "g": (lr ** 2).sum()}


def _compare(out, path, es, data):
    ll, ref_ll = out["loglik_chr"], es["loglik_chr"]
    assert np.all(np.abs(ll - ref_ll) <= 1e-9 * np.abs(ref_ll)), np.abs(ll - ref_ll) / np.abs(ref_ll)
    assert np.abs(out["rhoG"] - es["rhoG"]).max() < 1e-6
    assert np.abs(out["rhoZ"] - es["rhoZ"]).max() < 1e-6
    chr0 = np.array([ix[0] for ix in es["chrsI"]])
    assert np.abs(out["rhoG0"] - es["rhoG"][:, chr0]).max() < 1e-6
    assert np.abs(out["rhoZ0"] - es["rhoZ"][:, chr0]).max() < 1e-6
    sc = _stat_scales(data)
    for k in "abcdeg":     # sum_t w_t * posterior_t: error scale max|d posterior| * sum_t |w_t| with max|d posterior| <= 1e-9
        assert np.abs(out["stats"][k] - es["stats"][k]).max() <= 1e-9 * sc[k], (k, np.abs(out["stats"][k] - es["stats"][k]).max() / sc[k])
    assert np.array_equal(path, es["path"]), f"{np.count_nonzero(path != es['path'])} Viterbi states differ"


def test_config2_full_size_defaults_match_reference_c():
    """BASELINE configs[1] as benchmarked: E-step (log-likelihoods, posteriors, six sums) and Viterbi at the
    parameters reached after two EM iterations, solver at its defaults."""
    Z = 3
    data, _ = synth.make_data(1_000_000, Z=Z, maxCN=8, n_chr=23, seed=bench.SEED)
    p = bench.production_params(T.loadDefaultParameters, data, Z)
    g = p["genotypeParams"]
    sol = T.Solver(data, g, Z, 1e15, 1.0)          # defaults: no configure()
    cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=2, maxiterUpdate=1500, verbose=False, solver=sol)
    i = cp["n"].size - 1
    args = (cp["n"][i], cp["s"][:, i], cp["phi"][i], cp["var"][:, i], cp["piG"][:, i], cp["piZ"][:, i])
    out = sol.estep(*args, g, posteriors=True)
    tm = sol.timing()
    path = sol.viterbi(*args, g)
    sol.close()
    assert tm["n_chunks"] > 900 and tm["last_fwd_rounds"] >= 2          # the chunked scan really ran
    es = _reference_estep(data, g, Z, *args, 1e15, 1.0)
    _compare(out, path, es, data)


def test_config4_three_full_chromosomes_k125_match_reference_c():
    """BASELINE configs[3] (2 M SNPs, numClusters = 5) in its U = 25 variant: chromosomes 20-22 at full length."""
    Z = 5
    full, _ = synth.make_data(2_000_000, Z=Z, maxCN=8, n_chr=23, seed=20261019 + 4)
    sel = np.isin(full["chr"], (20, 21, 22))
    data = {k: v[sel] for k, v in full.items()}
    p = bench.production_params(T.loadDefaultParameters, data, Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    n, phi, s = 0.33, 2.02, np.array([0.004, 0.29, 0.42, 0.55, 0.68])
    var = g["var_0"] * 0.3
    args = (n, s, phi, var, g["piG_0"], sP["piZ_0"])
    sol = T.Solver(data, g, Z, 1e15, 1.0)
    out = sol.estep(*args, g, posteriors=True)
    tm = sol.timing()
    path = sol.viterbi(*args, g)
    sol.close()
    assert tm["n_chunks"] > 60
    es = _reference_estep(data, g, Z, *args, 1e15, 1.0)
    _compare(out, path, es, data)
