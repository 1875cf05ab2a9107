"""GPU parity tests of the Gaussian allele model (alleleEmissionModel = "Gaussian"):
bivariate-normal emission on (ref / tumDepth, logR) -> the same forward-backward / Viterbi kernels -> the sums
a, f, c, h, e, g -> the Gaussian coordinate descent.  Everything through the C ABI, against the CPU oracle.

Tolerances as for the binomial model (BASELINE.json north_star): Viterbi paths bit-exact, per-chromosome
log-likelihood 1e-9 relative, posteriors and converged parameters 1e-6 (asserted tighter where possible).
"""
import ctypes as C
import math

import numpy as np
import pytest

import titancna_b200 as T
from oracle import titan_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu


def _model(N, Z, maxCN, n_chr, seed, extended=False):
    data, _ = synth.make_data(N, Z=Z, maxCN=min(maxCN, 8), n_chr=n_chr, seed=seed, mean_seg=250)
    p = T.loadDefaultParameters(copyNumber=maxCN, numberClonalClusters=Z, alleleEmissionModel="Gaussian", data=data,
                                extendedTable=extended)
    return data, p


def _point(p, Z):
    g = p["genotypeParams"]
    U = g["rt"].size
    n, phi, s = 0.33, 2.05, np.array(synth.S_TRUE[:Z]) * 0.9 + 0.02
    var = g["var_0"] * np.linspace(0.6, 1.4, U)
    varR = g["varR_0"] * np.linspace(0.2, 0.5, U)
    return n, phi, s, var, varR


CASES = [
    # N, Z, maxCN, n_chr, txnExpLen, txnZstrength, chunk_len, warm, extended
    (6000, 3, 8, 3, 1e15, 1.0, 512, 160, False),
    (6000, 3, 8, 3, 1e15, 1.0, 0, 1, False),          # plain sequential recursion
    (5000, 1, 5, 2, 1e15, 5e5, 256, 64, False),
    (4000, 5, 8, 2, 1e9, 1e9, 300, 100, False),
    (3000, 2, 10, 2, 1e15, 1.0, 256, 64, True),       # 36 genotype states: two per lane
]


@pytest.mark.parametrize("N,Z,maxCN,n_chr,txn,zfac,chunk,warm,ext", CASES)
def test_gaussian_estep_matches_oracle(N, Z, maxCN, n_chr, txn, zfac, chunk, warm, ext):
    data, p = _model(N, Z, maxCN, n_chr, 21 + Z, ext)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    n, phi, s, var, varR = _point(p, Z)
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi)
    py = O.compute_py_gauss(data["ref"], data["tumDepth"], data["logR"], muR, muC, varR, var, g["corRho_0"], 1e-300)
    chrsI = O.chromosome_slices(data["chr"])
    es = O.e_step(data, chrsI, py, g["piG_0"], sP["piZ_0"], g["ZS"], g["ct"], Z, txn, zfac, engine="port", threads=8)
    st = O.sufficient_statistics_gauss(data["ref"], data["tumDepth"], data["logR"], es["rho"])
    sol = T.Solver(data, g, Z, txn, zfac)
    assert sol.gauss
    sol.configure(chunk, warm, 1e-10)
    out = sol.estep(n, s, phi, var, g["piG_0"], sP["piZ_0"], g, posteriors=True, varR=varR)
    sol.close()
    assert np.allclose(out["loglik_chr"], es["loglik_chr"], rtol=1e-9, atol=0)
    assert np.abs(out["loglik_chr"] - es["loglik_chr"]).max() <= 1e-11 * np.abs(es["loglik_chr"]).max()
    assert np.abs(out["rhoG"] - es["rhoG"]).max() < 1e-9
    assert np.abs(out["rhoZ"] - es["rhoZ"]).max() < 1e-9
    assert np.abs(out["muR"] - muR).max() == 0 and np.abs(out["muC"] - muC).max() == 0
    assert sorted(out["stats"]) == sorted("acefgh")
    for k in "acefgh":
        assert np.abs(out["stats"][k] - st[k]).max() <= 1e-9 * max(1.0, np.abs(st[k]).max()), k


@pytest.mark.parametrize("vit_mode", ["-1", "0", "2", "3", "4"])
@pytest.mark.parametrize("N,Z,maxCN,n_chr,txn,zfac", [(5000, 3, 8, 3, 1e15, 1.0), (3000, 1, 5, 2, 1e15, 5e5), (2500, 5, 8, 2, 1e9, 1e9)])
def test_gaussian_viterbi_bit_exact(N, Z, maxCN, n_chr, txn, zfac, vit_mode, monkeypatch):
    """Every forward kernel on the exact bivariate-normal log emission (R's association order, no pseudo-count, the
    correlation handed over by the caller) against the reference C on the oracle's emission matrix."""
    monkeypatch.setenv("TITAN_VIT_MODE", vit_mode)
    data, p = _model(N, Z, maxCN, n_chr, 31 + Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    n, phi, s, var, varR = _point(p, Z)
    cor = O.bivariate_covariance(data["ref"] / data["tumDepth"], data["logR"])["cor"]
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi)
    py = O.compute_py_gauss(data["ref"], data["tumDepth"], data["logR"], muR, muC, varR, var, cor, 0.0)
    lp = np.vectorize(math.log)(np.outer(g["piG_0"], sP["piZ_0"]).reshape(-1, order="F"))
    from oracle import refshim
    eng = "reference" if refshim.have_reference() else "port"
    want = np.concatenate([O.viterbi(lp, py[:, ix], g["ZS"], Z, data["posn"][ix], zfac * txn, txn, 0, engine=eng, ct=g["ct"])
                           for ix in O.chromosome_slices(data["chr"])])
    sol = T.Solver(data, g, Z, txn, zfac)
    got = sol.viterbi(n, s, phi, var, g["piG_0"], sP["piZ_0"], g, varR=varR, corRho=cor)
    sol.close()
    assert np.array_equal(got, want)
    assert len(np.unique(want)) > 3


@pytest.mark.parametrize("Z,maxCN", [(2, 5), (3, 8)])
def test_gaussian_em_matches_oracle(Z, maxCN):
    """runEMclonalCN + viterbiClonalCN of the Gaussian model end to end against the oracle's restatement of
    R/hmmClonal.R:8-368,370-485 (E-steps by the reference-identical C port)."""
    data, p = _model(8000, Z, maxCN, 3, 41 + Z)
    po = O.load_default_parameters(maxCN, Z, alleleEmissionModel="Gaussian", data=data)
    want = O.run_em_clonal_cn(data, po, txnExpLen=1e15, txnZstrength=1.0, maxiter=4, engine="port", threads=8)
    got = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=4, verbose=False)
    I = want["n"].size
    assert got["n"].size == I and I >= 3
    assert np.abs(got["loglik"][1:] - want["loglik"][1:]).max() <= 1e-9 * np.abs(want["loglik"][1:]).max()
    for k in ("n", "phi", "s", "var", "varR", "piG", "piZ", "muR", "muC"):
        assert np.abs(np.asarray(got[k]) - np.asarray(want[k])).max() < 1e-6, k
    assert np.abs(got["varR"][:, -1] - p["genotypeParams"]["varR_0"]).max() > 1e-4        # the allelic variances moved
    assert np.abs(got["rhoG"] - want["rhoG"]).max() < 1e-6 and np.abs(got["rhoZ"] - want["rhoZ"]).max() < 1e-6
    assert list(got["cd_sweeps"][1:]) == list(want["cd_sweeps"])
    # decode from the ORACLE's converged parameters on both sides: bit-exact paths (ties as in the reference)
    wpath = O.viterbi_clonal_cn(data, want, engine="port", threads=8)
    gpath = T.viterbiClonalCN(data, dict(want))
    assert np.array_equal(gpath, wpath)
    # and the product's own chain gives the same calls on this data
    assert np.mean(T.viterbiClonalCN(data, got) == wpath) > 0.999


def test_gaussian_fused_r_glue():
    """titan_solver_create_gaussian_R / titan_em_run_R (hyper with varR_0, alphaRHyper, betaRHyper, corRho_0) /
    titan_viterbi_gaussian_R through the .Call emulation against the host mirror."""
    from oracle import refshim
    from titancna_b200 import build as B
    glue = refshim.ShimLibrary(B.GLUE)
    assert glue.registered(("titan_solver_create_gaussian_R", "titan_viterbi_gaussian_R")) == {
        "titan_solver_create_gaussian_R": 9, "titan_viterbi_gaussian_R": 12}
    data, p = _model(5000, 2, 5, 3, 51)
    g, nP, pP, sP = p["genotypeParams"], p["normalParams"], p["ploidyParams"], p["cellPrevParams"]
    want = T.runEMclonalCN(data, p, txnExpLen=1e9, txnZstrength=1.0, maxiter=3, verbose=False)
    keep = []
    R = lambda v: glue._real(np.atleast_1d(np.asarray(v, float)), keep)
    off = T.chromosome_offsets(data["chr"]).astype(float)
    solver = glue.call("titan_solver_create_gaussian_R", R(off), R(data["posn"]), R(data["ref"]), R(data["tumDepth"]), R(data["logR"]),
                       R(1e9), R(1.0), R(g["ZS"]), R(2))
    hyper = {k: g[k] for k in ("rt", "ct", "ZS", "rn", "var_0", "alphaKHyper", "betaKHyper", "kappaGHyper", "piG_0",
                               "varR_0", "alphaRHyper", "betaRHyper", "corRho_0")}
    hyper.update({k: sP[k] for k in ("s_0", "alphaSHyper", "betaSHyper", "kappaZHyper", "piZ_0")})
    hyper.update({k: nP[k] for k in ("n_0", "alphaNHyper", "betaNHyper")})
    hyper.update({k: pP[k] for k in ("phi_0", "alphaPHyper", "betaPHyper")})
    fit = glue.list_to_dict(glue.call("titan_em_run_R", solver, glue.named_list(hyper, keep), R(3), R(1500), R(1), R(1), R(1), R(0.001)))
    for k in ("n", "phi", "loglik", "s", "var", "varR", "piG", "piZ"):
        assert np.array_equal(fit[k], want[k]), k
    i = want["n"].size - 1
    cor = O.bivariate_covariance(data["ref"] / data["tumDepth"], data["logR"])["cor"]
    vp = glue.call("titan_viterbi_gaussian_R", solver, R(fit["n"][i]), R(fit["s"][:, i]), R(fit["phi"][i]), R(fit["var"][:, i]),
                   R(fit["piG"][:, i - 1]), R(fit["piZ"][:, i - 1]), R(g["rt"]), R(g["ct"]), R(g["rn"]), R(fit["varR"][:, i]), R(cor))
    o = vp.contents
    path = np.ctypeslib.as_array(C.cast(o.data, C.POINTER(C.c_int)), shape=(o.length,)).copy()
    sol = T.Solver(data, g, 2, 1e9, 1.0)
    wpath = sol.viterbi(want["n"][i], want["s"][:, i], want["phi"][i], want["var"][:, i], want["piG"][:, i - 1],
                        want["piZ"][:, i - 1], g, varR=want["varR"][:, i], corRho=cor)
    sol.close()
    assert np.array_equal(path, wpath)
    # a Gaussian solver refuses hyper-parameters without the allelic-variance fields
    bad = {k: v for k, v in hyper.items() if k != "alphaRHyper"}
    with pytest.raises(refshim.RError, match="alphaRHyper"):
        glue.call("titan_em_run_R", solver, glue.named_list(bad, keep), R(3), R(1500), R(1), R(1), R(1), R(0.001))
    glue.lib.shim_free_all()
    glue.gc()


def test_gaussian_model_argument_errors():
    data, p = _model(600, 2, 5, 1, 61)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    bad = dict(data)
    bad["tumDepth"] = data["tumDepth"].copy()
    bad["ref"] = data["ref"].copy()
    bad["tumDepth"][7] = 0.0
    bad["ref"][7] = 0.0
    with pytest.raises(T.TitanError):                     # ref / tumDepth is NaN: the reference's sums would be NA
        T.Solver(bad, g, 2, 1e15, 1.0)
    g2 = dict(g)
    g2["corRho_0"] = None                                 # loadDefaultParameters() without data (R/utils.R:79-83)
    sol = T.Solver(data, g2, 2, 1e15, 1.0)
    with pytest.raises(ValueError, match="corRho_0"):
        sol.estep(0.3, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], g2)
    with pytest.raises(T.TitanError, match="corRho"):
        sol.estep(0.3, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], g, corRho=1.0)
    sol.close()
    # a binomial solver ignores the Gaussian fields: same result with and without them
    pb = T.loadDefaultParameters(5, 2, data=data)
    gb = pb["genotypeParams"]
    sol = T.Solver(data, gb, 2, 1e15, 1.0)
    a = sol.estep(0.3, sP["s_0"], 2.0, gb["var_0"], gb["piG_0"], sP["piZ_0"], gb)
    sol.close()
    assert sorted(a["stats"]) == sorted("abcdeg")
