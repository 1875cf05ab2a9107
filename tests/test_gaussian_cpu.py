"""CPU: the Gaussian allele model (alleleEmissionModel = "Gaussian") -- oracle restatement and host M-step.

The reference holds no fixture for this model (its vignette calls it "not fully implemented") and R cannot run here, so
the oracle's restatement is pinned by what CAN be checked without R:
  * the emission against an independent implementation of the bivariate normal density (scipy),
  * the update equations against numerical gradients of the expected complete-data log-likelihood they were derived from
    (the reference's own likelihoodFunc carries two slips -- log(1 / 2 * pi * ...) and an extra /2 on the quadratic
    terms, R/paramEstimation.R:321-324 -- restated as written, so it is NOT what the derivatives are checked against),
  * the product's C++ M-step against the oracle's.
"""
import ctypes
import math

import numpy as np
import pytest

import titancna_b200 as T
from titancna_b200 import _lib, build as B
from oracle import titan_oracle as O
from tests import synth


@pytest.fixture(scope="module", autouse=True)
def built():
    B.build()


def _setup(Z=2, maxCN=5, N=1500, seed=5):
    data, _ = synth.make_data(N, Z=Z, maxCN=maxCN, n_chr=2, seed=seed, mean_seg=100)
    po = O.load_default_parameters(maxCN, Z, alleleEmissionModel="Gaussian", data=data)
    rng = np.random.default_rng(seed)
    U = len(po["genotypeParams"]["rt"])
    rho = rng.dirichlet(np.full(U * Z, 0.05), size=N).T.reshape((U, Z, N), order="F")
    st = O.sufficient_statistics_gauss(data["ref"], data["tumDepth"], data["logR"], rho)
    return data, po, rho, st


def test_default_parameters_of_the_gaussian_model_match_oracle():
    data, po, _, _ = _setup()
    pp = T.loadDefaultParameters(5, 2, alleleEmissionModel="Gaussian", data=data)
    g, go = pp["genotypeParams"], po["genotypeParams"]
    assert g["alleleEmissionModel"] == go["alleleEmissionModel"] == "Gaussian"
    for k in ("varR_0", "alphaRHyper", "betaRHyper", "alphaKHyper", "rt"):
        assert np.allclose(g[k], go[k], rtol=1e-13, atol=0), k
    assert g["corRho_0"] == pytest.approx(go["corRho_0"], rel=1e-12)
    # the correlation the Viterbi wrapper re-derives (R/hmmClonal.R:536-538,561-567) is the same number up to rounding
    cor = O.bivariate_covariance(data["ref"] / data["tumDepth"], data["logR"])["cor"]
    assert cor == pytest.approx(go["corRho_0"], rel=1e-12)
    with pytest.raises(ValueError, match="binomial"):
        T.loadDefaultParameters(5, 2, alleleEmissionModel="Student")


def test_gaussian_emission_is_the_bivariate_normal_log_density():
    """computeBivariateNormalObslik / bivariateNormalpdflog (R/hmmClonal.R:525-559) against scipy's density."""
    from scipy.stats import multivariate_normal
    data, po, _, _ = _setup(N=300)
    g = po["genotypeParams"]
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], 0.3, np.array([0.01, 0.4]), g["ct"], 2.1)
    rng = np.random.default_rng(2)
    varR, var, rho = rng.uniform(0.005, 0.05, muR.shape[0]), rng.uniform(0.01, 0.2, muR.shape[0]), -0.23
    py = O.compute_py_gauss(data["ref"], data["tumDepth"], data["logR"], muR, muC, varR, var, rho, 0.0)
    x = np.stack([data["ref"] / data["tumDepth"], data["logR"]], axis=1)
    U, Z = muR.shape
    for k in (0, 3, U - 1):
        for z in range(Z):
            cov = rho * math.sqrt(varR[k] * var[k])
            want = multivariate_normal(mean=[muR[k, z], muC[k, z]], cov=[[varR[k], cov], [cov, var[k]]]).logpdf(x)
            assert np.allclose(py[k + z * U], want, rtol=1e-11, atol=1e-11)
    # NA handling of :557 and the pseudo-count of :158
    d0 = dict(data)
    d0["tumDepth"] = data["tumDepth"].copy()
    d0["ref"] = data["ref"].copy()
    d0["tumDepth"][5] = 0.0
    d0["ref"][5] = 0.0
    py0 = O.compute_py_gauss(d0["ref"], d0["tumDepth"], d0["logR"], muR, muC, varR, var, rho, 1e-300)
    assert np.all(py0[:, 5] == 1.0)


def _true_q(st, g, n, s, var, varR, phi):
    """Expected complete-data log-likelihood of the bivariate-normal model from the eight sums."""
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi)
    rho = g["corRho_0"]
    l0 = 1 / (2 * (1 - rho * rho))
    a, c, e, f, gg, h = (st[k] for k in "acefgh")
    v, vr = np.asarray(var)[:, None], np.asarray(varR)[:, None]
    q = -e * np.log(2 * math.pi * np.sqrt(v * vr * (1 - rho * rho))) - l0 * (
        (f - 2 * a * muR + e * muR ** 2) / vr + (gg - 2 * c * muC + e * muC ** 2) / v
        - 2 * rho * (h - a * muC - c * muR + e * muC * muR) / np.sqrt(v * vr))
    return float(q.sum())


def test_gaussian_update_equations_are_gradients_of_the_expected_log_likelihood():
    """R/paramEstimation.R:376-383 (n), :415-422 (s), :463-497 (var, varR), :510-515 (phi; its single-subscript
    indexing is exact only for one cluster, so phi is checked at Z = 1)."""
    data, po, rho_, st = _setup()
    g, nP, sP, pP = po["genotypeParams"], po["normalParams"], po["cellPrevParams"], po["ploidyParams"]
    rt, rn, ct, rho = g["rt"], g["rn"], g["ct"], g["corRho_0"]
    rng = np.random.default_rng(4)
    U = rt.size
    n, s, phi = 0.35, np.array([0.02, 0.45]), 1.8
    var, varR = rng.uniform(0.02, 0.1, U), rng.uniform(0.005, 0.03, U)
    hh = 1e-6
    lbeta = lambda x, a, b: (a - 1) * math.log(x) + (b - 1) * math.log(1 - x)
    Q = lambda n_, s_, var_, varR_, phi_: _true_q(st, g, n_, s_, var_, varR_, phi_)
    num = (Q(n + hh, s, var, varR, phi) + lbeta(n + hh, 2, 2) - Q(n - hh, s, var, varR, phi) - lbeta(n - hh, 2, 2)) / (2 * hh)
    assert O.deriv_n_gauss(n, s, var, varR, phi, rho, st, rt, rn, ct, 2, 2) == pytest.approx(num, rel=1e-6)
    for z in range(2):
        sp, sm = s.copy(), s.copy()
        sp[z] += hh
        sm[z] -= hh
        num = (Q(n, sp, var, varR, phi) + lbeta(sp[z], 2, 2) - Q(n, sm, var, varR, phi) - lbeta(sm[z], 2, 2)) / (2 * hh)
        assert O.deriv_s_gauss(s[z], z, n, var, varR, phi, rho, st, rt, rn, ct, 2, 2) == pytest.approx(num, rel=1e-6)
    # var of one copy-number level (all its states share the unknown; the prior is summed over them, :493-494)
    ind = np.flatnonzero(ct == 3)
    aK, bK = np.asarray(g["alphaKHyper"], float), np.asarray(g["betaKHyper"], float)
    v0 = 0.06

    def q_var(v):
        vv = var.copy()
        vv[ind] = v
        return Q(n, s, vv, varR, phi) + sum(O.invgammapdflog(v, aK[k], bK[k]) for k in ind)

    num = (q_var(v0 + hh) - q_var(v0 - hh)) / (2 * hh)
    got = O.deriv_var_gauss(v0, n, s, varR[ind], phi, rho, st["a"][ind], st["c"][ind], st["e"][ind], st["g"][ind], st["h"][ind],
                            rt[ind], rn, ct[ind], aK[ind], bK[ind], "logRatio")
    assert got == pytest.approx(num, rel=1e-5)
    # varR of one state (the caller swaps a <-> c and passes f as g, :189-190)
    k, kk = 4, slice(4, 5)
    aR, bR = np.asarray(g["alphaRHyper"], float), np.asarray(g["betaRHyper"], float)
    r0 = 0.012

    def q_varR(v):
        vv = varR.copy()
        vv[k] = v
        return Q(n, s, var, vv, phi) + O.invgammapdflog(v, aR[k], bR[k])

    num = (q_varR(r0 + 1e-7) - q_varR(r0 - 1e-7)) / 2e-7
    got = O.deriv_var_gauss(r0, n, s, var[k], phi, rho, st["c"][kk], st["a"][kk], st["e"][kk], st["f"][kk], st["h"][kk],
                            rt[kk], rn, ct[kk], aR[kk], bR[kk], "allelicRatio")
    assert got == pytest.approx(num, rel=1e-5)
    # phi at Z = 1
    data1, po1, _, st1 = _setup(Z=1, seed=9)
    g1 = po1["genotypeParams"]
    s1 = np.array([0.1])
    ig = lambda x: O.invgammapdflog(x, 20.0, 42.0)
    num = (_true_q(st1, g1, n, s1, var, varR, phi + hh) + ig(phi + hh) - _true_q(st1, g1, n, s1, var, varR, phi - hh) - ig(phi - hh)) / (2 * hh)
    got = O.deriv_phi_gauss(phi, n, s1, var, varR, g1["corRho_0"], st1, g1["rt"], g1["rn"], g1["ct"], 20.0, 42.0)
    assert got == pytest.approx(num, rel=1e-6)


def test_reference_objective_slips_are_restated_as_written():
    """likelihoodFunc's Gaussian branch (R/paramEstimation.R:321-327) evaluates log(1 / 2 * pi * sqrt(...)) = log(pi/2 * sqrt)
    and halves the two quadratic terms once more; with one state and unit sums the value can be written down by hand."""
    g = {"rt": np.array([0.7]), "rn": 0.5, "ct": np.array([3.0]), "corRho_0": 0.2, "alleleEmissionModel": "Gaussian",
         "kappaGHyper": np.array([2.0]), "alphaKHyper": np.array([3.0]), "betaKHyper": np.array([1.0]),
         "alphaRHyper": np.array([3.0]), "betaRHyper": np.array([1.0])}
    nP, sP, pP = {"alphaNHyper": 2.0, "betaNHyper": 2.0}, {"alphaSHyper": [2.0], "betaSHyper": [2.0], "kappaZHyper": np.array([2.0])}, \
        {"alphaPHyper": 20.0, "betaPHyper": 42.0}
    one = np.ones((1, 1))
    st = {"a": 0.6 * one, "c": 0.1 * one, "e": one, "f": 0.37 * one, "g": 0.02 * one, "h": 0.05 * one}
    n, s, var, varR, phi = 0.3, np.array([0.2]), np.array([0.05]), np.array([0.01]), 2.0
    got = O.likelihood_func_gauss(n, s, var, varR, phi, st, g, nP, sP, pP, np.array([1.0]), np.array([1.0]), 0.0, 0.0,
                                  "fixed", False, False)
    muR, muC = (m[0, 0] for m in O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi))
    rho = 0.2
    t1 = math.log(0.5 * math.pi * math.sqrt(0.05 * 0.01 * (1 - rho ** 2)))
    t2 = 1 / (2 * (1 - rho ** 2))
    t3 = (0.37 - 2 * 0.6 * muR + muR ** 2) / (2 * 0.01)
    t4 = (0.02 - 2 * 0.1 * muC + muC ** 2) / (2 * 0.05)
    t5 = 2 * rho * (0.05 - 0.6 * muC - 0.1 * muR + muC * muR) / math.sqrt(0.05 * 0.01)
    prior = O.invgammapdflog(0.05, 3.0, 1.0) + O.invgammapdflog(0.01, 3.0, 1.0)      # var + varR; pi priors are log(1) = 0 terms
    pr = O.prior_probs(n, s, phi, var, np.array([1.0]), np.array([1.0]), g, nP, sP, pP, "fixed", False, False, varR=varR)
    assert pr["var"] + pr["varR"] == pytest.approx(prior, rel=1e-14)
    assert got == pytest.approx(t1 - t2 * (t3 + t4 - t5) + pr["prior"], rel=1e-13)


@pytest.mark.parametrize("Z,maxCN", [(2, 5), (3, 8)])
def test_host_gaussian_mstep_matches_oracle(Z, maxCN):
    """titan_mstep_gaussian (C++ host code, titan_host.hpp) against the oracle's coordinate descent on the same sums."""
    data, po, rho, st = _setup(Z=Z, maxCN=maxCN, N=2000, seed=3 + Z)
    g, nP, sP, pP = po["genotypeParams"], po["normalParams"], po["cellPrevParams"], po["ploidyParams"]
    U = len(g["rt"])
    es = {"rho": rho, "rhoG": rho.sum(1), "rhoZ": rho.sum(0)}
    chr0 = np.array([ix[0] for ix in O.chromosome_slices(data["chr"])])
    want = O.m_step(data, es, 0.5, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], chr0, po, maxiterUpdate=1500,
                    varR_0=g["varR_0"])
    flat = np.concatenate([st[k].reshape(-1, order="F") for k in "afcheg"])       # slot order of titan_estep_out.stats
    keep = []
    from titancna_b200.hmm import _hyper
    h = _hyper(T.loadDefaultParameters(maxCN, Z, alleleEmissionModel="Gaussian", data=data), keep)
    out_s, out_var, out_varR, out_piG, out_piZ = np.zeros(Z), np.zeros(U), np.zeros(U), np.zeros(U), np.zeros(Z)
    n, phi, prior = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    sweeps = ctypes.c_int()
    P = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_lib.c_double_p)
    rG0 = np.ascontiguousarray(es["rhoG"][:, chr0].T)          # [nChr][U]
    rZ0 = np.ascontiguousarray(es["rhoZ"][:, chr0].T)
    s0, v0, r0, pg, pz = (np.ascontiguousarray(a, dtype=np.float64) for a in (sP["s_0"], g["var_0"], g["varR_0"], g["piG_0"], sP["piZ_0"]))
    rc = T.lib().titan_mstep_gaussian(ctypes.byref(h), U, Z, chr0.size, 1500, 1, 1, 1, P(flat), P(rG0), P(rZ0), 0.5, P(s0), P(v0),
                                      P(r0), 2.0, P(pg), P(pz), ctypes.byref(n), P(out_s), P(out_var), P(out_varR),
                                      ctypes.byref(phi), P(out_piG), P(out_piZ), ctypes.byref(sweeps), ctypes.byref(prior))
    assert rc == 0, T.lib().titan_last_error()
    assert sweeps.value == want["sweeps"]
    assert abs(n.value - want["n"]) < 1e-11 and abs(phi.value - want["phi"]) < 1e-11
    assert np.abs(out_s - want["s"]).max() < 1e-11
    assert np.abs(out_var / want["var"] - 1).max() < 1e-10 and np.abs(out_varR / want["varR"] - 1).max() < 1e-10
    assert np.abs(out_piG - want["piG"]).max() < 1e-15 and np.abs(out_piZ - want["piZ"]).max() < 1e-15
    pr = O.prior_probs(want["n"], want["s"], want["phi"], want["var"], want["piG"], want["piZ"], g, nP, sP, pP, varR=want["varR"])
    assert abs(prior.value - pr["prior"]) < 1e-9 * abs(pr["prior"])
    # the binomial entry refuses nothing silently: a Gaussian hyper set without its extra fields is an error
    h2 = _hyper(T.loadDefaultParameters(maxCN, Z), keep)
    h2.varR_0 = None
    rc = T.lib().titan_mstep_gaussian(ctypes.byref(h2), U, Z, chr0.size, 1500, 1, 1, 1, P(flat), P(rG0), P(rZ0), 0.5, P(s0), P(v0),
                                      P(r0), 2.0, P(pg), P(pz), ctypes.byref(n), P(out_s), P(out_var), P(out_varR),
                                      ctypes.byref(phi), P(out_piG), P(out_piZ), ctypes.byref(sweeps), ctypes.byref(prior))
    assert rc == _lib.ERR_ARG and b"varR_0" in T.lib().titan_last_error()


def test_gaussian_solver_without_a_device_fails_loudly():
    if T.device_count() > 0:
        pytest.skip("a device is present")
    data, _, _, _ = _setup(N=200)
    p = T.loadDefaultParameters(5, 2, alleleEmissionModel="Gaussian", data=data)
    with pytest.raises(T.TitanError) as e:
        T.runEMclonalCN(data, p, maxiter=2, verbose=False)
    assert e.value.code == _lib.ERR_NO_DEVICE
