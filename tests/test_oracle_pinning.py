"""CPU: the oracle against the reference's fixture and against the reference's own object code."""
import math
import os

import numpy as np
import pytest

from oracle import refshim, titan_oracle as O
from tests import synth


def _fixture_inputs(golden, col=2):
    g = {k[2:]: golden[k] for k in golden if k.startswith("g_")}
    muR, muC = O.clonal_two_component_mixture(g["rt"], float(g["rn"][0]), golden["cp_n"][col], golden["cp_s"][:, col],
                                              g["ct"], golden["cp_phi"][col])
    py = O.compute_py_c(golden["data_ref"], golden["data_tumDepth"], golden["data_logR"], muR, muC,
                        golden["cp_var"][:, col])
    lp = np.log(np.outer(golden["cp_piG"][:, col], golden["cp_piZ"][:, col]).reshape(-1, order="F"))
    return g, muR, muC, py, lp


def test_mixture_means_match_fixture(golden):
    g = {k[2:]: golden[k] for k in golden if k.startswith("g_")}
    for col in range(3):
        muR, muC = O.clonal_two_component_mixture(g["rt"], float(g["rn"][0]), golden["cp_n"][col],
                                                  golden["cp_s"][:, col], g["ct"], golden["cp_phi"][col])
        assert np.abs(muR - golden["cp_muR"][:, :, col]).max() <= 1e-15
        assert np.abs(muC - golden["cp_muC"][:, :, col]).max() <= 1e-15


def test_port_estep_reproduces_fixture_posteriors(golden):
    """E-step identity pinned by data/EMresults.rda: tolerance 1e-13 absolute."""
    g, muR, muC, py, lp = _fixture_inputs(golden)
    txn = float(golden["cp_txnExpLen"][0])
    zst = float(golden["cp_txnZstrength"][0]) * txn
    rho, ll = O.fwd_back(lp, py, g["ZS"], 2, golden["data_posn"].astype(float), zst, txn, engine="port")
    r = np.exp(rho).reshape((12, 2, -1), order="F")
    assert np.abs(r.sum(1) - golden["cp_rhoG"]).max() < 1e-13
    assert np.abs(r.sum(0) - golden["cp_rhoZ"]).max() < 1e-13
    assert ll == pytest.approx(float(golden["ref_loglik"][0]), rel=0, abs=0)   # bit-identical to the reference C
    path = O.viterbi(lp, py, g["ZS"], 2, golden["data_posn"].astype(float), zst, txn, engine="port")
    assert np.array_equal(path, golden["ref_path"].astype(np.int32))
    assert 1 + np.count_nonzero(np.diff(path)) == 26          # SURVEY.md 8c known answer


def test_emission_numpy_and_c_agree_bitwise(golden):
    g, muR, muC, py, lp = _fixture_inputs(golden)
    py2 = O.compute_py(golden["data_ref"], golden["data_tumDepth"], golden["data_logR"], muR, muC, golden["cp_var"][:, 2])
    assert np.array_equal(py, py2)


@pytest.mark.skipif(not refshim.have_reference(), reason="oracle/_ref not built (no reference checkout)")
@pytest.mark.parametrize("txn,zst,Z,maxCN", [(1e15, 1e15, 3, 8), (1e9, 1e18, 2, 5), (1e15, 5e20, 1, 8), (1e5, 1e5, 2, 3)])
def test_port_is_bit_identical_to_reference_object_code(txn, zst, Z, maxCN):
    data, truth = synth.make_data(1500, Z=Z, maxCN=maxCN, n_chr=2, seed=7, mean_seg=150)
    p = O.load_default_parameters(maxCN, Z)
    g = p["genotypeParams"]
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], 0.4, p["cellPrevParams"]["s_0"], g["ct"], 2.2)
    py = O.compute_py_c(data["ref"], data["tumDepth"], data["logR"], muR, muC, g["var_0"])
    lp = np.log(np.outer(g["piG_0"], p["cellPrevParams"]["piZ_0"]).reshape(-1, order="F"))
    for idx in O.chromosome_slices(data["chr"]):
        a = O.fwd_back(lp, py[:, idx], g["ZS"], Z, data["posn"][idx], zst, txn, engine="port")
        b = O.fwd_back(lp, py[:, idx], g["ZS"], Z, data["posn"][idx], zst, txn, engine="reference", ct=g["ct"])
        assert np.array_equal(a[0], b[0]) and a[1] == b[1]
        pa = O.viterbi(lp, py[:, idx], g["ZS"], Z, data["posn"][idx], zst, txn, engine="port")
        pb = O.viterbi(lp, py[:, idx], g["ZS"], Z, data["posn"][idx], zst, txn, engine="reference", ct=g["ct"])
        assert np.array_equal(pa, pb)


@pytest.mark.skipif(not refshim.have_reference(), reason="oracle/_ref not built")
def test_transition_matrix_restatement():
    ref = refshim.reference()
    p = O.load_default_parameters(5, 3)
    zs = p["genotypeParams"]["ZS"]
    U, K = zs.size, zs.size * 3
    for rG, rZ in [(1 - 2.5e-12, 1 - 2.5e-12), (0.9, 0.7), (1 - 2.5e-12, 1.0), (0.5, 0.5)]:
        A = ref.transition_matrix(K, U, zs, rG, rZ)
        B = np.zeros(K * K)
        O.port().port_transition(O._p(B), K, U, O._p(np.ascontiguousarray(zs)), rG, rZ, 0)
        assert np.array_equal(A, B.reshape((K, K), order="F"))
        assert np.allclose(A.sum(1), 1.0, atol=1e-14)


def test_registered_routines_and_errors():
    if not refshim.have_reference():
        pytest.skip("oracle/_ref not built")
    ref = refshim.reference()
    assert ref.registered() == {"getPositionOverlapC": 3, "fwd_backC_clonalCN": 9, "viterbiC_clonalCN": 9}
    with pytest.raises(refshim.RError, match="obslik must be"):
        ref.fwd_back(np.zeros(4), np.zeros((3, 5)), np.zeros(2), np.arange(2.0), 2, np.arange(5.0), 1e15, 1e15)
    out = ref.position_overlap([5, 50, 500], [1, 40], [10, 60])
    assert out.tolist() == [1.0, 2.0, 0.0]


def test_zeroin_finds_roots_like_brent():
    f = lambda x: math.cos(x) - x
    r = O.uniroot(f, 0.0, 1.0)
    assert abs(r - 0.7390851332151607) < 1e-15
    with pytest.raises(O.RootError):
        O.uniroot(lambda x: x * x + 1, -1, 1)


def test_mstep_ten_sweeps_and_sane_trajectory(golden):
    """M-step is not pinned by any reference fixture (it was written by an older package version with
    another pi update and kappaG); check the structural facts of SURVEY.md 7.2-6 -- exactly 10
    coordinate-descent sweeps per M-step -- and that the man-page run (man/TitanCNA-package.Rd:48-76)
    lands in the same neighbourhood as the old-version fixture."""
    data = {"chr": golden["data_chr"], "posn": golden["data_posn"].astype(float), "ref": golden["data_ref"].astype(float),
            "tumDepth": golden["data_tumDepth"].astype(float), "logR": golden["data_logR"]}
    p = O.load_default_parameters(5, 2, skew=0.1)
    p["genotypeParams"]["alphaKHyper"][:] = 500
    p["ploidyParams"]["phi_0"] = 1.5
    cp = O.run_em_clonal_cn(data, p, txnExpLen=1e9, txnZstrength=1e9, maxiter=3, maxiterUpdate=500)
    assert cp["cd_sweeps"] == [10] * (len(cp["n"]) - 1)
    assert np.all(np.diff(cp["loglik"][1:]) > 0)
    assert np.abs(cp["n"][1:] - golden["cp_n"][1:]).max() < 0.03
    assert np.abs(cp["phi"][1:] - golden["cp_phi"][1:]).max() < 0.03
    assert np.abs(cp["s"][:, 1:] - golden["cp_s"][:, 1:]).max() < 0.03


def test_mstep_derivatives_are_gradients_of_the_objective(golden):
    """The update equations (R/paramEstimation.R:359-437,499-530) must be the partial derivatives of
    likelihoodFunc (:287-356): guards the restatement of both against transcription slips."""
    data = {"chr": golden["data_chr"], "posn": golden["data_posn"].astype(float), "ref": golden["data_ref"].astype(float),
            "tumDepth": golden["data_tumDepth"].astype(float), "logR": golden["data_logR"]}
    p = O.load_default_parameters(5, 2, skew=0.1)
    g, nP, sP, pP = p["genotypeParams"], p["normalParams"], p["cellPrevParams"], p["ploidyParams"]
    rng = np.random.default_rng(3)
    rho = rng.dirichlet(np.ones(24), size=400).T.reshape((12, 2, 400), order="F")
    st = O.sufficient_statistics(data["ref"][:400], data["tumDepth"][:400], data["logR"][:400], rho)
    n, s, var, phi, h = 0.4, np.array([0.01, 0.4]), g["var_0"].copy(), 1.6, 1e-6
    F = lambda n, s, var, phi: O.likelihood_func(n, s, var, phi, st, g, nP, sP, pP, g["piG_0"], sP["piZ_0"], 0.0, 0.0,
                                                 "map", True, True)
    num = (F(n + h, s, var, phi) - F(n - h, s, var, phi)) / (2 * h)
    assert O.deriv_n(n, s, var, phi, st, g["rt"], g["rn"], g["ct"], 2, 2) == pytest.approx(num, rel=1e-6)
    for z in range(2):
        sp, sm = s.copy(), s.copy()
        sp[z] += h
        sm[z] -= h
        num = (F(n, sp, var, phi) - F(n, sm, var, phi)) / (2 * h)
        assert O.deriv_s(s[z], z, n, var, phi, st, g["rt"], g["rn"], g["ct"], 2, 2) == pytest.approx(num, rel=1e-6)
    num = (F(n, s, var, phi + h) - F(n, s, var, phi - h)) / (2 * h)
    assert O.deriv_phi(phi, n, s, var, st, g["rt"], g["rn"], g["ct"], 20, 42) == pytest.approx(num, rel=1e-6)


def _fixture_iteration(golden, col):
    """Parameters of column `col` of the fixture's convergeParams and what the reference stored for column col + 1."""
    g = {k[2:]: golden[k] for k in golden if k.startswith("g_")}
    g["rn"] = float(g["rn"][0])
    g["alleleEmissionModel"] = "binomial"
    Z = golden["cp_s"].shape[0]
    nP = {"n_0": 0.5, "alphaNHyper": 2.0, "betaNHyper": 2.0}
    sP = {"s_0": golden["cp_s"][:, 0], "alphaSHyper": np.full(Z, 2.0), "betaSHyper": np.full(Z, 2.0),
          "kappaZHyper": golden["s_kappaZHyper"]}
    pP = {"phi_0": 1.5, "alphaPHyper": 20.0, "betaPHyper": 42.0}
    return g, nP, sP, pP


@pytest.mark.parametrize("col", [0, 1, 2])
def test_fixture_loglik_trajectory_pins_estep_and_priors(golden, col):
    """data/EMresults.rda stores loglik[i] = sum of the chromosome log-likelihoods of the E-step run with column i-1
    + priorProbs(column i) (R/hmmClonal.R:218,292-311) for three EM iterations.  The oracle's E-step on column i-1 plus
    its restatement of priorProbs (R/hmmClonal.R:646-665,715-783: beta, inverse-gamma and Dirichlet log densities) on the
    STORED column i reproduces every one of them: the forward-backward log-likelihood is pinned at three different
    parameter sets and the prior terms at three, by a reference-held vector.  (The M-step itself stays unpinned: the
    stored columns come from an older package version whose coordinate descent differs)"""
    g, nP, sP, pP = _fixture_iteration(golden, col)
    data = {"chr": golden["data_chr"], "posn": golden["data_posn"].astype(float), "ref": golden["data_ref"].astype(float),
            "tumDepth": golden["data_tumDepth"].astype(float), "logR": golden["data_logR"]}
    Z = golden["cp_s"].shape[0]
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], golden["cp_n"][col], golden["cp_s"][:, col], g["ct"],
                                              golden["cp_phi"][col])
    assert np.abs(muR - golden["cp_muR"][:, :, col]).max() < 1e-15 and np.abs(muC - golden["cp_muC"][:, :, col]).max() < 1e-15
    py = O.compute_py_c(data["ref"], data["tumDepth"], data["logR"], muR, muC, golden["cp_var"][:, col])
    es = O.e_step(data, O.chromosome_slices(data["chr"]), py, golden["cp_piG"][:, col], golden["cp_piZ"][:, col], g["ZS"],
                  g["ct"], Z, float(golden["cp_txnExpLen"][0]), float(golden["cp_txnZstrength"][0]), engine="port")
    c1 = col + 1
    pr = O.prior_probs(golden["cp_n"][c1], golden["cp_s"][:, c1], golden["cp_phi"][c1], golden["cp_var"][:, c1],
                       golden["cp_piG"][:, c1], golden["cp_piZ"][:, c1], g, nP, sP, pP)
    want = float(golden["cp_loglik"][c1])
    assert abs((es["loglik"] + pr["prior"]) - want) <= 1e-13 * abs(want)
    # the product's host-side priors (titan_host.hpp:prior_probs through titan_mstep's out_prior is tied to its own
    # M-step output, so evaluate the C++ densities through the oracle-equivalent decomposition instead)
    assert pr["prior"] == pytest.approx(pr["s"] + pr["n"] + pr["phi"] + pr["var"] + pr["piG"] + pr["piZ"], rel=1e-15)
