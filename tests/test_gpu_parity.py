"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances (BASELINE.json north_star): Viterbi paths bit-exact; per-chromosome log-likelihood
within 1e-9 relative; posteriors and converged EM parameters within 1e-6.  The asserts below are
tighter than required wherever the implementation allows.
"""
import math

import numpy as np
import pytest

import titancna_b200 as T
from oracle import titan_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9      # north_star: per-chromosome log-likelihood, relative
POST_ATOL = 1e-6    # north_star: posteriors / parameters, absolute


def _fixture(golden, col=2):
    g = {k[2:]: golden[k] for k in golden if k.startswith("g_")}
    muR, muC = O.clonal_two_component_mixture(g["rt"], float(g["rn"][0]), golden["cp_n"][col], golden["cp_s"][:, col],
                                              g["ct"], golden["cp_phi"][col])
    py = O.compute_py_c(golden["data_ref"], golden["data_tumDepth"], golden["data_logR"], muR, muC,
                        golden["cp_var"][:, col])
    lp = np.log(np.outer(golden["cp_piG"][:, col], golden["cp_piZ"][:, col]).reshape(-1, order="F"))
    return g, py, lp


def test_library_loaded_and_device_present():
    assert T.device_count() >= 1
    assert b"sm_100a" in T.lib().titan_version()


def test_legacy_fwd_back_reproduces_reference_fixture(golden):
    """.Call-compatible entry on data/EMresults.rda: stored rhoG/rhoZ are the known answer."""
    g, py, lp = _fixture(golden)
    txn = float(golden["cp_txnExpLen"][0])
    zst = float(golden["cp_txnZstrength"][0]) * txn
    out = T.fwd_back_clonalCN(lp, py, g["ct"], g["ZS"], 2, golden["data_posn"].astype(float), zst, txn, 0)
    r = np.exp(out["rho"]).reshape((12, 2, -1), order="F")
    assert np.abs(r.sum(1) - golden["cp_rhoG"]).max() < 1e-9
    assert np.abs(r.sum(0) - golden["cp_rhoZ"]).max() < 1e-9
    ref_ll = float(golden["ref_loglik"][0])
    assert abs(out["loglik"] - ref_ll) <= 1e-11 * abs(ref_ll)
    # production transition settings on the same data, against the reference's own C
    out2 = T.fwd_back_clonalCN(lp, py, g["ct"], g["ZS"], 2, golden["data_posn"].astype(float), 1e15, 1e15, 0)
    ref_ll2 = float(golden["ref_loglik_1e15"][0])
    assert abs(out2["loglik"] - ref_ll2) <= 1e-11 * abs(ref_ll2)


def test_legacy_viterbi_bit_exact_on_fixture(golden):
    g, py, lp = _fixture(golden)
    txn = float(golden["cp_txnExpLen"][0])
    zst = float(golden["cp_txnZstrength"][0]) * txn
    posn = golden["data_posn"].astype(float)
    path = T.viterbi_clonalCN(lp, py, g["ct"], g["ZS"], 2, posn, zst, txn, 0)
    assert np.array_equal(path, golden["ref_path"].astype(np.int32))
    path2 = T.viterbi_clonalCN(lp, py, g["ct"], g["ZS"], 2, posn, 1e15, 1e15, 0)
    assert np.array_equal(path2, golden["ref_path_1e15"].astype(np.int32))


def _oracle_estep(data, p, n, s, phi, var, piG, piZ, txn, zfac):
    g = p["genotypeParams"]
    Z = len(s)
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi)
    py = O.compute_py_c(data["ref"], data["tumDepth"], data["logR"], muR, muC, var)
    chrsI = O.chromosome_slices(data["chr"])
    es = O.e_step(data, chrsI, py, piG, piZ, g["ZS"], g["ct"], Z, txn, zfac, engine="port", threads=8)
    es["stats"] = O.sufficient_statistics(data["ref"], data["tumDepth"], data["logR"], es["rho"])
    es["py"], es["chrsI"] = py, chrsI
    return es


CASES = [
    # N, Z, maxCN, n_chr, txnExpLen, txnZstrength, chunk_len, warm
    (6000, 3, 8, 3, 1e15, 1.0, 512, 160),      # production transition settings, K=75
    (6000, 3, 8, 3, 1e15, 1.0, 0, 1),          # plain sequential recursion
    (5000, 1, 8, 2, 1e15, 5e5, 256, 64),       # vignette zStrength: 1-rhoZ == 0 exactly at short range
    (4000, 5, 8, 2, 1e9, 1e9, 300, 100),       # K=125, man-page settings
    (3000, 2, 5, 4, 1e5, 1.0, 128, 32),        # short txnExpLen: non-sticky transitions
    (3000, 4, 3, 1, 1e15, 1.0, 64, 16),        # smallest genotype table (U=6)
]


@pytest.mark.parametrize("N,Z,maxCN,n_chr,txn,zfac,chunk,warm", CASES)
def test_fused_estep_matches_oracle(N, Z, maxCN, n_chr, txn, zfac, chunk, warm):
    data, truth = synth.make_data(N, Z=Z, maxCN=maxCN, n_chr=n_chr, seed=11 + Z, mean_seg=300)
    p = T.loadDefaultParameters(copyNumber=maxCN, numberClonalClusters=Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    n, phi, s = 0.35, 2.1, np.array(synth.S_TRUE[:Z]) * 0.9 + 0.02
    var = g["var_0"] * np.linspace(0.5, 1.5, g["var_0"].size)
    es = _oracle_estep(data, p, n, s, phi, var, g["piG_0"], sP["piZ_0"], txn, zfac)
    sol = T.Solver(data, g, Z, txn, zfac)
    sol.configure(chunk, warm, 1e-10)
    out = sol.estep(n, s, phi, var, g["piG_0"], sP["piZ_0"], g, posteriors=True)
    tm = sol.timing()
    sol.close()
    assert np.allclose(out["loglik_chr"], es["loglik_chr"], rtol=LL_RTOL, atol=0)
    assert np.abs(out["loglik_chr"] - es["loglik_chr"]).max() <= 1e-11 * np.abs(es["loglik_chr"]).max()
    assert np.abs(out["rhoG"] - es["rhoG"]).max() < 1e-9 and np.abs(out["rhoG"] - es["rhoG"]).max() < POST_ATOL
    assert np.abs(out["rhoZ"] - es["rhoZ"]).max() < 1e-9
    chr0 = np.array([ix[0] for ix in es["chrsI"]])
    assert np.abs(out["rhoG0"] - es["rhoG"][:, chr0]).max() < 1e-9
    assert np.abs(out["rhoZ0"] - es["rhoZ"][:, chr0]).max() < 1e-9
    for k in "abcdeg":
        ref = es["stats"][k]
        assert np.abs(out["stats"][k] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), k
    assert tm["n_chunks"] >= n_chr


@pytest.mark.parametrize("vit_mode", ["-1", "0", "2", "3", "4"])
@pytest.mark.parametrize("N,Z,maxCN,n_chr,txn,zfac", [
    (6000, 3, 8, 3, 1e15, 1.0), (5000, 1, 8, 2, 1e15, 5e5), (4000, 5, 8, 2, 1e9, 1e9), (3000, 2, 5, 4, 1e5, 1.0),
    (2000, 8, 4, 2, 1e15, 1.0), (1500, 8, 8, 2, 1e15, 1.0), (1500, 6, 8, 1, 1e12, 3.0)])
def test_fused_viterbi_bit_exact(N, Z, maxCN, n_chr, txn, zfac, vit_mode, monkeypatch):
    """Every forward kernel (auto choice, shared-memory dense scan, register-sliced dense scan,
    cluster-sliced scan) against the reference C.  Systematic near-ties included: with
    hetBaselineSkew = 0 the diploid-HET states of different clusters have identical emissions
    (SURVEY.md 7.2-2)."""
    monkeypatch.setenv("TITAN_VIT_MODE", vit_mode)
    data, truth = synth.make_data(N, Z=Z, maxCN=maxCN, n_chr=n_chr, seed=5 + Z, mean_seg=250)
    for skew in (None, 0.0):
        p = T.loadDefaultParameters(copyNumber=maxCN, numberClonalClusters=Z, hetBaselineSkew=skew)
        g, sP = p["genotypeParams"], p["cellPrevParams"]
        n, phi = 0.3, 2.0
        s = np.array(synth.S_TRUE[:Z]) * 0.95 + 0.01
        muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"], n, s, g["ct"], phi)
        py = O.compute_py_c(data["ref"], data["tumDepth"], data["logR"], muR, muC, g["var_0"])
        lp = np.vectorize(math.log)(np.outer(g["piG_0"], sP["piZ_0"]).reshape(-1, order="F"))
        want = np.empty(N, dtype=np.int32)
        for idx in O.chromosome_slices(data["chr"]):
            want[idx] = O.viterbi(lp, py[:, idx], g["ZS"], Z, data["posn"][idx], zfac * txn, txn, engine="port")
        sol = T.Solver(data, g, Z, txn, zfac)
        got = sol.viterbi(n, s, phi, g["var_0"], g["piG_0"], sP["piZ_0"], g)
        sol.close()
        assert np.array_equal(got, want), f"first mismatch at {np.flatnonzero(got != want)[:5]}"


def test_em_matches_oracle_on_reference_fixture_data(golden):
    """Full runEMclonalCN + viterbiClonalCN on the chr2 fixture with the man-page settings
    (man/TitanCNA-package.Rd:48-76): converged parameters within 1e-6, same iteration count,
    identical Viterbi calls."""
    data = {"chr": golden["data_chr"], "posn": golden["data_posn"].astype(float), "ref": golden["data_ref"].astype(float),
            "tumDepth": golden["data_tumDepth"].astype(float), "logR": golden["data_logR"]}
    kw = dict(txnExpLen=1e9, txnZstrength=1e9, maxiter=3, maxiterUpdate=500)
    po = O.load_default_parameters(5, 2, skew=0.1)
    po["genotypeParams"]["alphaKHyper"][:] = 500
    po["ploidyParams"]["phi_0"] = 1.5
    want = O.run_em_clonal_cn(data, po, engine="port", **kw)
    pg = T.loadDefaultParameters(copyNumber=5, numberClonalClusters=2, skew=0.1)
    pg["genotypeParams"]["alphaKHyper"][:] = 500
    pg["ploidyParams"]["phi_0"] = 1.5
    got = T.runEMclonalCN(data, pg, verbose=False, **kw)
    assert got["n"].shape == want["n"].shape
    assert list(got["cd_sweeps"][1:]) == want["cd_sweeps"]
    for k in ("n", "phi", "s", "var", "piG", "piZ", "muR", "muC"):
        assert np.abs(np.asarray(got[k]) - np.asarray(want[k])).max() < POST_ATOL, k
    assert np.allclose(got["loglik"][1:], want["loglik"][1:], rtol=LL_RTOL, atol=0)
    assert np.abs(got["rhoG"] - want["rhoG"]).max() < POST_ATOL
    assert np.abs(got["rhoZ"] - want["rhoZ"]).max() < POST_ATOL
    wp = O.viterbi_clonal_cn(data, want, engine="port")
    gp = T.viterbiClonalCN(data, got)
    assert np.array_equal(gp, wp)


def test_em_production_settings_synthetic():
    """scripts/R_scripts/titanCNA.R:243-256 settings on a sub-sampled synthetic genome."""
    N, Z = 12000, 2
    data, truth = synth.make_data(N, Z=Z, maxCN=8, n_chr=4, seed=99, mean_seg=400)
    kw = dict(txnExpLen=1e15, txnZstrength=1.0, maxiter=4, maxiterUpdate=1500)

    def params(mod):
        p = mod(8, Z, data=data)
        p["genotypeParams"]["alphaKHyper"][:] = 10000
        p["genotypeParams"]["betaKHyper"][:] = 25
        p["ploidyParams"]["phi_0"] = 2.0
        p["normalParams"]["n_0"] = 0.5
        return p

    want = O.run_em_clonal_cn(data, params(O.load_default_parameters), engine="port", threads=8, **kw)
    got = T.runEMclonalCN(data, params(T.loadDefaultParameters), verbose=False, **kw)
    assert got["n"].shape == want["n"].shape
    for k in ("n", "phi", "s", "var"):
        assert np.abs(np.asarray(got[k]) - np.asarray(want[k])).max() < POST_ATOL, k
    assert np.allclose(got["loglik"][1:], want["loglik"][1:], rtol=LL_RTOL, atol=0)
    gpath = T.viterbiClonalCN(data, got)
    wpath = O.viterbi_clonal_cn(data, want, engine="port", threads=8)
    assert np.array_equal(gpath, wpath)
    # segment calls bit-exact (north_star): decode + segmenter on both paths
    from oracle import results_oracle as RO
    segs = T.outputTitanSegments(T.outputTitanResults(data, got, gpath, correctResults=False)["results"], "synthetic", got)
    wsegs = RO.output_segments(RO.output_results({k: np.asarray(v).tolist() for k, v in data.items()}, want, wpath), "synthetic")
    assert segs["Length.snp."].size == len(wsegs)
    for k in ("Start_Position.bp.", "End_Position.bp.", "Length.snp.", "TITAN_state", "Copy_Number", "MinorCN", "MajorCN"):
        assert np.array_equal(np.asarray(segs[k], float), np.array([w[k] for w in wsegs], float)), k
    assert list(segs["TITAN_call"]) == [w["TITAN_call"] for w in wsegs]


def test_chunked_scan_equals_sequential_recursion_at_scale():
    """Size-independent property at a size the O(N K^2) oracle cannot reach: the chunk-parallel scan
    and the plain sequential recursion are the same function."""
    N, Z = 400000, 3
    data, truth = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=20261020)
    p = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    sol = T.Solver(data, g, Z, 1e15, 1.0)
    args = (0.5, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    sol.configure(0, 1, 1e-10)
    seq = sol.estep(*args, posteriors=True)
    sol.configure(512, 160, 1e-10)
    par = sol.estep(*args, posteriors=True)
    tm = sol.timing()
    sol.close()
    assert tm["n_chunks"] > 23 and tm["last_fwd_rounds"] >= 2
    assert np.allclose(par["loglik_chr"], seq["loglik_chr"], rtol=1e-12, atol=0)
    assert np.abs(par["rhoG"] - seq["rhoG"]).max() < 1e-8
    assert np.abs(par["rhoZ"] - seq["rhoZ"]).max() < 1e-8
    # a statistic is sum_t w_t * posterior_t: its natural error scale is max|d posterior| * sum_t |w_t|
    lb = np.abs(np.asarray([x for x in (data["ref"], data["tumDepth"] - data["ref"], data["logR"])]))
    scale = {"a": lb[0].sum(), "b": lb[1].sum(), "c": lb[2].sum(), "d": 12.0 * N, "e": float(N), "g": (lb[2] ** 2).sum()}
    for k in "abcdeg":
        assert np.allclose(par["stats"][k], seq["stats"][k], rtol=1e-9, atol=1e-9 * scale[k]), k
    # posterior marginals are distributions; the e statistic counts every SNP once
    assert np.abs(par["rhoG"].sum(0) - 1).max() < 1e-12
    assert np.abs(par["rhoZ"].sum(0) - 1).max() < 1e-12
    assert abs(par["stats"]["e"].sum() - N) < 1e-6
    assert abs(par["stats"]["a"].sum() - data["ref"].sum()) < 1e-6 * data["ref"].sum()


def test_edge_cases_and_errors():
    # chromosomes of length 1 and 2 next to a normal one
    data, _ = synth.make_data(600, Z=2, maxCN=5, n_chr=3, seed=3, mean_seg=100)
    keep = np.r_[0:1, np.flatnonzero(data["chr"] == 2)[:2], np.flatnonzero(data["chr"] == 3)]
    data = {k: v[keep] for k, v in data.items()}
    p = T.loadDefaultParameters(copyNumber=5, numberClonalClusters=2)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    es = _oracle_estep(data, p, 0.4, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], 1e15, 1.0)
    sol = T.Solver(data, g, 2, 1e15, 1.0)
    sol.configure(50, 10, 1e-10)
    out = sol.estep(0.4, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], g, posteriors=True)
    path = sol.viterbi(0.4, sP["s_0"], 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    sol.close()
    assert np.allclose(out["loglik_chr"], es["loglik_chr"], rtol=LL_RTOL, atol=0)
    assert np.abs(out["rhoG"] - es["rhoG"]).max() < 1e-9
    lp = np.vectorize(math.log)(np.outer(g["piG_0"], sP["piZ_0"]).reshape(-1, order="F"))
    want = np.concatenate([O.viterbi(lp, es["py"][:, ix], g["ZS"], 2, data["posn"][ix], 1e15, 1e15) for ix in es["chrsI"]])
    assert np.array_equal(path, want)
    # error behaviour mirrors the reference's error() text / refuses what it does not cover
    with pytest.raises(T.TitanError, match="obslik must be"):
        T.fwd_back_clonalCN(np.zeros(4), np.zeros((3, 5)), np.zeros(2), np.arange(2.0), 2, np.arange(5.0), 1e15, 1e15)
    with pytest.raises(T.TitanError, match="positive"):
        T.Solver(data, g, 2, 0.0, 1.0)


@pytest.mark.parametrize("env", [{}, {"TITAN_PROBES": "0"}, {"TITAN_SPEC_FINAL": "0"}, {"TITAN_PROBE_WINDOW": "0"},
                                 {"TITAN_ROUND_BATCH": "1"}, {"TITAN_VIT_MODE": "2"}, {"TITAN_VIT_MODE": "3"},
                                 {"TITAN_VIT_MODE": "4"}])
def test_kernel_variants_agree(env, monkeypatch):
    """Every execution strategy (re-runs from the neighbour's vector only or with the probe-based prediction,
    speculative final pass on or off, one or several cycles per host synchronisation, every Viterbi kernel) is the
    same function of the inputs: compare each against the plain sequential recursion on a genome whose boundaries
    need several verification cycles (two nearly identical clusters: slow forgetting)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    N, Z = 60000, 3
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=5, seed=77, mean_seg=2000)
    p = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    s = np.array([0.01, 0.30, 0.32])          # two nearly identical clusters: slow forgetting
    sol = T.Solver(data, g, Z, 1e15, 1.0)
    args = (0.4, s, 2.0, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    sol.configure(0, 1, 1e-10)
    seq = sol.estep(*args, posteriors=True)
    sol.configure(256, 64, 1e-10)
    par = sol.estep(*args, posteriors=True)
    tm = sol.timing()
    p1 = sol.viterbi(*args)
    sol.close()
    monkeypatch.setenv("TITAN_VIT_MODE", "0")
    ref_sol = T.Solver(data, g, Z, 1e15, 1.0)
    p0 = ref_sol.viterbi(*args)
    ref_sol.close()
    assert np.array_equal(p0, p1)                     # every Viterbi forward kernel: identical paths
    assert tm["last_fwd_rounds"] >= 2
    assert np.allclose(par["loglik_chr"], seq["loglik_chr"], rtol=1e-11, atol=0)
    assert np.abs(par["rhoG"] - seq["rhoG"]).max() < 1e-7
    assert np.abs(par["rhoZ"] - seq["rhoZ"]).max() < 1e-7
    for k in "abcdeg":
        assert np.allclose(par["stats"][k], seq["stats"][k], rtol=1e-8, atol=1e-6), k


def test_r_call_glue_is_a_drop_in_for_the_reference_so(golden):
    """The .Call boundary itself: the same ctypes driver (oracle/refshim.py) loads the reference's own
    TitanCNA shared object and this repo's r_glue.c build, registers both through R_init_TitanCNA,
    looks the routines up by name, and passes identical SEXPs to both."""
    from oracle import refshim
    from titancna_b200 import build as B
    glue = refshim.ShimLibrary(B.GLUE)
    g, py, lp = _fixture(golden)
    posn = golden["data_posn"].astype(float)
    sl = slice(0, 4000)
    got = glue.fwd_back(lp, py[:, sl], g["ct"], g["ZS"], 2, posn[sl], 1e18, 1e9, 0)
    gpath = glue.viterbi(lp, py[:, sl], g["ct"], g["ZS"], 2, posn[sl], 1e18, 1e9, 0)
    if refshim.have_reference():
        want = refshim.reference().fwd_back(lp, py[:, sl], g["ct"], g["ZS"], 2, posn[sl], 1e18, 1e9, 0)
        wpath = refshim.reference().viterbi(lp, py[:, sl], g["ct"], g["ZS"], 2, posn[sl], 1e18, 1e9, 0)
    else:
        rho, ll = O.fwd_back(lp, py[:, sl], g["ZS"], 2, posn[sl], 1e18, 1e9)
        want, wpath = {"rho": rho, "loglik": ll}, O.viterbi(lp, py[:, sl], g["ZS"], 2, posn[sl], 1e18, 1e9)
    assert got["rho"].shape == want["rho"].shape
    assert abs(got["loglik"] - want["loglik"]) <= 1e-11 * abs(want["loglik"])
    assert np.abs(np.exp(got["rho"]) - np.exp(want["rho"])).max() < 1e-9
    assert gpath.dtype == np.int32 and np.array_equal(gpath, wpath)


@pytest.mark.parametrize("mode", ["exome", "haplotype"])
def test_baseline_config5_value_distributions(mode):
    """BASELINE configs[4]: exome-like positions (SNPs clustered in ~200 bp intervals with gaps up to 1e6 bp, so
    the transition terms span their whole range) and 10X haplotype-block counts (block sums: depths in the
    thousands).  E-step, EM parameters and Viterbi path against the oracle."""
    N, Z = 5000, 3
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=3, seed=31, mean_seg=300, exome=(mode == "exome"),
                              haplotype_bins=(mode == "haplotype"), bin_size=2_000_000)
    assert mode != "haplotype" or data["tumDepth"].max() > 1000
    kw = dict(txnExpLen=1e15, txnZstrength=1.0, maxiter=2, maxiterUpdate=1500)

    def params(mod):
        p = mod(8, Z, data=data)
        p["genotypeParams"]["alphaKHyper"][:] = 10000
        p["genotypeParams"]["betaKHyper"][:] = 25
        return p

    want = O.run_em_clonal_cn(data, params(O.load_default_parameters), engine="port", threads=8, **kw)
    got = T.runEMclonalCN(data, params(T.loadDefaultParameters), verbose=False, **kw)
    assert got["n"].shape == want["n"].shape
    for k in ("n", "phi", "s", "var"):
        assert np.abs(np.asarray(got[k]) - np.asarray(want[k])).max() < POST_ATOL, k
    assert np.allclose(got["loglik"][1:], want["loglik"][1:], rtol=LL_RTOL, atol=0)
    assert np.abs(got["rhoG"] - want["rhoG"]).max() < POST_ATOL
    assert np.array_equal(T.viterbiClonalCN(data, got), O.viterbi_clonal_cn(data, want, engine="port", threads=8))


def _legacy_case(N, U_zs, Z, outlier, K=None, seed=0):
    rng = np.random.default_rng(seed)
    zs = np.asarray(U_zs, float)
    K = (zs.size * Z + (1 if outlier else 0)) if K is None else K
    posn = np.cumsum(rng.integers(1, 5000, size=N)).astype(float)
    truth = np.repeat(rng.integers(0, K, size=N // 40 + 1), 40)[:N]
    py = rng.normal(-6.0, 1.5, size=(K, N))
    py[truth, np.arange(N)] += 5.0                       # a sharp, piecewise-constant signal
    lp = np.log(rng.dirichlet(np.full(K, 5.0)))
    return lp, py, zs, posn


@pytest.mark.parametrize("name,zs,Z,outlier,K,txn,zst", [
    ("outlier state", [0, 1, 2, -1, 4, 5, 6, 7, 8, 9, 10, 11], 2, 1, None, 1e15, 1e15),
    ("outlier, one cluster", [0, 1, 2, -1, 4, 5], 1, 1, None, 1e9, 1e18),
    ("repeated zygosity keys", [0, 1, 1, -1, 2, 2, 3], 3, 0, None, 1e15, 5e20),
    ("K not a multiple of numClust", [0, 1, 2, -1, 4, 5], 2, 0, 13, 1e6, 1e6),
    ("40 unit states", list(range(3)) + [-1] + list(range(4, 40)), 2, 0, None, 1e15, 1e15),
])
def test_generic_compatibility_path_matches_reference_semantics(name, zs, Z, outlier, K, txn, zst):
    """Layouts the structured kernels do not cover go through the dense kernels: outlier state with the
    C-level semantics of src/fwd_backC_clonalCN.c:224-247, repeated zygosity keys, odd K."""
    lp, py, zs, posn = _legacy_case(1200, zs, Z, outlier, K)
    want_rho, want_ll = O.fwd_back(lp, py, zs, Z, posn, zst, txn, outlier, engine="port")
    want_path = O.viterbi(lp, py, zs, Z, posn, zst, txn, outlier, engine="port")
    got = T.fwd_back_clonalCN(lp, py, np.zeros(zs.size), zs, Z, posn, zst, txn, outlier)
    assert abs(got["loglik"] - want_ll) <= 1e-11 * abs(want_ll), name
    assert np.abs(np.exp(got["rho"]) - np.exp(want_rho)).max() < 1e-9, name
    path = T.viterbi_clonalCN(lp, py, np.zeros(zs.size), zs, Z, posn, zst, txn, outlier)
    assert np.array_equal(path, want_path), name


def test_dense_kernel_cross_checks_structured_kernels(monkeypatch, golden):
    """The dense kernel assumes no structure at all: same inputs through both implementations."""
    g, py, lp = _fixture(golden)
    posn = golden["data_posn"].astype(float)
    a = T.fwd_back_clonalCN(lp, py, g["ct"], g["ZS"], 2, posn, 1e15, 1e15, 0)
    pa = T.viterbi_clonalCN(lp, py, g["ct"], g["ZS"], 2, posn, 1e15, 1e15, 0)
    monkeypatch.setenv("TITAN_FORCE_DENSE", "1")
    b = T.fwd_back_clonalCN(lp, py, g["ct"], g["ZS"], 2, posn, 1e15, 1e15, 0)
    pb = T.viterbi_clonalCN(lp, py, g["ct"], g["ZS"], 2, posn, 1e15, 1e15, 0)
    assert abs(a["loglik"] - b["loglik"]) <= 1e-12 * abs(b["loglik"])
    assert np.abs(np.exp(a["rho"]) - np.exp(b["rho"])).max() < 1e-9
    assert np.array_equal(pa, pb) and np.array_equal(pa, golden["ref_path_1e15"].astype(np.int32))


def test_correct_results_reenters_the_estep_like_the_reference():
    """outputTitanResults(correctResults=TRUE, recomputeLogLik=TRUE) -- the reference's default, SURVEY.md 3.4:
    removeEmptyClusters prunes a cluster, rescales n / s, and runs runEMclonalCN(maxiter=1, fixed normal, no s /
    ploidy update) with the NEW cluster count (R/utils.R:1500-1522).  The same corrected parameters go through the
    oracle's EM; log-likelihood, means and posteriors must agree to the north_star tolerances."""
    from oracle import results_oracle as RO
    N, Z = 5000, 3
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=3, seed=21, mean_seg=300)
    p = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=Z, data=data)
    cp = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=3, verbose=False)
    path = T.viterbiClonalCN(data, cp)
    cp["txn_exp_len"], cp["txn_z_strength"] = 1e15, 1.0          # the fields R/utils.R:1514-1515 reads
    share = [np.count_nonzero((path - 1) // 25 == z) / N for z in range(Z)]
    thr = float(np.sort(share)[1]) * 0.999 + 1e-9                # prune exactly the least populated cluster(s)
    out = T.outputTitanResults(data, cp, path, proportionThreshold=thr, proportionThresholdClonal=0.0, verbose=False)
    got = out["convergeParams"]
    rows = RO.output_results({k: np.asarray(v).tolist() for k, v in data.items()}, cp, path)
    want_cp, want_rows = RO.remove_empty_clusters(cp, rows, thr, 0.0)
    newZ = want_cp["s"].shape[0]
    assert 1 <= newZ < Z and got["s"].shape[0] == newZ
    it = want_cp["n"].size - 1
    po = {"genotypeParams": dict(cp["genotypeParams"]), "ploidyParams": dict(cp["ploidyParams"]),
          "normalParams": dict(cp["normalParams"]), "cellPrevParams": dict(want_cp["cellPrevParams"])}
    po["genotypeParams"].update(var_0=cp["var"][:, it], piG_0=cp["piG"][:, it])
    po["normalParams"]["n_0"] = float(want_cp["n"][it])
    po["ploidyParams"]["phi_0"] = float(cp["phi"][it])
    po["cellPrevParams"].update(s_0=want_cp["s"][:, it], piZ_0=np.asarray(want_cp["piZ"])[:newZ, it])
    want = O.run_em_clonal_cn(data, po, txnExpLen=1e15, txnZstrength=1.0, maxiter=1, normalEstimateMethod="fixed",
                              estimateS=False, estimatePloidy=False, engine="port", threads=8)
    assert abs(got["loglik"][it] - want["loglik"][-1]) <= LL_RTOL * abs(want["loglik"][-1])
    assert np.array_equal(got["loglik"][:it], cp["loglik"][:it])
    for k in ("muC", "muR"):
        assert np.abs(got[k][:, :, it] - want[k][:, :, 1]).max() < POST_ATOL, k
    assert got["rhoZ"].shape == (newZ, N) and got["rhoG"].shape == (25, N)
    assert np.abs(got["rhoG"] - want["rhoG"]).max() < POST_ATOL
    assert np.abs(got["rhoZ"] - want["rhoZ"]).max() < POST_ATOL
    for k in ("ClonalCluster", "CellularPrevalence"):
        a, b = np.asarray(out["corrResults"][k], float), np.array([r[k] for r in want_rows], float)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), k
    assert np.nanmax(out["corrResults"]["ClonalCluster"]) <= newZ


def test_sweep_with_selection_on_one_gpu():
    """TitanCNA.snakefile:8-52 + selectSolution.R in one process: every (ploidy, numClusters) solution through
    runEMclonalCN + viterbiClonalCN + result table + S_Dbw, then the reference's selection rule."""
    data, _ = synth.make_data(6000, Z=2, maxCN=8, n_chr=4, seed=31, mean_seg=300)
    out, info = T.run_sweep(data, 0, 1, ploidies=(2, 3), clusters=(1, 2), maxiter=3, txnExpLen=1e15, txnZstrength=1.0)
    assert len(out) == 4 and sorted((o["ploidy_0"], o["numClusters"]) for o in out) == [(2, 1), (2, 2), (3, 1), (3, 2)]
    for o in out:
        assert np.isfinite(o["loglik"]) and np.isfinite(o["sdbw"]) and o["sdbw"] > 0
        assert int(np.sum(o["segs"]["Length.snp."])) == 6000
    best = T.selectSolution(out)
    assert (best["ploidy_0"], best["numClusters"]) in {(o["ploidy_0"], o["numClusters"]) for o in out}
    from oracle import selection_oracle as SO
    want = SO.select_solution(out)
    assert (best["ploidy_0"], best["numClusters"]) == (want["ploidy_0"], want["numClusters"])


def test_full_size_viterbi_properties(monkeypatch):
    """BASELINE configs[1] size (1 M SNPs, 23 chromosomes, K = 75), where the O(N K^2) oracle cannot go: the three
    forward kernels and the tiled back-trace are the same function, and the decode of the genome is the
    concatenation of the decodes of its chromosomes (the HMM restarts at every chromosome, R/hmmClonal.R:470-483)."""
    N, Z = 1_000_000, 3
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=20261020)
    p = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=Z, data=data)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    args = (0.45, sP["s_0"], 2.04, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    paths = {}
    for mode in ("3", "2", "0"):
        monkeypatch.setenv("TITAN_VIT_MODE", mode)
        sol = T.Solver(data, g, Z, 1e15, 1.0)
        paths[mode] = sol.viterbi(*args)
        sol.close()
    assert np.array_equal(paths["3"], paths["0"]) and np.array_equal(paths["2"], paths["0"])
    full = paths["3"]
    assert full.min() >= 1 and full.max() <= 25 * Z and 1 + np.count_nonzero(np.diff(full)) > 23
    monkeypatch.setenv("TITAN_VIT_MODE", "-1")
    for c in (1, 16, 23):                                       # longest, a mid-sized and the last chromosome
        idx = np.flatnonzero(data["chr"] == c)
        sub = {k: v[idx] for k, v in data.items()}
        sol = T.Solver(sub, g, Z, 1e15, 1.0)
        alone = sol.viterbi(*args)
        sol.close()
        assert np.array_equal(alone, full[idx]), c


def test_readme_flow_from_allele_count_file(tmp_path):
    """The whole per-sample flow of scripts/R_scripts/titanCNA.R:204-285 through the mirrored API: allele-count file
    -> staging -> EM -> Viterbi -> result table -> segments -> corrected copy numbers -> parameter file."""
    src, _ = synth.make_data(8000, Z=2, maxCN=8, n_chr=5, seed=44, mean_seg=300)
    rng = np.random.default_rng(0)
    order = rng.permutation(8000)                                  # the file need not be sorted
    flip = rng.random(8000) < 0.5                                  # loadAlleleCounts symmetrises the counts itself
    refc = np.where(flip, src["tumDepth"] - src["ref"], src["ref"]).astype(int)
    f = tmp_path / "ac.txt"
    with open(f, "w") as fh:
        fh.write("chr\tposition\tref\trefCount\tNref\tNrefCount\n")
        for t in order:
            fh.write(f"chr{src['chr'][t]}\t{int(src['posn'][t])}\tA\t{refc[t]}\tC\t{int(src['tumDepth'][t]) - refc[t]}\n")
    counts = T.loadAlleleCounts(str(f))
    assert np.array_equal(counts["ref"], src["ref"]) and np.array_equal(counts["tumDepth"], src["tumDepth"])
    bins = {"chr": counts["chr"], "start": counts["posn"], "end": counts["posn"], "log2": src["logR"] / np.log(2)}
    counts["logR"] = T.staging.logRfromLog2(T.getPositionOverlap(counts["chr"], counts["posn"], bins))
    assert np.allclose(counts["logR"], src["logR"], rtol=1e-12, atol=1e-15)
    data = T.filterData(counts, chrs=[str(c) for c in range(1, 23)], minDepth=10, maxDepth=1000)
    p = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=2, data=data)
    cp = T.runEMclonalCN(data, p, maxiter=3, txnExpLen=1e15, txnZstrength=1.0, verbose=False)
    path = T.viterbiClonalCN(data, cp)
    out = T.outputTitanResults(data, cp, path, recomputeLogLik=False, verbose=False)
    segs = T.outputTitanSegments(out["corrResults"], "sample", out["convergeParams"])
    cpc = out["convergeParams"]
    corr = T.correctIntegerCN(out["corrResults"], segs, 1 - cpc["n"][-1], cpc["phi"][-1], chrs=[str(c) for c in range(1, 6)])
    fit = T.outputModelParameters(cpc, corr["cn"], str(tmp_path / "sample_cluster2.params.txt"))
    text = (tmp_path / "sample_cluster2.params.txt").read_text().splitlines()
    assert text[0].startswith("Normal contamination estimate:\t") and text[-1].startswith("S_Dbw validity index (Both):\t")
    assert np.isfinite(fit["S_Dbw"]) and int(np.sum(segs["Length.snp."])) == data["posn"].size
    assert "Corrected_Copy_Number" in corr["segs"] and len(corr["cn"]["Corrected_Call"]) == data["posn"].size


def test_estep_is_bitwise_reproducible_at_scale():
    """The E-step is a fixed-order computation (no floating-point atomics): repeated calls with the same
    parameters return bit-identical statistics, with and without the N-sized posterior outputs, on a grid large
    enough that several CTAs share every SM.  Any difference is a race (this caught shared memory being reused
    over live mbarrier words)."""
    N, Z = 400000, 3
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=20261020)
    p = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=Z)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    sol = T.Solver(data, g, Z, 1e15, 1.0)
    args = (0.45, sP["s_0"], 2.04, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    for chunk in (1024, 512):
        sol.configure(chunk, 192, 1e-8)
        ref = None
        for r in range(6):
            out = sol.estep(*args, posteriors=(r % 2 == 0))
            key = np.concatenate([out["stats"][k].ravel() for k in "abcdeg"] + [out["loglik_chr"], out["rhoG0"].ravel()])
            assert np.isfinite(key).all()
            if ref is None:
                ref = key
            assert np.array_equal(ref, key), (chunk, r)
    sol.close()


@pytest.mark.gpu
def test_speculative_final_pass_is_bit_identical_to_a_single_final_pass(monkeypatch):
    """The final pass of a chunk reads only that chunk's own verified inputs, so running it speculatively on a second
    stream while the verification cycles still run (and redoing the chunks that ran again) must not change a bit."""
    data, _ = synth.make_data(300000, Z=3, maxCN=8, n_chr=23, seed=bench_seed(), mean_seg=3000)
    p = T.loadDefaultParameters(copyNumber=8, numberClonalClusters=3, data=data)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    args = (0.42, np.array([0.01, 0.30, 0.33]), 2.03, g["var_0"], g["piG_0"], sP["piZ_0"], g)
    keys = {}
    for spec in ("1", "0"):
        monkeypatch.setenv("TITAN_SPEC_FINAL", spec)
        sol = T.Solver(data, g, 3, 1e15, 1.0)
        out = sol.estep(*args, posteriors=True)
        tm = sol.timing()
        sol.close()
        assert tm["last_fwd_rounds"] >= 3                   # at least one real cycle: the speculative pass had work to overlap
        keys[spec] = np.concatenate([out["stats"][k].ravel() for k in "abcdeg"] + [out["loglik_chr"], out["rhoG"].ravel(), out["rhoZ"].ravel()])
    assert np.array_equal(keys["1"], keys["0"])


def bench_seed():
    import bench
    return bench.SEED


@pytest.mark.gpu
def test_fused_r_glue_returns_the_runEMclonalCN_list_and_an_int_path():
    """Level 2 of the drop-in: csrc/r_glue_fused.c through the .Call emulation -- titan_solver_create_R (external
    pointer + finalizer), titan_em_run_R (the named list of R/hmmClonal.R:344-357, iteration axis trimmed to 1:i) and
    titan_viterbi_R (INTSXP[N]) -- against the host mirror runEMclonalCN / viterbiClonalCN on the same data."""
    import ctypes as C
    from oracle import refshim
    from titancna_b200 import build as B
    glue = refshim.ShimLibrary(B.GLUE)
    data, _ = synth.make_data(6000, Z=2, maxCN=5, n_chr=3, seed=11, mean_seg=300)
    p = T.loadDefaultParameters(copyNumber=5, numberClonalClusters=2, data=data)
    g, nP, pP, sP = p["genotypeParams"], p["normalParams"], p["ploidyParams"], p["cellPrevParams"]
    want = T.runEMclonalCN(data, p, txnExpLen=1e9, txnZstrength=1.0, maxiter=4, verbose=False)
    wpath = T.viterbiClonalCN(data, want)
    keep = []
    R = lambda v: glue._real(np.atleast_1d(np.asarray(v, float)), keep)
    off = T.chromosome_offsets(data["chr"]).astype(float)
    solver = glue.call("titan_solver_create_R", R(off), R(data["posn"]), R(data["ref"]), R(data["tumDepth"]), R(data["logR"]),
                       R(1e9), R(1.0), R(g["ZS"]), R(2))
    assert solver.contents.type == refshim.EXTPTRSXP and solver.contents.data
    hyper = {k: g[k] for k in ("rt", "ct", "ZS", "rn", "var_0", "alphaKHyper", "betaKHyper", "kappaGHyper", "piG_0")}
    hyper.update({k: sP[k] for k in ("s_0", "alphaSHyper", "betaSHyper", "kappaZHyper", "piZ_0")})
    hyper.update({k: nP[k] for k in ("n_0", "alphaNHyper", "betaNHyper")})
    hyper.update({k: pP[k] for k in ("phi_0", "alphaPHyper", "betaPHyper")})
    out = glue.call("titan_em_run_R", solver, glue.named_list(hyper, keep), R(4), R(1500), R(1), R(1), R(1), R(0.001))
    fit = glue.list_to_dict(out)
    assert list(fit) == ["n", "s", "var", "varR", "phi", "piG", "piZ", "muR", "muC", "loglik", "rhoG", "rhoZ"]
    U, Z, N, I = len(g["rt"]), 2, data["posn"].size, want["n"].size
    assert fit["n"].shape == (I,) and fit["s"].shape == (Z, I) and fit["var"].shape == (U, I) and fit["muR"].shape == (U * Z, I)
    assert fit["rhoG"].shape == (U, N) and fit["rhoZ"].shape == (Z, N)
    for k in ("n", "phi", "loglik", "s", "var", "piG", "piZ"):
        assert np.array_equal(fit[k], want[k]), k
    assert np.array_equal(fit["muR"].reshape((U, Z, I), order="F"), want["muR"])
    assert np.array_equal(fit["rhoG"], want["rhoG"]) and np.array_equal(fit["rhoZ"], want["rhoZ"])
    i = I - 1
    vp = glue.call("titan_viterbi_R", solver, R(fit["n"][i]), R(fit["s"][:, i]), R(fit["phi"][i]), R(fit["var"][:, i]),
                   R(fit["piG"][:, i - 1]), R(fit["piZ"][:, i - 1]), R(g["rt"]), R(g["ct"]), R(g["rn"]))
    o = vp.contents
    assert o.type == refshim.INTSXP and o.length == N
    path = np.ctypeslib.as_array(C.cast(o.data, C.POINTER(C.c_int)), shape=(N,)).copy()
    assert np.array_equal(path, wpath)
    with pytest.raises(refshim.RError, match="must be a numeric vector"):
        bad = dict(hyper)
        del bad["kappaGHyper"]
        glue.call("titan_em_run_R", solver, glue.named_list(bad, keep), R(4), R(1500), R(1), R(1), R(1), R(0.001))
    glue.lib.shim_free_all()
    assert glue.gc() >= 1                                   # the finalizer destroys the solver ...
    with pytest.raises(refshim.RError, match="not a live titan solver"):      # ... a stale handle is refused
        stale = glue.lib.shim_wrap(refshim.EXTPTRSXP, 1, None, -1, -1)
        glue.call("titan_viterbi_R", stale, *[R(0.0) for _ in range(9)])
    glue.lib.shim_free_all()


@pytest.mark.gpu
@pytest.mark.parametrize("chunk,warm,vit_mode", [(400, 128, "-1"), (0, 1, "0"), (400, 128, "4")])
def test_config4_as_named_36_genotype_states_k180(chunk, warm, vit_mode, monkeypatch):
    """BASELINE configs[3] as written: maxCN = 10 -> the 36-state genotype table (two genotypes per lane in the
    forward-backward kernels), numClusters = 5, K = 180.  The reference's C is generic in K, so its port referees the
    E-step (log-likelihoods, posteriors, the six sums) and the Viterbi path bit for bit."""
    monkeypatch.setenv("TITAN_VIT_MODE", vit_mode)      # auto = warp-reduction kernel with two genotypes per lane; 0 = dense scan
    N, Z = 2400, 5
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=2, seed=41, mean_seg=200)
    p = T.loadDefaultParameters(copyNumber=10, numberClonalClusters=Z, extendedTable=True)
    g, sP = p["genotypeParams"], p["cellPrevParams"]
    assert g["rt"].size == 36 and g["ct"][-1] == 10 and np.count_nonzero(g["ZS"] == -1) == 1
    n, phi, s = 0.35, 2.1, np.array(synth.S_TRUE[:Z]) * 0.9 + 0.02
    var = g["var_0"] * np.linspace(0.5, 1.5, g["var_0"].size)
    es = _oracle_estep(data, p, n, s, phi, var, g["piG_0"], sP["piZ_0"], 1e15, 1.0)
    sol = T.Solver(data, g, Z, 1e15, 1.0)
    sol.configure(chunk, warm, 1e-10)
    out = sol.estep(n, s, phi, var, g["piG_0"], sP["piZ_0"], g, posteriors=True)
    got = sol.viterbi(n, s, phi, var, g["piG_0"], sP["piZ_0"], g)
    sol.close()
    assert np.allclose(out["loglik_chr"], es["loglik_chr"], rtol=LL_RTOL, atol=0)
    assert np.abs(out["rhoG"] - es["rhoG"]).max() < 1e-9 and np.abs(out["rhoZ"] - es["rhoZ"]).max() < 1e-9
    for k in "abcdeg":
        ref = es["stats"][k]
        assert np.abs(out["stats"][k] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), k
    lp = np.vectorize(math.log)(np.outer(g["piG_0"], sP["piZ_0"]).reshape(-1, order="F"))
    want = np.empty(N, dtype=np.int32)
    for idx in O.chromosome_slices(data["chr"]):
        want[idx] = O.viterbi(lp, es["py"][:, idx], g["ZS"], Z, data["posn"][idx], 1e15, 1e15, engine="port")
    assert got.max() <= 180 and np.array_equal(got, want), f"first mismatch at {np.flatnonzero(got != want)[:5]}"


@pytest.mark.gpu
def test_config4_as_named_em_matches_oracle_em():
    """runEMclonalCN on the 36-state table (K = 108 at three clusters) against the oracle's EM restatement."""
    data, _ = synth.make_data(3000, Z=3, maxCN=8, n_chr=2, seed=43, mean_seg=250)
    p = T.loadDefaultParameters(copyNumber=10, numberClonalClusters=3, data=data, extendedTable=True)
    got = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=3, verbose=False)
    want = O.run_em_clonal_cn(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=3, engine="port", threads=8)
    for k in ("n", "phi", "s", "var"):
        assert np.allclose(got[k], want[k], rtol=0, atol=1e-6), k
    assert np.allclose(got["loglik"][1:], want["loglik"][1:], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("col", [0, 1, 2])
def test_fused_estep_reproduces_the_fixture_loglik_trajectory(golden, col):
    """data/EMresults.rda: loglik[i] = E-step log-likelihood at the parameters of column i-1 + priorProbs(column i)
    (R/hmmClonal.R:218,311).  The fused E-step through the C ABI, at each of the three stored parameter sets, plus the
    priors of the stored next column, gives the stored value to 1e-11 relative (contract: 1e-9)."""
    g = {k[2:]: golden[k] for k in golden if k.startswith("g_")}
    g["rn"] = float(g["rn"][0])
    Z = golden["cp_s"].shape[0]
    data = {"chr": golden["data_chr"], "posn": golden["data_posn"].astype(float), "ref": golden["data_ref"].astype(float),
            "tumDepth": golden["data_tumDepth"].astype(float), "logR": golden["data_logR"]}
    sol = T.Solver(data, g, Z, float(golden["cp_txnExpLen"][0]), float(golden["cp_txnZstrength"][0]))
    out = sol.estep(golden["cp_n"][col], golden["cp_s"][:, col], golden["cp_phi"][col], golden["cp_var"][:, col],
                    golden["cp_piG"][:, col], golden["cp_piZ"][:, col], g, posteriors=(col == 2))
    sol.close()
    nP = {"alphaNHyper": 2.0, "betaNHyper": 2.0}
    sP = {"alphaSHyper": np.full(Z, 2.0), "betaSHyper": np.full(Z, 2.0), "kappaZHyper": golden["s_kappaZHyper"]}
    pP = {"alphaPHyper": 20.0, "betaPHyper": 42.0}
    c1 = col + 1
    pr = O.prior_probs(golden["cp_n"][c1], golden["cp_s"][:, c1], golden["cp_phi"][c1], golden["cp_var"][:, c1],
                       golden["cp_piG"][:, c1], golden["cp_piZ"][:, c1], g, nP, sP, pP)
    want = float(golden["cp_loglik"][c1])
    assert abs(out["loglik"] + pr["prior"] - want) <= 1e-11 * abs(want)
    assert np.abs(out["muR"] - golden["cp_muR"][:, :, col]).max() < 1e-15
    if col == 2:                                      # the stored marginals are those of the last E-step
        assert np.abs(out["rhoG"] - golden["cp_rhoG"]).max() < 1e-9 and np.abs(out["rhoZ"] - golden["cp_rhoZ"]).max() < 1e-9


@pytest.mark.gpu
def test_em_without_posterior_download_gives_the_same_trajectory():
    """runEMclonalCN(posteriors=False) skips the one N-sized device-to-host copy (what the sweep does): identical
    parameter trajectory bit for bit, rhoG / rhoZ None, and the result tables / segments still come out."""
    data, _ = synth.make_data(20000, Z=2, maxCN=5, n_chr=4, seed=77, mean_seg=400)
    p = T.loadDefaultParameters(copyNumber=5, numberClonalClusters=2, data=data)
    a = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=4, verbose=False)
    b = T.runEMclonalCN(data, p, txnExpLen=1e15, txnZstrength=1.0, maxiter=4, verbose=False, posteriors=False)
    assert b["rhoG"] is None and b["rhoZ"] is None and a["rhoG"].shape == (len(p["genotypeParams"]["rt"]), 20000)
    for k in ("n", "phi", "s", "var", "piG", "piZ", "loglik", "muR", "muC"):
        assert np.array_equal(a[k], b[k]), k
    path = T.viterbiClonalCN(data, b)
    assert np.array_equal(path, T.viterbiClonalCN(data, a))
    res = T.outputTitanResults(data, b, path, recomputeLogLik=False, verbose=False)
    assert len(res["corrResults"]["TITANstate"]) == 20000
