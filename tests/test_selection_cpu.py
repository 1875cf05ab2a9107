"""CPU: S_Dbw validity index, *.params.txt fields and the solution-selection rule --
the numpy mirror against the loop restatement of R/utils.R:604-768,978-1063 and selectSolution.R:58-137."""
import math

import numpy as np
import pytest

from titancna_b200 import results as R
from titancna_b200 import selection as S
from oracle import results_oracle as RO
from oracle import selection_oracle as SO
from tests import synth


def _table(Z, seed, N=2500):
    rng = np.random.default_rng(seed)
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=3, seed=seed, mean_seg=40)
    path = np.repeat(rng.integers(1, 25 * Z + 1, size=N // 9 + 1), 9)[:N].astype(np.int32)
    s_last = np.sort(rng.uniform(0.01, 0.6, size=Z))[::-1].copy()
    cp = {"s": np.column_stack([np.linspace(0.1, 0.5, Z), s_last]), "var": np.full((25, 2), 0.02),
          "n": np.array([0.5, 0.31234567]), "phi": np.array([2.0, 2.0456789]), "loglik": np.array([-np.inf, -2142676.4127]),
          "muR": rng.uniform(0.5, 1, size=(25, Z, 2)), "muC": rng.normal(0, 0.4, size=(25, Z, 2)),
          "genotypeParams": {"alleleEmissionModel": "binomial"}, "useOutlierState": False, "symmetric": True}
    tab = R.outputTitanResults(data, cp, path, correctResults=False)["results"]
    rows = RO.output_results({k: np.asarray(v).tolist() for k, v in data.items()}, cp, path)
    return cp, tab, rows


@pytest.mark.parametrize("Z,seed", [(1, 1), (2, 2), (3, 3)])
@pytest.mark.parametrize("method", ["Halkidi", "Tong"])
@pytest.mark.parametrize("data_type", ["LogRatio", "AllelicRatio"])
def test_sdbw_index_matches_loop_restatement(Z, seed, method, data_type):
    cp, tab, rows = _table(Z, seed)
    got = S.computeSDbwIndex(tab, data_type=data_type, S_Dbw_method=method)
    want = SO.compute_sdbw_index(rows, data_type=data_type, method=method)
    for k in ("scat", "dens_bw", "S_DbwIndex"):
        assert math.isfinite(want[k]) and got[k] == pytest.approx(want[k], rel=1e-10, abs=1e-12), (k, got, want)
    with pytest.raises(ValueError, match="data.type"):
        S.computeSDbwIndex(tab, data_type="Depth")
    with pytest.raises(ValueError, match="S_Dbw.method"):
        S.computeSDbwIndex(tab, S_Dbw_method="Other")


def test_model_parameter_file_fields():
    """The lines selectSolution.R greps for (formatParams, :58-72) with R's signif()/sprintf formatting."""
    cp, tab, rows = _table(3, 5)
    out = S.outputModelParameters(cp, tab)
    lines = out["lines"]
    assert lines[0] == "Normal contamination estimate:\t0.3123"
    assert lines[1] == "Average tumour ploidy estimate:\t2.046"
    prev = " ".join(S._rnum(1 - v, 4) for v in np.sort(cp["s"][:, 1]))
    assert lines[2] == "Clonal cluster cellular prevalence Z=3:\t" + prev
    assert lines[3].startswith("AllelicRatio binomial means for clonal cluster Z=1:\t") and len(lines[3].split("\t")[1].split(" ")) == 25
    assert "Number of iterations:\t2" in lines and "Log likelihood:\t-2142680" in lines      # signif(, 6) of -2142676.4127
    lr = SO.compute_sdbw_index(rows, data_type="LogRatio", method="Tong")
    ar = SO.compute_sdbw_index(rows, data_type="AllelicRatio", method="Tong")
    assert out["S_Dbw"] == pytest.approx(lr["dens_bw"] + ar["dens_bw"] + lr["scat"] + ar["scat"], rel=1e-10)
    assert lines[-1] == "S_Dbw validity index (Both):\t%0.4f " % out["S_Dbw"]
    assert [ln.split("\t")[0] for ln in lines[-9:]] == [f"S_Dbw {w} ({d}):" for d in ("LogRatio", "AllelicRatio", "Both")
                                                        for w in ("dens.bw", "scat", "validity index")]
    assert S._rnum(1e-5, 4) == "1e-05" and S._rnum(float("nan"), 4) == "NA" and S._rnum(123456.789, 4) == "123500"


def test_model_parameter_file_of_the_gaussian_allele_model():
    """R/utils.R:1006,1018-1023: the means line names the allele model and the allelic variances get their own line
    between the means and the logR variances."""
    cp, tab, _ = _table(2, 7)
    U = np.asarray(cp["var"]).shape[0]
    cp = dict(cp)
    cp["genotypeParams"] = dict(cp.get("genotypeParams") or {}, alleleEmissionModel="Gaussian")
    cp["varR"] = np.tile(np.linspace(0.011, 0.019, U)[:, None], (1, np.asarray(cp["var"]).shape[1]))
    lines = S.outputModelParameters(cp, tab)["lines"]
    heads = [ln.split("\t")[0] for ln in lines]
    assert heads[3] == "AllelicRatio Gaussian means for clonal cluster Z=1:"
    k = heads.index("AllelicRatio Gaussian variance:")
    assert heads[k + 1] == "logRatio Gaussian variance:" and heads[k - 1].startswith("logRatio Gaussian means for clonal cluster Z=2")
    assert lines[k].split("\t")[1].split(" ")[0] == "0.011" and len(lines[k].split("\t")[1].split(" ")) == U
    binom = S.outputModelParameters(dict(cp, genotypeParams={"alleleEmissionModel": "binomial"}), tab)["lines"]
    assert "AllelicRatio Gaussian variance:" not in [ln.split("\t")[0] for ln in binom] and len(binom) == len(lines) - 1


def _records(ll2, ll3, ll4, sdbw):
    out = []
    for p, lls in ((2, ll2), (3, ll3), (4, ll4)):
        for z, ll in enumerate(lls, start=1):
            out.append({"ploidy_0": p, "numClusters": z, "loglik": ll, "sdbw": sdbw[(p, z)]})
    return out


def test_solution_selection_rule():
    sd = {(p, z): 1.0 + 0.1 * ((p * 7 + z * 3) % 5) for p in (2, 3, 4) for z in (1, 2, 3)}
    cases = [
        ([-100.0, -99.0, -98.0], [-97.0, -96.0, -95.0], [-120.0, -119.0, -118.0]),   # ploidy 2 wins through its 5 % bonus
        ([-100.0, -99.0, -98.0], [-90.0, -89.0, -88.0], [-120.0, -119.0, -118.0]),   # ploidy 3 clearly better
        ([-100.0, -99.0, -98.0], [-96.0, -80.0, -95.0], [-80.0, -93.9, -70.0]),      # split vote: 4 wins two rows
        ([-100.0, -99.0, -98.0], [], [-80.0, -93.9, -99.0]),                         # no ploidy-3 run at all
    ]
    for ll2, ll3, ll4 in cases:
        rec = _records(ll2, ll3, ll4, sd)
        got, want = S.selectSolution(rec), SO.select_solution(rec)
        assert (got["ploidy_0"], got["numClusters"]) == (want["ploidy_0"], want["numClusters"])
    assert S.selectSolution(_records(*cases[0], sd))["ploidy_0"] == 2
    assert S.selectSolution(_records(*cases[1], sd))["ploidy_0"] == 3
    assert S.selectSolution(_records(*cases[2], sd))["ploidy_0"] == 4
    # homozygous-deletion veto (:81-90): the otherwise best solution is discarded
    rec = _records(*cases[1], sd)
    best = S.selectSolution(rec)
    for r in rec:
        r["segs"] = {"TITAN_call": np.array(["HET", "DLOH"]), "Length.snp.": np.array([5000, 400])}
    veto = next(r for r in rec if (r["ploidy_0"], r["numClusters"]) == (best["ploidy_0"], best["numClusters"]))
    veto["segs"] = {"TITAN_call": np.array(["HET", "HOMD", "DLOH"]), "Length.snp.": np.array([5000, 1600, 400])}
    got, want = S.selectSolution(rec), SO.select_solution(rec)
    assert (got["ploidy_0"], got["numClusters"]) == (want["ploidy_0"], want["numClusters"]) != (best["ploidy_0"], best["numClusters"])
    with pytest.raises(ValueError):
        S.selectSolution([r for r in rec if r["ploidy_0"] == 2])
