"""CPU: decode + segmenter -- the host mirror against the loop restatement."""
import numpy as np
import pytest

import titancna_b200 as T
from titancna_b200 import results as R
from oracle import results_oracle as RO
from tests import synth


def _fake_cp(Z, U, s):
    return {"s": np.column_stack([np.linspace(0.1, 0.5, Z), np.asarray(s, float)]), "var": np.zeros((U, 2)),
            "useOutlierState": False, "symmetric": True}


@pytest.mark.parametrize("Z,s", [(1, [0.02]), (3, [0.4, 0.01, 0.2]), (5, [0.5, 0.1, 0.3, 0.2, 0.05])])
def test_results_and_segments_match_loop_restatement(Z, s):
    rng = np.random.default_rng(Z)
    data, _ = synth.make_data(4000, Z=Z, maxCN=8, n_chr=4, seed=Z, mean_seg=60)
    U, N = 25, data["posn"].size
    path = np.repeat(rng.integers(1, U * Z + 1, size=N // 7 + 1), 7)[:N].astype(np.int32)
    cp = _fake_cp(Z, U, s)
    res = R.outputTitanResults(data, cp, path, correctResults=False)["results"]
    rows = RO.output_results({k: np.asarray(v).tolist() for k, v in data.items()}, cp, path)
    for k in ("CopyNumber", "TITANstate", "ClonalCluster", "CellularPrevalence", "AllelicRatio", "LogRatio"):
        a = np.asarray(res[k], float)
        b = np.array([r[k] for r in rows], float)
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        assert np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-13, atol=1e-15), k   # numpy SIMD vs libm exp/log2: 1 ulp before the log2(exp(x)) cancellation
    assert list(res["TITANcall"]) == [r["TITANcall"] for r in rows]
    segs = R.outputTitanSegments(res, "test", cp)
    want = RO.output_segments(rows, "test")
    assert len(want) == segs["Length.snp."].size and segs["Length.snp."].sum() == N
    for k in ("Start_Position.bp.", "End_Position.bp.", "Length.snp.", "Median_Ratio", "Median_logR", "TITAN_state",
              "Copy_Number", "MinorCN", "MajorCN", "Clonal_Cluster", "Cellular_Prevalence"):
        a = np.asarray(segs[k], float)
        b = np.array([w[k] for w in want], float)
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), k
    assert list(segs["TITAN_call"]) == [w["TITAN_call"] for w in want]
    assert list(segs["Chromosome"]) == [w["Chromosome"] for w in want]


def test_segments_of_the_reference_fixture_path(golden):
    """Known answer from SURVEY.md 8c: the fixture's Viterbi path has 26 runs of the joint state; the
    segmenter runs on the GENOTYPE state only, so it may merge runs that differ in cluster alone."""
    path = golden["ref_path"].astype(np.int32)
    data = {"chr": golden["data_chr"], "posn": golden["data_posn"].astype(float), "ref": golden["data_ref"].astype(float),
            "tumDepth": golden["data_tumDepth"].astype(float), "logR": golden["data_logR"]}
    cp = {"s": golden["cp_s"], "var": golden["cp_var"], "useOutlierState": False, "symmetric": True}
    res = R.outputTitanResults(data, cp, path, correctResults=False)["results"]
    segs = R.outputTitanSegments(res, "chr2", cp)
    joint_runs = 1 + np.count_nonzero(np.diff(path))
    geno_runs = 1 + np.count_nonzero(np.diff((path - 1) % 12))
    assert joint_runs == 26 and segs["Length.snp."].size == geno_runs <= 26
    assert segs["Length.snp."].sum() == path.size
    assert set(res["TITANcall"]) <= {"HOMD", "DLOH", "NLOH", "HET", "ALOH", "GAIN", "ASCNA", "BCNA", "UBCNA"}
    het = (res["TITANcall"] == "HET") & (res["CopyNumber"] == 2)
    assert np.all(np.isnan(res["ClonalCluster"][het]))


def test_decode_tables_and_quirks():
    d = R.decoupleMegaVar(np.array([1, 25, 26, 50, 75]), 25)
    assert d["G"].tolist() == [1, 25, 1, 25, 25] and d["Z"].tolist() == [1, 1, 2, 2, 3]
    dec = R.decodeLOH(np.arange(25))
    for g in range(25):
        call, cn = RO.decode_loh(g)
        assert dec["G"][g] == call and dec["CN"][g] == cn
        assert R.getMajorMinorCN(g) == {"majorCN": RO._MM[g][0], "minorCN": RO._MM[g][1]}
        assert sum(RO._MM[g]) == cn


def _cluster_cp(Z, U, s_last, n_last=0.3, N=10):
    rng = np.random.default_rng(7)
    return {"s": np.column_stack([np.linspace(0.1, 0.5, Z), np.asarray(s_last, float)]), "var": np.zeros((U, 2)),
            "n": np.array([0.5, n_last]), "phi": np.array([2.0, 2.1]), "piZ": rng.random((Z, 2)), "rhoZ": rng.random((Z, N)),
            "muC": rng.random((U, Z, 2)), "muR": rng.random((U, Z, 2)), "loglik": np.array([-np.inf, -5.0]),
            "cellPrevParams": {"s_0": np.linspace(0.1, 0.5, Z), "alphaSHyper": np.full(Z, 2.0), "betaSHyper": np.full(Z, 2.0),
                               "kappaZHyper": np.full(Z, 2.0), "piZ_0": np.full(Z, 1 / Z)},
            "useOutlierState": False, "symmetric": True}


@pytest.mark.parametrize("name,Z,s_last,weights,thr", [
    ("all kept", 3, [0.4, 0.01, 0.2], [0.3, 0.3, 0.4], 0.05),
    ("clonal cluster below its own threshold", 3, [0.05, 0.3, 0.6], [0.03, 0.5, 0.47], 0.001),
    ("middle cluster empty: reassigned to the right", 3, [0.3, 0.1, 0.2], [0.5, 0.01, 0.49], 0.05),
    ("last cluster empty: reassigned to the left", 4, [0.1, 0.2, 0.3, 0.4], [0.4, 0.3, 0.29, 0.01], 0.05),
    ("first two empty", 4, [0.4, 0.3, 0.2, 0.1], [0.01, 0.02, 0.5, 0.47], 0.05),
    ("unsorted prevalences", 5, [0.5, 0.1, 0.3, 0.2, 0.05], [0.3, 0.02, 0.3, 0.03, 0.35], 0.05),
    ("nothing populated", 3, [0.2, 0.1, 0.3], [0.02, 0.02, 0.02], 0.05),
    ("single cluster", 1, [0.02], [0.9], 0.05)])
def test_remove_empty_clusters_matches_loop_restatement(name, Z, s_last, weights, thr):
    """R/utils.R:1420-1498 (the part of correctResults=TRUE before the E-step re-entry)."""
    rng = np.random.default_rng(len(name))
    N, U = 3000, 25
    data, _ = synth.make_data(N, Z=max(Z, 1), maxCN=8, n_chr=3, seed=3, mean_seg=50)
    w = np.asarray(weights, float)
    lab = rng.choice(np.arange(Z + 1), size=N, p=np.append(w, 1 - w.sum()))      # label Z -> diploid HET (no cluster)
    geno = np.where(lab == Z, 4, rng.choice([1, 2, 3, 5, 6, 9, 14, 25], size=N))  # unit state 4 = HET
    path = (geno + np.minimum(lab, Z - 1) * U).astype(np.int32)
    cp = _cluster_cp(Z, U, s_last, N=N)
    out = R.outputTitanResults(data, cp, path, correctResults=True, proportionThreshold=thr, recomputeLogLik=False)
    rows = RO.output_results({k: np.asarray(v).tolist() for k, v in data.items()}, cp, path)
    want_cp, want_rows = RO.remove_empty_clusters(cp, rows, thr, 0.05)
    got_cp, got = out["convergeParams"], out["corrResults"]
    for k in ("ClonalCluster", "CellularPrevalence"):
        a, b = np.asarray(got[k], float), np.array([r[k] for r in want_rows], float)
        assert np.array_equal(np.isnan(a), np.isnan(b)), (name, k)
        assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), (name, k)
    for k in ("s", "n", "piZ", "rhoZ", "muC", "muR"):
        assert np.array_equal(np.asarray(got_cp[k]), np.asarray(want_cp[k])), (name, k)
    for k in cp["cellPrevParams"]:
        assert np.array_equal(got_cp["cellPrevParams"][k], want_cp["cellPrevParams"][k]), (name, k)
    # the inputs are untouched (R passes by value) and the uncorrected table is returned next to the corrected one
    assert np.array_equal(np.asarray(cp["s"])[:, 1], np.asarray(s_last, float))
    base = R.outputTitanResults(data, cp, path, correctResults=False)
    assert base["corrResults"] is None and base["convergeParams"] is cp
    a, b = np.asarray(out["results"]["ClonalCluster"]), np.asarray(base["results"]["ClonalCluster"])
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    if name == "all kept":
        assert got_cp["s"].shape[0] == Z and got_cp["n"][1] == 1 - (1 - 0.01) * (1 - 0.3)
    if name == "nothing populated":
        assert got_cp["s"].shape == (1, 2) and got_cp["s"][0, 1] == 0.0 and got_cp["n"][1] == 1 - 0.1


def test_recompute_loglik_without_transition_fields_is_an_error_not_undefined_behaviour():
    """R/utils.R:1514-1515 reads convergeParams$txn_exp_len / $txn_z_strength, which nothing sets: the reference
    passes NULL to its C code.  The mirror refuses instead."""
    data, _ = synth.make_data(500, Z=2, maxCN=8, n_chr=2, seed=3, mean_seg=50)
    cp = _cluster_cp(2, 25, [0.1, 0.3], N=500)
    path = np.full(500, 1, dtype=np.int32)
    with pytest.raises(ValueError, match="txn_exp_len"):
        R.outputTitanResults(data, cp, path)            # reference defaults: correctResults=TRUE, recomputeLogLik=TRUE


@pytest.mark.parametrize("gender,purity,max_auto", [("male", 0.7, None), ("female", 0.7, None), ("male", 0.1, None),
                                                     ("female", 0.45, 6), ("male", 0.9, 5)])
def test_correct_integer_cn_matches_loop_restatement(gender, purity, max_auto):
    """R/utils.R:1175-1340 (scripts/R_scripts/titanCNA.R:278-281): copy numbers re-derived from the corrected log
    ratio for high-level amplifications, homozygous deletions, chrX of males and calls that contradict the ratio."""
    rng = np.random.default_rng(11)
    N, Z, U = 3000, 2, 25
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=23, seed=17, mean_seg=40)
    data = dict(data)
    data["chr"] = np.where(np.asarray(data["chr"]) == 23, "X", np.asarray(data["chr"]).astype(str))
    data["logR"] = data["logR"] + rng.normal(0, 0.5, size=N)            # push some calls against their log ratio
    path = np.repeat(rng.integers(1, U * Z + 1, size=N // 6 + 1), 6)[:N].astype(np.int32)
    cp = _cluster_cp(Z, U, [0.05, 0.4], N=N)
    tab = R.outputTitanResults(data, cp, path, correctResults=False)["results"]
    segs = R.outputTitanSegments(tab, "s", cp)
    rows = RO.output_results({k: np.asarray(v).tolist() for k, v in data.items()}, cp, path)
    seg_rows = RO.output_segments(rows, "s")
    got = R.correctIntegerCN(tab, segs, purity, 2.3, maxCNtoCorrect_autosomes=max_auto, gender=gender)
    w_cn, w_segs = RO.correct_integer_cn(rows, seg_rows, purity, 2.3, max_auto=max_auto, gender=gender)
    for table, want, keys in ((got["cn"], w_cn, ("logR_Copy_Number", "Corrected_logR", "Corrected_Ratio", "Corrected_Copy_Number")),
                              (got["segs"], w_segs, ("logR_Copy_Number", "Corrected_logR", "Corrected_Ratio", "Corrected_Copy_Number",
                                                     "Corrected_MajorCN", "Corrected_MinorCN"))):
        for k in keys:
            a, b = np.asarray(table[k], float), np.array([r[k] for r in want], float)
            assert np.array_equal(np.isnan(a), np.isnan(b)), (k, gender)
            assert np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-12, atol=1e-12), (k, gender)
        assert list(table["Corrected_Call"]) == [r["Corrected_Call"] for r in want]
    changed = np.count_nonzero(got["segs"]["Corrected_Copy_Number"] != np.asarray(segs["Copy_Number"], float))
    assert (changed > 0) == (purity >= 0.2)
    assert "Corrected_Copy_Number" not in tab and "Corrected_Copy_Number" not in segs          # inputs untouched
    if gender == "male" and purity >= 0.2:
        x = np.asarray(got["cn"]["Chr"]).astype(str) == "X"
        assert x.any() and np.all(np.isnan(got["cn"]["Corrected_Ratio"][x]))
        assert set(got["cn"]["Corrected_Call"][x]) <= {"HETD", "NEUT", "GAIN", "AMP", "HLAMP"}
    # the S_Dbw index prefers the corrected copy number when the table has it (R/utils.R:631-637)
    from titancna_b200 import selection as S
    assert S.computeSDbwIndex(got["cn"], data_type="LogRatio")["S_DbwIndex"] > 0


@pytest.mark.parametrize("Z,s_last", [(1, [0.2]), (2, [0.1, 0.45])])
def test_subclone_profiles_match_loop_restatement(Z, s_last):
    """R/utils.R:1526-1577 through outputTitanResults(subcloneProfiles=TRUE), the production setting."""
    rng = np.random.default_rng(Z)
    N, U = 1500, 25
    data, _ = synth.make_data(N, Z=Z, maxCN=8, n_chr=3, seed=9, mean_seg=40)
    path = np.repeat(rng.integers(1, U * Z + 1, size=N // 5 + 1), 5)[:N].astype(np.int32)
    cp = _cluster_cp(Z, U, s_last, N=N)
    tab = R.outputTitanResults(data, cp, path, correctResults=False)["results"]
    rows = RO.output_results({k: np.asarray(v).tolist() for k, v in data.items()}, cp, path)
    want = RO.subclone_profiles(rows)
    for k in want[0]:
        if k.endswith("TITANcall"):
            assert list(tab[k]) == [w[k] for w in want], k
        else:
            a, b = np.asarray(tab[k], float), np.array([w[k] for w in want], float)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=0, atol=1e-15), k
    assert ("Subclone2.Prevalence" in tab) == (Z == 2)
    three = R.outputTitanResults(data, _cluster_cp(3, U, [0.1, 0.2, 0.3], N=N), np.minimum(path, 75), correctResults=False)["results"]
    assert "Subclone1.CopyNumber" not in three                       # more than two clusters: no profiles (:943-948)
