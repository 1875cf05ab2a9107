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
    res = R.outputTitanResults(data, cp, path)
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
    res = R.outputTitanResults(data, cp, path)
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
    with pytest.raises(NotImplementedError):
        R.outputTitanResults({}, {"useOutlierState": False, "symmetric": True}, [], correctResults=True)
