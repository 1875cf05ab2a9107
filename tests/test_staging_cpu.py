"""CPU: input staging.  loadAlleleCounts + filterData are pinned against the reference's own
fixture: data/EMresults.rda `data` was built from inst/extdata/test_alleleCounts_chr2.txt by exactly these two
functions (vignettes/TitanCNA.Rnw:93-120); tests/golden/allelecounts_chr2.npz holds that file's numeric columns."""
import os

import numpy as np
import pytest

from titancna_b200 import staging as G

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def counts():
    z = np.load(os.path.join(HERE, "golden", "allelecounts_chr2.npz"))
    return {"chr": z["chr"].astype(str), "position": z["position"], "ref": np.full(z["chr"].size, "A"),
            "refCount": z["refCount"], "Nref": np.full(z["chr"].size, "X"), "NrefCount": z["NrefCount"]}


def test_load_and_filter_reproduce_the_reference_fixture(counts, golden):
    data = G.loadAlleleCounts(counts)
    assert data["posn"].size == counts["position"].size == 19999 and np.all(np.diff(data["posn"]) >= 0)
    assert np.array_equal(data["ref"], np.maximum(counts["refCount"], counts["NrefCount"]).astype(float)[np.argsort(counts["position"], kind="stable")])
    assert np.array_equal(data["tumDepth"], data["refOriginal"] + data["nonRef"])
    fx_posn = golden["data_posn"]
    idx = np.searchsorted(data["posn"], fx_posn)
    assert np.array_equal(data["posn"][idx], fx_posn)                       # every fixture SNP comes from the file
    assert np.array_equal(data["ref"][idx], golden["data_ref"].astype(float))            # symmetric reference count
    assert np.array_equal(data["tumDepth"][idx], golden["data_tumDepth"].astype(float))
    # logR as the vignette attached it (GC correction itself is out of scope): filterData must keep exactly the
    # fixture's rows -- depth window 10..200, missing logR dropped
    logR = np.full(data["posn"].size, np.nan)
    logR[idx] = golden["data_logR"]
    data["logR"] = logR
    out = G.filterData(data, chrs=[str(c) for c in range(1, 23)] + ["X", "Y"], minDepth=10, maxDepth=200)
    for k, ref in (("posn", "data_posn"), ("ref", "data_ref"), ("tumDepth", "data_tumDepth"), ("logR", "data_logR")):
        assert np.array_equal(np.asarray(out[k], float), np.asarray(golden[ref], float)), k
    assert set(out["chr"]) == {"2"}
    tight = G.filterData(data, minDepth=30, maxDepth=60)
    assert tight["posn"].size and tight["tumDepth"].min() >= 30 and tight["tumDepth"].max() <= 60


def test_file_loader_sorting_and_format_check(tmp_path):
    good = tmp_path / "ac.txt"
    good.write_text("chr\tposition\tref\trefCount\tNref\tNrefCount\n"
                    "chr10\t500\tA\t7\tC\t9\nchrX\t40\tG\t5\tT\t1\nchr2\t900\tA\t3\tC\t8\nchr2\t100\tA\t6\tG\t6\n"
                    "chrUn_gl000220\t5\tA\t1\tC\t1\nchr1\t77\tT\t2\tC\t0\n")
    d = G.loadAlleleCounts(str(good))
    assert d["chr"].tolist() == ["1", "2", "2", "10", "X"] and d["posn"].tolist() == [77, 100, 900, 500, 40]
    assert d["ref"].tolist() == [2, 6, 8, 9, 5] and d["refOriginal"].tolist() == [2, 6, 3, 7, 5]
    assert G.loadAlleleCounts(str(good), symmetric=False)["ref"].tolist() == [2, 6, 3, 7, 5]
    assert G.loadAlleleCounts(str(good), genomeStyle="UCSC")["chr"].tolist() == ["chr1", "chr2", "chr2", "chr10", "chrX"]
    bad = tmp_path / "bad.txt"
    bad.write_text("chr\tposition\tref\trefCount\tNref\tNrefCount\n1\t10\tA\t0.5\tC\t2\n")
    with pytest.raises(ValueError, match="required specifications"):
        G.loadAlleleCounts(str(bad))
    with pytest.raises(ValueError, match="filename or data.frame"):
        G.loadAlleleCounts(42)


def test_filter_options_and_centromeres():
    data = {"chr": np.array(["1"] * 5 + ["2"] * 3 + ["3"]), "posn": np.array([10, 20, 30, 40, 50, 5, 15, 25, 7]),
            "ref": np.arange(9.0), "refOriginal": np.arange(9.0), "nonRef": np.ones(9), "tumDepth": np.full(9, 50.0),
            "logR": np.array([0.1, np.nan, 0.2, 0.3, 0.1, 0.0, 0.2, 0.1, 0.4])}
    out = G.filterData(data, chrs=["1", "2", "3"], map=np.array([1, 1, 0.5, 1, 1, 1, np.nan, 1, 1.0]), mapThres=0.9,
                       centromeres={"Chr": ["chr1"], "Start": [38], "End": [45]}, centromere_flankLength=2)
    # NA logR, low / NA map score, centromere 36..47 on chr1, single-SNP chromosome 3
    assert out["chr"].tolist() == ["1", "1", "2", "2"] and out["posn"].tolist() == [10, 50, 5, 25]
    white = G.filterData(data, positionList={"chr": ["1", "2"], "posn": [30, 15]})
    assert white["posn"].tolist() == [30, 15]


def test_position_overlap_against_naive_scan():
    rng = np.random.default_rng(3)
    for max_len in (350, 1500):                                 # disjoint bins, then overlapping intervals
        start = np.sort(rng.choice(np.arange(1, 100000, 400), size=200, replace=False))
        end = start + rng.integers(0, max_len, size=200)
        val = rng.normal(size=200)
        dv = {"chr": np.where(np.arange(200) % 2 == 0, "1", "2"), "start": start, "end": end, "logR": val}
        chr_ = np.where(rng.random(3000) < 0.5, "1", "2")
        posn = rng.integers(1, 100010, size=3000)
        got = G.getPositionOverlap(chr_, posn, dv)
        want = np.full(3000, np.nan)
        for q in range(3000):
            for j in range(200):
                if dv["chr"][j] == chr_[q] and start[j] <= posn[q] <= end[j]:
                    want[q] = val[j]                           # last hit wins
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
        assert np.count_nonzero(~np.isnan(got)) > 300
    # overlapping intervals: the later interval overrides the earlier one
    ov = {"chr": ["1", "1"], "start": [10, 15], "end": [20, 30], "logR": [1.0, 2.0]}
    assert G.getPositionOverlap(["1"] * 4, [9, 12, 17, 30], ov).tolist()[1:] == [1.0, 2.0, 2.0]
    assert np.isnan(G.getPositionOverlap(["1"], [9], ov)[0])
    assert np.allclose(G.logRfromLog2([0.0, 1.0, -1.0]), [0.0, np.log(2), -np.log(2)])


def test_c_abi_position_overlap_matches_the_reference_object_code():
    """titan_get_position_overlap against the reference's own getPositionOverlapC (oracle/_ref), including its
    integer truncation of the coordinates and first-hit rule."""
    from oracle import refshim
    rng = np.random.default_rng(5)
    start = np.sort(rng.integers(1, 5000, size=60)).astype(float) + 0.7
    stop = start + rng.integers(0, 40, size=60)
    posn = rng.integers(1, 5100, size=800).astype(float) + 0.2
    want = refshim.reference().position_overlap(posn, start, stop)
    got = G.position_overlap_index(posn, start, stop)
    assert np.array_equal(got, np.asarray(want, float)) and got.max() > 0
