"""Regenerates tests/golden/emresults_chr2.npz from the reference checkout.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

Contents
  data_*          the reference fixture's `data` object (data/EMresults.rda; chr2, N=15,623)
  cp_*            its `convergeParams` (init + 3 EM iterations, U=12, Z=2) incl. rhoG/rhoZ
  ref_loglik      fwd_backC_clonalCN log-likelihood from the reference's own C (oracle/_ref)
  ref_path        viterbiC_clonalCN path from the reference's own C, both run with the
                  fixture's column-3 parameters
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import rdx, refshim, titan_oracle as O  # noqa: E402

REF = os.environ.get("TITAN_REFERENCE", "/root/reference")


def main():
    d = rdx.load_rda(os.path.join(REF, "data", "EMresults.rda"))
    data, cp = d["data"], d["convergeParams"]
    g = cp["genotypeParams"]
    out = {}
    for k in ("posn", "ref", "tumDepth"):
        assert np.all(data[k] == np.round(data[k]))
        out["data_" + k] = data[k].astype(np.int32)
    out["data_logR"] = data["logR"]
    out["data_chr"] = data["chr"].astype(np.int32)
    for k in ("n", "s", "var", "phi", "piG", "piZ", "muR", "muC", "loglik", "rhoG", "rhoZ"):
        out["cp_" + k] = cp[k]
    out["cp_txnExpLen"] = cp["txnExpLen"]
    out["cp_txnZstrength"] = cp["txnZstrength"]
    for k in ("rt", "rn", "ZS", "ct", "var_0", "alphaKHyper", "betaKHyper", "kappaGHyper"):
        out["g_" + k] = np.asarray(g[k], dtype=np.float64)
    out["s_kappaZHyper"] = cp["cellPrevParams"]["kappaZHyper"]

    c = 2  # column 3 (1-based): the parameters the stored rhoG/rhoZ were computed with
    muR, muC = O.clonal_two_component_mixture(g["rt"], g["rn"][0], cp["n"][c], cp["s"][:, c], g["ct"], cp["phi"][c])
    py = O.compute_py(data["ref"], data["tumDepth"], data["logR"], muR, muC, cp["var"][:, c])
    lp = np.log(np.outer(cp["piG"][:, c], cp["piZ"][:, c]).reshape(-1, order="F"))
    txn = float(cp["txnExpLen"][0])
    zst = float(cp["txnZstrength"][0]) * txn
    ref = refshim.reference()
    fb = ref.fwd_back(lp, py, g["ct"], g["ZS"], 2, data["posn"], zst, txn, 0)
    out["ref_loglik"] = np.array([fb["loglik"]])
    out["ref_path"] = ref.viterbi(lp, py, g["ct"], g["ZS"], 2, data["posn"], zst, txn, 0).astype(np.int16)
    # production-style transition settings on the same data (txnExpLen=1e15, txnZstrength=1)
    fb2 = ref.fwd_back(lp, py, g["ct"], g["ZS"], 2, data["posn"], 1e15, 1e15, 0)
    out["ref_loglik_1e15"] = np.array([fb2["loglik"]])
    out["ref_path_1e15"] = ref.viterbi(lp, py, g["ct"], g["ZS"], 2, data["posn"], 1e15, 1e15, 0).astype(np.int16)
    dst = os.path.join(ROOT, "tests", "golden", "emresults_chr2.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    staging()


def staging():
    """tests/golden/allelecounts_chr2.npz: the numeric columns of inst/extdata/test_alleleCounts_chr2.txt, the file the
    fixture's `data` object was built from (vignettes/TitanCNA.Rnw:93-120) -- pins loadAlleleCounts / filterData."""
    cols = {"chr": [], "position": [], "refCount": [], "NrefCount": []}
    with open(os.path.join(REF, "inst", "extdata", "test_alleleCounts_chr2.txt")) as fh:
        next(fh)
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if len(f) < 6 or not f[0]:
                continue
            cols["chr"].append(int(f[0])); cols["position"].append(int(f[1]))
            cols["refCount"].append(int(f[3])); cols["NrefCount"].append(int(f[5]))
    dst = os.path.join(ROOT, "tests", "golden", "allelecounts_chr2.npz")
    np.savez_compressed(dst, **{k: np.asarray(v, dtype=np.int32) for k, v in cols.items()})
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
