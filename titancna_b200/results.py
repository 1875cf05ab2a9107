"""outputTitanResults / outputTitanSegments -- host-side mirror of R/utils.R:772-976,1065-1135,1342-1401.

The step immediately after the Viterbi decode: joint state ->
(genotype, cluster), call labels, per-SNP table, run-length segments of the genotype state per
chromosome with their medians.  O(N) host table work in the reference (R) and here (vectorised numpy);
tables are dicts of equal-length numpy columns.  Quirks of the reference are preserved on purpose:
the cluster relabelling `sortS$ix[Zclust]` (:891-894) and the crossed MinorCN/MajorCN assignment
(:1107-1108 with the inverted field names of getMajorMinorCN :1342-1401).
Not covered: correctResults=TRUE (removeEmptyClusters, R/utils.R:1420-1524), subclone profiles, file output.
"""
from __future__ import annotations

import numpy as np

# symmetric genotype table, state index 0..24 (R/utils.R:788-862)
_CALLS = np.array(["HOMD", "DLOH", "NLOH", "HET", "ALOH", "GAIN", "ALOH", "ASCNA", "BCNA", "ALOH", "ASCNA", "UBCNA",
                   "ALOH", "ASCNA", "UBCNA", "BCNA", "ALOH", "ASCNA", "UBCNA", "UBCNA", "ALOH", "ASCNA", "UBCNA",
                   "UBCNA", "BCNA"])
_CN = np.array([0, 1, 2, 2, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8])
# getMajorMinorCN(state)$majorCN / $minorCN exactly as named in the reference (the names are inverted there)
_MAJOR_FIELD = np.array([0, 0, 0, 1, 0, 1, 0, 1, 2, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 4])
_MINOR_FIELD = _CN - _MAJOR_FIELD


def decoupleMegaVar(G, K, useOutlierState=False):
    """R/utils.R:772-785: joint state (1-based) -> dict(G = unit state 1..K, Z = cluster 1..)."""
    if useOutlierState:
        raise NotImplementedError("useOutlierState=TRUE is not covered")
    G = np.asarray(G, dtype=np.int64)
    newG = G % K
    newG[newG == 0] = K
    return {"G": newG, "Z": -(-G // K)}


def decodeLOH(G, symmetric=True):
    """R/utils.R:788-862 (symmetric table): state 0..24 -> dict(G = call label, CN = copy number)."""
    if not symmetric:
        raise NotImplementedError("symmetric=FALSE is deprecated in the reference (R/utils.R:16-18)")
    G = np.asarray(G, dtype=np.int64)
    ok = (G >= 0) & (G < _CALLS.size)
    calls = np.where(ok, _CALLS[np.clip(G, 0, _CALLS.size - 1)], "NA")
    cn = np.where(ok, _CN[np.clip(G, 0, _CN.size - 1)], -1).astype(np.float64)
    cn[~ok] = np.nan
    return {"G": calls, "CN": cn}


def getMajorMinorCN(state, symmetric=True):
    """R/utils.R:1342-1401, field names as in the reference."""
    return {"majorCN": int(_MAJOR_FIELD[state]), "minorCN": int(_MINOR_FIELD[state])}


def outputTitanResults(data, convergeParams, optimalPath, correctResults=False, **_ignored):
    """R/utils.R:865-976 with correctResults=FALSE: the per-SNP result table."""
    if correctResults:
        raise NotImplementedError("correctResults=TRUE (removeEmptyClusters) is not covered")
    cp = convergeParams
    if cp.get("useOutlierState") is None:
        raise ValueError("convergeParams does not contain element: useOutlierState.")
    if cp.get("symmetric") is None:
        raise ValueError("convergeParams does not contain element: symmetric.")
    s_all = np.asarray(cp["s"], float).reshape(np.asarray(cp["s"]).shape[0], -1)
    K = np.asarray(cp["var"]).shape[0]
    i = s_all.shape[1] - 1
    part = decoupleMegaVar(optimalPath, K)
    G = part["G"] - 1
    ix = np.argsort(s_all[:, i], kind="stable")          # sort(..., index.return=TRUE)$ix, 0-based
    s = s_all[ix, i]
    Zclust = (ix + 1)[part["Z"] - 1].astype(np.float64)  # sortS$ix[Zclust]
    dec = decodeLOH(G, cp["symmetric"])
    calls, CN = dec["G"], dec["CN"]
    Zclust[(calls == "HET") & (CN == 2)] = np.nan        # diploid HET positions do not have clusters
    Sout = np.full(Zclust.size, np.nan)
    ok = ~np.isnan(Zclust) & (Zclust > 0)
    Sout[ok] = s[Zclust[ok].astype(np.int64) - 1]
    ref0 = np.asarray(data.get("refOriginal", data["ref"]), float)
    depth = np.asarray(data["tumDepth"], float)
    return {"Chr": np.asarray(data["chr"]), "Position": np.asarray(data["posn"]), "RefCount": ref0, "Depth": depth,
            "AllelicRatio": ref0 / depth, "LogRatio": np.log2(np.exp(np.asarray(data["logR"], float))),
            "CopyNumber": CN, "TITANstate": G, "TITANcall": calls, "ClonalCluster": Zclust,
            "CellularPrevalence": 1 - Sout}


def outputTitanSegments(results, id, convergeParams=None):
    """R/utils.R:1065-1135: run-length segments of TITANstate per chromosome."""
    state = np.asarray(results["TITANstate"])
    chr_ = np.asarray(results["Chr"])
    n = state.size
    brk = np.ones(n, dtype=bool)
    brk[1:] = (state[1:] != state[:-1]) | (chr_[1:] != chr_[:-1])
    start = np.flatnonzero(brk)
    end = np.append(start[1:], n) - 1
    ar = np.maximum(results["AllelicRatio"], 1 - np.asarray(results["AllelicRatio"]))
    lr = np.asarray(results["LogRatio"])
    med_ar = np.array([np.round(np.nanmedian(ar[a:b + 1]), 6) for a, b in zip(start, end)])
    med_lr = np.array([np.round(np.nanmedian(lr[a:b + 1]), 6) for a, b in zip(start, end)])
    st = state[start].astype(np.int64)
    return {"Sample": np.full(start.size, id), "Chromosome": chr_[start].astype(str),
            "Start_Position.bp.": np.asarray(results["Position"])[start], "End_Position.bp.": np.asarray(results["Position"])[end],
            "Length.snp.": end - start + 1, "Median_Ratio": med_ar, "Median_logR": med_lr, "TITAN_state": st,
            "TITAN_call": np.asarray(results["TITANcall"])[start], "Copy_Number": np.asarray(results["CopyNumber"])[start],
            "MinorCN": _MAJOR_FIELD[st], "MajorCN": _MINOR_FIELD[st],      # crossed on purpose (:1107-1108)
            "Clonal_Cluster": np.asarray(results["ClonalCluster"])[start],
            "Cellular_Prevalence": np.asarray(results["CellularPrevalence"])[start]}
