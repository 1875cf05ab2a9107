"""Input staging -- host-side mirror of loadAlleleCounts / filterData / removeCentromere / getPositionOverlap
(R/utils.R:151-195, 273-329, 464-478) and of the log-ratio convention of the driver
(scripts/R_scripts/titanCNA.R:218-219): the step in front of the hot path.

Produces the `data` object runEMclonalCN / titan_solver_create consume: SoA columns chr, posn, ref, refOriginal,
nonRef, tumDepth (+ logR), sorted 1..22, X, Y and by position.  Tables are dicts of equal-length numpy columns.
Not covered: extractAlleleReadCounts (BAM pile-up), correctReadDepth (GC / mappability loess on WIG files),
haplotype (10X) loaders.
"""
from __future__ import annotations

import numpy as np

from ._lib import c_double_p, check, lib

__all__ = ["loadAlleleCounts", "filterData", "removeCentromere", "getPositionOverlap", "setGenomeStyle", "logRfromLog2"]

_ORDER = {**{str(i): i for i in range(1, 23)}, "X": 23, "Y": 24, "MT": 25, "M": 25}


def setGenomeStyle(x, genomeStyle="NCBI", filterExtraChr=True):
    """R/utils.R:480-494 for the two styles the workflow uses: NCBI ("1", "X") and UCSC ("chr1", "chrX").
    Returns (names in the requested style, keep mask of the autosomes / sex / mitochondrial chromosomes)."""
    if genomeStyle not in ("NCBI", "UCSC"):
        raise ValueError("setGenomeStyle: genomeStyle must be 'NCBI' or 'UCSC'")
    bare = np.array([s[3:] if s.lower().startswith("chr") else s for s in np.asarray(x).astype(str)])
    bare = np.where(bare == "M", "MT", bare) if genomeStyle == "NCBI" else np.where(bare == "MT", "M", bare)
    keep = np.array([b in _ORDER for b in bare]) if filterExtraChr else np.ones(bare.size, bool)
    names = bare if genomeStyle == "NCBI" else np.array(["chr" + b for b in bare])
    return names, keep


def _chr_rank(names):
    bare = [s[3:] if s.lower().startswith("chr") else s for s in names]
    return np.array([_ORDER.get(b, 100) for b in bare])


def loadAlleleCounts(inCounts, symmetric=True, genomeStyle="NCBI", sep="\t", header=True):
    """R/utils.R:151-195.  `inCounts`: path of the six-column allele-count file (chr, position, ref, refCount, Nref,
    NrefCount) or a table (dict of columns in that order).  Chromosomes are sorted 1..22, X, Y, positions ascending
    within each; ref = max(refCount, NrefCount) in symmetric mode."""
    if isinstance(inCounts, str):
        import pandas as pd
        df = pd.read_csv(inCounts, sep=sep, header=0 if header else None)
        if df.shape[1] < 6 or not all(np.issubdtype(df.iloc[:, c].dtype, np.integer) for c in (1, 3, 5)):
            raise ValueError("loadAlleleCounts: Input counts file format does not \n        \t\tmatch required specifications.")
        cols = [df.iloc[:, c].to_numpy() for c in range(6)]
    elif isinstance(inCounts, dict):
        cols = [np.asarray(v) for v in list(inCounts.values())[:6]]
        if len(cols) < 6:
            raise ValueError("loadAlleleCounts: Input counts file format does not \n        \t\tmatch required specifications.")
    else:
        raise ValueError("loadAlleleCounts: Must provide a filename or data.frame \n    \t\tto inCounts")
    names, keep = setGenomeStyle(cols[0], genomeStyle)
    names = names[keep]
    posn, ref0, nref = (np.asarray(cols[c])[keep] for c in (1, 3, 5))
    order = np.lexsort((posn, _chr_rank(names)))          # orderSeqlevels(X.is.sexchrom = TRUE), then position (stable)
    names, posn = names[order], posn[order]
    ref0, nref = ref0[order].astype(np.float64), nref[order].astype(np.float64)
    return {"chr": names, "posn": posn, "ref": np.maximum(ref0, nref) if symmetric else ref0.copy(),
            "refOriginal": ref0, "nonRef": nref, "tumDepth": ref0 + nref}


def removeCentromere(data, centromere, flankLength=0):
    """R/utils.R:316-329.  `centromere`: dict with Chr, Start, End columns."""
    keep = np.ones(len(data["chr"]), bool)
    for c, a, b in zip(np.asarray(centromere["Chr"]).astype(str), centromere["Start"], centromere["End"]):
        keep &= ~((np.asarray(data["chr"]).astype(str) == c) & (data["posn"] >= a - flankLength) & (data["posn"] <= b + flankLength))
    return {k: np.asarray(v)[keep] for k, v in data.items()}


def filterData(data, chrs=None, minDepth=10, maxDepth=200, positionList=None, map=None, mapThres=0.9,
               centromeres=None, centromere_flankLength=0):
    """R/utils.R:273-310: depth window, non-missing logR, mappability, position white-list, centromeres,
    chromosome white-list (chromosomes left with a single SNP are dropped)."""
    n = len(data["chr"])
    chr_ = np.asarray(data["chr"]).astype(str)
    keep = (np.asarray(data["tumDepth"]) <= maxDepth) & (np.asarray(data["tumDepth"]) >= minDepth)
    keep &= ~np.isnan(np.asarray(data["logR"], float))
    if map is not None:
        with np.errstate(invalid="ignore"):
            keep &= np.asarray(map, float) >= mapThres            # NA map score: which() drops it
    if positionList is not None:
        white = set(zip(np.asarray(positionList["chr"]).astype(str), np.asarray(positionList["posn"]).tolist()))
        keep &= np.array([(c, p) in white for c, p in zip(chr_, np.asarray(data["posn"]).tolist())])
    out = {k: np.asarray(v)[keep] for k, v in data.items() if len(v) == n}
    if centromeres is not None:
        style = "UCSC" if (out["chr"].size and str(out["chr"][0]).lower().startswith("chr")) else "NCBI"
        cen = dict(centromeres)
        cen["Chr"] = setGenomeStyle(cen["Chr"], style, filterExtraChr=False)[0]
        out = removeCentromere(out, cen, centromere_flankLength)
    if chrs is not None:
        sel = np.isin(np.asarray(out["chr"]).astype(str), np.asarray(chrs).astype(str))
        out = {k: v[sel] for k, v in out.items()}
        names, counts = np.unique(out["chr"].astype(str), return_counts=True)
        sel = np.isin(out["chr"].astype(str), names[counts > 1])
        out = {k: v[sel] for k, v in out.items()}
    return out


def getPositionOverlap(chr, posn, dataVal):
    """R/utils.R:464-478: the value (4th column of `dataVal`: chr, start, end, logR) of the interval that contains
    each position, NA where none does; with overlapping intervals the last one wins, as the reference's
    `hitVal[from(hits)] <- logR[to(hits)]` does.  Per chromosome the interval search is the C-ABI routine
    titan_get_position_overlap (src/getPositionOverlapC.c:19-50: first containing interval, integer compare)
    when the intervals of a chromosome do not overlap, a sorted sweep otherwise."""
    cols = list(dataVal.values())
    dchr, start, end, val = (np.asarray(cols[0]).astype(str), np.asarray(cols[1], float), np.asarray(cols[2], float),
                             np.asarray(cols[3], float))
    chr_ = np.asarray(chr).astype(str)
    posn = np.asarray(posn, float)
    out = np.full(posn.size, np.nan)
    for c in np.unique(chr_):
        q = np.flatnonzero(chr_ == c)
        d = np.flatnonzero(dchr == c)
        if q.size == 0 or d.size == 0:
            continue
        order = np.argsort(start[d], kind="stable")
        s_, e_ = start[d][order], end[d][order]
        if np.all(s_[1:] > e_[:-1]):                              # disjoint: binary search
            k = np.searchsorted(s_, posn[q], side="right") - 1
            hit = (k >= 0) & (posn[q] <= e_[np.maximum(k, 0)])
            out[q[hit]] = val[d][order][k[hit]]
        else:                                                     # overlapping intervals: last containing one wins
            for j in range(d.size):
                inside = (posn[q] >= start[d[j]]) & (posn[q] <= end[d[j]])
                out[q[inside]] = val[d[j]]
    return out


def position_overlap_index(posn, start, stop):
    """The reference's own C routine behind the C ABI (src/getPositionOverlapC.c:19-50): 1-based index of the first
    interval [start, stop] that contains each position (0 if none)."""
    posn, start, stop = (np.ascontiguousarray(v, dtype=np.float64) for v in (posn, start, stop))
    out = np.zeros(posn.size, dtype=np.float64)
    check(lib().titan_get_position_overlap(posn.ctypes.data_as(c_double_p), posn.size, start.ctypes.data_as(c_double_p),
                                           stop.ctypes.data_as(c_double_p), start.size, out.ctypes.data_as(c_double_p)))
    return out


def logRfromLog2(log2ratio):
    """scripts/R_scripts/titanCNA.R:219 `data$logR <- log(2^logR)`: corrected log2 ratios -> natural log."""
    return np.log(np.power(2.0, np.asarray(log2ratio, float)))
