"""Loop-by-loop restatement of the reference's decode + segmenter for the parity tests.

TEST INFRASTRUCTURE ONLY.  Follows R/utils.R:772-785 (decoupleMegaVar), :788-862 (decodeLOH, symmetric
table), :865-913 (outputTitanResults core columns, correctResults=FALSE), :1065-1135
(outputTitanSegments) and :1342-1401 (getMajorMinorCN) statement by statement, including the
reference's quirks (cluster relabelling through sortS$ix; crossed Minor/Major assignment).
Parity unpinned: the reference holds no golden table for these functions.
"""
import math

import numpy as np


def decouple(G, K):
    newG, newZ = [], []
    for g in G:
        r = g % K
        newG.append(K if r == 0 else r)
        newZ.append(int(math.ceil(g / K)))
    return newG, newZ


def decode_loh(g):
    if g == 0:
        return "HOMD", 0
    if g == 1:
        call = "DLOH"
    elif g == 2:
        call = "NLOH"
    elif g in (4, 6, 9, 12, 16, 20):
        call = "ALOH"
    elif g == 3:
        call = "HET"
    elif g == 5:
        call = "GAIN"
    elif g in (7, 10, 13, 17, 21):
        call = "ASCNA"
    elif g in (8, 15, 24):
        call = "BCNA"
    elif g in (11, 14, 18, 19, 22, 23):
        call = "UBCNA"
    else:
        return "NA", float("nan")
    if g == 1:
        cn = 1
    elif 2 <= g <= 3:
        cn = 2
    elif 4 <= g <= 5:
        cn = 3
    elif 6 <= g <= 8:
        cn = 4
    elif 9 <= g <= 11:
        cn = 5
    elif 12 <= g <= 15:
        cn = 6
    elif 16 <= g <= 19:
        cn = 7
    else:
        cn = 8
    return call, cn


_MM = {0: (0, 0), 1: (0, 1), 2: (0, 2), 3: (1, 1), 4: (0, 3), 5: (1, 2), 6: (0, 4), 7: (1, 3), 8: (2, 2), 9: (0, 5),
       10: (1, 4), 11: (2, 3), 12: (0, 6), 13: (1, 5), 14: (2, 4), 15: (3, 3), 16: (0, 7), 17: (1, 6), 18: (2, 5),
       19: (3, 4), 20: (0, 8), 21: (1, 7), 22: (2, 6), 23: (3, 5), 24: (4, 4)}     # state -> (majorCN, minorCN) fields


def output_results(data, cp, path):
    s_all = np.asarray(cp["s"], float)
    K = np.asarray(cp["var"]).shape[0]
    i = s_all.shape[1] - 1
    newG, newZ = decouple([int(p) for p in path], K)
    order = sorted(range(s_all.shape[0]), key=lambda z: (s_all[z, i], z))      # sort(..., index.return = TRUE)
    s_sorted = [s_all[z, i] for z in order]
    rows = []
    ref0 = data.get("refOriginal", data["ref"])
    for t in range(len(path)):
        G = newG[t] - 1
        zc = order[newZ[t] - 1] + 1                                             # sortS$ix[Zclust]
        call, cn = decode_loh(G)
        if call == "HET" and cn == 2:
            zc = None
        sout = s_sorted[zc - 1] if zc is not None and zc > 0 else float("nan")
        rows.append({"Chr": data["chr"][t], "Position": data["posn"][t], "AllelicRatio": ref0[t] / data["tumDepth"][t],
                     "LogRatio": math.log2(math.exp(data["logR"][t])), "CopyNumber": cn, "TITANstate": G, "TITANcall": call,
                     "ClonalCluster": float("nan") if zc is None else float(zc), "CellularPrevalence": 1 - sout})
    return rows


def output_segments(rows, sample):
    segs = []
    start = 0
    n = len(rows)
    while start < n:
        end = start
        while end + 1 < n and rows[end + 1]["Chr"] == rows[start]["Chr"] and rows[end + 1]["TITANstate"] == rows[start]["TITANstate"]:
            end += 1
        seg = rows[start:end + 1]
        ar = [max(r["AllelicRatio"], 1 - r["AllelicRatio"]) for r in seg]
        major_field, minor_field = _MM[seg[0]["TITANstate"]]
        segs.append({"Sample": sample, "Chromosome": str(seg[0]["Chr"]), "Start_Position.bp.": seg[0]["Position"],
                     "End_Position.bp.": seg[-1]["Position"], "Length.snp.": len(seg),
                     "Median_Ratio": round(float(np.median(ar)), 6), "Median_logR": round(float(np.median([r["LogRatio"] for r in seg])), 6),
                     "TITAN_state": seg[0]["TITANstate"], "TITAN_call": seg[0]["TITANcall"], "Copy_Number": seg[0]["CopyNumber"],
                     "MinorCN": major_field, "MajorCN": minor_field,           # :1107-1108 cross-assignment
                     "Clonal_Cluster": seg[0]["ClonalCluster"], "Cellular_Prevalence": seg[0]["CellularPrevalence"]})
        start = end + 1
    return segs
