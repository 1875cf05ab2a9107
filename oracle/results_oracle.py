"""Loop-by-loop restatement of the reference's decode + segmenter for the parity tests.

TEST INFRASTRUCTURE ONLY.  Follows R/utils.R:772-785 (decoupleMegaVar), :788-862 (decodeLOH, symmetric
table), :865-913 (outputTitanResults core columns, correctResults=FALSE), :1065-1135
(outputTitanSegments), :1175-1340 (correctIntegerCN, logRbasedCN, allelicRatioBasedCN), :1342-1401
(getMajorMinorCN) and :1420-1498 (removeEmptyClusters, without its
recomputeLogLik tail: the tests re-enter the oracle's own EM for that) statement by statement, including
the reference's quirks (cluster relabelling through sortS$ix; crossed Minor/Major assignment; `s` re-sorted
but piZ/rhoZ/mu subset in stored order; in-place relabelling loop).
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


def remove_empty_clusters(cp, rows, proportion_threshold=0.001, proportion_threshold_clonal=0.05):
    """R/utils.R:1420-1498 on a list of row dicts.  Returns (new cp dict of plain lists / arrays, new rows)."""
    rows = [dict(r) for r in rows]
    s = [list(map(float, r)) for r in np.asarray(cp["s"], float).reshape(np.asarray(cp["s"]).shape[0], -1)]
    Z = len(s)
    n = [float(v) for v in np.atleast_1d(cp["n"])]
    clust = {}
    for cl in range(1, Z + 1):
        cnt = 0
        for r in rows:
            if r["ClonalCluster"] == cl:
                cnt += 1
        frac = cnt / len(rows)
        clust[cl] = None if (frac < proportion_threshold or (frac < proportion_threshold_clonal and cl == 1)) else cl
    k = len(s[0]) - 1
    s = [s[z] for z in sorted(range(Z), key=lambda z: (s[z][k], z))]
    out = {"piZ": None, "rhoZ": None, "muC": None, "muR": None}
    sP = {key: list(np.atleast_1d(v)) for key, v in (cp.get("cellPrevParams") or {}).items()}
    kept = [cl for cl in range(1, Z + 1) if clust[cl] is not None]
    if kept:
        purity = (1 - s[kept[0] - 1][k]) * (1 - n[k])
        n[k] = 1 - purity
        s = [s[cl - 1] for cl in kept]
        first = s[0][k]
        for row in s:
            row[k] = 1 - (1 - row[k]) / (1 - first)
        idx = [cl - 1 for cl in kept]
        for key in ("piZ", "rhoZ"):
            if cp.get(key) is not None:
                out[key] = np.asarray(cp[key])[idx]
        for key in ("muC", "muR"):
            if cp.get(key) is not None:
                out[key] = np.asarray(cp[key])[:, idx]
        sP = {key: [v[j] if j < len(v) else float("nan") for j in idx] for key, v in sP.items()}
        for new, cl in enumerate(kept, start=1):
            clust[cl] = new
        for cl in range(1, Z + 1):
            target = None
            if clust[cl] is None:
                right = [c for c in range(1, Z + 1) if clust[c] is not None and str(c) > str(cl)]
                left = [c for c in range(1, Z + 1) if clust[c] is not None and str(c) < str(cl)]
                if right:
                    target = clust[right[0]]
                elif left:
                    target = clust[left[-1]]
            else:
                target = clust[cl]
            if target is None:
                continue
            hit = [r for r in rows if r["ClonalCluster"] == cl]
            for r in hit:
                r["CellularPrevalence"] = 1 - s[target - 1][k]
            for r in hit:
                r["ClonalCluster"] = float(target)
    else:
        n[k] = 1 - s[0][k]
        s = [s[0]]
        s[0][k] = 0.0
        sP = {key: [v[0] if len(v) else float("nan")] for key, v in sP.items()}
        for key in out:
            out[key] = cp.get(key)
        for r in rows:
            if r["TITANcall"] != "HET":
                r["CellularPrevalence"] = 1
                r["ClonalCluster"] = 1.0
    new_cp = dict(cp)
    new_cp.update({"s": np.array(s), "n": np.array(n), "cellPrevParams": {key: np.array(v) for key, v in sP.items()}})
    for key, v in out.items():
        if v is not None:
            new_cp[key] = v
    return new_cp, rows


def _isna(v):
    return v is None or v != v


def logr_based_cn(x, purity, ploidy, cell_prev, cn=2):
    if _isna(x):
        return float("nan")
    cp = 1.0 if _isna(cell_prev) else cell_prev
    ct = (2.0 ** x) * (cn * (1 - purity) + purity * ploidy * (cn / 2)) - cn * (1 - purity) - cn * purity * (1 - cp)
    ct = ct / (purity * cp)
    return max(ct, 1 / 2 ** 6)


def allelic_ratio_based_cn(x, ct, purity, cell_prev, rn=0.5, cn=2):
    if _isna(x) or _isna(ct):
        return float("nan")
    cp = 1.0 if _isna(cell_prev) else cell_prev
    total = (1 - purity) * cn + purity * (1 - cp) * cn + purity * cp * ct
    rt = (x * total - ((1 - purity) * rn * cn + purity * (1 - cp) * rn * cn)) / (purity * cp * ct)
    return max(min(rt, 1.0), 0.0)


def _rint(v):
    return float(round(v)) if not _isna(v) else float("nan")      # Python round(): half to even, like R


def correct_integer_cn(cn_rows, seg_rows, purity, ploidy, max_auto=None, max_x=None, correct_homd=True,
                       min_purity=0.2, gender="male", chrs=None):
    """R/utils.R:1175-1312 on lists of row dicts (per-SNP table rows, segment rows)."""
    chrs = [str(c) for c in (list(range(1, 23)) + ["X"] if chrs is None else chrs)]
    cn_rows, seg_rows = [dict(r) for r in cn_rows], [dict(r) for r in seg_rows]
    autos = [c for c in chrs if "X" not in c and "Y" not in c]
    xs = [c for c in chrs if "X" in c]
    if max_auto is None:
        max_auto = max(r["Copy_Number"] for r in seg_rows if str(r["Chromosome"]) in autos and not _isna(r["Copy_Number"]))
    if max_x is None and gender == "female" and xs:
        max_x = max(r["Copy_Number"] for r in seg_rows if str(r["Chromosome"]) in xs and not _isna(r["Copy_Number"]))
    names = ["HOMD", "HETD", "NEUT", "GAIN", "AMP", "HLAMP"] + ["HLAMP"] * 1000
    names_x = ["HETD", "NEUT", "GAIN", "AMP", "HLAMP"] + ["HLAMP"] * 1000
    for rows, ckey, lkey, rkey, pkey, nkey in ((seg_rows, "Chromosome", "Median_logR", "Median_Ratio", "Cellular_Prevalence", "Copy_Number"),
                                               (cn_rows, "Chr", "LogRatio", "AllelicRatio", "CellularPrevalence", "CopyNumber")):
        for r in rows:
            c = str(r[ckey])
            r["logR_Copy_Number"] = r["Corrected_logR"] = r["Corrected_Ratio"] = float("nan")
            if c in chrs:
                r["logR_Copy_Number"] = logr_based_cn(r[lkey], purity, ploidy, r[pkey], 2)
                r["Corrected_logR"] = math.log2(r["logR_Copy_Number"] / ploidy)
                r["Corrected_Ratio"] = allelic_ratio_based_cn(r[rkey], r["logR_Copy_Number"], purity, r[pkey])
            if gender == "male" and c in xs:
                r["logR_Copy_Number"] = logr_based_cn(r[lkey], purity, ploidy, r[pkey], 1)
                r["Corrected_logR"] = math.log2(r["logR_Copy_Number"] / (ploidy / 2))
                r["Corrected_Ratio"] = float("nan")
            r["Corrected_Copy_Number"] = float(r[nkey]) if not _isna(r[nkey]) else float("nan")
    for r in seg_rows:
        r["Corrected_MajorCN"], r["Corrected_MinorCN"] = float(r["MajorCN"]), float(r["MinorCN"])
    if purity >= min_purity:
        touched = set()
        for rows, ckey, nkey, is_seg in ((seg_rows, "Chromosome", "Copy_Number", True), (cn_rows, "Chr", "CopyNumber", False)):
            for idx, r in enumerate(rows):
                c = str(r[ckey])
                hit = False
                if c in chrs and not _isna(r[nkey]) and r[nkey] >= max_auto:
                    hit = True
                if correct_homd and c in chrs and not _isna(r[nkey]) and r[nkey] == 0:
                    hit = True
                if not is_seg and c in chrs and _isna(r[nkey]):
                    hit = True
                if gender == "male" and c in xs:
                    hit = True
                if hit:
                    r["Corrected_Copy_Number"] = _rint(r["logR_Copy_Number"])
                    if is_seg:
                        touched.add(idx)
        for idx, r in enumerate(seg_rows):
            rl, cc = _rint(r["logR_Copy_Number"]), r["Corrected_Copy_Number"]
            if _isna(rl) or _isna(cc):
                continue
            if ((rl < ploidy and cc > ploidy) or (rl > ploidy and cc < ploidy)) and abs(rl - cc) > 2:
                r["Corrected_Copy_Number"] = rl
                touched.add(idx)
                for q in cn_rows:
                    if str(q["Chr"]) == str(r["Chromosome"]) and r["Start_Position.bp."] <= q["Position"] <= r["End_Position.bp."]:
                        q["Corrected_Copy_Number"] = rl
        for idx in touched:
            r = seg_rows[idx]
            if _isna(r["Corrected_Ratio"]) or _isna(r["Corrected_Copy_Number"]):
                r["Corrected_MajorCN"] = r["Corrected_MinorCN"] = float("nan")
            else:
                r["Corrected_MajorCN"] = _rint(r["Corrected_Ratio"] * r["Corrected_Copy_Number"])
                r["Corrected_MinorCN"] = _rint((1 - r["Corrected_Ratio"]) * r["Corrected_Copy_Number"])
    for rows, ckey, nkey in ((seg_rows, "Chromosome", "Copy_Number"), (cn_rows, "Chr", "CopyNumber")):
        for r in rows:
            cc = r["Corrected_Copy_Number"]
            r["Corrected_Call"] = None if _isna(cc) else names[int(cc)]
            if gender == "male" and xs:
                if str(r[ckey]) in xs and not _isna(cc):
                    r["Corrected_Call"] = names_x[int(cc)]
            elif max_x is not None and str(r[ckey]) in xs and not _isna(r[nkey]) and r[nkey] >= max_x:
                r["Corrected_Call"] = "HLAMP"
    return cn_rows, seg_rows


def subclone_profiles(rows):
    """R/utils.R:1526-1577 on row dicts -> list of dicts with the Subclone1.* / Subclone2.* fields."""
    clust = [0 if _isna(r["ClonalCluster"]) else r["ClonalCluster"] for r in rows]
    ncl = int(max(clust)) if clust else 0
    pairs = []
    for r in rows:
        key = (r["ClonalCluster"], r["CellularPrevalence"])
        if not any((a == key[0] or (_isna(a) and _isna(key[0]))) and (b == key[1] or (_isna(b) and _isna(key[1]))) for a, b in pairs):
            pairs.append(key)
    prev = {cl: [p for c, p in pairs if c == cl] for cl in (1, 2)}
    out = []
    for r in rows:
        o = {}
        if ncl == 0:
            o = {"Subclone1.CopyNumber": r["CopyNumber"], "Subclone1.TITANcall": r["TITANcall"], "Subclone1.Prevalence": "NA"}
        elif ncl == 1:
            o = {"Subclone1.CopyNumber": r["CopyNumber"], "Subclone1.TITANcall": r["TITANcall"], "Subclone1.Prevalence": prev[1][0]}
        elif ncl == 2:
            two = r["ClonalCluster"] == 2
            o = {"Subclone2.CopyNumber": r["CopyNumber"], "Subclone2.TITANcall": r["TITANcall"], "Subclone2.Prevalence": prev[2][0],
                 "Subclone1.CopyNumber": 2 if two else r["CopyNumber"], "Subclone1.TITANcall": "HET" if two else r["TITANcall"],
                 "Subclone1.Prevalence": prev[1][0] - prev[2][0]}
        out.append(o)
    return out
