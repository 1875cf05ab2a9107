"""outputTitanResults / outputTitanSegments -- host-side mirror of R/utils.R:772-976,1065-1135,1342-1401.

The step immediately after the Viterbi decode: joint state ->
(genotype, cluster), call labels, per-SNP table, run-length segments of the genotype state per
chromosome with their medians.  O(N) host table work in the reference (R) and here (vectorised numpy);
tables are dicts of equal-length numpy columns.  Quirks of the reference are preserved on purpose:
the cluster relabelling `sortS$ix[Zclust]` (:891-894) and the crossed MinorCN/MajorCN assignment
(:1107-1108 with the inverted field names of getMajorMinorCN :1342-1401).
correctResults=TRUE (removeEmptyClusters, R/utils.R:1420-1524) is covered, including the hidden re-entry into the
E-step (recomputeLogLik=TRUE -> runEMclonalCN(maxiter=1, fixed normal, no s / ploidy update)).
correctIntegerCN (:1175-1340), the step between the segments and the parameter file in the production driver
(scripts/R_scripts/titanCNA.R:278-281), is covered as well.
Subclone profiles (:1526-1577) for one or two clusters are covered.
Not covered: posterior-probability columns, file output.
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


def outputTitanResults(data, convergeParams, optimalPath, filename=None, is_haplotypeData=False, posteriorProbs=False,
                       subcloneProfiles=True, correctResults=True, proportionThreshold=0.05,
                       proportionThresholdClonal=0.05, recomputeLogLik=True, rerunViterbi=False, verbose=True):
    """R/utils.R:865-976.  Returns dict(results = per-SNP table, corrResults = the table after removeEmptyClusters
    (None with correctResults=False), convergeParams = the -- possibly corrected -- parameter list), as the
    reference's list(results, corrResults, convergeParams).  Tables are dicts of equal-length numpy columns."""
    if filename is not None or posteriorProbs or is_haplotypeData:
        raise NotImplementedError("file output, posterior columns and haplotype columns are not covered")
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
    outmat = {"Chr": np.asarray(data["chr"]), "Position": np.asarray(data["posn"]), "RefCount": ref0, "Depth": depth,
              "AllelicRatio": ref0 / depth, "LogRatio": np.log2(np.exp(np.asarray(data["logR"], float))),
              "CopyNumber": CN, "TITANstate": G, "TITANcall": calls, "ClonalCluster": Zclust,
              "CellularPrevalence": 1 - Sout}
    corrmat = None
    if correctResults:
        corr = removeEmptyClusters(data, cp, outmat, proportionThreshold=proportionThreshold,
                                   proportionThresholdClonal=proportionThresholdClonal,
                                   recomputeLogLik=recomputeLogLik, verbose=verbose)
        corrmat, cp = corr["results"], corr["convergeParams"]
    if subcloneProfiles and s_all.shape[0] <= 2:           # numClust is taken BEFORE the correction (:875,:944)
        target = corrmat if correctResults else outmat     # the columns land on the table that is written out
        target.update(getSubcloneProfiles(target))
    return {"results": outmat, "corrResults": corrmat, "convergeParams": cp}


def getSubcloneProfiles(titanResults):
    """R/utils.R:1526-1577: per-subclone copy number / call / prevalence columns for one or two clonal clusters
    (Subclone1.*, Subclone2.*).  With two clusters, positions of cluster 2 are diploid HET in subclone 1 and the
    prevalence of subclone 1 is the difference of the two cluster prevalences."""
    cc = np.asarray(titanResults["ClonalCluster"], float)
    prev = np.asarray(titanResults["CellularPrevalence"], float)
    cn = np.asarray(titanResults["CopyNumber"], float)
    calls = np.asarray(titanResults["TITANcall"], dtype=object)
    with np.errstate(invalid="ignore"):
        ncl = int(np.nanmax(np.where(np.isnan(cc), 0, cc))) if cc.size else 0

    def prevalence_of(cl):                                  # unique(cbind(Cluster, Prevalence)) rows of that cluster
        v = np.unique(prev[cc == cl])
        return v

    n = cc.size
    out = {}
    if ncl == 0:
        out["Subclone1.CopyNumber"], out["Subclone1.TITANcall"] = cn.copy(), calls.copy()
        out["Subclone1.Prevalence"] = np.full(n, "NA", dtype=object)
    elif ncl == 1:
        p1 = prevalence_of(1)
        out["Subclone1.CopyNumber"], out["Subclone1.TITANcall"] = cn.copy(), calls.copy()
        out["Subclone1.Prevalence"] = np.resize(p1, n).astype(float) if p1.size else np.full(n, np.nan)
    elif ncl == 2:
        p2, p1 = prevalence_of(2), prevalence_of(1)
        d = p1 - p2 if p1.size == p2.size else np.resize(p1, max(p1.size, p2.size)) - np.resize(p2, max(p1.size, p2.size))
        out["Subclone2.CopyNumber"], out["Subclone2.TITANcall"] = cn.copy(), calls.copy()
        out["Subclone2.Prevalence"] = np.resize(p2, n).astype(float)
        c1, t1 = cn.copy(), calls.copy()
        c1[cc == 2] = 2
        t1[cc == 2] = "HET"
        out["Subclone1.CopyNumber"], out["Subclone1.TITANcall"] = c1, t1
        out["Subclone1.Prevalence"] = np.resize(d, n).astype(float)
    else:
        raise ValueError("getSubcloneProfiles: more than 2 clonal clusters (outputTitanResults only calls it for <= 2)")
    return out


def removeEmptyClusters(data, convergeParams, results, proportionThreshold=0.001, proportionThresholdClonal=0.05,
                        recomputeLogLik=True, verbose=True, solver=None):
    """R/utils.R:1420-1524.  Drops clonal clusters that hold too small a share of the SNPs, re-expresses normal
    contamination and cellular prevalences relative to the first kept cluster, relabels the result table and --
    recomputeLogLik -- re-enters the E-step once with the corrected parameters.  Inputs are not modified.
    Reference quirks kept: `s` is re-sorted by its last column while piZ / rhoZ / muR / muC are subset in their
    stored order (:1434-1448); the relabelling loop edits the table in place, cluster by cluster (:1454-1470);
    with no populated cluster the first row of the sorted `s` is kept whatever the table holds (:1474-1477)."""
    cp = dict(convergeParams)
    res = {k: np.array(v, copy=True) for k, v in results.items()}
    s = np.array(np.asarray(cp["s"], float).reshape(np.asarray(cp["s"]).shape[0], -1), copy=True)
    Z = s.shape[0]
    n = np.array(np.atleast_1d(cp["n"]), dtype=float, copy=True)
    cc = res["ClonalCluster"] = np.asarray(res["ClonalCluster"], float)
    prev = res["CellularPrevalence"] = np.asarray(res["CellularPrevalence"], float)
    nrow = cc.size
    clust = list(range(1, Z + 1))                          # NA -> None
    for cl in range(1, Z + 1):
        frac = np.count_nonzero(cc == cl) / nrow
        if frac < proportionThreshold or (frac < proportionThresholdClonal and cl == 1):
            clust[cl - 1] = None
    k = s.shape[1] - 1
    s = s[np.argsort(s[:, k], kind="stable")]              # order(..., decreasing = FALSE)
    sP = {key: np.array(np.atleast_1d(v), copy=True) for key, v in (cp.get("cellPrevParams") or {}).items()}

    def sub(v, idx):                                       # `[`(v, idx): NA beyond the end
        v = np.asarray(v)
        out = np.full(len(idx), np.nan) if v.dtype.kind == "f" else np.zeros(len(idx), dtype=v.dtype)
        for q, j in enumerate(idx):
            if j < v.size:
                out[q] = v[j]
        return out

    if any(c is not None for c in clust):
        keep = [j for j, c in enumerate(clust) if c is not None]
        purity = (1 - s[keep[0], k]) * (1 - n[k])
        n[k] = 1 - purity
        s = s[keep]
        s[:, k] = 1 - (1 - s[:, k]) / (1 - s[0, k])
        for key, axis in (("piZ", 0), ("rhoZ", 0), ("muC", 1), ("muR", 1)):
            if cp.get(key) is not None:
                cp[key] = np.take(np.asarray(cp[key]), keep, axis=axis)
        sP = {key: sub(v, keep) for key, v in sP.items()}
        lab = 0
        for j in range(Z):
            if clust[j] is not None:
                lab += 1
                clust[j] = lab
        names = [str(j + 1) for j in range(Z)]
        for cl in range(1, Z + 1):                         # in place, in this order
            rows = cc == cl
            if clust[cl - 1] is None:
                right = [j for j in range(Z) if clust[j] is not None and names[j] > str(cl)]   # character comparison (:1458)
                left = [j for j in range(Z) if clust[j] is not None and names[j] < str(cl)]
                if right:
                    prev[rows] = 1 - s[clust[right[0]] - 1, k]
                    cc[rows] = clust[right[0]]
                elif left:
                    prev[rows] = 1 - s[clust[left[-1]] - 1, k]
                    cc[rows] = clust[left[-1]]
            else:
                prev[rows] = 1 - s[clust[cl - 1] - 1, k]
                cc[rows] = clust[cl - 1]
    else:
        n[k] = 1 - s[0, k]
        s = s[:1]
        s[:, k] = 0.0
        sP = {key: sub(v, [0]) for key, v in sP.items()}
        not_het = np.asarray(res["TITANcall"]) != "HET"
        prev[not_het] = 1
        cc[not_het] = 1
    cp["s"], cp["n"], cp["cellPrevParams"] = s, n, sP

    if recomputeLogLik:
        if cp.get("txn_exp_len") is None or cp.get("txn_z_strength") is None:
            raise ValueError("removeEmptyClusters(recomputeLogLik=TRUE) reads convergeParams$txn_exp_len and $txn_z_strength "
                             "(R/utils.R:1514-1515), which runEMclonalCN does not set: the reference passes NULL and its C "
                             "code then reads a zero-length vector (undefined behaviour).  Set both fields, or call with "
                             "recomputeLogLik=FALSE as scripts/R_scripts/titanCNA.R:266 does.")
        from .hmm import runEMclonalCN
        it = n.size - 1
        newZ = s.shape[0]
        new = {"genotypeParams": dict(cp["genotypeParams"]), "ploidyParams": dict(cp["ploidyParams"]),
               "normalParams": dict(cp["normalParams"]), "cellPrevParams": dict(sP)}
        new["genotypeParams"]["var_0"] = np.asarray(cp["var"], float)[:, it]
        if cp.get("varR") is not None:
            new["genotypeParams"]["varR_0"] = np.asarray(cp["varR"], float)[:, it]
        new["genotypeParams"]["piG_0"] = np.asarray(cp["piG"], float)[:, it]
        new["normalParams"]["n_0"] = float(n[it])
        new["ploidyParams"]["phi_0"] = float(np.atleast_1d(cp["phi"])[it])
        new["cellPrevParams"]["s_0"] = s[:, it].copy()
        new["cellPrevParams"]["piZ_0"] = np.asarray(cp["piZ"], float)[:newZ, it]
        p = runEMclonalCN(data, new, maxiter=1, txnExpLen=float(cp["txn_exp_len"]), txnZstrength=float(cp["txn_z_strength"]),
                          useOutlierState=False, normalEstimateMethod="fixed", estimateS=False, estimatePloidy=False,
                          verbose=verbose, solver=solver if getattr(solver, "Z", None) == newZ else None)
        ll = np.array(np.atleast_1d(cp["loglik"]), dtype=float, copy=True)
        ll[it] = p["loglik"][-1]
        cp["loglik"] = ll
        for key in ("muC", "muR"):
            m = np.array(cp[key], dtype=float, copy=True)
            m[:, :, it] = p[key][:, :, 1] if p[key].shape[1] == m.shape[1] else p[key][:, :1, 1]   # R recycles U values over the clusters
            cp[key] = m
        cp["rhoZ"], cp["rhoG"] = p["rhoZ"], p["rhoG"]
    return {"convergeParams": cp, "results": res}


def _segment_medians(v, seg_id, nseg):
    """round(median(v[segment], na.rm = TRUE), 6) for every segment at once (one sort instead of one per segment)."""
    v = np.asarray(v, float)
    ok = ~np.isnan(v)
    order = np.lexsort((v[ok], seg_id[ok]))
    sv, sid = v[ok][order], seg_id[ok][order]
    cnt = np.bincount(sid, minlength=nseg)
    first = np.cumsum(cnt) - cnt
    out = np.full(nseg, np.nan)
    has = cnt > 0
    lo = first[has] + (cnt[has] - 1) // 2
    hi = first[has] + cnt[has] // 2
    out[has] = (sv[lo] + sv[hi]) / 2
    return np.round(out, 6)


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
    seg_id = np.cumsum(brk) - 1
    med_ar, med_lr = _segment_medians(ar, seg_id, start.size), _segment_medians(lr, seg_id, start.size)
    st = state[start].astype(np.int64)
    return {"Sample": np.full(start.size, id), "Chromosome": chr_[start].astype(str),
            "Start_Position.bp.": np.asarray(results["Position"])[start], "End_Position.bp.": np.asarray(results["Position"])[end],
            "Length.snp.": end - start + 1, "Median_Ratio": med_ar, "Median_logR": med_lr, "TITAN_state": st,
            "TITAN_call": np.asarray(results["TITANcall"])[start], "Copy_Number": np.asarray(results["CopyNumber"])[start],
            "MinorCN": _MAJOR_FIELD[st], "MajorCN": _MINOR_FIELD[st],      # crossed on purpose (:1107-1108)
            "Clonal_Cluster": np.asarray(results["ClonalCluster"])[start],
            "Cellular_Prevalence": np.asarray(results["CellularPrevalence"])[start]}


_CN_NAMES = ["HOMD", "HETD", "NEUT", "GAIN", "AMP", "HLAMP"]
_CN_NAMES_X = ["HETD", "NEUT", "GAIN", "AMP", "HLAMP"]


def logRbasedCN(x, purity, ploidyT, cellPrev=None, cn=2):
    """R/utils.R:1314-1327: total copies implied by a log2 ratio at the given purity / ploidy / cellular prevalence."""
    x = np.asarray(x, float)
    cp = np.ones_like(x) if cellPrev is None else np.where(np.isnan(np.asarray(cellPrev, float)), 1.0, np.asarray(cellPrev, float))
    ct = (np.power(2.0, x) * (cn * (1 - purity) + purity * ploidyT * (cn / 2)) - cn * (1 - purity) - cn * purity * (1 - cp))
    ct = ct / (purity * cp)
    return np.where(np.isnan(ct), np.nan, np.maximum(ct, 1 / 2 ** 6))


def allelicRatioBasedCN(x, ct, purity, cellPrev=None, rn=0.5, cn=2):
    """R/utils.R:1329-1340: tumour allelic fraction implied by an observed ratio."""
    x, ct = np.asarray(x, float), np.asarray(ct, float)
    cp = np.ones_like(x) if cellPrev is None else np.where(np.isnan(np.asarray(cellPrev, float)), 1.0, np.asarray(cellPrev, float))
    total = (1 - purity) * cn + purity * (1 - cp) * cn + purity * cp * ct
    with np.errstate(invalid="ignore", divide="ignore"):
        rt = (x * total - ((1 - purity) * rn * cn + purity * (1 - cp) * rn * cn)) / (purity * cp * ct)
    return np.where(np.isnan(rt), np.nan, np.clip(rt, 0.0, 1.0))


def correctIntegerCN(cn, segs, purity, ploidy, maxCNtoCorrect_autosomes=None, maxCNtoCorrect_X=None, correctHOMD=True,
                     minPurityToCorrect=0.2, gender="male", chrs=None):
    """R/utils.R:1175-1312: copy numbers re-derived from the purity / ploidy corrected log ratio where the HMM's
    state space cannot represent them (>= the largest modelled copy number, homozygous deletions, chrX of males,
    calls that contradict the log ratio).  Returns dict(cn = per-SNP table, segs = segment table) with the
    Corrected_* columns added; the inputs are not modified.  (scripts/R_scripts/titanCNA.R:278-281.)"""
    chrs = [str(c) for c in (list(range(1, 23)) + ["X"] if chrs is None else chrs)]
    cn = {k: np.array(v, copy=True) for k, v in cn.items()}
    segs = {k: np.array(v, copy=True) for k, v in segs.items()}
    seg_ratio = "Median_HaplotypeRatio" if "Median_HaplotypeRatio" in segs else "Median_Ratio"
    cn_ratio = "HaplotypeRatio" if "HaplotypeRatio" in cn else "AllelicRatio"
    autosomes = [c for c in chrs if "X" not in c and "Y" not in c]
    chr_x = [c for c in chrs if "X" in c]
    s_chr, c_chr = np.asarray(segs["Chromosome"]).astype(str), np.asarray(cn["Chr"]).astype(str)
    s_in, c_in = np.isin(s_chr, chrs), np.isin(c_chr, chrs)
    s_x, c_x = np.isin(s_chr, chr_x), np.isin(c_chr, chr_x)
    s_cn = np.asarray(segs["Copy_Number"], float)
    c_cn = np.asarray(cn["CopyNumber"], float)
    if maxCNtoCorrect_autosomes is None:
        maxCNtoCorrect_autosomes = np.nanmax(s_cn[np.isin(s_chr, autosomes)])
    if maxCNtoCorrect_X is None and gender == "female" and chr_x:
        maxCNtoCorrect_X = np.nanmax(s_cn[s_x])

    def corrected(lr, ratio, prev, inside, on_x):
        lr_cn = np.where(inside, logRbasedCN(lr, purity, ploidy, prev, cn=2), np.nan)
        with np.errstate(invalid="ignore", divide="ignore"):
            c_lr = np.log2(lr_cn / ploidy)
        c_ratio = np.where(inside, allelicRatioBasedCN(ratio, lr_cn, purity, prev, rn=0.5, cn=2), np.nan)
        if gender == "male" and chr_x:
            lr_cn = np.where(on_x, logRbasedCN(lr, purity, ploidy, prev, cn=1), lr_cn)
            with np.errstate(invalid="ignore", divide="ignore"):
                c_lr = np.where(on_x, np.log2(lr_cn / (ploidy / 2)), c_lr)
            c_ratio = np.where(on_x, np.nan, c_ratio)
        return lr_cn, c_lr, c_ratio

    segs["logR_Copy_Number"], segs["Corrected_logR"], segs["Corrected_Ratio"] = corrected(
        segs["Median_logR"], segs[seg_ratio], segs["Cellular_Prevalence"], s_in, s_x)
    cn["logR_Copy_Number"], cn["Corrected_logR"], cn["Corrected_Ratio"] = corrected(
        cn["LogRatio"], cn[cn_ratio], cn["CellularPrevalence"], c_in, c_x)
    s_corr, c_corr = s_cn.copy(), c_cn.copy()
    s_major, s_minor = np.asarray(segs["MajorCN"], float).copy(), np.asarray(segs["MinorCN"], float).copy()
    with np.errstate(invalid="ignore"):
        s_round, c_round = np.round(segs["logR_Copy_Number"]), np.round(cn["logR_Copy_Number"])
        if purity >= minPurityToCorrect:
            touched = s_in & (s_cn >= maxCNtoCorrect_autosomes)
            c_touch = c_in & (c_cn >= maxCNtoCorrect_autosomes)
            if correctHOMD:
                touched |= s_in & (s_cn == 0)
                c_touch |= c_in & (c_cn == 0)
            c_touch |= c_in & np.isnan(c_cn)
            if gender == "male" and chr_x:
                touched |= s_x
                c_touch |= c_x
            s_corr[touched] = s_round[touched]
            c_corr[c_touch] = c_round[c_touch]
            opp = (((s_round < ploidy) & (s_corr > ploidy)) | ((s_round > ploidy) & (s_corr < ploidy))) & (np.abs(s_round - s_corr) > 2)
            s_corr[opp] = s_round[opp]
            c_pos = np.asarray(cn["Position"], float)
            for j in np.flatnonzero(opp):                       # bins inside a re-called segment follow it (last hit wins)
                inside = (c_chr == s_chr[j]) & (c_pos >= segs["Start_Position.bp."][j]) & (c_pos <= segs["End_Position.bp."][j])
                c_corr[inside] = s_corr[j]
            touched |= opp
            s_major[touched] = np.round(segs["Corrected_Ratio"][touched] * s_corr[touched])
            s_minor[touched] = np.round((1 - segs["Corrected_Ratio"][touched]) * s_corr[touched])

    def calls(corr, on_x, titan_cn):
        names = np.array(_CN_NAMES + ["HLAMP"] * 1000, dtype=object)
        names_x = np.array(_CN_NAMES_X + ["HLAMP"] * 1000, dtype=object)
        out = np.full(corr.size, None, dtype=object)
        ok = ~np.isnan(corr)
        out[ok] = names[corr[ok].astype(np.int64)]
        if gender == "male" and chr_x:
            sel = on_x & ok
            out[sel] = names_x[corr[sel].astype(np.int64)]
        elif maxCNtoCorrect_X is not None:
            with np.errstate(invalid="ignore"):
                out[on_x & (titan_cn >= maxCNtoCorrect_X)] = "HLAMP"
        return out

    segs["Corrected_Copy_Number"], segs["Corrected_MajorCN"], segs["Corrected_MinorCN"] = s_corr, s_major, s_minor
    segs["Corrected_Call"] = calls(s_corr, s_x, s_cn)
    cn["Corrected_Copy_Number"] = c_corr
    cn["Corrected_Call"] = calls(c_corr, c_x, c_cn)
    return {"cn": cn, "segs": segs}
