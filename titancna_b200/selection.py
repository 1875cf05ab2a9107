"""Solution selection for the (ploidy x numClusters) sweep -- host-side mirror of the reference's
computeSDbwIndex / sdbw.density (R/utils.R:604-768), outputModelParameters + printSDbw (:978-1063, :1403-1417)
and the selection rule of scripts/R_scripts/selectSolution.R:58-137.

O(N) host work on the result table that follows the hot path; numpy, like the reference's R.  Tables are dicts
of equal-length numpy columns (titancna_b200.results).  Reference quirks kept on purpose:
  * AllelicRatio branch: `st[x[, which(TITANstate == "HET")]] <- NA` compares a numeric column with "HET" and so
    never removes anything (:646); only rows without a clonal cluster fall into the catch-all flat state 4;
  * `stdev[i]` holds a VARIANCE (:667) and avgStdev = sqrt(sum of variances) / K (:681);
  * the union of two clusters is a SET union (duplicate values collapse, :704) while its confidence interval
    uses the length of that set (:757);
  * a single-member cluster has NA variance and drops out of every na.rm sum.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["computeSDbwIndex", "outputModelParameters", "selectSolution"]


def _var(x):
    x = x[~np.isnan(x)]
    return float(np.var(x, ddof=1)) if x.size > 1 else float("nan")


def _sd(x):
    v = _var(x)
    return math.sqrt(v) if v == v else float("nan")


def _scale(x):
    """R's scale(): centre by the mean, divide by the sample standard deviation (NA ignored)."""
    x = np.asarray(x, float)
    return (x - np.nanmean(x)) / _sd(x)


def _density(x, avg_stdev, st_dev=None, method="Halkidi", centroid=None, centroid_method="median"):
    """sdbw.density (R/utils.R:745-768): members within a radius of the centroid."""
    if centroid is None:
        centroid = np.nanmedian(x) if centroid_method == "median" else np.nanmean(x)
    if method == "Halkidi":
        radius = avg_stdev
    else:
        if st_dev is None:
            st_dev = _sd(x)
        radius = 1.96 * (st_dev / math.sqrt(x.size))
    with np.errstate(invalid="ignore"):
        return int(np.count_nonzero(np.abs(x - centroid) <= radius))


def computeSDbwIndex(x, centroid_method="median", data_type="LogRatio", use_corrected_cn=True,
                     S_Dbw_method="Halkidi", symmetric=True):
    """R/utils.R:604-742 -> dict(S_DbwIndex, dens_bw, scat)."""
    if data_type not in ("LogRatio", "AllelicRatio", "HaplotypeRatio"):
        raise ValueError("computeSDbwIndex: data.type must be either 'LogRatio' or 'AllelicRatio'")
    if S_Dbw_method not in ("Halkidi", "Tong"):
        raise ValueError("computeSDbwIndex: S_Dbw.method must be either 'Halkidi' or 'Tong'")
    cn_col = "Corrected_Copy_Number" if (use_corrected_cn and "Corrected_Copy_Number" in x) else "CopyNumber"
    clonal = np.asarray(x["ClonalCluster"], float)
    if data_type == "LogRatio":
        cn = np.asarray(x[cn_col], float) + 1
        cn[cn == 3] = np.nan                                     # all CN = 2 positions
        flat = (clonal - 1) * np.nanmax(cn) + cn
        flat[np.isnan(flat)] = 3
        val = _scale(x[data_type])
    else:
        st = np.asarray(x["TITANstate"], float) + 1             # the "HET" filter of :646 never matches
        flat = (clonal - 1) * np.nanmax(st) + st
        flat[np.isnan(flat)] = 4 if symmetric else 5
        r = np.asarray(x[data_type], float)
        val = _scale(np.maximum(r, 1 - r))
    clust = np.unique(flat)
    K, N = clust.size, flat.size
    members = [val[flat == c] for c in clust]
    var_d = _var(val)
    stdev = np.array([_var(m) for m in members])
    if S_Dbw_method == "Halkidi":
        scat_ci = stdev / var_d
    else:
        scat_ci = np.array([((N - m.size) / N) * (v / var_d) for m, v in zip(members, stdev)])
    avg_stdev = math.sqrt(np.nansum(stdev)) / K
    dens = np.full((K, K), np.nan)
    own = [_density(m, avg_stdev, method=S_Dbw_method, centroid_method=centroid_method) for m in members]
    med = [float(np.nanmedian(m)) for m in members]
    sds = [_sd(m) for m in members]
    # union(xci, xcj) is a SET: work on the sorted distinct values of every cluster (computed once) and on the
    # pairwise overlap counts |Ci n Cj| (one membership matrix product) instead of K^2 sorts of N values
    uniq = [np.unique(m[~np.isnan(m)]) for m in members]
    has_na = [bool(np.isnan(m).any()) for m in members]
    allv, inv = np.unique(np.concatenate(uniq), return_inverse=True)
    common = np.zeros((K, K))
    off = np.cumsum([0] + [u.size for u in uniq])
    for a in range(0, allv.size, 1 << 18):                       # chunked so the matrix stays small
        b = min(a + (1 << 18), allv.size)
        M = np.zeros((b - a, K), dtype=np.float32)
        for c in range(K):
            ids = inv[off[c]:off[c + 1]]
            ids = ids[(ids >= a) & (ids < b)]
            M[ids - a, c] = 1
        common += M.T.astype(np.float64) @ M.astype(np.float64)
    for i in range(K):
        for j in range(K):
            if i == j:
                continue
            ni, nj = members[i].size, members[j].size
            n_union = uniq[i].size + uniq[j].size - int(common[i, j]) + (1 if (has_na[i] or has_na[j]) else 0)
            if S_Dbw_method == "Halkidi":
                cij = (med[i] + med[j]) / 2
                radius = avg_stdev
            else:
                lam = 0.7
                with np.errstate(invalid="ignore", divide="ignore"):
                    cij = lam * ((nj * med[i] + ni * med[j]) / (ni + nj)) + (1 - lam) * \
                        np.float64(own[i] * med[i] + own[j] * med[j]) / np.float64(own[i] + own[j])
                radius = 1.96 * (((sds[i] + sds[j]) / 2) / math.sqrt(n_union))
            d = 0
            if cij == cij and radius == radius:
                # |x - cij| <= radius evaluated exactly as the reference does, on the candidates a (slightly widened)
                # binary search leaves
                pad = 4 * np.spacing(abs(cij) + abs(radius))
                parts = []
                for u in (uniq[i], uniq[j]):
                    w = u[np.searchsorted(u, cij - radius - pad, "left"):np.searchsorted(u, cij + radius + pad, "right")]
                    parts.append(w[np.abs(w - cij) <= radius])
                d = parts[0].size + parts[1].size - np.intersect1d(parts[0], parts[1], assume_unique=True).size
            mx = max(own[i], own[j])
            dens[i, j] = d / (0.1 if mx == 0 else mx)
    with np.errstate(invalid="ignore", divide="ignore"):
        scat = np.nansum(scat_ci) / np.float64(K if S_Dbw_method == "Halkidi" else K - 1)
        dens_bw = np.nansum(dens) / np.float64(K * (K - 1))
    return {"S_DbwIndex": float(scat + dens_bw), "dens_bw": float(dens_bw), "scat": float(scat)}


def _rnum(v, digits):
    """as.character(signif(v, digits)): shortest decimal of the rounded value, up to 15 significant digits."""
    v = float(v)
    if v != v:
        return "NA"
    if math.isinf(v):
        return "Inf" if v > 0 else "-Inf"
    return format(float(f"{v:.{digits}g}"), ".15g")


def outputModelParameters(convergeParams, results, filename=None, S_Dbw_scale=1, S_Dbw_method="Tong",
                          S_Dbw_useCorrectedCN=True):
    """R/utils.R:978-1063.  Returns dict(dens_bw, scat, S_Dbw, lines); `lines` is the text of the *.params.txt
    file the snakemake workflow collects (written to `filename` if given)."""
    cp = convergeParams
    s_all = np.asarray(cp["s"], float).reshape(np.asarray(cp["s"]).shape[0], -1)
    Z, i = s_all.shape[0], s_all.shape[1] - 1
    s = np.sort(s_all[:, i], kind="stable")
    ratio = "HaplotypeRatio" if "HaplotypeRatio" in results else "AllelicRatio"
    model = (cp.get("genotypeParams") or {}).get("alleleEmissionModel", "binomial")
    lines = [f"Normal contamination estimate:\t{_rnum(np.atleast_1d(cp['n'])[i], 4)}",
             f"Average tumour ploidy estimate:\t{_rnum(np.atleast_1d(cp['phi'])[i], 4)}",
             f"Clonal cluster cellular prevalence Z={Z}:\t" + " ".join(_rnum(1 - v, 4) for v in s)]
    muR, muC = np.asarray(cp["muR"], float), np.asarray(cp["muC"], float)
    for j in range(Z):
        lines.append(f"{ratio} {model} means for clonal cluster Z={j + 1}:\t" + " ".join(_rnum(v, 4) for v in muR[:, j, i]))
        lines.append(f"logRatio Gaussian means for clonal cluster Z={j + 1}:\t" +
                     " ".join(_rnum(math.log2(math.exp(v)), 4) for v in muC[:, j, i]))
    if model == "Gaussian":
        lines.append(f"{ratio} Gaussian variance:\t" + " ".join(_rnum(v, 4) for v in np.asarray(cp["varR"], float)[:, i]))
    lines.append("logRatio Gaussian variance:\t" + " ".join(_rnum(v, 4) for v in np.asarray(cp["var"], float)[:, i]))
    lines.append(f"Number of iterations:\t{np.atleast_1d(cp['phi']).size}")
    lines.append(f"Log likelihood:\t{_rnum(np.atleast_1d(cp['loglik'])[i], 6)}")
    kw = dict(centroid_method="median", use_corrected_cn=S_Dbw_useCorrectedCN, S_Dbw_method=S_Dbw_method,
              symmetric=cp.get("symmetric", True))
    lr = computeSDbwIndex(results, data_type="LogRatio", **kw)
    ar = computeSDbwIndex(results, data_type=ratio, **kw)
    both = {k: lr[k] + ar[k] for k in lr}
    for d, name in ((lr, "LogRatio"), (ar, ratio), (both, "Both")):
        lines.append("S_Dbw dens.bw (%s):\t%0.4f " % (name, d["dens_bw"]))
        lines.append("S_Dbw scat (%s):\t%0.4f " % (name, d["scat"]))
        lines.append("S_Dbw validity index (%s):\t%0.4f " % (name, S_Dbw_scale * d["dens_bw"] + d["scat"]))
    if filename is not None:
        with open(filename, "w") as fh:
            fh.write("\n".join(lines) + "\n")
    return {"dens_bw": both["dens_bw"], "scat": both["scat"], "S_Dbw": S_Dbw_scale * both["dens_bw"] + both["scat"],
            "lines": lines}


def selectSolution(solutions, threshold=0.05, homd_snp_thres=1500, homd_snp_prop=0.02):
    """scripts/R_scripts/selectSolution.R:58-137 on in-memory records.  `solutions`: dicts with ploidy_0 (2, 3 or 4),
    numClusters, loglik, sdbw and optionally segs (outputTitanSegments table).  Ploidy 2 must be present.  Returns
    the chosen record: the ploidy whose solutions most often have the best log-likelihood (ploidy 2 gets a bonus of
    threshold * |loglik|), then the smallest S_Dbw within it; solutions with large homozygous deletions are vetoed."""
    groups = {}
    for sol in solutions:
        r = dict(sol)
        segs = r.get("segs")
        if segs is not None and np.any(np.asarray(segs["TITAN_call"]) == "HOMD"):
            ln = np.asarray(segs["Length.snp."], float)
            homd = np.asarray(segs["TITAN_call"]) == "HOMD"
            if ln[homd].sum() / ln.sum() > homd_snp_prop or np.count_nonzero(ln[homd] > homd_snp_thres) > 0:
                r["loglik"], r["sdbw"] = -math.inf, math.inf
        groups.setdefault(int(r["ploidy_0"]), []).append(r)
    if 2 not in groups:
        raise ValueError("selectSolution: ploidy 2 solutions are required (--ploidyRun2)")
    for g in groups.values():
        g.sort(key=lambda r: str(r["numClusters"]))             # list.files(): character order of "_clusterN"
    if 3 not in groups and 4 not in groups:
        raise ValueError("phi3Dir or phi4Dir must be provided.")
    n2 = len(groups[2])

    def col(p):                                                  # cbind() recycles a shorter / missing column
        g = groups.get(p)
        if not g:
            return [math.nan] * n2
        return [g[k % len(g)]["loglik"] for k in range(n2)]

    c2 = [r["loglik"] + abs(r["loglik"]) * threshold for r in groups[2]]
    c3, c4 = col(3), col(4)
    votes = []
    for k in range(n2):
        row = [c2[k], c3[k], c4[k]]
        best, arg = None, None
        for q, v in enumerate(row):                              # which.max: first maximum, NA skipped
            if v == v and (best is None or v > best):
                best, arg = v, q + 1
        if arg is not None:
            votes.append(arg)
    counts = {v: votes.count(v) for v in sorted(set(votes))}     # table(): names ascending
    ploidy_ind = max(counts, key=lambda v: (counts[v], -v))      # which.max: first of the largest counts
    opt = groups[{1: 2, 2: 3, 3: 4}[ploidy_ind]]
    pick = None
    for r in opt:                                                # which.min: first minimum, NA skipped
        if r["sdbw"] == r["sdbw"] and (pick is None or r["sdbw"] < pick["sdbw"]):
            pick = r
    return pick
