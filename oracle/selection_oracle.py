"""Loop restatement of the reference's S_Dbw validity index and solution-selection rule.

TEST INFRASTRUCTURE ONLY.  Follows R/utils.R:604-742 (computeSDbwIndex), :745-768 (sdbw.density) and
scripts/R_scripts/selectSolution.R:58-137 statement by statement with plain Python lists (no numpy
vectorisation), including the reference's quirks: the "HET" filter of the allelic-ratio branch never
matches (:646), stdev[] holds variances (:667), union() is a set union (:704).
Parity unpinned: the reference holds no golden value for these functions.
"""
import math


def _isna(v):
    return v is None or v != v


def _mean(xs):
    xs = [v for v in xs if not _isna(v)]
    return sum(xs) / len(xs)


def _var(xs):
    xs = [v for v in xs if not _isna(v)]
    if len(xs) < 2:
        return float("nan")
    m = sum(xs) / len(xs)
    return sum((v - m) ** 2 for v in xs) / (len(xs) - 1)


def _sd(xs):
    v = _var(xs)
    return math.sqrt(v) if v == v else float("nan")


def _median(xs):
    xs = sorted(v for v in xs if not _isna(v))
    n = len(xs)
    return xs[n // 2] if n % 2 else (xs[n // 2 - 1] + xs[n // 2]) / 2


def sdbw_density(x, avg_stdev, st_dev=None, method="Halkidi", centroid=None, centroid_method="median"):
    if centroid is None:
        centroid = _median(x) if centroid_method == "median" else _mean(x)
    if method == "Halkidi":
        radius = avg_stdev
    else:
        if st_dev is None:
            st_dev = _sd(x)
        radius = 1.96 * (st_dev / math.sqrt(len(x)))
    count = 0
    for v in x:
        if _isna(v) or _isna(centroid) or _isna(radius):
            continue
        if abs(v - centroid) <= radius:
            count += 1
    return count


def compute_sdbw_index(rows, centroid_method="median", data_type="LogRatio", use_corrected_cn=True,
                       method="Halkidi", symmetric=True):
    cn_col = "Corrected_Copy_Number" if (use_corrected_cn and "Corrected_Copy_Number" in rows[0]) else "CopyNumber"
    flat, raw = [], []
    if data_type == "LogRatio":
        cn = []
        for r in rows:
            c = r[cn_col] + 1
            cn.append(None if c == 3 else c)
        mx = max(c for c in cn if c is not None)
        for r, c in zip(rows, cn):
            z = r["ClonalCluster"]
            flat.append(3 if (c is None or _isna(z)) else (z - 1) * mx + c)
            raw.append(r[data_type])
    else:
        st = [r["TITANstate"] + 1 for r in rows]
        mx = max(st)
        for r, s_ in zip(rows, st):
            z = r["ClonalCluster"]
            flat.append((4 if symmetric else 5) if _isna(z) else (z - 1) * mx + s_)
            raw.append(max(r[data_type], 1 - r[data_type]))
    m, sd = _mean(raw), _sd(raw)
    val = [(v - m) / sd for v in raw]
    clust = sorted(set(flat))
    K, N = len(clust), len(rows)
    members = [[v for v, f in zip(val, flat) if f == c] for c in clust]
    var_d = _var(val)
    stdev, scat_ci = [], []
    for mem in members:
        v = _var(mem)
        stdev.append(v)
        scat_ci.append(v / var_d if method == "Halkidi" else ((N - len(mem)) / N) * (v / var_d))
    avg_stdev = math.sqrt(sum(v for v in stdev if not _isna(v))) / K
    total = 0.0
    for i in range(K):
        di = sdbw_density(members[i], avg_stdev, method=method, centroid_method=centroid_method)
        for j in range(K):
            if i == j:
                continue
            dj = sdbw_density(members[j], avg_stdev, method=method, centroid_method=centroid_method)
            union = []
            seen = set()
            for v in members[i] + members[j]:
                if v not in seen:
                    seen.add(v)
                    union.append(v)
            ci, cj = _median(members[i]), _median(members[j])
            ni, nj = len(members[i]), len(members[j])
            std = (_sd(members[i]) + _sd(members[j])) / 2
            if method == "Halkidi":
                cij = (ci + cj) / 2
            else:
                lam = 0.7
                den = di + dj
                second = (di * ci + dj * cj) / den if den != 0 else float("nan")
                cij = lam * ((nj * ci + ni * cj) / (ni + nj)) + (1 - lam) * second
            dij = sdbw_density(union, avg_stdev, st_dev=std, method=method, centroid=cij, centroid_method=centroid_method)
            mxd = max(di, dj)
            total += dij / (0.1 if mxd == 0 else mxd)
    ssum = sum(v for v in scat_ci if not _isna(v))
    div = K if method == "Halkidi" else K - 1
    scat = ssum / div if div != 0 else (float("nan") if ssum == 0 else math.copysign(float("inf"), ssum))
    dens = total / (K * (K - 1)) if K > 1 else float("nan")
    return {"S_DbwIndex": scat + dens, "dens_bw": dens, "scat": scat}


def select_solution(records, threshold=0.05, homd_snp_thres=1500, homd_snp_prop=0.02):
    by = {2: [], 3: [], 4: []}
    for rec in records:
        r = dict(rec)
        segs = r.get("segs")
        if segs is not None:
            tot = homd = 0
            big = False
            for call, ln in zip(segs["TITAN_call"], segs["Length.snp."]):
                tot += ln
                if call == "HOMD":
                    homd += ln
                    big = big or ln > homd_snp_thres
            if homd > 0 and (homd / tot > homd_snp_prop or big):
                r["loglik"], r["sdbw"] = -math.inf, math.inf
        by[int(r["ploidy_0"])].append(r)
    for p in by:
        by[p].sort(key=lambda r: str(r["numClusters"]))
    votes = {}
    for k in range(len(by[2])):
        cand = [by[2][k]["loglik"] + abs(by[2][k]["loglik"]) * threshold]
        for p in (3, 4):
            cand.append(by[p][k % len(by[p])]["loglik"] if by[p] else float("nan"))
        arg = None
        for q in range(3):
            if cand[q] == cand[q] and (arg is None or cand[q] > cand[arg]):
                arg = q
        votes[arg + 1] = votes.get(arg + 1, 0) + 1
    winner = None
    for v in sorted(votes):
        if winner is None or votes[v] > votes[winner]:
            winner = v
    group = by[{1: 2, 2: 3, 3: 4}[winner]]
    best = None
    for r in group:
        if r["sdbw"] == r["sdbw"] and (best is None or r["sdbw"] < best["sdbw"]):
            best = r
    return best
