"""CPU oracle for TitanCNA's E-step hot path -- a restatement of the reference's R-side numerics.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under titancna_b200/ may import it.

Every function cites the reference lines (relative to /root/reference) it follows.  The native
part of the path (dense log-space forward-backward, Viterbi) is delegated to either
  * oracle/_ref/libtitanref.so  -- the reference's own C, built unmodified (refshim.py), or
  * oracle/libtitan_port.so     -- this repo's C restatement (titan_port.c),
selected with `engine=`.  tests/test_oracle_pinning.py checks the two against each other and
against the reference's fixture data/EMresults.rda (extracted into tests/golden/).

Pinning status: emission + forward-backward are pinned by the reference
fixture; M-step / EM trajectory / Viterbi path are "parity unpinned" by any reference-held
golden vector (the reference has none) -- they are pinned only against the reference's C
object code (Viterbi) and restated from source (M-step; R's uniroot = Brent zeroin, R core
`stats`, not vendored in /root/reference, restated below from its published algorithm).

Deliberate deviations from R (no R interpreter exists here to compare against):
  * lgamma() is the host libm's (glibc), tabulated at integer arguments; R uses its own nmath
    lgammafn.  lchoose(N, x) in the statistic `d` (R/paramEstimation.R:36) is taken as the
    same lgamma bracket (mathematically identical).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libtitan_port.so")
EPS = float(np.finfo(np.float64).eps)
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


_port = None


def port():
    """oracle/libtitan_port.so (built by oracle/Makefile)."""
    global _port
    if _port is None:
        L = C.CDLL(PORT_SO)
        L.port_fwd_back.restype = C.c_int
        L.port_fwd_back.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_double,
                                    C.c_double, C.c_int, _dp, _dp]
        L.port_viterbi.restype = C.c_int
        L.port_viterbi.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_double,
                                   C.c_double, C.c_int, C.POINTER(C.c_int)]
        L.port_transition.restype = None
        L.port_transition.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int]
        L.port_rho.restype = C.c_double
        L.port_rho.argtypes = [C.c_double] * 3
        L.port_lgamma_table.restype = None
        L.port_lgamma_table.argtypes = [C.c_int, _dp]
        L.port_emission.restype = None
        L.port_emission.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int,
                                    C.c_double, _dp]
        _port = L
    return _port


# ----------------------------------------------------------------------------------------
# parameters  (R/utils.R:7-148, 578-597; R/hmmClonal.R:667-671)
# ----------------------------------------------------------------------------------------
def estimate_dirichlet_params_map(kappa):
    """R/hmmClonal.R:667-671"""
    kappa = np.asarray(kappa, dtype=np.float64)
    return (kappa - 1) / (np.sum(kappa) - kappa.size)


def setup_clonal_parameters(Z, sPriorStrength=2):
    """R/utils.R:578-597"""
    s_0 = np.arange(1, Z + 1, dtype=np.float64) / 10
    s_0[0] = 0.001
    return {
        "s_0": s_0,
        "alphaSHyper": np.full(Z, float(sPriorStrength)),
        "betaSHyper": np.full(Z, float(sPriorStrength)),
        "kappaZHyper": np.full(Z, 2.0),
    }


# symmetric genotype table, R/utils.R:24-26,42-43 (rt[0] is rn)
_RT = [None, 1, 1, 1 / 2, 1, 2 / 3, 1, 3 / 4, 2 / 4, 1, 4 / 5, 3 / 5, 1, 5 / 6, 4 / 6, 3 / 6, 1,
       6 / 7, 5 / 7, 4 / 7, 1, 7 / 8, 6 / 8, 5 / 8, 4 / 8]
_CT = [0, 1, 2, 2, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8]


def load_default_parameters(copyNumber=5, numberClonalClusters=1, skew=0.0, hetBaselineSkew=None,
                            alleleEmissionModel="binomial", symmetric=True, data=None):
    """R/utils.R:7-148.  `data` is a dict with ref, tumDepth, logR (or None)."""
    if copyNumber < 3 or copyNumber > 8:
        raise ValueError("loadDefaultParameters: Fewer than 3 or more than 8 copies are being "
                         "specified. Please use minimum 3 or maximum 8 'copyNumber'.")
    if alleleEmissionModel not in ("binomial", "Gaussian"):
        raise ValueError('loadDefaultParameters: alleleEmissionModel must be either "binomial" or "Gaussian".')
    rn = 0.5
    rt = np.array([rn] + _RT[1:], dtype=np.float64)
    rt = rt + skew
    rt[rt > 1] = 1
    rt[rt < (rn + skew)] = rn + skew
    if hetBaselineSkew is not None:
        hetARshift = hetBaselineSkew + 0.5
    elif data is not None:
        hetARshift = float(np.nanmedian(np.asarray(data["ref"], float) / np.asarray(data["tumDepth"], float)))
    else:
        hetARshift = 0.55
    hetState = np.array([4, 9, 16, 25])
    rt[hetState - 1] = hetARshift
    ZS = np.arange(25, dtype=np.float64)
    ct = np.array(_CT, dtype=np.float64)
    ZS[hetState[0] - 1] = -1
    rn = rn + skew
    ind = ct <= copyNumber
    hetState = hetState[hetState <= ind.sum()]
    rt, ZS, ct = rt[ind], ZS[ind], ct[ind]
    K = rt.size
    kappa_base = 100.0
    corRho_0 = None
    if data is not None and len(data.get("ref", [])) > 0 and len(data.get("logR", [])) > 0:
        ar = np.asarray(data["ref"], float) / np.asarray(data["tumDepth"], float)
        corRho_0 = float(np.corrcoef(ar, np.asarray(data["logR"], float))[0, 1])
    if data is not None:
        betaK = 25.0
        alphaK = betaK / float(np.var(np.asarray(data["logR"], float), ddof=1))
        betaR = 25.0
        ar = np.asarray(data["ref"], float) / np.asarray(data["tumDepth"], float)
        alphaR = betaR / float(np.nanvar(ar, ddof=1))
    else:
        alphaK, betaK, alphaR, betaR = 10000.0, 25.0, 10000.0, 25.0
    g = {
        "rt": rt, "rn": rn, "ZS": ZS, "ct": ct, "corRho_0": corRho_0,
        "var_0": np.full(K, 1 / 20), "alphaKHyper": np.full(K, alphaK), "betaKHyper": np.full(K, betaK),
        "alleleEmissionModel": alleleEmissionModel,
        "varR_0": np.full(K, 1 / 20), "alphaRHyper": np.full(K, alphaR), "betaRHyper": np.full(K, betaR),
    }
    kap = np.full(K, kappa_base) + 1
    kap[hetState - 1] = kappa_base * 5
    kap[ct == 0] = kappa_base / 50
    g["kappaGHyper"] = kap
    g["piG_0"] = estimate_dirichlet_params_map(kap)
    g["outlierVar"] = 10000.0
    g["symmetric"] = symmetric
    sP = setup_clonal_parameters(numberClonalClusters)
    sP["piZ_0"] = estimate_dirichlet_params_map(sP["kappaZHyper"])
    return {
        "genotypeParams": g,
        "ploidyParams": {"phi_0": 2.0, "alphaPHyper": 20.0, "betaPHyper": 42.0},
        "normalParams": {"n_0": 0.5, "alphaNHyper": 2.0, "betaNHyper": 2.0},
        "cellPrevParams": sP,
    }


# ----------------------------------------------------------------------------------------
# emissions  (R/hmmClonal.R:491-523, 579-634)
# ----------------------------------------------------------------------------------------
def clonal_two_component_mixture(rt, rn, n, s, ct, phi):
    """R/hmmClonal.R:579-603 -> (muR, muC), each U x Z.  Same association order as the R."""
    rt = np.asarray(rt, float)[:, None]
    ct = np.asarray(ct, float)[:, None]
    s = np.atleast_1d(np.asarray(s, float))[None, :]
    cn = 2.0
    num = n * rn * cn + (1 - n) * s * rn * cn + (1 - n) * (1 - s) * rt * ct
    tot = n * cn + (1 - n) * s * cn + (1 - n) * (1 - s) * ct
    ploidy = n * cn + (1 - n) * phi
    # np.log on a small array: route through math.log so it is the host libm's log
    muC = np.vectorize(math.log)(tot / ploidy)
    return num / tot, muC


_lgam_tab = np.zeros(0)


def lgamma_table(nmax):
    global _lgam_tab
    if _lgam_tab.size <= nmax:
        t = np.zeros(int(nmax) + 1)
        port().port_lgamma_table(int(nmax), _p(t))
        _lgam_tab = t
    return _lgam_tab


def binomial_log_norm(x, depth):
    """lgamma(N+1) - lgamma(k+1) - lgamma(N-k+1), R/hmmClonal.R:618-619 (left-to-right)."""
    x = np.asarray(x, float)
    depth = np.asarray(depth, float)
    xi, di = x.astype(np.int64), depth.astype(np.int64)
    if np.all(xi == x) and np.all(di == depth) and np.all(xi >= 0) and np.all(di >= xi):
        tab = lgamma_table(int(di.max()) if di.size else 0)
        return (tab[di] - tab[xi]) - tab[di - xi]
    lg = np.vectorize(math.lgamma)
    return (lg(depth + 1) - lg(x + 1)) - lg(depth - x + 1)


def compute_py(ref, depth, logR, muR, muC, var, pseudo=1e-300):
    """K x N log emission matrix, joint index j = u + z*U (R/hmmClonal.R:160-171).
    Pure numpy statement of binomialpdflog + normalpdflog (:617-634)."""
    ref, depth, logR = (np.asarray(a, float) for a in (ref, depth, logR))
    U, Z = muR.shape
    mur = muR.reshape(-1, order="F")[:, None]
    muc = muC.reshape(-1, order="F")[:, None]
    v = np.tile(np.asarray(var, float), Z)[:, None]
    log = np.vectorize(math.log)
    c = binomial_log_norm(ref, depth)[None, :]
    pyR = c + (ref[None, :] * log(mur) + (depth - ref)[None, :] * log(1 - mur))
    cN = log(1 / (np.sqrt(v) * math.sqrt(2 * math.pi)))
    pyC = cN + (-((logR[None, :] - muc) ** 2) / (2 * v))
    return (pyR + pyC) + pseudo


def compute_py_c(ref, depth, logR, muR, muC, var, pseudo=1e-300):
    """Same as compute_py through oracle/titan_port.c:port_emission (fast; bit-identical)."""
    ref, depth, logR = (np.ascontiguousarray(a, dtype=np.float64) for a in (ref, depth, logR))
    U, Z = muR.shape
    N = ref.size
    lbin = np.ascontiguousarray(binomial_log_norm(ref, depth))
    mur = np.ascontiguousarray(muR.reshape(-1, order="F"))
    muc = np.ascontiguousarray(muC.reshape(-1, order="F"))
    var = np.ascontiguousarray(var, dtype=np.float64)
    py = np.empty((N, U * Z))
    port().port_emission(_p(ref), _p(depth), _p(logR), _p(lbin), N, _p(mur), _p(muc), _p(var), U, Z,
                         pseudo, _p(py))
    return py.T  # K x N view, Fortran-contiguous


def bivariate_covariance(x1, x2):
    """getBivariateCovariance, R/hmmClonal.R:561-567: sample covariance matrix (n-1 denominator) and the
    correlation derived from it (covar[1,2] / sqrt(covar[1,1] * covar[2,2]))."""
    x1, x2 = np.asarray(x1, float), np.asarray(x2, float)
    ok = ~(np.isnan(x1) | np.isnan(x2))
    a, b = x1[ok], x2[ok]
    da, db = a - a.mean(), b - b.mean()
    nm1 = a.size - 1
    c11, c22, c12 = float(np.dot(da, da)) / nm1, float(np.dot(db, db)) / nm1, float(np.dot(da, db)) / nm1
    return {"covar": np.array([[c11, c12], [c12, c22]]), "cor": c12 / math.sqrt(c11 * c22)}


def compute_py_gauss(ref, depth, logR, muR, muC, varR, var, corRho, pseudo=0.0):
    """K x N log emission matrix of the Gaussian allele model: computeBivariateNormalObslik +
    bivariateNormalpdflog (R/hmmClonal.R:525-559), called as in :153-158 / :269-274 with x1 = ref / tumDepth,
    x2 = logR, var1 = varR (allelic), var2 = var (logR).  Every operation in R's association order;
    `pseudo` is the `py <- py + pseudoCounts` of :158,:274 (the Viterbi wrapper :444-447 adds nothing)."""
    ref, depth, logR = (np.asarray(a, float) for a in (ref, depth, logR))
    U, Z = muR.shape
    with np.errstate(invalid="ignore", divide="ignore"):
        x1, x2 = ref / depth, logR
    varR, var = np.asarray(varR, float), np.asarray(var, float)
    rho = float(corRho)
    py = np.empty((U * Z, x1.size))
    for k in range(U):
        var1, var2 = float(varR[k]), float(var[k])
        c = math.log((2 * math.pi) * math.sqrt((var1 * var2) * (1 - rho * rho)))
        l0 = 1 / (2 * (1 - rho * rho))
        sq = math.sqrt(var1 * var2)
        for z in range(Z):
            d1, d2 = x1 - muR[k, z], x2 - muC[k, z]
            l1 = (d1 * d1) / var1
            l2 = (d2 * d2) / var2
            l3 = (((2 * rho) * d1) * d2) / sq
            y = (-c) - (l0 * ((l1 + l2) - l3))
            y[np.isnan(y)] = 1.0                         # :557
            py[k + z * U] = y + pseudo
    return py


def sufficient_statistics_gauss(ref, depth, logR, rho):
    """a,c,e,f,g,h of R/paramEstimation.R:27-42 (Gaussian allele model), each U x Z; b and d stay zero."""
    ref, depth, logR = (np.asarray(a, float) for a in (ref, depth, logR))
    r = ref / depth
    ein = lambda w: np.einsum("kzt,t->kz", rho, w)
    z = np.zeros(rho.shape[:2])
    return {"a": ein(r), "b": z, "c": ein(logR), "d": z.copy(), "e": rho.sum(axis=2), "f": ein(r * r),
            "g": ein(logR ** 2), "h": ein(r * logR)}


# ----------------------------------------------------------------------------------------
# native recursions through one of the two CPU engines
# ----------------------------------------------------------------------------------------
def fwd_back(log_pi, py, ZS, Z, posn, zstrength, txnlen, outlier=0, engine="port", ct=None):
    """One chromosome: -> (log posterior K x T, loglik).  `.Call("fwd_backC_clonalCN")` contract."""
    py = np.asarray(py, float)
    K, T = py.shape
    if engine == "reference":
        from . import refshim
        r = refshim.reference().fwd_back(log_pi, py, ct if ct is not None else np.zeros(K // Z), ZS, Z,
                                         posn, zstrength, txnlen, outlier)
        return r["rho"], r["loglik"]
    pyf = np.ascontiguousarray(py.T)            # [T][K] == column-major K x T
    lp = np.ascontiguousarray(log_pi, dtype=np.float64)
    zs = np.ascontiguousarray(ZS, dtype=np.float64)
    ps = np.ascontiguousarray(posn, dtype=np.float64)
    g = np.empty((T, K))
    ll = C.c_double(0)
    rc = port().port_fwd_back(_p(lp), _p(pyf), K, T, _p(zs), int(Z), _p(ps), float(zstrength), float(txnlen),
                              int(outlier), _p(g), C.byref(ll))
    if rc != 0:
        raise RuntimeError("port_fwd_back failed")
    return g.T, ll.value


def viterbi(log_pi, py, ZS, Z, posn, zstrength, txnlen, outlier=0, engine="port", ct=None):
    """One chromosome: -> int32 path[T] (1-based).  `.Call("viterbiC_clonalCN")` contract."""
    py = np.asarray(py, float)
    K, T = py.shape
    if engine == "reference":
        from . import refshim
        return refshim.reference().viterbi(log_pi, py, ct if ct is not None else np.zeros(K // Z), ZS, Z,
                                           posn, zstrength, txnlen, outlier)
    pyf = np.ascontiguousarray(py.T)
    lp = np.ascontiguousarray(log_pi, dtype=np.float64)
    zs = np.ascontiguousarray(ZS, dtype=np.float64)
    ps = np.ascontiguousarray(posn, dtype=np.float64)
    path = np.empty(T, dtype=np.int32)
    rc = port().port_viterbi(_p(lp), _p(pyf), K, T, _p(zs), int(Z), _p(ps), float(zstrength), float(txnlen),
                             int(outlier), path.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != 0:
        raise RuntimeError("port_viterbi failed")
    return path


def chromosome_slices(chr_):
    """chrsI of R/hmmClonal.R:112-123: index sets in order of first appearance (unique())."""
    chr_ = np.asarray(chr_)
    seen, out = {}, []
    for c in chr_:
        if c not in seen:
            seen[c] = True
            out.append(np.flatnonzero(chr_ == c))
    return out


def _pmap(fn, items, threads):
    if threads and threads > 1 and len(items) > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:   # ctypes releases the GIL
            return list(ex.map(fn, items))
    return [fn(it) for it in items]


# ----------------------------------------------------------------------------------------
# E-step + sufficient statistics  (R/hmmClonal.R:191-234; R/paramEstimation.R:27-44)
# ----------------------------------------------------------------------------------------
def e_step(data, chrsI, py, piG, piZ, ZS, ct, Z, txnExpLen, txnZstrength, engine="port", threads=1):
    """-> dict(loglik (sum over chromosomes), loglik_chr, rho (U x Z x N posterior), rhoG, rhoZ)."""
    U = piG.size
    piGiZi = np.outer(piG, piZ).reshape(-1, order="F")          # :196, genotype fastest
    log_pi = np.vectorize(math.log)(piGiZi)
    posn = np.asarray(data["posn"], float)

    def one(idx):
        return fwd_back(log_pi, py[:, idx], ZS, Z, posn[idx], txnZstrength * txnExpLen, txnExpLen,
                        0, engine=engine, ct=ct)

    res = _pmap(one, chrsI, threads)
    N = posn.size
    logrho = np.empty((U * Z, N))
    for idx, (g, _) in zip(chrsI, res):
        logrho[:, idx] = g
    ll_chr = np.array([r[1] for r in res])
    rho = np.exp(logrho).reshape((U, Z, N), order="F")          # :230
    return {"loglik": float(np.sum(ll_chr)), "loglik_chr": ll_chr, "rho": rho,
            "rhoZ": rho.sum(axis=0), "rhoG": rho.sum(axis=1)}     # :233-234


def sufficient_statistics(ref, depth, logR, rho):
    """a,b,c,d,e,g of R/paramEstimation.R:27-42 (binomial model), each U x Z."""
    ref, depth, logR = (np.asarray(a, float) for a in (ref, depth, logR))
    lch = binomial_log_norm(ref, depth)
    ein = lambda w: np.einsum("kzt,t->kz", rho, w)
    return {"a": ein(ref), "b": ein(depth - ref), "c": ein(logR), "d": ein(lch),
            "e": rho.sum(axis=2), "g": ein(logR ** 2)}


# ----------------------------------------------------------------------------------------
# priors, objective, M-step  (R/hmmClonal.R:646-665,715-783; R/paramEstimation.R:75-547)
# ----------------------------------------------------------------------------------------
def _lbeta(a, b):
    return math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)


def betapdflog(x, a, b):
    """R/hmmClonal.R:646-650"""
    return -_lbeta(a, b) + (a - 1) * math.log(x) + (b - 1) * math.log(1 - x)


def invgammapdflog(x, a, b):
    """R/hmmClonal.R:652-657"""
    return (a * math.log(b) - math.lgamma(a)) + ((-a - 1) * math.log(x) + (-b / x))


def dirichletpdflog(x, k):
    """R/hmmClonal.R:659-665"""
    x = np.asarray(x, float)
    k = np.asarray(k, float)
    c = math.lgamma(float(np.sum(k))) - float(np.sum([math.lgamma(v) for v in k]))
    return c + float(np.sum((k - 1) * np.log(x)))


def prior_probs(n, s, phi, var, piG, piZ, gParams, nParams, sParams, pParams,
                normalEstimateMethod="map", estimateS=True, estimatePloidy=True, varR=None):
    """R/hmmClonal.R:715-783; prior_gamma_varR (:767-775) only for the Gaussian allele model."""
    out = {"piG": dirichletpdflog(piG, gParams["kappaGHyper"]),
           "piZ": dirichletpdflog(piZ, sParams["kappaZHyper"])}
    ps = 0.0
    if estimateS:
        for z in range(len(s)):
            ps += betapdflog(s[z], sParams["alphaSHyper"][z], sParams["betaSHyper"][z])
    out["s"] = ps
    out["n"] = betapdflog(n, nParams["alphaNHyper"], nParams["betaNHyper"]) if normalEstimateMethod == "map" else 0.0
    ct = np.asarray(gParams["ct"])
    pv = 0.0
    for lvl in _unique_in_order(ct):               # first state of each copy-number level (:746-756)
        k = int(np.flatnonzero(ct == lvl)[0])
        pv += invgammapdflog(var[k], gParams["alphaKHyper"][k], gParams["betaKHyper"][k])
    out["var"] = pv
    out["phi"] = invgammapdflog(phi, pParams["alphaPHyper"], pParams["betaPHyper"]) if estimatePloidy else 0.0
    pvr = 0.0
    if gParams.get("alleleEmissionModel", "binomial") == "Gaussian":
        for k in range(len(var)):
            pvr += invgammapdflog(varR[k], gParams["alphaRHyper"][k], gParams["betaRHyper"][k])
    out["varR"] = pvr
    out["prior"] = out["s"] + out["n"] + out["phi"] + out["var"] + out["varR"] + out["piG"] + out["piZ"]
    return out


def _unique_in_order(v):
    seen, out = set(), []
    for x in v:
        if x not in seen:
            seen.add(x)
            out.append(x)
    return out


def likelihood_func(n, s, var, phi, st, gParams, nParams, sParams, pParams, piG, piZ, likInitG, likInitZ,
                    normalEstimateMethod, estimateS, estimatePloidy):
    """R/paramEstimation.R:287-356 (binomial branch :331-338)."""
    muR, muC = clonal_two_component_mixture(gParams["rt"], gParams["rn"], n, s, gParams["ct"], phi)
    U, Z = muR.shape
    lik = 0.0
    for z in range(Z):
        for k in range(U):
            lik = lik + st["d"][k, z] + st["a"][k, z] * math.log(muR[k, z]) + st["b"][k, z] * math.log(1 - muR[k, z]) \
                + st["e"][k, z] * math.log(1 / math.sqrt(2 * math.pi * var[k])) \
                - (st["g"][k, z] - 2 * st["c"][k, z] * muC[k, z] + st["e"][k, z] * muC[k, z] ** 2) / (2 * var[k])
    pr = prior_probs(n, s, phi, var, piG, piZ, gParams, nParams, sParams, pParams,
                     normalEstimateMethod, estimateS, estimatePloidy)
    return lik + pr["prior"] + likInitG + likInitZ + 0.0


def deriv_n(n, s, var, phi, st, rt, rn, ct, alphaN, betaN):
    """R/paramEstimation.R:359-399 (binomial branch)."""
    muR, muC = clonal_two_component_mixture(rt, rn, n, s, ct, phi)
    cn = 2.0
    tot = 0.0
    for z in range(len(s)):
        den = n * cn + (1 - n) * s[z] * cn + (1 - n) * (1 - s[z]) * ct
        dmuC = (cn - s[z] * cn - (1 - s[z]) * ct) / den - (cn - phi) / (n * cn + (1 - n) * phi)
        dmuR = ((1 - s[z]) * cn * ct * (rn - rt)) / (den ** 2)
        for k in range(len(ct)):
            t1 = st["a"][k, z] * dmuR[k] / muR[k, z] - st["b"][k, z] * dmuR[k] / (1 - muR[k, z])
            t2 = (st["c"][k, z] - st["e"][k, z] * muC[k, z]) * dmuC[k] / var[k]
            tot = tot + (t1 + t2)
    return tot + ((alphaN - 1) / n - (betaN - 1) / (1 - n))


def deriv_s(sz, z, n, var, phi, st, rt, rn, ct, alphaS, betaS):
    """R/paramEstimation.R:401-437 for cluster z (binomial branch)."""
    muR, muC = clonal_two_component_mixture(rt, rn, n, [sz], ct, phi)
    cn = 2.0
    den = n * cn + (1 - n) * sz * cn + (1 - n) * (1 - sz) * ct
    dmuC = ((1 - n) * cn - (1 - n) * ct) / den
    dmuR = ((1 - n) * cn * ct * (rn - rt)) / (den ** 2)
    tot = 0.0
    for k in range(len(ct)):
        t1 = st["a"][k, z] * dmuR[k] / muR[k, 0] - st["b"][k, z] * dmuR[k] / (1 - muR[k, 0])
        t2 = (st["c"][k, z] - st["e"][k, z] * muC[k, 0]) * dmuC[k] / var[k]
        tot = tot + (t1 + t2)
    return tot + ((alphaS - 1) / sz - (betaS - 1) / (1 - sz))


def var_exact(n, s, phi, st, rt, rn, ct, ck, alphaK, betaK):
    """R/paramEstimation.R:442-461"""
    _, muC = clonal_two_component_mixture(rt, rn, n, s, ct, phi)
    t1 = 0.0
    t2 = 0.0
    for k in np.flatnonzero(ct == ck):
        for z in range(len(s)):
            t1 = t1 + -(st["g"][k, z] - 2 * st["c"][k, z] * muC[k, z] + st["e"][k, z] * muC[k, z] ** 2)
            t2 = t2 + -(st["e"][k, z])
    t1 = t1 - 2 * betaK
    t2 = t2 - 2 * (alphaK + 1)
    return t1 / t2


def deriv_phi(phi, n, s, var, st, rt, rn, ct, alphaP, betaP):
    """R/paramEstimation.R:499-530 (binomial branch)."""
    _, muC = clonal_two_component_mixture(rt, rn, n, s, ct, phi)
    cn = 2.0
    dmuC = -(1 - n) / (n * cn + (1 - n) * phi)
    tot = 0.0
    for z in range(len(s)):
        for k in range(len(ct)):
            t1 = st["c"][k, z] / var[k]
            t2 = st["e"][k, z] * muC[k, z] / var[k]
            tot = tot + (t1 - t2) * dmuC
    return tot + ((-alphaP - 1) / phi + betaP / phi ** 2)


# ---- Gaussian allele model (alleleEmissionModel = "Gaussian"): the same functions' other branches -------------
# The reference flags this model "not fully implemented" (vignettes/TitanCNA.Rnw:123); its objective carries an
# operator-precedence slip (log(1 / 2 * pi * ...), R/paramEstimation.R:321) and the ploidy derivative indexes the
# K x Z matrices with a single subscript (:512-513, i.e. column 1 for every cluster).  Both are restated as written.
def likelihood_func_gauss(n, s, var, varR, phi, st, gParams, nParams, sParams, pParams, piG, piZ, likInitG, likInitZ,
                          normalEstimateMethod, estimateS, estimatePloidy):
    """R/paramEstimation.R:287-356, Gaussian branch :313-329."""
    muR, muC = clonal_two_component_mixture(gParams["rt"], gParams["rn"], n, s, gParams["ct"], phi)
    U, Z = muR.shape
    rho = float(gParams["corRho_0"])
    a, c, e, f, g, h = (st[k] for k in "acefgh")
    lik = 0.0
    for z in range(Z):
        for k in range(U):
            term1 = e[k, z] * math.log(((1 / 2) * math.pi) * math.sqrt((var[k] * varR[k]) * (1 - rho * rho)))
            term2 = 1 / (2 * (1 - rho * rho))
            term3 = ((f[k, z] - (2 * a[k, z]) * muR[k, z]) + e[k, z] * (muR[k, z] * muR[k, z])) / (2 * varR[k])
            term4 = ((g[k, z] - (2 * c[k, z]) * muC[k, z]) + e[k, z] * (muC[k, z] * muC[k, z])) / (2 * var[k])
            term5 = ((2 * rho) * (((h[k, z] - a[k, z] * muC[k, z]) - c[k, z] * muR[k, z])
                                  + (e[k, z] * muC[k, z]) * muR[k, z])) / math.sqrt(var[k] * varR[k])
            lik = (lik + term1) - term2 * ((term3 + term4) - term5)
    pr = prior_probs(n, s, phi, var, piG, piZ, gParams, nParams, sParams, pParams,
                     normalEstimateMethod, estimateS, estimatePloidy, varR=varR)
    return lik + pr["prior"] + likInitG + likInitZ + 0.0


def deriv_n_gauss(n, s, var, varR, phi, rho, st, rt, rn, ct, alphaN, betaN):
    """R/paramEstimation.R:359-399, Gaussian branch :376-383."""
    muR, muC = clonal_two_component_mixture(rt, rn, n, s, ct, phi)
    cn = 2.0
    a, c, e = st["a"], st["c"], st["e"]
    tot = 0.0
    for z in range(len(s)):
        den = n * cn + (1 - n) * s[z] * cn + (1 - n) * (1 - s[z]) * ct
        dmuC = (cn - s[z] * cn - (1 - s[z]) * ct) / den - (cn - phi) / (n * cn + (1 - n) * phi)
        dmuR = ((1 - s[z]) * cn * ct * (rn - rt)) / (den ** 2)
        for k in range(len(ct)):
            term0 = 1 / (1 - rho * rho)
            rR, rC = a[k, z] - e[k, z] * muR[k, z], c[k, z] - e[k, z] * muC[k, z]
            term1 = rR * dmuR[k] / varR[k]
            term2 = rC * dmuC[k] / var[k]
            term3 = rho * ((rC * dmuR[k]) + (rR * dmuC[k])) / math.sqrt(varR[k] * var[k])
            tot = tot + term0 * (term1 + term2 - term3)
    return tot + ((alphaN - 1) / n - (betaN - 1) / (1 - n))


def deriv_s_gauss(sz, z, n, var, varR, phi, rho, st, rt, rn, ct, alphaS, betaS):
    """R/paramEstimation.R:401-437 for cluster z, Gaussian branch :415-422."""
    muR, muC = clonal_two_component_mixture(rt, rn, n, [sz], ct, phi)
    cn = 2.0
    den = n * cn + (1 - n) * sz * cn + (1 - n) * (1 - sz) * ct
    dmuC = ((1 - n) * cn - (1 - n) * ct) / den
    dmuR = ((1 - n) * cn * ct * (rn - rt)) / (den ** 2)
    a, c, e = st["a"], st["c"], st["e"]
    tot = 0.0
    for k in range(len(ct)):
        term0 = 1 / (1 - rho * rho)
        rR, rC = a[k, z] - e[k, z] * muR[k, 0], c[k, z] - e[k, z] * muC[k, 0]
        term1 = rR * dmuR[k] / varR[k]
        term2 = rC * dmuC[k] / var[k]
        term3 = rho * ((rC * dmuR[k]) + (rR * dmuC[k])) / math.sqrt(varR[k] * var[k])
        tot = tot + term0 * (term1 + term2 - term3)
    return tot + ((alphaS - 1) / sz - (betaS - 1) / (1 - sz))


def deriv_var_gauss(v, n, s, var2, phi, rho, a, c, e, g, h, rt, rn, ct, alphaK, betaK, kind):
    """clonalCNDerivativeVarUpdateEqn, R/paramEstimation.R:463-497.  `a, c, e, g, h` are the row subsets the caller
    hands over (|ind| x Z); for kind == "allelicRatio" the caller has already swapped a <-> c and passes f as g
    (:189-190).  The prior term is summed over every state of the subset (:493-494); R's sum() accumulates in
    long double."""
    muR, muC = clonal_two_component_mixture(rt, rn, n, s, ct, phi)
    mus1, mus2 = (muC, muR) if kind == "logRatio" else (muR, muC)
    var2 = np.atleast_1d(np.asarray(var2, float))
    K, Z = len(alphaK), len(s)
    dlik = 0.0
    for z in range(Z):
        for k in range(K):
            term1 = e[k, z] / (2 * v)
            term2 = 1 / (2 * (1 - rho * rho))
            term3 = ((g[k, z] - (2 * c[k, z]) * mus1[k, z]) + e[k, z] * (mus1[k, z] * mus1[k, z])) / (v * v)
            term4 = (rho * (((h[k, z] - c[k, z] * mus2[k, z]) - a[k, z] * mus1[k, z])
                            + (e[k, z] * mus1[k, z]) * mus2[k, z])) / (v * math.sqrt(v * var2[k]))
            dlik = (dlik + (-term1)) + (term2 * (term3 - term4))
    dinv = np.longdouble(0)
    for k in range(K):
        dinv = dinv + np.longdouble((-alphaK[k] - 1) / v + betaK[k] / (v * v))
    return dlik + float(dinv)


def deriv_phi_gauss(phi, n, s, var, varR, rho, st, rt, rn, ct, alphaP, betaP):
    """R/paramEstimation.R:499-530, Gaussian branch :510-515: a[k], c[k], e[k], mus$R[k], mus$C[k] are single
    subscripts into K x Z matrices, i.e. column 1, whatever z is."""
    muR, muC = clonal_two_component_mixture(rt, rn, n, s, ct, phi)
    cn = 2.0
    dmuC = -(1 - n) / (n * cn + (1 - n) * phi)
    a, c, e = st["a"], st["c"], st["e"]
    tot = 0.0
    for z in range(len(s)):
        for k in range(len(ct)):
            term0 = 1 / (1 - rho * rho)
            term1 = (c[k, 0] - e[k, 0] * muC[k, 0]) * dmuC / var[k]
            term2 = rho * ((a[k, 0] - e[k, 0] * muR[k, 0]) * dmuC) / math.sqrt(varR[k] * var[k])
            tot = tot + term0 * (term1 - term2)
    return tot + ((-alphaP - 1) / phi + betaP / phi ** 2)


class RootError(RuntimeError):
    pass


def r_zeroin(f, ax, bx, fa, fb, tol, maxit):
    """Brent's zeroin as in R core src/library/stats/src/zeroin.c (R_zeroin2) -- the algorithm
    behind stats::uniroot (call sites R/paramEstimation.R:135,155,223).  Not vendored in
    /root/reference; restated from the published Forsythe-Malcolm-Moler / netlib C version."""
    a, b, c, fc = ax, bx, ax, fa
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    it = maxit + 1
    while it > 0:
        it -= 1
        prev_step = b - a
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2 * EPS * abs(b) + tol / 2
        new_step = (c - b) / 2
        if abs(new_step) <= tol_act or fb == 0.0:
            return b
        if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                t1 = fb / fa
                p = cb * t1
                q = 1.0 - t1
            else:
                q = fa / fc
                t1 = fb / fc
                t2 = fb / fa
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0))
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if p > 0:
                q = -q
            else:
                p = -p
            if p < (0.75 * cb * q - abs(tol_act * q) / 2) and p < abs(prev_step * q / 2):
                new_step = p / q
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0 else -tol_act
        a, fa = b, fb
        b += new_step
        fb = f(b)
        if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
            c, fc = a, fa
    return b


def uniroot(f, lower, upper, tol=1e-15, maxiter=1000):
    """stats::uniroot front end: evaluates the end points, truncates infinities to +-DBL_MAX,
    raises when the end points are NA or do not bracket (the reference's tryCatch then keeps the
    previous iterate, R/paramEstimation.R:134-139)."""
    big = float(np.finfo(np.float64).max)
    fl, fu = float(f(lower)), float(f(upper))
    if math.isnan(fl) or math.isnan(fu):
        raise RootError("f.lower/f.upper is NA")
    fl = max(min(fl, big), -big)
    fu = max(min(fu, big), -big)
    sgn = lambda v: int(v > 0) - int(v < 0)
    if not (sgn(fl) * sgn(fu) <= 0):
        raise RootError("f() values at end points not of opposite sign")

    def g(x):
        v = f(x)
        if math.isnan(v):
            raise RootError("f returned NaN")
        return max(min(v, big), -big)

    return r_zeroin(g, lower, upper, fl, fu, tol, maxiter)


def update_parameters(n_0, s_0, var_0, phi_0, st, likInitG, likInitZ, piG, piZ, gParams, nParams, sParams,
                      pParams, normalEstimateMethod="map", estimateS=True, estimatePloidy=True,
                      maxiter=500, miniter=10):
    """Coordinate descent of R/paramEstimation.R:75-283 (binomial model).
    Order per sweep: n -> s_1..s_Z -> var (per copy-number level) -> phi; the *last* sweep is
    returned (:272), and the loop runs while ((!converged && i < maxiter) || i <= miniter)."""
    rt, rn, ct = np.asarray(gParams["rt"], float), gParams["rn"], np.asarray(gParams["ct"], float)
    Z = len(s_0)
    n_prev, s_prev, var_prev, phi_prev = float(n_0), np.array(s_0, float), np.array(var_0, float), float(phi_0)
    F = lambda n, s, var, phi: likelihood_func(n, s, var, phi, st, gParams, nParams, sParams, pParams, piG, piZ,
                                               likInitG, likInitZ, normalEstimateMethod, estimateS, estimatePloidy)
    objfun = [F(n_prev, s_prev, var_prev, phi_prev)]
    converged, i = False, 1
    lo, hi = EPS, 1 - EPS
    n, s, var, phi = n_prev, s_prev.copy(), var_prev.copy(), phi_prev
    while ((not converged) and i < maxiter) or i <= miniter:
        i += 1
        if normalEstimateMethod == "map":
            try:
                n = uniroot(lambda v: deriv_n(v, s_prev, var_prev, phi_prev, st, rt, rn, ct,
                                              nParams["alphaNHyper"], nParams["betaNHyper"]), lo, hi)
            except (RootError, ZeroDivisionError, ValueError):
                n = n_prev
        else:
            n = n_prev
        s = s_prev.copy()
        if estimateS:
            for z in range(Z):
                try:
                    s[z] = uniroot(lambda v: deriv_s(v, z, n, var_prev, phi_prev, st, rt, rn, ct,
                                                     sParams["alphaSHyper"][z], sParams["betaSHyper"][z]), lo, hi)
                except (RootError, ZeroDivisionError, ValueError):
                    s[z] = s_prev[z]
        var = np.zeros_like(var_prev)
        for lvl in _unique_in_order(ct):
            ind = np.flatnonzero(ct == lvl)
            var[ind] = var_exact(n, s, phi_prev, st, rt, rn, ct, lvl, gParams["alphaKHyper"][ind[0]],
                                 gParams["betaKHyper"][ind[0]])
        if estimatePloidy:
            try:
                phi = uniroot(lambda v: deriv_phi(v, n, s, var, st, rt, rn, ct, pParams["alphaPHyper"],
                                                  pParams["betaPHyper"]), EPS, 10.0)
            except (RootError, ZeroDivisionError, ValueError):
                phi = phi_prev
        else:
            phi = phi_prev
        objfun.append(F(n, s, var, phi))
        n_prev, s_prev, var_prev, phi_prev = n, s.copy(), var.copy(), phi
        cur, prv = objfun[-1], objfun[-2]
        if not math.isnan(cur) and (abs(cur - prv) / abs(cur)) <= 0.001:
            converged = True
    return {"n": n, "s": s, "var": var, "phi": phi, "maxFind": i, "maxF": objfun[-1], "sweeps": i - 1}


def update_parameters_gauss(n_0, s_0, var_0, varR_0, phi_0, st, likInitG, likInitZ, piG, piZ, gParams, nParams, sParams,
                            pParams, normalEstimateMethod="map", estimateS=True, estimatePloidy=True,
                            maxiter=500, miniter=10):
    """Coordinate descent of R/paramEstimation.R:75-283, Gaussian allele model: per sweep n -> s_1..s_Z -> var per
    copy-number level (uniroot on (eps, 5), :165-183) -> varR per genotype state (:185-200, with the NEW var[k]) ->
    phi; failed roots keep the previous iterate; the last sweep is returned."""
    rt, rn, ct = np.asarray(gParams["rt"], float), gParams["rn"], np.asarray(gParams["ct"], float)
    rho = float(gParams["corRho_0"])
    aK, bK = np.asarray(gParams["alphaKHyper"], float), np.asarray(gParams["betaKHyper"], float)
    aR, bR = np.asarray(gParams["alphaRHyper"], float), np.asarray(gParams["betaRHyper"], float)
    Z, U = len(s_0), len(rt)
    n_prev, s_prev, phi_prev = float(n_0), np.array(s_0, float), float(phi_0)
    var_prev, varR_prev = np.array(var_0, float), np.array(varR_0, float)
    F = lambda n, s, var, varR, phi: likelihood_func_gauss(n, s, var, varR, phi, st, gParams, nParams, sParams, pParams,
                                                           piG, piZ, likInitG, likInitZ, normalEstimateMethod,
                                                           estimateS, estimatePloidy)
    objfun = [F(n_prev, s_prev, var_prev, varR_prev, phi_prev)]
    converged, i = False, 1
    lo, hi = EPS, 1 - EPS
    n, s, var, varR, phi = n_prev, s_prev.copy(), var_prev.copy(), varR_prev.copy(), phi_prev
    bad = (RootError, ZeroDivisionError, ValueError)
    while ((not converged) and i < maxiter) or i <= miniter:
        i += 1
        if normalEstimateMethod == "map":
            try:
                n = uniroot(lambda v: deriv_n_gauss(v, s_prev, var_prev, varR_prev, phi_prev, rho, st, rt, rn, ct,
                                                    nParams["alphaNHyper"], nParams["betaNHyper"]), lo, hi)
            except bad:
                n = n_prev
        else:
            n = n_prev
        s = s_prev.copy()
        if estimateS:
            for z in range(Z):
                try:
                    s[z] = uniroot(lambda v: deriv_s_gauss(v, z, n, var_prev, varR_prev, phi_prev, rho, st, rt, rn, ct,
                                                           sParams["alphaSHyper"][z], sParams["betaSHyper"][z]), lo, hi)
                except bad:
                    s[z] = s_prev[z]
        var = np.zeros_like(var_prev)
        for lvl in _unique_in_order(ct):
            ind = np.flatnonzero(ct == lvl)
            try:
                var[ind] = uniroot(lambda v: deriv_var_gauss(v, n, s, varR_prev[ind], phi_prev, rho, st["a"][ind], st["c"][ind],
                                                             st["e"][ind], st["g"][ind], st["h"][ind], rt[ind], rn, ct[ind],
                                                             aK[ind], bK[ind], "logRatio"), EPS, 5.0)
            except bad:
                var[ind] = var_prev[ind]
        varR = np.zeros_like(varR_prev)
        for k in range(U):
            kk = slice(k, k + 1)
            try:
                varR[k] = uniroot(lambda v: deriv_var_gauss(v, n, s, var[k], phi_prev, rho, st["c"][kk], st["a"][kk],
                                                            st["e"][kk], st["f"][kk], st["h"][kk], rt[kk], rn, ct[kk],
                                                            aR[kk], bR[kk], "allelicRatio"), EPS, 5.0)
            except bad:
                varR[k] = varR_prev[k]
        if estimatePloidy:
            try:
                phi = uniroot(lambda v: deriv_phi_gauss(v, n, s, var, varR, rho, st, rt, rn, ct, pParams["alphaPHyper"],
                                                        pParams["betaPHyper"]), EPS, 10.0)
            except bad:
                phi = phi_prev
        else:
            phi = phi_prev
        objfun.append(F(n, s, var, varR, phi))
        n_prev, s_prev, var_prev, varR_prev, phi_prev = n, s.copy(), var.copy(), varR.copy(), phi
        cur, prv = objfun[-1], objfun[-2]
        if not math.isnan(cur) and (abs(cur - prv) / abs(cur)) <= 0.001:
            converged = True
    return {"n": n, "s": s, "var": var, "varR": varR, "phi": phi, "maxFind": i, "maxF": objfun[-1], "sweeps": i - 1}


def estimate_mix_weights(rho_at_chr0, kappa):
    """R/paramEstimation.R:533-547: note sum() over the whole sub-matrix (Appendix B #3)."""
    kappa = np.asarray(kappa, float)
    tot = float(np.sum(rho_at_chr0))
    return (tot + kappa - 1) / (tot + np.sum(kappa) - kappa.size)


def m_step(data, es, n_0, s_0, phi_0, var_0, piG_0, piZ_0, chrPos_0, params, normalEstimateMethod="map",
           estimateS=True, estimatePloidy=True, maxiterUpdate=1500, stats=None, varR_0=None):
    """estimateClonalCNParamsMap, R/paramEstimation.R:10-70."""
    g, nP, sP, pP = (params[k] for k in ("genotypeParams", "normalParams", "cellPrevParams", "ploidyParams"))
    gauss = g.get("alleleEmissionModel", "binomial") == "Gaussian"
    suff = sufficient_statistics_gauss if gauss else sufficient_statistics
    st = stats if stats is not None else suff(data["ref"], data["tumDepth"], data["logR"], es["rho"])
    rhoG0, rhoZ0 = es["rhoG"][:, chrPos_0], es["rhoZ"][:, chrPos_0]
    log = np.vectorize(math.log)
    likInitG = float(np.sum(log(np.asarray(piG_0, float)) @ rhoG0))        # :43
    likInitZ = float(np.sum(log(np.asarray(piZ_0, float)) @ rhoZ0))        # :44
    if gauss:
        out = update_parameters_gauss(n_0, s_0, var_0, varR_0, phi_0, st, likInitG, likInitZ, piG_0, piZ_0, g, nP, sP, pP,
                                      normalEstimateMethod, estimateS, estimatePloidy, maxiter=maxiterUpdate, miniter=10)
    else:
        out = update_parameters(n_0, s_0, var_0, phi_0, st, likInitG, likInitZ, piG_0, piZ_0, g, nP, sP, pP,
                                normalEstimateMethod, estimateS, estimatePloidy, maxiter=maxiterUpdate, miniter=10)
        out["varR"] = None if varR_0 is None else np.array(varR_0, float)      # :212 varR <- varR_prev
    out["piZ"] = estimate_mix_weights(rhoZ0, sP["kappaZHyper"])
    out["piG"] = estimate_mix_weights(rhoG0, g["kappaGHyper"])
    out["stats"] = st
    out["likInitG"], out["likInitZ"] = likInitG, likInitZ
    return out


# ----------------------------------------------------------------------------------------
# EM driver + Viterbi wrapper  (R/hmmClonal.R:8-368, 370-485)
# ----------------------------------------------------------------------------------------
def run_em_clonal_cn(data, params, txnExpLen=1e15, txnZstrength=5e5, maxiter=15, maxiterUpdate=1500,
                     pseudoCounts=1e-300, normalEstimateMethod="map", estimateS=True, estimatePloidy=True,
                     likChangeThreshold=0.001, engine="port", threads=1, verbose=False):
    """runEMclonalCN (either allele model, useOutlierState=FALSE).  Returns the convergeParams
    list of R/hmmClonal.R:344-367 as a dict of arrays with an iteration axis last."""
    g, nP, sP, pP = (params[k] for k in ("genotypeParams", "normalParams", "cellPrevParams", "ploidyParams"))
    gauss = g.get("alleleEmissionModel", "binomial") == "Gaussian"
    ref, depth, logR = (np.asarray(data[k], float) for k in ("ref", "tumDepth", "logR"))
    Z, U, N = len(sP["s_0"]), len(g["rt"]), ref.size
    chrsI = chromosome_slices(data["chr"])
    chrPos_0 = np.array([ix[0] for ix in chrsI])
    maxit = maxiter + 1
    n = np.zeros(maxit); phi = np.zeros(maxit); loglik = np.zeros(maxit)
    s = np.zeros((Z, maxit)); var = np.zeros((U, maxit)); piG = np.zeros((U, maxit)); piZ = np.zeros((Z, maxit))
    muR = np.zeros((U, Z, maxit)); muC = np.zeros((U, Z, maxit))
    varR = np.zeros((U, maxit))
    i = 0                                   # 0-based column (R's i = 1)
    loglik[i] = -np.inf
    s[:, i], var[:, i], phi[i], n[i] = sP["s_0"], g["var_0"], pP["phi_0"], nP["n_0"]
    varR[:, i] = g.get("varR_0", np.full(U, 1 / 20))
    piG[:, i], piZ[:, i] = g["piG_0"], sP["piZ_0"]
    muR[:, :, i], muC[:, :, i] = clonal_two_component_mixture(g["rt"], g["rn"], n[i], s[:, i], g["ct"], phi[i])

    def emissions(col):
        if gauss:                                                   # :149-158, :265-274
            return compute_py_gauss(ref, depth, logR, muR[:, :, col], muC[:, :, col], varR[:, col], var[:, col],
                                    g["corRho_0"], pseudoCounts)
        return compute_py_c(ref, depth, logR, muR[:, :, col], muC[:, :, col], var[:, col], pseudoCounts)

    py = emissions(i)
    converged = False
    es = None
    sweeps = []
    while not converged and (i + 1) < maxit:
        i += 1
        es = e_step(data, chrsI, py, piG[:, i - 1], piZ[:, i - 1], g["ZS"], g["ct"], Z, txnExpLen, txnZstrength,
                    engine=engine, threads=threads)
        loglik[i] = es["loglik"]
        mo = m_step(data, es, n[i - 1], s[:, i - 1], phi[i - 1], var[:, i - 1], piG[:, i - 1], piZ[:, i - 1],
                    chrPos_0, params, normalEstimateMethod, estimateS, estimatePloidy, maxiterUpdate,
                    varR_0=varR[:, i - 1])
        sweeps.append(mo["sweeps"])
        n[i], s[:, i], var[:, i], phi[i], piZ[:, i], piG[:, i] = mo["n"], mo["s"], mo["var"], mo["phi"], mo["piZ"], mo["piG"]
        varR[:, i] = mo["varR"]
        muR[:, :, i], muC[:, :, i] = clonal_two_component_mixture(g["rt"], g["rn"], n[i], s[:, i], g["ct"], phi[i])
        py = emissions(i)
        pr = prior_probs(n[i], s[:, i], phi[i], var[:, i], piG[:, i], piZ[:, i], g, nP, sP, pP,
                         normalEstimateMethod, estimateS, estimatePloidy, varR=varR[:, i])
        loglik[i] = loglik[i] + pr["prior"]
        if verbose:
            print(f"EM iter {i}: loglik={loglik[i]:.4f} n={n[i]:.4f} phi={phi[i]:.4f} s={s[:, i]}")
        if (abs(loglik[i] - loglik[i - 1]) / abs(loglik[i])) <= likChangeThreshold and loglik[i] >= loglik[i - 1]:
            converged = True
        elif loglik[i] < loglik[i - 1]:
            converged = True
            i -= 1
    m = i + 1
    return {"n": n[:m], "s": s[:, :m], "var": var[:, :m], "varR": varR[:, :m], "phi": phi[:m], "piG": piG[:, :m], "piZ": piZ[:, :m],
            "muR": muR[:, :, :m], "muC": muC[:, :, :m], "loglik": loglik[:m],
            "rhoG": es["rhoG"] if es else None, "rhoZ": es["rhoZ"] if es else None,
            "txnExpLen": txnExpLen, "txnZstrength": txnZstrength, "useOutlierState": False,
            "genotypeParams": g, "ploidyParams": pP, "normalParams": nP, "cellPrevParams": sP,
            "symmetric": g.get("symmetric", True), "cd_sweeps": sweeps}


def viterbi_clonal_cn(data, cp, engine="port", threads=1):
    """viterbiClonalCN, R/hmmClonal.R:370-485: emissions from the LAST column i, initial
    distribution from column i-1 (Appendix B #2)."""
    g = cp["genotypeParams"]
    Z = cp["s"].shape[0]
    i = cp["s"].shape[1] - 1
    ref, depth, logR = (np.asarray(data[k], float) for k in ("ref", "tumDepth", "logR"))
    if g.get("alleleEmissionModel", "binomial") == "Gaussian":
        # :439-447: convergeParams carries no corRho_0 element, so computeBivariateNormalObslik receives NULL and
        # falls back to the correlation of getBivariateCovariance(x1, x2) (:536-538); no pseudoCounts are added
        cor = bivariate_covariance(ref / depth, logR)["cor"]
        py = compute_py_gauss(ref, depth, logR, cp["muR"][:, :, i], cp["muC"][:, :, i], cp["varR"][:, i], cp["var"][:, i],
                              cor, 0.0)
    else:
        py = compute_py_c(ref, depth, logR, cp["muR"][:, :, i], cp["muC"][:, :, i], cp["var"][:, i], 1e-300)
    piGiZi = np.outer(cp["piG"][:, i - 1], cp["piZ"][:, i - 1]).reshape(-1, order="F")
    log_pi = np.vectorize(math.log)(piGiZi)
    chrsI = chromosome_slices(data["chr"])
    posn = np.asarray(data["posn"], float)
    zst = cp["txnZstrength"] * cp["txnExpLen"]

    def one(idx):
        return viterbi(log_pi, py[:, idx], g["ZS"], Z, posn[idx], zst, cp["txnExpLen"], 0, engine=engine, ct=g["ct"])

    parts = _pmap(one, chrsI, threads)
    out = np.empty(posn.size, dtype=np.int32)
    for idx, p in zip(chrsI, parts):
        out[idx] = p
    return out
