"""loadDefaultParameters / setupClonalParameters -- host-side mirror of R/utils.R:7-148,578-597.

Same names, argument meaning and error text as the reference; R lists become dicts of numpy
arrays.  (O(U) table code; nothing here is N-sized.)
"""
from __future__ import annotations

import numpy as np


def estimateDirichletParamsMap(kappa):
    """R/hmmClonal.R:667-671"""
    kappa = np.asarray(kappa, dtype=np.float64)
    return (kappa - 1) / (np.sum(kappa) - kappa.size)


def setupClonalParameters(Z, sPriorStrength=2):
    """R/utils.R:578-597"""
    s_0 = np.arange(1, Z + 1, dtype=np.float64) / 10
    s_0[0] = 0.001   # first cluster should be clonally dominant
    return {"s_0": s_0,
            "alphaSHyper": np.full(Z, float(sPriorStrength)),
            "betaSHyper": np.full(Z, float(sPriorStrength)),
            "kappaZHyper": np.ones(Z) + 1}


def _symmetric_table(rn, max_copies=8):
    """rt / ct of the symmetric genotype table for total copies 0..max_copies (R/utils.R:24-26,42-43 list 0..8):
    for c copies the major-allele fractions c/c, (c-1)/c, ... down to 1/2.  The same rule continued to 9 and 10
    copies gives the 36-state table of BASELINE configs[3] (ct 9: 1, 8/9 .. 5/9; ct 10: 1, .9 .. .5)."""
    rt, ct, balanced = [rn, 1.0], [0, 1], []
    for c in range(2, max_copies + 1):
        for major in range(c, (c + 1) // 2 - 1, -1):
            rt.append(major / c)
            ct.append(c)
            if 2 * major == c:
                balanced.append(len(rt))          # 1-based index of the balanced (HET) state
    return np.array(rt, dtype=np.float64), np.array(ct, dtype=np.float64), np.array(balanced)


def dataMoments(data):
    """The data-dependent numbers loadDefaultParameters derives (R/utils.R:32-40,80-95): median and variance of the allelic
    ratio, variance of logR, their correlation.  O(N) with a sort: a sweep over (ploidy, numClusters) computes them once
    and hands them to every loadDefaultParameters(..., moments=) call instead of 15 times."""
    ref, depth, logR = (np.asarray(data[k], float) for k in ("ref", "tumDepth", "logR"))
    ar = ref / depth
    return {"hetARshift": float(np.nanmedian(ar)),
            "corRho_0": float(np.corrcoef(ar, logR)[0, 1]) if ref.size > 0 and logR.size > 0 else None,
            "var_logR": float(np.var(logR, ddof=1)), "var_ar": float(np.nanvar(ar, ddof=1))}


def loadDefaultParameters(copyNumber=5, numberClonalClusters=1, skew=0, hetBaselineSkew=None,
                          alleleEmissionModel="binomial", symmetric=True, data=None, extendedTable=False, moments=None):
    """R/utils.R:7-148.  `data` (optional) is a dict with ref, tumDepth, logR.
    `extendedTable=True` (no R counterpart) lifts the reference's limit of 8 copies to 10: 36 genotype states.
    `moments=dataMoments(data)` (no R counterpart) reuses the data-dependent numbers across calls."""
    if data is not None and moments is None:
        moments = dataMoments(data)
    if copyNumber < 3 or copyNumber > (10 if extendedTable else 8):
        raise ValueError("loadDefaultParameters: Fewer than 3 or more than 8 copies are \n             "
                         "being specified. Please use minimum 3 or maximum 8 'copyNumber'.")
    if alleleEmissionModel not in ("binomial", "Gaussian"):
        raise ValueError('loadDefaultParameters: alleleEmissionModel must be either "binomial" or "Gaussian".')
    rn = 0.5
    rt, ct, hetState = _symmetric_table(rn, 10 if extendedTable else 8)     # hetState = c(4, 9, 16, 25[, 36])
    rt = rt + skew
    rt[rt > 1] = 1
    rt[rt < (rn + skew)] = rn + skew
    if hetBaselineSkew is not None:
        hetARshift = hetBaselineSkew + 0.5
    elif data is not None:
        hetARshift = moments["hetARshift"]
    else:
        hetARshift = 0.55
    rt[hetState - 1] = hetARshift
    ZS = np.arange(rt.size, dtype=np.float64)
    ZS[hetState[0] - 1] = -1          # diploid heterozygous marker (R/utils.R:67)
    rn = rn + skew
    ind = ct <= copyNumber
    hetState = hetState[hetState <= int(ind.sum())]
    rt, ZS, ct = rt[ind], ZS[ind], ct[ind]
    K = rt.size
    kappaGHyper_base = 100.0
    corRho_0 = None
    have_data = data is not None
    if have_data and len(data.get("ref", ())) > 0 and len(data.get("logR", ())) > 0:
        corRho_0 = moments["corRho_0"]
    if have_data:
        betaK = 25.0
        alphaK = betaK / moments["var_logR"]
        betaR = 25.0
        alphaR = betaR / moments["var_ar"]
    else:
        alphaK, betaK, alphaR, betaR = 10000.0, 25.0, 10000.0, 25.0
    genotypeParams = {
        "rt": rt, "rn": rn, "ZS": ZS, "ct": ct, "corRho_0": corRho_0,
        "var_0": np.full(K, 1 / 20), "alphaKHyper": np.full(K, alphaK), "betaKHyper": np.full(K, betaK),
        "alleleEmissionModel": alleleEmissionModel,
        "varR_0": np.full(K, 1 / 20), "alphaRHyper": np.full(K, alphaR), "betaRHyper": np.full(K, betaR),
    }
    kappa = np.full(K, kappaGHyper_base) + 1
    kappa[hetState - 1] = kappaGHyper_base * 5
    kappa[ct == 0] = kappaGHyper_base / 50
    genotypeParams["kappaGHyper"] = kappa
    genotypeParams["piG_0"] = estimateDirichletParamsMap(kappa)
    genotypeParams["outlierVar"] = 10000.0
    genotypeParams["symmetric"] = symmetric
    sParams = setupClonalParameters(Z=numberClonalClusters)
    sParams["piZ_0"] = estimateDirichletParamsMap(sParams["kappaZHyper"])
    return {"genotypeParams": genotypeParams,
            "ploidyParams": {"phi_0": 2.0, "alphaPHyper": 20.0, "betaPHyper": 42.0},
            "normalParams": {"n_0": 0.5, "alphaNHyper": 2.0, "betaNHyper": 2.0},
            "cellPrevParams": sParams}
