"""runEMclonalCN / viterbiClonalCN -- host-side mirror of R/hmmClonal.R:8-485 over the C ABI.

Same names, arguments, defaults, return fields and error text as the reference's R functions
(`data` and `params` are dicts of numpy arrays where R uses lists).  All N-sized work happens on the
GPU behind include/titancna_b200.h; only O(U*Z) numbers come back per EM iteration.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (DataDesc, EmOpts, EmOut, EstepOut, Hyper, Params, Timing, TitanError, c_double_p, check, f64, lib)


def chromosome_offsets(chr_):
    """Contiguous runs of `chr` -> offsets [nChr+1] (R/hmmClonal.R:112-123 builds which(chr==c);
    loadAlleleCounts sorts by chromosome, R/utils.R:183-194, so runs are contiguous)."""
    chr_ = np.asarray(chr_)
    if chr_.size == 0:
        return np.zeros(1, dtype=np.int64)
    change = np.flatnonzero(chr_[1:] != chr_[:-1]) + 1
    starts = np.concatenate([[0], change, [chr_.size]]).astype(np.int64)
    labels = chr_[starts[:-1]]
    if len(set(labels.tolist())) != labels.size:
        raise ValueError("data$chr must be sorted so that every chromosome is one contiguous run")
    return starts


class Solver:
    """Device-resident data set + scratch for one (data, U, Z, txnExpLen, txnZstrength)."""

    def __init__(self, data, genotypeParams, Z, txnExpLen, txnZstrength, device=None):
        for k in ("chr", "posn", "ref", "tumDepth", "logR"):
            if data.get(k) is None:
                raise ValueError("data must contain named list elements: chr, posn, ref, tumDepth, \n"
                                 "             logR. See loadDataFromFile()")
        L = lib()
        if device is not None:
            check(L.titan_set_device(int(device)))
        self.starts = chromosome_offsets(data["chr"])
        self.nChr = self.starts.size - 1
        self.N = int(np.asarray(data["posn"]).size)
        self.U = int(np.asarray(genotypeParams["rt"]).size)
        self.Z = int(Z)
        self.K = self.U * self.Z
        # the C ABI takes bare pointers: every column must really have N entries, every per-genotype vector U
        for k in ("chr", "ref", "tumDepth", "logR"):
            if np.asarray(data[k]).size != self.N:
                raise ValueError(f"data${k} has {np.asarray(data[k]).size} entries, data$posn has {self.N}")
        for k in ("ct", "ZS"):
            if np.asarray(genotypeParams[k]).size < self.U:
                raise ValueError(f"genotypeParams${k} has fewer entries than genotypeParams$rt ({self.U})")
        keep = []
        d = DataDesc()
        d.N, d.nChr = self.N, self.nChr
        d.chr_start = self.starts.ctypes.data_as(_lib.c_i64_p)
        for name, key in (("posn", "posn"), ("ref", "ref"), ("tumDepth", "tumDepth"), ("logR", "logR")):
            arr, ptr = f64(data[key])
            keep.append(arr)
            setattr(d, name, ptr)
        d.txnExpLen, d.txnZstrength = float(txnExpLen), float(txnZstrength)
        d.U, d.Z = self.U, self.Z
        zs, d.ZS = f64(genotypeParams["ZS"])
        model = genotypeParams.get("alleleEmissionModel") or "binomial"        # R/hmmClonal.R:392-394
        if model not in ("binomial", "Gaussian"):
            raise ValueError('loadDefaultParameters: alleleEmissionModel must be either "binomial" or "Gaussian".')
        self.gauss = model == "Gaussian"
        d.alleleModel = 1 if self.gauss else 0
        h = C.c_void_p()
        check(L.titan_solver_create(C.byref(d), C.byref(h)))
        self._h = h
        self.nChrOut = self.nChr          # whole-genome chromosome count once set_comm() shards the solution

    def close(self):
        if getattr(self, "_h", None):
            lib().titan_solver_destroy(self._h)
            self._h = None

    __del__ = close

    def configure(self, chunk_len=1024, warmup_len=192, rel_tol=1e-8, max_rounds=64):
        check(lib().titan_solver_configure(self._h, int(chunk_len), int(warmup_len), float(rel_tol), int(max_rounds)))

    def set_stream(self, cuda_stream_ptr):
        check(lib().titan_solver_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_comm(self, comm, nChr_total, chr_global):
        """This solver holds chromosomes `chr_global` (ascending indices) of an `nChr_total`-chromosome data set whose
        other chromosomes live on the other ranks of `comm` (a _lib.Comm): estep() / runEMclonalCN then return
        whole-genome statistics on every rank (one NCCL all-reduce per E-step inside the library)."""
        if comm is None:
            check(lib().titan_solver_set_comm(self._h, None, 0, None))
            self.nChrOut = self.nChr
            return
        idx = np.ascontiguousarray(chr_global, dtype=np.int32)
        if idx.size != self.nChr:
            raise ValueError("chr_global must name every chromosome this solver holds")
        check(lib().titan_solver_set_comm(self._h, comm._h, int(nChr_total), idx.ctypes.data_as(C.POINTER(C.c_int32))))
        self.nChrOut = int(nChr_total)
        self._comm = comm                 # keep the communicator alive as long as the solver uses it

    def timing(self):
        t = Timing()
        check(lib().titan_solver_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def _params(self, n, s, phi, var, piG, piZ, g, keep, varR=None, corRho=None):
        p = Params()
        p.n, p.phi, p.rn = float(n), float(phi), float(g["rn"])
        cols = [("s", s), ("var", var), ("piG", piG), ("piZ", piZ), ("rt", g["rt"]), ("ct", g["ct"])]
        if self.gauss:
            cols.append(("varR", g["varR_0"] if varR is None else varR))
            rho = g.get("corRho_0") if corRho is None else corRho
            if rho is None:
                raise ValueError("genotypeParams$corRho_0 is NULL: the Gaussian allele model needs "
                                 "loadDefaultParameters(data=) (R/utils.R:79-83)")
            p.corRho = float(rho)
        for name, val in cols:
            arr, ptr = f64(np.atleast_1d(val))
            if name in ("var", "varR", "rt", "ct", "piG") and arr.size < self.U:
                raise ValueError(f"{name} has {arr.size} entries, the solver has {self.U} genotype states")
            keep.append(arr)
            setattr(p, name, ptr)
        return p

    def estep(self, n, s, phi, var, piG, piZ, genotypeParams, posteriors=False, varR=None, corRho=None):
        """One E-step -> dict(loglik_chr, stats{a,b,c,d,e,g}[U,Z], rhoG0, rhoZ0, [rhoG, rhoZ], muR, muC).
        Gaussian allele model: stats{a,f,c,h,e,g} (R/paramEstimation.R:30-32), `varR` / `corRho` default to
        genotypeParams$varR_0 / corRho_0."""
        keep = []
        p = self._params(n, s, phi, var, piG, piZ, genotypeParams, keep, varR, corRho)
        U, Z, K, nChr, N = self.U, self.Z, self.K, self.nChrOut, self.N
        ll = np.zeros(nChr); st = np.zeros(6 * K); g0 = np.zeros(nChr * U); z0 = np.zeros(nChr * Z)
        muR = np.zeros(K); muC = np.zeros(K)
        o = EstepOut()
        o.loglik_chr, o.stats = ll.ctypes.data_as(c_double_p), st.ctypes.data_as(c_double_p)
        o.rhoG0, o.rhoZ0 = g0.ctypes.data_as(c_double_p), z0.ctypes.data_as(c_double_p)
        o.muR, o.muC = muR.ctypes.data_as(c_double_p), muC.ctypes.data_as(c_double_p)
        if posteriors:
            rG = np.zeros((N, U)); rZ = np.zeros((N, Z))
            o.rhoG, o.rhoZ = rG.ctypes.data_as(c_double_p), rZ.ctypes.data_as(c_double_p)
        check(lib().titan_estep(self._h, C.byref(p), C.byref(o)))
        out = {"loglik_chr": ll, "loglik": float(np.sum(ll)),
               "stats": {nm: st[i * K:(i + 1) * K].reshape((U, Z), order="F")
                         for i, nm in enumerate("afcheg" if self.gauss else "abcdeg")},
               "stats_flat": st,
               "rhoG0": g0.reshape((U, nChr), order="F"), "rhoZ0": z0.reshape((Z, nChr), order="F"),
               "muR": muR.reshape((U, Z), order="F"), "muC": muC.reshape((U, Z), order="F")}
        if posteriors:
            out["rhoG"], out["rhoZ"] = rG.T, rZ.T          # U x N, Z x N like R
        return out

    def viterbi(self, n, s, phi, var, piG, piZ, genotypeParams, varR=None, corRho=None):
        keep = []
        p = self._params(n, s, phi, var, piG, piZ, genotypeParams, keep, varR, corRho)
        path = np.zeros(self.N, dtype=np.int32)
        check(lib().titan_viterbi(self._h, C.byref(p), path.ctypes.data_as(C.POINTER(C.c_int32))))
        return path


def _hyper(params, keep):
    g, pP, nP, sP = (params[k] for k in ("genotypeParams", "ploidyParams", "normalParams", "cellPrevParams"))
    h = Hyper()
    for name, val in (("rt", g["rt"]), ("ct", g["ct"]), ("ZS", g["ZS"]), ("var_0", g["var_0"]),
                      ("alphaKHyper", g["alphaKHyper"]), ("betaKHyper", g["betaKHyper"]),
                      ("kappaGHyper", g["kappaGHyper"]), ("piG_0", g["piG_0"]), ("s_0", sP["s_0"]),
                      ("alphaSHyper", sP["alphaSHyper"]), ("betaSHyper", sP["betaSHyper"]),
                      ("kappaZHyper", sP["kappaZHyper"]), ("piZ_0", sP["piZ_0"])):
        arr, ptr = f64(np.atleast_1d(val))
        keep.append(arr)
        setattr(h, name, ptr)
    h.rn = float(g["rn"])
    h.n_0, h.alphaNHyper, h.betaNHyper = float(nP["n_0"]), float(nP["alphaNHyper"]), float(nP["betaNHyper"])
    h.phi_0, h.alphaPHyper, h.betaPHyper = float(pP["phi_0"]), float(pP["alphaPHyper"]), float(pP["betaPHyper"])
    gauss = g.get("alleleEmissionModel", "binomial") == "Gaussian"
    if gauss and any(g.get(k) is None for k in ("varR_0", "alphaRHyper", "betaRHyper", "corRho_0")):
        raise ValueError("genotypeParams must contain varR_0, alphaRHyper, betaRHyper and corRho_0 for the Gaussian "
                         "allele model. See loadDefaultParameters(data=).")
    for name in ("varR_0", "alphaRHyper", "betaRHyper"):       # binomial model: varR_0 is only carried along (:212)
        if g.get(name) is not None:
            arr, ptr = f64(np.atleast_1d(g[name]))
            keep.append(arr)
            setattr(h, name, ptr)
    if gauss:
        h.corRho_0 = float(g["corRho_0"])
    return h


def _check_params(params):
    """argument checks of R/hmmClonal.R:15-54 (same messages)"""
    if any(params.get(k) is None for k in ("genotypeParams", "ploidyParams", "normalParams", "cellPrevParams")):
        raise ValueError("params must include genotypeParams, ploidyParams, normalParams, cellPrevParams.")
    g, pP, nP, sP = (params[k] for k in ("genotypeParams", "ploidyParams", "normalParams", "cellPrevParams"))
    if any(g.get(k) is None for k in ("rt", "ct", "rn", "ZS", "var_0", "alphaKHyper", "betaKHyper", "kappaGHyper")):
        raise ValueError("genotypeParams must contain named list elements: rt, rn, ct, ZS, \n"
                         "             var_0, alphaKHyper, betaKHyper, kappaGHyper.\n"
                         "             See loadDefaultParameters().")
    if any(pP.get(k) is None for k in ("phi_0", "alphaPHyper", "betaPHyper")):
        raise ValueError("ploidyParams must contain named list elements: phi_0, alphaPHyper, \n"
                         "             betaPHyper. See loadDefaultParameters()")
    if any(nP.get(k) is None for k in ("n_0", "alphaNHyper", "betaNHyper")):
        raise ValueError("normalParams must contain named list elements: n_0, alphaNHyper, \n"
                         "             betaNHyper. See loadDefaultParameters()")
    if any(sP.get(k) is None for k in ("s_0", "kappaZHyper", "alphaSHyper", "betaSHyper")):
        raise ValueError("sParams (clonal parameters) must contain named list elements: s_0, \n"
                         "             kappaSHyper, alphaSHyper, betaSHyper. \n"
                         "             See setupClonalParameters().")


def runEMclonalCN(data, params, txnExpLen=1e15, txnZstrength=5e5, maxiter=15, maxiterUpdate=1500,
                  pseudoCounts=1e-300, normalEstimateMethod="map", estimateS=True, estimatePloidy=True,
                  useOutlierState=False, likChangeThreshold=0.001, verbose=True, solver=None,
                  chunk_len=None, warmup_len=None, rel_tol=None, return_timing=False, posteriors=True):
    """R/hmmClonal.R:8-368.  Returns the convergeParams list as a dict (iteration axis last).
    Extra keyword arguments (solver reuse, chunking knobs) have no R counterpart.  `posteriors=False` leaves
    rhoG / rhoZ (8 (U + Z) N bytes, the only N-sized output) on the device and returns None for them: nothing in
    the sweep -> outputTitanResults(posteriorProbs=FALSE) -> selectSolution chain reads them."""
    _check_params(params)
    g, pP, nP, sP = (params[k] for k in ("genotypeParams", "ploidyParams", "normalParams", "cellPrevParams"))
    if useOutlierState:
        raise TitanError(_lib.ERR_UNSUPPORTED, "useOutlierState=TRUE is not covered (no shipped driver enables it)")
    Z, U = len(sP["s_0"]), len(g["rt"])
    own = solver is None
    if own:
        solver = Solver(data, g, Z, txnExpLen, txnZstrength)
    try:
        if chunk_len is not None or warmup_len is not None or rel_tol is not None:
            solver.configure(1024 if chunk_len is None else chunk_len, 192 if warmup_len is None else warmup_len,
                             1e-8 if rel_tol is None else rel_tol)
        N, K, M = solver.N, U * Z, int(maxiter) + 1
        keep = []
        h = _hyper(params, keep)
        o = EmOpts(int(maxiter), int(maxiterUpdate), int(normalEstimateMethod == "map"), int(bool(estimateS)),
                   int(bool(estimatePloidy)), float(likChangeThreshold), int(bool(posteriors)), int(bool(verbose)))
        bufs = {"n": np.zeros(M), "phi": np.zeros(M), "loglik": np.zeros(M), "s": np.zeros((M, Z)),
                "piZ": np.zeros((M, Z)), "var": np.zeros((M, U)), "piG": np.zeros((M, U)),
                "muR": np.zeros((M, K)), "muC": np.zeros((M, K)), "varR": np.zeros((M, U))}
        if posteriors:
            bufs["rhoG"], bufs["rhoZ"] = np.zeros((N, U)), np.zeros((N, Z))
        sweeps = np.zeros(M, dtype=np.int32)
        out = EmOut()
        for k, a in bufs.items():
            setattr(out, k, a.ctypes.data_as(c_double_p))
        out.cd_sweeps = sweeps.ctypes.data_as(C.POINTER(C.c_int32))
        check(lib().titan_em_run(solver._h, C.byref(h), C.byref(o), C.byref(out)))
        i = out.iters
        res = {"n": bufs["n"][:i].copy(), "s": bufs["s"][:i].T.copy(), "var": bufs["var"][:i].T.copy(),
               "varR": bufs["varR"][:i].T.copy(),
               "phi": bufs["phi"][:i].copy(), "piG": bufs["piG"][:i].T.copy(), "piZ": bufs["piZ"][:i].T.copy(),
               "muR": bufs["muR"][:i].reshape((i, Z, U)).transpose(2, 1, 0).copy(),
               "muC": bufs["muC"][:i].reshape((i, Z, U)).transpose(2, 1, 0).copy(),
               "loglik": bufs["loglik"][:i].copy(), "rhoG": bufs["rhoG"].T if posteriors else None,
               "rhoZ": bufs["rhoZ"].T if posteriors else None,
               "txnExpLen": txnExpLen, "txnZstrength": txnZstrength, "useOutlierState": useOutlierState,
               "genotypeParams": g, "ploidyParams": pP, "normalParams": nP, "cellPrevParams": sP,
               "symmetric": g.get("symmetric", True), "cd_sweeps": sweeps[:i].copy()}
        if return_timing:
            res["timing"] = dict(solver.timing(), estep_ms=out.estep_ms, mstep_ms=out.mstep_ms, reduce_ms=out.reduce_ms)
        return res
    finally:
        if own:
            solver.close()


def viterbiClonalCN(data, convergeParams, genotypeParams=None, solver=None):
    """R/hmmClonal.R:370-485 -> integer vector length N, 1-based joint state index.
    Emission parameters from the last column i, initial distribution from column i-1 (:402,:466-467)."""
    cp = convergeParams
    if genotypeParams is None:
        if cp.get("genotypeParams"):
            genotypeParams = cp["genotypeParams"]
        else:
            raise ValueError("convergeParams does not contain list element: genotypeParams. \n"
                             "                 Please specify one to viterbiClonalCN() or add the element to \n"
                             "                 convergeParams.")
    if cp.get("txnExpLen") is None or cp.get("txnZstrength") is None or cp.get("useOutlierState") is None:
        raise ValueError("convergeParams does not contain list elements: txnExpLen, \n"
                         "             txnZstrength and/or useOutlierState. ")
    if cp["useOutlierState"]:
        raise TitanError(_lib.ERR_UNSUPPORTED, "useOutlierState=TRUE is not covered (no shipped driver enables it)")
    s = np.asarray(cp["s"], float).reshape(np.asarray(cp["s"]).shape[0], -1)
    Z, i = s.shape
    i -= 1
    var = np.asarray(cp["var"], float)
    own = solver is None
    if own:
        solver = Solver(data, genotypeParams, Z, float(np.ravel(cp["txnExpLen"])[0]),
                        float(np.ravel(cp["txnZstrength"])[0]))
    try:
        varR = corRho = None
        if solver.gauss:
            # R/hmmClonal.R:439-447 hands computeBivariateNormalObslik `convergeParams$corRho_0`, an element the
            # convergeParams list does not have (:344-367): NULL, so the correlation is re-derived from the data by
            # getBivariateCovariance (:536-538,561-567) as covar[1,2] / sqrt(covar[1,1] * covar[2,2])
            varR = np.asarray(cp["varR"], float)[:, i]
            x1, x2 = np.asarray(data["ref"], float) / np.asarray(data["tumDepth"], float), np.asarray(data["logR"], float)
            ok = ~(np.isnan(x1) | np.isnan(x2))
            d1, d2 = x1[ok] - x1[ok].mean(), x2[ok] - x2[ok].mean()
            nm1 = d1.size - 1
            c11, c22, c12 = float(np.dot(d1, d1)) / nm1, float(np.dot(d2, d2)) / nm1, float(np.dot(d1, d2)) / nm1
            corRho = c12 / np.sqrt(c11 * c22)
        return solver.viterbi(np.ravel(cp["n"])[i], s[:, i], np.ravel(cp["phi"])[i], var[:, i],
                              np.asarray(cp["piG"], float)[:, i - 1], np.asarray(cp["piZ"], float)[:, i - 1],
                              genotypeParams, varR=varR, corRho=corRho)
    finally:
        if own:
            solver.close()


def fwd_back_clonalCN(log_piGiZi, py, copyNumKey, zygosityKey, numClust, positions, zStrength, txnLen, useOutlier=0):
    """.Call("fwd_backC_clonalCN", ...) through the C ABI -> dict(rho = K x T log posterior, loglik)."""
    py = np.asarray(py, dtype=np.float64)
    if py.ndim != 2:
        raise ValueError("py must be a K x T matrix")
    K, T = py.shape
    lp, lpp = f64(log_piGiZi)
    pyf, pyp = f64(py.T)                      # [T][K] == column-major K x T
    ct, ctp = f64(copyNumKey)
    zs, zsp = f64(zygosityKey)
    ps, psp = f64(positions)
    if lp.size != K or ps.size != T:
        raise TitanError(_lib.ERR_ARG, f"fwd_backC_clonalCN: The obslik must be {lp.size}-by-{ps.size} dimension.")
    rho = np.zeros((T, K))
    ll = C.c_double(0)
    check(lib().titan_fwd_back_clonalCN(lpp, K, pyp, T, ctp, ct.size, zsp, zs.size, int(numClust), psp,
                                        float(zStrength), float(txnLen), int(useOutlier),
                                        rho.ctypes.data_as(c_double_p), C.byref(ll)))
    return {"rho": rho.T, "loglik": ll.value}


def viterbi_clonalCN(log_piGiZi, py, copyNumKey, zygosityKey, numClust, positions, zStrength, txnLen, useOutlier=0):
    """.Call("viterbiC_clonalCN", ...) through the C ABI -> int32 path[T], 1-based."""
    py = np.asarray(py, dtype=np.float64)
    if py.ndim != 2:
        raise ValueError("py must be a K x T matrix")
    K, T = py.shape
    lp, lpp = f64(log_piGiZi)
    pyf, pyp = f64(py.T)
    ct, ctp = f64(copyNumKey)
    zs, zsp = f64(zygosityKey)
    ps, psp = f64(positions)
    if lp.size != K or ps.size != T:
        raise TitanError(_lib.ERR_ARG, f"viterbiC_clonalCN: The obslik must be {lp.size}-by-{ps.size} dimension.")
    path = np.zeros(T, dtype=np.int32)
    check(lib().titan_viterbi_clonalCN(lpp, K, pyp, T, ctp, ct.size, zsp, zs.size, int(numClust), psp,
                                       float(zStrength), float(txnLen), int(useOutlier),
                                       path.ctypes.data_as(C.POINTER(C.c_int))))
    return path
