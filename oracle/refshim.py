"""ctypes driver for shared objects compiled against oracle/shim/R.h.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
--impl reference legs).  Nothing under titancna_b200/ imports this module.

It plays the role of R's `.Call("name", ...)`: the library's `R_init_TitanCNA`
registers its routine table (reference: src/register.c:8-18), routines are then
looked up *by string* exactly as R/hmmClonal.R:209,478 do, and invoked through a
setjmp trampoline so that `error()` (Rf_error) comes back as a Python exception.

The same driver loads either
  * oracle/_ref/libtitanref.so  - the reference's unmodified src/*.c, or
  * titancna_b200/csrc/TitanCNA_shim.so - this repo's drop-in `.Call` glue,
which is how the drop-in claim is tested without an R installation.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP = 13, 14, 16, 19, 22

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libtitanref.so")


class _Sexp(C.Structure):
    pass


_Sexp._fields_ = [
    ("type", C.c_int),
    ("length", C.c_int),
    ("data", C.c_void_p),
    ("dim", C.POINTER(_Sexp)),
    ("names", C.POINTER(_Sexp)),
]
SEXP = C.POINTER(_Sexp)


class RError(RuntimeError):
    """Raised when the callee invoked error()/Rf_error()."""


class ShimLibrary:
    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.shim_wrap.restype = SEXP
        L.shim_wrap.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.shim_call9.restype = SEXP
        L.shim_call9.argtypes = [C.c_void_p] + [SEXP] * 9
        L.shim_call3.restype = SEXP
        L.shim_call3.argtypes = [C.c_void_p] + [SEXP] * 3
        L.shim_last_error.restype = C.c_char_p
        L.shim_lookup.restype = C.c_void_p
        L.shim_lookup.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.shim_free_all.restype = None
        if hasattr(L, "shim_calln"):      # glue builds carry the generic trampoline + external pointers
            L.shim_calln.restype = SEXP
            L.shim_calln.argtypes = [C.c_void_p, C.c_int, C.POINTER(SEXP)]
            L.shim_named_list.restype = SEXP
            L.shim_named_list.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(SEXP)]
            L.shim_run_finalizers.restype = C.c_int
        L.R_init_TitanCNA.restype = None
        L.R_init_TitanCNA.argtypes = [C.c_void_p]
        L.R_init_TitanCNA(None)  # library.dynam -> R_init_<pkg>

    # -- argument marshalling -------------------------------------------------
    def _real(self, x, keep, dim=None):
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel(order="F"))
        keep.append(a)
        nr, nc = (dim if dim is not None else (-1, -1))
        return self.lib.shim_wrap(REALSXP, a.size, a.ctypes.data_as(C.c_void_p), nr, nc)

    def routine(self, name: str):
        n = C.c_int(-1)
        fn = self.lib.shim_lookup(name.encode(), C.byref(n))
        if not fn:
            raise KeyError(f"{name} is not a registered .Call routine of {self.path}")
        return fn, n.value

    # -- generic .Call (any registered routine; arguments are SEXPs built by _real / named_list) ----
    def call(self, name, *args):
        fn, nargs = self.routine(name)
        if nargs != len(args):
            raise TypeError(f".Call({name!r}) takes {nargs} arguments, got {len(args)}")
        arr = (SEXP * len(args))(*args)
        out = self.lib.shim_calln(fn, len(args), arr)
        if not out:
            msg = self.lib.shim_last_error().decode(errors="replace")
            self.lib.shim_free_all()
            raise RError(msg)
        return out

    def named_list(self, items, keep):
        """dict name -> numeric vector/scalar  ->  VECSXP with a names attribute"""
        names = (C.c_char_p * len(items))(*[k.encode() for k in items])
        elts = (SEXP * len(items))(*[self._real(np.atleast_1d(v), keep) for v in items.values()])
        keep.append((names, elts))
        return self.lib.shim_named_list(len(items), names, elts)

    @staticmethod
    def list_to_dict(out):
        """named VECSXP of REALSXPs -> dict of numpy arrays (matrices in R's column-major shape)"""
        o = out.contents
        assert o.type == VECSXP
        elts = C.cast(o.data, C.POINTER(SEXP))
        names = C.cast(o.names.contents.data, C.POINTER(SEXP))
        res = {}
        for i in range(o.length):
            nm = C.cast(names[i].contents.data, C.c_char_p).value.decode()
            e = elts[i].contents
            a = np.ctypeslib.as_array(C.cast(e.data, C.POINTER(C.c_double)), shape=(e.length,)).copy()
            if e.dim:
                d = C.cast(e.dim.contents.data, C.POINTER(C.c_int))
                a = a.reshape((int(d[0]), int(d[1])), order="F")
            res[nm] = a
        return res

    def gc(self):
        """run the finalizers of every external pointer made so far (R's garbage collection of dropped solvers)"""
        return self.lib.shim_run_finalizers()

    def registered(self, names=("getPositionOverlapC", "fwd_backC_clonalCN", "viterbiC_clonalCN")):
        out = {}
        for nm in names:
            try:
                out[nm] = self.routine(nm)[1]
            except KeyError:
                pass
        return out

    def _call9(self, name, log_pi, py, ct, zs, Z, posn, zstrength, txnlen, outlier):
        fn, nargs = self.routine(name)
        assert nargs == 9
        keep = []
        py = np.asarray(py, dtype=np.float64)
        if py.ndim != 2:
            raise ValueError("py must be a K x T matrix")
        args = [
            self._real(log_pi, keep),
            self._real(py, keep, dim=py.shape),
            self._real(ct, keep),
            self._real(zs, keep),
            self._real([Z], keep),
            self._real(posn, keep),
            self._real(np.atleast_1d(zstrength), keep),
            self._real(np.atleast_1d(txnlen), keep),
            self._real([outlier], keep),
        ]
        out = self.lib.shim_call9(fn, *args)
        if not out:
            msg = self.lib.shim_last_error().decode(errors="replace")
            self.lib.shim_free_all()
            raise RError(msg)
        return out, keep

    # -- the three registered routines -----------------------------------------
    def fwd_back(self, log_pi, py, ct, zs, Z, posn, zstrength, txnlen, outlier=0):
        """.Call("fwd_backC_clonalCN", ...) -> dict(rho=K x T log posterior, loglik)."""
        out, keep = self._call9("fwd_backC_clonalCN", log_pi, py, ct, zs, Z, posn,
                                zstrength, txnlen, outlier)
        try:
            o = out.contents
            assert o.type == VECSXP and o.length == 2
            elts = C.cast(o.data, C.POINTER(SEXP))
            names = C.cast(o.names.contents.data, C.POINTER(SEXP))
            got = [C.cast(names[i].contents.data, C.c_char_p).value.decode() for i in range(2)]
            assert got == ["rho", "loglik"], got
            g = elts[0].contents
            dims = C.cast(g.dim.contents.data, C.POINTER(C.c_int))
            K, T = int(dims[0]), int(dims[1])
            rho = np.ctypeslib.as_array(C.cast(g.data, C.POINTER(C.c_double)), shape=(K * T,)).copy()
            ll = C.cast(elts[1].contents.data, C.POINTER(C.c_double))[0]
            return {"rho": rho.reshape((K, T), order="F"), "loglik": float(ll)}
        finally:
            self.lib.shim_free_all()

    def viterbi(self, log_pi, py, ct, zs, Z, posn, zstrength, txnlen, outlier=0):
        """.Call("viterbiC_clonalCN", ...) -> int32 path[T], 1-based."""
        out, keep = self._call9("viterbiC_clonalCN", log_pi, py, ct, zs, Z, posn,
                                zstrength, txnlen, outlier)
        try:
            o = out.contents
            assert o.type == INTSXP
            return np.ctypeslib.as_array(C.cast(o.data, C.POINTER(C.c_int)), shape=(o.length,)).copy()
        finally:
            self.lib.shim_free_all()

    def position_overlap(self, posn, start, stop):
        fn, nargs = self.routine("getPositionOverlapC")
        assert nargs == 3
        keep = []
        out = self.lib.shim_call3(fn, self._real(posn, keep), self._real(start, keep),
                                  self._real(stop, keep))
        try:
            if not out:
                raise RError(self.lib.shim_last_error().decode(errors="replace"))
            o = out.contents
            return np.ctypeslib.as_array(C.cast(o.data, C.POINTER(C.c_double)), shape=(o.length,)).copy()
        finally:
            self.lib.shim_free_all()

    def transition_matrix(self, K, U, zs, rhoG, rhoZ, outlier=0):
        """preparePositionSpecificMatrix (fwd_backC_clonalCN.c:215-301) -> K x K, A[i,j]=P(i->j).
        Only available on the reference build (non-static helper)."""
        f = self.lib.preparePositionSpecificMatrix
        f.restype = None
        f.argtypes = [C.POINTER(C.c_double), C.c_uint, C.c_uint, C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_double, C.c_double, C.c_uint, C.c_uint]
        A = np.zeros(K * K)
        ct = np.zeros(U)
        zs = np.ascontiguousarray(zs, dtype=np.float64)
        f(A.ctypes.data_as(C.POINTER(C.c_double)), K, U, ct.ctypes.data_as(C.POINTER(C.c_double)),
          zs.ctypes.data_as(C.POINTER(C.c_double)), rhoG, rhoZ, outlier, 0)
        return A.reshape((K, K), order="F")


_ref = None


def reference() -> ShimLibrary:
    """The reference's own C, built by oracle/Makefile into oracle/_ref/."""
    global _ref
    if _ref is None:
        _ref = ShimLibrary(REF_SO)
    return _ref


def have_reference() -> bool:
    return os.path.exists(REF_SO)
