"""ctypes binding of include/titancna_b200.h (the C ABI the R glue binds as well).

The library is loaded from titancna_b200/csrc/libtitancna_b200.so.  There is no Python or CPU
implementation behind it: if the shared object is missing, or no CUDA device is usable, the calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libtitancna_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)


class TitanError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[titancna_b200 error {code}] {msg}")
        self.code = code
        self.msg = msg


ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NUMERIC = 1, 2, 3, 4, 5


class DataDesc(C.Structure):
    _fields_ = [("N", C.c_int64), ("nChr", C.c_int32), ("chr_start", c_i64_p), ("posn", c_double_p),
                ("ref", c_double_p), ("tumDepth", c_double_p), ("logR", c_double_p), ("txnExpLen", C.c_double),
                ("txnZstrength", C.c_double), ("U", C.c_int32), ("Z", C.c_int32), ("ZS", c_double_p),
                ("alleleModel", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("n", C.c_double), ("phi", C.c_double), ("s", c_double_p), ("var", c_double_p),
                ("piG", c_double_p), ("piZ", c_double_p), ("rt", c_double_p), ("ct", c_double_p),
                ("rn", C.c_double), ("varR", c_double_p), ("corRho", C.c_double)]


class EstepOut(C.Structure):
    _fields_ = [("loglik_chr", c_double_p), ("stats", c_double_p), ("rhoG0", c_double_p), ("rhoZ0", c_double_p),
                ("rhoG", c_double_p), ("rhoZ", c_double_p), ("muR", c_double_p), ("muC", c_double_p)]


class Hyper(C.Structure):
    _fields_ = [("rt", c_double_p), ("ct", c_double_p), ("ZS", c_double_p), ("rn", C.c_double),
                ("var_0", c_double_p), ("alphaKHyper", c_double_p), ("betaKHyper", c_double_p),
                ("kappaGHyper", c_double_p), ("piG_0", c_double_p),
                ("s_0", c_double_p), ("alphaSHyper", c_double_p), ("betaSHyper", c_double_p),
                ("kappaZHyper", c_double_p), ("piZ_0", c_double_p),
                ("n_0", C.c_double), ("alphaNHyper", C.c_double), ("betaNHyper", C.c_double),
                ("phi_0", C.c_double), ("alphaPHyper", C.c_double), ("betaPHyper", C.c_double),
                ("varR_0", c_double_p), ("alphaRHyper", c_double_p), ("betaRHyper", c_double_p),
                ("corRho_0", C.c_double)]


class EmOpts(C.Structure):
    _fields_ = [("maxiter", C.c_int32), ("maxiterUpdate", C.c_int32), ("normal_map", C.c_int32),
                ("estimateS", C.c_int32), ("estimatePloidy", C.c_int32), ("likChangeThreshold", C.c_double),
                ("want_posteriors", C.c_int32), ("verbose", C.c_int32)]


class EmOut(C.Structure):
    _fields_ = [("iters", C.c_int32), ("n", c_double_p), ("phi", c_double_p), ("loglik", c_double_p),
                ("s", c_double_p), ("piZ", c_double_p), ("var", c_double_p), ("piG", c_double_p),
                ("muR", c_double_p), ("muC", c_double_p), ("rhoG", c_double_p), ("rhoZ", c_double_p),
                ("cd_sweeps", C.POINTER(C.c_int32)), ("estep_ms", C.c_double), ("mstep_ms", C.c_double),
                ("reduce_ms", C.c_double), ("varR", c_double_p)]


class Timing(C.Structure):
    _fields_ = [("last_estep_ms", C.c_double), ("last_fwd_ms", C.c_double), ("last_bwd_ms", C.c_double),
                ("last_reduce_ms", C.c_double), ("last_fwd_r0_ms", C.c_double), ("last_bwd_r0_ms", C.c_double),
                ("last_viterbi_ms", C.c_double),
                ("last_fwd_rounds", C.c_int32), ("last_bwd_rounds", C.c_int32),
                ("last_fwd_chunk_runs", C.c_int64), ("last_bwd_chunk_runs", C.c_int64),
                ("kernel_launches", C.c_int64), ("n_chunks", C.c_int32), ("engine", C.c_int32),
                ("last_emission_ms", C.c_double), ("last_rounds_ms", C.c_double)]


EXPORTS = [
    "titan_last_error", "titan_version", "titan_device_count", "titan_set_device", "titan_get_device",
    "titan_fwd_back_clonalCN", "titan_viterbi_clonalCN", "titan_get_position_overlap",
    "titan_solver_create", "titan_solver_destroy", "titan_solver_configure", "titan_solver_set_stream",
    "titan_estep", "titan_viterbi", "titan_em_run", "titan_mstep", "titan_mstep_gaussian", "titan_solver_timing",
    "titan_release_cached_memory", "titan_measure_fp64_peak",
    "titan_solver_shape", "titan_comm_unique_id", "titan_comm_create", "titan_comm_destroy", "titan_comm_info", "titan_solver_set_comm",
]

_lib = None


def lib():
    """Load the shared object (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TitanError(ERR_NO_DEVICE, f"{LIB_PATH} is missing: run `python -m titancna_b200.build` "
                                        f"(the CUDA extension is mandatory; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.titan_last_error.restype = C.c_char_p
    L.titan_version.restype = C.c_char_p
    L.titan_device_count.argtypes = [c_int_p]
    L.titan_set_device.argtypes = [C.c_int]
    L.titan_get_device.argtypes = [c_int_p]
    legacy = [c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, C.c_int,
              c_double_p, C.c_double, C.c_double, C.c_int]
    L.titan_fwd_back_clonalCN.argtypes = legacy + [c_double_p, c_double_p]
    L.titan_viterbi_clonalCN.argtypes = legacy + [c_int_p]
    L.titan_get_position_overlap.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p]
    L.titan_solver_create.argtypes = [C.POINTER(DataDesc), C.POINTER(C.c_void_p)]
    L.titan_solver_destroy.argtypes = [C.c_void_p]
    L.titan_solver_destroy.restype = None
    L.titan_solver_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
    L.titan_solver_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.titan_estep.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(EstepOut)]
    L.titan_viterbi.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(C.c_int32)]
    L.titan_em_run.argtypes = [C.c_void_p, C.POINTER(Hyper), C.POINTER(EmOpts), C.POINTER(EmOut)]
    L.titan_mstep.argtypes = [C.POINTER(Hyper), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              c_double_p, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p, C.c_double,
                              c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                              c_double_p, c_int_p, c_double_p]
    L.titan_mstep_gaussian.argtypes = [C.POINTER(Hyper), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       c_double_p, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p, c_double_p,
                                       C.c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                       c_double_p, c_double_p, c_double_p, c_int_p, c_double_p]
    L.titan_solver_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    L.titan_measure_fp64_peak.argtypes = [c_double_p, c_double_p]
    L.titan_solver_shape.argtypes = [C.c_void_p, c_i64_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.titan_comm_unique_id.argtypes = [C.c_char_p]
    L.titan_comm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.titan_comm_destroy.argtypes = [C.c_void_p]
    L.titan_comm_destroy.restype = None
    L.titan_comm_info.argtypes = [C.c_void_p, c_int_p, c_int_p, c_int_p]
    L.titan_solver_set_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
    for name in EXPORTS:
        if name not in ("titan_last_error", "titan_version", "titan_solver_destroy", "titan_comm_destroy"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise TitanError(rc, lib().titan_last_error().decode(errors="replace"))


def f64(a):
    """contiguous float64 copy/view + its pointer (keep the array alive while the pointer is used)"""
    arr = np.ascontiguousarray(a, dtype=np.float64)
    return arr, arr.ctypes.data_as(c_double_p)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().titan_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def current_device() -> int:
    """CUDA device of the calling thread (titan_get_device)."""
    d = C.c_int(0)
    check(lib().titan_get_device(C.byref(d)))
    return d.value


def measure_fp64_peak():
    """DFMA throughput of the current device in TFLOP/s (titan_measure_fp64_peak) and the launch's milliseconds."""
    tf, ms = C.c_double(0), C.c_double(0)
    check(lib().titan_measure_fp64_peak(C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


class Comm:
    """One NCCL communicator over the ranks that share a chromosome-sharded solution (include/titancna_b200.h section 3)."""

    ID_BYTES = 128

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(Comm.ID_BYTES)
        check(lib().titan_comm_unique_id(buf))
        return buf.raw

    def __init__(self, unique_id: bytes, rank: int, world: int):
        if len(unique_id) != Comm.ID_BYTES:
            raise ValueError("unique_id must be the 128 bytes of Comm.unique_id() drawn on rank 0")
        h = C.c_void_p()
        check(lib().titan_comm_create(C.create_string_buffer(unique_id, Comm.ID_BYTES), int(rank), int(world), C.byref(h)))
        self._h, self.rank, self.world = h, int(rank), int(world)

    def nccl_version(self) -> int:
        v = C.c_int(0)
        check(lib().titan_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None):
            lib().titan_comm_destroy(self._h)
            self._h = None

    __del__ = close
