"""ctypes binding of libbmpc.so -- the same C ABI (include/bmpc.h) a Julia host would `ccall`.

There is deliberately no fallback: if the CUDA library is missing or no CUDA device is
present, the first use raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# BMPC_LIB: another build of the same library (instrumented study builds, tools/studies); never a CPU implementation
LIB_PATH = os.environ.get("BMPC_LIB") or os.path.join(HERE, "libbmpc.so")

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
STATUS_OPTIMAL, STATUS_ITERATION_LIMIT, STATUS_INFEASIBLE = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Dims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("N", "nu", "ny", "nd", "nxhat", "Hp", "Hc", "neps", "shared_model",
                                         "max_iter", "device", "team")] + [("tol", C.c_double)]


class Softness(C.Structure):
    _fields_ = [(k, c_double_p) for k in ("C_umin", "C_umax", "C_dumin", "C_dumax", "C_ymin", "C_ymax",
                                          "c_xmin", "c_xmax")]


class StepIO(C.Structure):
    _fields_ = ([("xhat0", C.c_void_p), ("lastu0", C.c_void_p), ("ry", C.c_void_p), ("Rhat_y", C.c_void_p),
                 ("Rhat_u", C.c_void_p), ("d0", C.c_void_p), ("Dhat0", C.c_void_p), ("Ztilde", C.c_void_p),
                 ("u", C.c_void_p), ("J", C.c_void_p), ("status", C.c_void_p), ("iters", C.c_void_p),
                 ("device_ptrs", C.c_int32), ("sync", C.c_int32), ("resident", C.c_int32), ("host_mapped", C.c_int32),
                 ("y0m", C.c_void_p), ("Yhat_s", C.c_void_p), ("kkt", C.c_void_p)])


class MheDims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("N", "nu", "nym", "nd", "nxhat", "He", "neps", "direct", "shared_model",
                                         "max_iter", "device", "reserved")] + [("tol", C.c_double)]


class Info(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("Yhat0", "U0", "xhat0end", "F", "qtilde", "r")]


class BmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbmpc error {code}: {msg}")
        self.code = code


_lib = None

# every symbol include/bmpc.h declares (checked by tests/test_abi.py)
SYMBOLS = ["bmpc_last_error", "bmpc_version", "bmpc_create", "bmpc_destroy", "bmpc_set_stream", "bmpc_set_model",
           "bmpc_set_predmat", "bmpc_set_weights", "bmpc_set_oppoints", "bmpc_set_constraints", "bmpc_step",
           "bmpc_getinfo", "bmpc_set_estimator", "bmpc_set_estimator_cov", "bmpc_get_cov", "bmpc_set_state", "bmpc_get_state", "bmpc_set_gather", "bmpc_launch_info", "bmpc_launch_count",
           "bmhe_create", "bmhe_destroy", "bmhe_set_predmat", "bmhe_set_cov", "bmhe_set_constraints", "bmhe_reset",
           "bmhe_correct", "bmhe_update", "bmhe_update_solve", "bmhe_set_stream", "bmhe_launch_count",
           "bmpc_set_gather_pull", "bmpc_gather_epoch", "bmpc_gather_pull", "bmpc_gather_timed_out",
           "bmpc_set_custom", "bmpc_set_custom_bounds", "bmpc_get_states", "bmpc_set_weights_dense"]


def lib():
    """Load libbmpc.so (raises if it was not built: there is no Python/CPU implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python -m __graft_entry__` or "
            "`python modelpredictivecontrol.jl_b200/build.py` (nvcc, sm_100a). No CPU fallback exists.")
    L = C.CDLL(LIB_PATH)
    L.bmpc_last_error.restype = C.c_char_p
    L.bmpc_version.restype = C.c_int
    L.bmpc_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Dims), c_int32_p]
    L.bmpc_destroy.argtypes = [C.c_void_p]
    L.bmpc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.bmpc_set_model.argtypes = [C.c_void_p] + [c_double_p] * 9 + [C.c_double]
    L.bmpc_set_predmat.argtypes = [C.c_void_p] + [c_double_p] * 13
    L.bmpc_set_weights.argtypes = [C.c_void_p, c_double_p, C.c_int32, c_double_p]
    L.bmpc_set_weights_dense.argtypes = [C.c_void_p, c_double_p, C.c_int32, c_double_p, C.c_int32]
    L.bmpc_set_oppoints.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.bmpc_set_constraints.argtypes = [C.c_void_p] + [c_double_p] * 8 + [C.POINTER(Softness)]
    L.bmpc_step.argtypes = [C.c_void_p, C.POINTER(StepIO)]
    L.bmpc_getinfo.argtypes = [C.c_void_p, C.POINTER(Info)]
    L.bmpc_launch_info.argtypes = [C.c_void_p, c_int32_p]
    L.bmpc_set_estimator.argtypes = [C.c_void_p] + [c_double_p] * 7 + [C.c_int32]
    L.bmpc_set_estimator_cov.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
    L.bmpc_get_cov.argtypes = [C.c_void_p, c_double_p]
    L.bmpc_set_state.argtypes = [C.c_void_p, c_double_p]
    L.bmpc_get_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.bmpc_set_gather.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]
    L.bmpc_launch_count.argtypes = [C.c_void_p]
    L.bmpc_launch_count.restype = C.c_int64
    L.bmhe_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(MheDims)]
    L.bmhe_destroy.argtypes = [C.c_void_p]
    L.bmhe_set_predmat.argtypes = [C.c_void_p] + [c_double_p] * 8
    L.bmhe_set_cov.argtypes = [C.c_void_p] + [c_double_p] * 5 + [C.c_double]
    L.bmhe_set_constraints.argtypes = [C.c_void_p] + [c_double_p] * 9
    L.bmhe_reset.argtypes = [C.c_void_p]
    L.bmhe_correct.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p,
                               c_int32_p, c_double_p, c_double_p]
    L.bmhe_update.argtypes = [C.c_void_p, c_double_p]
    L.bmhe_update_solve.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                    c_int32_p, c_int32_p, c_double_p, c_double_p]
    L.bmpc_get_states.argtypes = [C.c_void_p, c_double_p]
    L.bmpc_set_custom.argtypes = [C.c_void_p, C.c_int32] + [c_double_p] * 7
    L.bmpc_set_custom_bounds.argtypes = [C.c_void_p] + [c_double_p] * 4
    L.bmhe_set_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.bmpc_set_gather_pull.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                       c_int32_p, C.c_int32]
    L.bmpc_gather_epoch.argtypes = [C.c_void_p]
    L.bmpc_gather_epoch.restype = C.c_int64
    L.bmpc_gather_pull.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.bmpc_gather_timed_out.argtypes = [C.c_void_p]
    L.bmhe_launch_count.argtypes = [C.c_void_p]
    L.bmhe_launch_count.restype = C.c_int64
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise BmpcError(rc, lib().bmpc_last_error().decode())


def dptr(a):
    """double* of a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def colmajor(a):
    """(N, rows, cols) numpy batch -> instance-major buffer of COLUMN-major matrices (Julia layout)."""
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))
