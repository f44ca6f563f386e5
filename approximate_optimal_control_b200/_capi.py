"""ctypes binding of libb200pmp.so (C ABI declared in include/b200pmp.h).

There is NO CPU fallback: if the library is missing and cannot be built, or no CUDA device is
visible when a compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

# ---- enums (include/b200pmp.h) ---------------------------------------------------------------------
SYS_FLATQUAD, SYS_ORBITS, SYS_LQR_STATEPENALTY, SYS_GENERATED = 0, 1, 2, 3
METHOD_RK4, METHOD_TSIT5, METHOD_DOPRI5 = 0, 1, 2
F32, F64 = 0, 1
INDEP_TIME, INDEP_VALUE = 0, 1
EVENT_NONE, EVENT_VALUE_GE, EVENT_VXX_NORM, EVENT_LQR_VALUE_LE, EVENT_NN_VALUE_UCB_LE = 0, 1, 2, 3, 4
RESULT_OK, RESULT_EVENT, RESULT_MAX_STEPS, RESULT_DT_MIN, RESULT_NAN = 0, 1, 2, 3, 4
ABI_VERSION = 5
FLAG_USTAR_LITERAL = 1
FLAG_VXX = 2
FLAG_VXX_SYMMETRIC = 4
FLAG_USTAR_SMOOTH = 8

METHODS = {"rk4": METHOD_RK4, "tsit5": METHOD_TSIT5, "dopri5": METHOD_DOPRI5}


class Problem(C.Structure):
    _fields_ = [("system_id", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
                ("u_lo", C.c_double * 2), ("u_hi", C.c_double * 2), ("params", C.c_double * 16)]
    # python-side only: the library that serves this problem (None = libb200pmp.so; a codegen plug-in otherwise)
    lib_path = None


class Opts(C.Structure):
    _fields_ = [("method", C.c_int32), ("dtype", C.c_int32), ("indep", C.c_int32),
                ("adaptive", C.c_int32), ("max_steps", C.c_int32), ("force_dtmin", C.c_int32),
                ("event", C.c_int32), ("save_dense", C.c_int32),
                ("t0", C.c_double), ("t1", C.c_double), ("dt0", C.c_double),
                ("rtol", C.c_double), ("atol", C.c_double), ("dtmin", C.c_double),
                ("dtmax", C.c_double), ("safety", C.c_double), ("factormin", C.c_double),
                ("factormax", C.c_double), ("event_level", C.c_double),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class Dense(C.Structure):
    _fields_ = [("ts", C.c_void_p), ("x", C.c_void_p), ("t", C.c_void_p), ("v", C.c_void_p),
                ("vx", C.c_void_p), ("vxx", C.c_void_p)]


class NnCostate(C.Structure):
    _fields_ = [("n_nets", C.c_int32), ("n_hidden_layers", C.c_int32), ("width", C.c_int32), ("reserved", C.c_int32),
                ("weights", C.c_void_p), ("x_mean", C.c_double * 16), ("x_std", C.c_double * 16),
                ("v_mean", C.c_double), ("v_std", C.c_double), ("n_sigma", C.c_double), ("sigma_max", C.c_double)]


EXPORTS = ("pmp_abi_version", "pmp_last_error", "pmp_state_rows", "pmp_state_rows_flags", "pmp_device_count",
           "pmp_solve_workspace_bytes", "pmp_solve_batch_cuda", "pmp_solve_batch_push_cuda", "pmp_solve_jvp_cuda", "pmp_solve_launch_count", "pmp_solve_host", "pmp_solve_host_lqr", "pmp_solve_host_plan", "pmp_host_release", "pmp_host_probe_copy", "pmp_select_count_cuda", "pmp_select_write_cuda", "pmp_comm_unique_id", "pmp_comm_init", "pmp_comm_destroy", "pmp_allgather_endpoints",
           "pmp_rhs_batch_cuda", "pmp_rhs_batch_cuda_flags", "pmp_ustar_batch_cuda", "pmp_ustar_batch_cuda_flags", "pmp_interp_batch_cuda", "pmp_forward_batch_cuda", "pmp_forward_nn_batch_cuda", "pmp_nn_value_batch_cuda",
           "pmp_fma_peak_cuda")

_lib = None
_libs = {}


class PmpError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (building first if stale/missing) libb200pmp.so.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("B200PMP_LIB") or _build.LIB_PATH  # B200PMP_LIB: a tuning variant built by scripts/build_variant.py
    if os.environ.get("B200PMP_LIB") is None and (os.environ.get("B200PMP_REBUILD") or _build._stale()):
        # missing, or built from other sources than the ones in the tree (content hash): rebuild rather than run old kernels
        path = _build.build(force=bool(os.environ.get("B200PMP_REBUILD")))
    _lib = load(path)
    return _lib


def lib_of(problem: "Problem") -> C.CDLL:
    """The library that serves `problem`: libb200pmp.so for the compiled systems, the codegen plug-in that was
    built for it otherwise (codegen.plugin_problem sets problem.lib_path)."""
    path = getattr(problem, "lib_path", None)
    return lib() if path is None else load(path)


def load(path: str) -> C.CDLL:
    """dlopen a library exporting the C ABI of include/b200pmp.h and declare its signatures."""
    if path in _libs:
        return _libs[path]
    try:
        L = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise PmpError(f"cannot load {path}: {e}. There is no CPU fallback for the PMP path.") from e
    L.pmp_last_error.restype = C.c_char_p
    L.pmp_state_rows.argtypes = [C.POINTER(Problem)]
    L.pmp_solve_workspace_bytes.restype = C.c_int64
    L.pmp_solve_workspace_bytes.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64]
    L.pmp_solve_launch_count.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64]
    L.pmp_solve_batch_cuda.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64, C.c_void_p,
                                       C.c_void_p, C.POINTER(Dense), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.pmp_solve_batch_push_cuda.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_int32,
                                            C.c_void_p, C.c_void_p]
    L.pmp_comm_unique_id.argtypes = [C.c_void_p]
    L.pmp_comm_init.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p]
    L.pmp_comm_destroy.argtypes = [C.c_void_p]
    L.pmp_allgather_endpoints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    L.pmp_solve_jvp_cuda.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64] + [C.c_void_p] * 8
    L.pmp_solve_host.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64] + [C.c_void_p] * 11 + [C.c_int32]
    if hasattr(L, "pmp_solve_host_lqr"):
        L.pmp_solve_host_lqr.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64] + [C.c_void_p] * 11 + [C.c_int32]
    L.pmp_solve_host_plan.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64, C.c_int32, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]
    L.pmp_host_release.argtypes = []
    if True:
        L.pmp_select_count_cuda.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                            C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.pmp_select_write_cuda.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 8
    if hasattr(L, "pmp_host_probe_copy"):
        L.pmp_host_probe_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    L.pmp_rhs_batch_cuda.argtypes = [C.POINTER(Problem), C.c_int32, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    L.pmp_state_rows_flags.argtypes = [C.POINTER(Problem), C.c_int32]
    L.pmp_rhs_batch_cuda_flags.argtypes = [C.POINTER(Problem), C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
    L.pmp_ustar_batch_cuda.argtypes = [C.POINTER(Problem), C.c_int32, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    L.pmp_ustar_batch_cuda_flags.argtypes = [C.POINTER(Problem), C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    L.pmp_interp_batch_cuda.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.pmp_forward_batch_cuda.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
    L.pmp_forward_nn_batch_cuda.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.POINTER(NnCostate), C.c_int64] + [C.c_void_p] * 9
    L.pmp_nn_value_batch_cuda.argtypes = [C.POINTER(Problem), C.POINTER(NnCostate), C.c_int32, C.c_int64] + [C.c_void_p] * 5
    L.pmp_fma_peak_cuda.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.POINTER(C.c_double), C.c_void_p]
    if L.pmp_abi_version() != ABI_VERSION:
        raise PmpError(f"{path}: ABI version {L.pmp_abi_version()} != {ABI_VERSION}; rebuild")
    _libs[path] = L
    return L


def check(rc: int, L: C.CDLL = None) -> int:
    if rc < 0:
        msg = (L or lib()).pmp_last_error().decode()
        if L is None and not msg:  # the failing call may have gone to a plug-in library
            msg = "; ".join(m for m in (l.pmp_last_error().decode() for l in _libs.values()) if m)
        raise PmpError(f"libb200pmp error {rc}: {msg}")
    return rc


def make_opts(*, method=METHOD_TSIT5, dtype=F32, indep=INDEP_TIME, adaptive=1, max_steps=128,
              force_dtmin=1, event=EVENT_NONE, save_dense=0, t0=0.0, t1=-3.0, dt0=-0.1, rtol=1e-5,
              atol=1e-5, dtmin=0.0005, dtmax=0.5, safety=0.9, factormin=0.2, factormax=10.0,
              event_level=0.0, flags=0) -> Opts:
    """Defaults = the reference's flatquad solve (pontryagin_utils.py:496-525,
    flatquad_landing_experiment.py:368-382); controller constants are diffrax's PIDController defaults."""
    if isinstance(method, str):
        method = METHODS[method.lower()]
    return Opts(int(method), int(dtype), int(indep), int(adaptive), int(max_steps), int(force_dtmin),
                int(event), int(save_dense), float(t0), float(t1), float(dt0), float(rtol),
                float(atol), float(dtmin), float(dtmax), float(safety), float(factormin),
                float(factormax), float(event_level), int(flags), 0)
