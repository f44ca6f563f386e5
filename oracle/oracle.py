"""ctypes front end of the CPU ORACLE (oracle/libpmp_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module; nothing under
approximate_optimal_control_b200/ does.  The structs mirror include/b200pmp.h so that the oracle and
the CUDA library are driven with byte-identical problem / option records.

Batches are numpy arrays in the product's struct-of-arrays layout: y[row, i] with
rows = (x[0..nx), t, v, vx[0..nx)) -- the reference's state dict order
(/root/reference pontryagin_utils.py:454-458).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpmp_oracle.so")

SYS_FLATQUAD, SYS_ORBITS, SYS_LQR = 0, 1, 2
RK4, TSIT5, DOPRI5 = 0, 1, 2
F32, F64 = 0, 1
INDEP_TIME, INDEP_VALUE = 0, 1
EVENT_NONE, EVENT_VALUE_GE, EVENT_VXX_NORM, EVENT_LQR_VALUE_LE, EVENT_NN_VALUE_UCB_LE = 0, 1, 2, 3, 4
FLAG_USTAR_LITERAL, FLAG_VXX = 1, 2


class Problem(C.Structure):
    _fields_ = [("system_id", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
                ("u_lo", C.c_double * 2), ("u_hi", C.c_double * 2), ("params", C.c_double * 16)]


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


def build(force: bool = False) -> str:
    """Compile oracle/libpmp_oracle.so with the committed Makefile (g++, OpenMP)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f))
                for f in ("pmp_oracle.cpp", "pmp_oracle.hpp", "opcount.hpp", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE, "-B", "libpmp_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.pmp_oracle_last_error.restype = C.c_char_p
        L.pmp_oracle_solve_batch.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64,
                                             C.c_void_p, C.c_void_p, C.POINTER(Dense), C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.pmp_oracle_rhs_batch.argtypes = [C.POINTER(Problem), C.c_int32, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_int]
        L.pmp_oracle_ustar_batch.argtypes = [C.POINTER(Problem), C.c_int32, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int]
        L.pmp_oracle_interp_batch.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int32, C.c_void_p]
        L.pmp_oracle_opcount.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().pmp_oracle_num_threads())


# ---- the three systems with the reference's constants ---------------------------------------------
def flatquad_problem() -> Problem:
    """flatquad_landing_experiment.py:269-273, :313-315, :355-358."""
    m, g, r = 20.0, 9.81, 0.5
    I = m * (r / 2) ** 2
    umax = m * g * 1.2 / 2
    p = Problem(system_id=SYS_FLATQUAD, nx=6, nu=2)
    p.u_lo[:] = [0.0, 0.0]
    p.u_hi[:] = [umax, umax]
    scales = [0.3, np.deg2rad(30.0), 0.5, np.deg2rad(120.0)]
    vals = [m, g, r, I] + [1.0 / s ** 2 for s in scales]
    for i, v in enumerate(vals):
        p.params[i] = v
    return p


def orbits_problem() -> Problem:
    """experiment_orbits.py:24-40, :83-94."""
    p = Problem(system_id=SYS_ORBITS, nx=2, nu=1)
    p.u_lo[0], p.u_hi[0] = -0.2, 0.2
    p.params[0], p.params[1] = 0.1, 10.0
    return p


def lqr_problem(u_lo=-1.0, u_hi=1.0, a=0.01) -> Problem:
    """experiment_lqr_statepenalty.py:23-38, :89."""
    p = Problem(system_id=SYS_LQR, nx=2, nu=1)
    p.u_lo[0], p.u_hi[0] = u_lo, u_hi
    p.params[0], p.params[1] = a, 1.0
    return p


def make_opts(method=TSIT5, dtype=F32, indep=INDEP_TIME, adaptive=1, max_steps=128, force_dtmin=1,
              event=EVENT_NONE, save_dense=0, t0=0.0, t1=-3.0, dt0=-0.1, rtol=1e-5, atol=1e-5,
              dtmin=0.0005, dtmax=0.5, safety=0.9, factormin=0.2, factormax=10.0,
              event_level=0.0, flags=0) -> Opts:
    """Defaults = the flatquad solve (pontryagin_utils.py:496-525, flatquad_landing_experiment.py:368-382)."""
    return Opts(method, dtype, indep, adaptive, max_steps, force_dtmin, event, save_dense, t0, t1,
                dt0, rtol, atol, dtmin, dtmax, safety, factormin, factormax, event_level, flags, 0)


def _np_dtype(dtype):
    return np.float32 if dtype == F32 else np.float64


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}: {lib().pmp_oracle_last_error().decode()}")


def solve_batch(problem: Problem, opts: Opts, y_f: np.ndarray, n_threads: int = 0,
                model_pass1: bool = False):
    """y_f [NS, n] -> dict(y_end [NS,n], n_accepted, n_steps, result [, ts, x, t, v, vx])."""
    dt = _np_dtype(opts.dtype)
    y_f = np.ascontiguousarray(y_f, dtype=dt)
    ns, n = y_f.shape
    nx = problem.nx
    vxx = bool(opts.flags & FLAG_VXX)
    assert ns == 2 * nx + 2 + (nx * nx if vxx else 0)
    out = {"y_end": np.empty((ns, n), dt), "n_accepted": np.empty(n, np.int32),
           "n_steps": np.empty(n, np.int32), "result": np.empty(n, np.int32)}
    dense = None
    if opts.save_dense:
        s1 = opts.max_steps + 1
        out.update(ts=np.empty((n, s1), dt), x=np.empty((n, s1, nx), dt), t=np.empty((n, s1), dt),
                   v=np.empty((n, s1), dt), vx=np.empty((n, s1, nx), dt))
        if vxx:
            out["vxx"] = np.empty((n, s1, nx, nx), dt)
        dense = Dense(*(out[k].ctypes.data for k in ("ts", "x", "t", "v", "vx")), out["vxx"].ctypes.data if vxx else None)
    _check(lib().pmp_oracle_solve_batch(C.byref(problem), C.byref(opts), n, y_f.ctypes.data,
                                        out["y_end"].ctypes.data,
                                        C.byref(dense) if dense is not None else None,
                                        out["n_accepted"].ctypes.data, out["n_steps"].ctypes.data,
                                        out["result"].ctypes.data, int(n_threads), int(model_pass1)))
    return out


def solve_jvp(problem: Problem, opts: Opts, y_f: np.ndarray, dy_f: np.ndarray) -> dict:
    """fp64: endpoint and its directional derivative along dy_f (forward-mode dual numbers through the whole solve, the
    step control decided on primal values as jax's jvp of diffeqsolve does).  y_f, dy_f [NS, n]."""
    y_f = np.ascontiguousarray(y_f, dtype=np.float64)
    dy_f = np.ascontiguousarray(dy_f, dtype=np.float64)
    ns, n = y_f.shape
    out = {"y_end": np.empty((ns, n)), "dy_end": np.empty((ns, n)), "n_accepted": np.empty(n, np.int32),
           "n_steps": np.empty(n, np.int32), "result": np.empty(n, np.int32)}
    L = lib()
    L.pmp_oracle_solve_jvp.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64] + [C.c_void_p] * 7
    _check(L.pmp_oracle_solve_jvp(C.byref(problem), C.byref(opts), n, y_f.ctypes.data, dy_f.ctypes.data, out["y_end"].ctypes.data,
                                  out["dy_end"].ctypes.data, out["n_accepted"].ctypes.data, out["n_steps"].ctypes.data,
                                  out["result"].ctypes.data))
    return out


def rhs_batch(problem: Problem, dtype: int, y: np.ndarray, model_pass1: bool = False) -> np.ndarray:
    dt = _np_dtype(dtype)
    y = np.ascontiguousarray(y, dtype=dt)
    dy = np.empty_like(y)
    _check(lib().pmp_oracle_rhs_batch(C.byref(problem), dtype, y.shape[1], y.ctypes.data,
                                      dy.ctypes.data, int(model_pass1)))
    return dy


def ustar_batch(problem: Problem, dtype: int, x: np.ndarray, lam: np.ndarray,
                model_pass1: bool = False) -> np.ndarray:
    dt = _np_dtype(dtype)
    x = np.ascontiguousarray(x, dtype=dt)
    lam = np.ascontiguousarray(lam, dtype=dt)
    u = np.empty((problem.nu, x.shape[1]), dt)
    _check(lib().pmp_oracle_ustar_batch(C.byref(problem), dtype, x.shape[1], x.ctypes.data,
                                        lam.ctypes.data, u.ctypes.data, int(model_pass1)))
    return u


def interp_batch(problem: Problem, opts: Opts, y0: np.ndarray, t1: np.ndarray, target: np.ndarray, mode: int = 0) -> np.ndarray:
    """Dense output on one step per query (diffrax Solution.evaluate): y0 [NS, nq] node at the start of the
    step, t1 [nq] end of the step, target [nq] query time (mode 0) or value level (mode 1) -> [NS, nq]."""
    dt = _np_dtype(opts.dtype)
    y0 = np.ascontiguousarray(y0, dtype=dt)
    t1 = np.ascontiguousarray(t1, dtype=dt)
    target = np.ascontiguousarray(target, dtype=dt)
    out = np.empty_like(y0)
    _check(lib().pmp_oracle_interp_batch(C.byref(problem), C.byref(opts), y0.shape[1], y0.ctypes.data, t1.ctypes.data,
                                         target.ctypes.data, int(mode), out.ctypes.data))
    return out


def solve_batch_to_times(problem: Problem, y_f: np.ndarray, tq: np.ndarray, rtol=1e-12, atol=1e-12, t1=-3.0,
                         max_steps=4096, chunk=256) -> np.ndarray:
    """Converged fp64 state of trajectory i at ITS OWN time tq[i] (between t0 = 0 and t1): a tight Tsit5 solve with
    SaveAt(steps) output, then the interpolant on the step that contains tq[i] (its error, O(h^5) with h ~ 5e-3, is far
    below any tolerance the tests use).  y_f [NS, n], tq [n] -> [NS, n].  Test infrastructure (dense-node parity)."""
    y_f = np.ascontiguousarray(y_f, dtype=np.float64)
    tq = np.asarray(tq, dtype=np.float64)
    ns, n = y_f.shape
    nx = problem.nx
    opts = make_opts(dtype=F64, rtol=rtol, atol=atol, dtmin=0.0, max_steps=max_steps, t1=t1, save_dense=1)
    sgn = 1.0 if t1 > 0 else -1.0
    out = np.empty((ns, n))
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        d = solve_batch(problem, opts, y_f[:, a:b])
        assert (d["result"] == 0).all(), "truth solve ran out of steps"
        m = b - a
        i = np.arange(m)
        key = sgn * d["ts"]                       # increasing along the slots, +inf padding
        s = np.array([np.searchsorted(key[j], sgn * tq[a + j], side="right") - 1 for j in range(m)])
        s = np.clip(s, 0, d["n_accepted"] - 1)
        node = np.concatenate([d["x"][i, s].T, d["t"][i, s][None], d["v"][i, s][None], d["vx"][i, s].T], 0)
        out[:, a:b] = interp_batch(problem, opts, node, d["ts"][i, s + 1], tq[a:b], 0)
    return out


def ustar_smooth(problem: Problem, x: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """flatquad, fp64: u_star_2d(..., smooth=True) (pontryagin_utils.py:217-254); x, lam [6, n] -> u [2, n]."""
    x = np.ascontiguousarray(x, np.float64); lam = np.ascontiguousarray(lam, np.float64)
    u = np.empty((2, x.shape[1]))
    L = lib()
    L.pmp_oracle_ustar_smooth.argtypes = [C.POINTER(Problem), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    _check(L.pmp_oracle_ustar_smooth(C.byref(problem), x.shape[1], x.ctypes.data, lam.ctypes.data, u.ctypes.data))
    return u


def vxx_rhs(problem: Problem, x, lam, vxx):
    """flatquad: (fx, flam, gx, glam, vxx_dot) of pontryagin_utils.py:460-467 for one state (fp64)."""
    x = np.ascontiguousarray(x, np.float64); lam = np.ascontiguousarray(lam, np.float64)
    vxx = np.ascontiguousarray(vxx, np.float64)
    jac = np.zeros((4, 6, 6)); out = np.zeros((6, 6))
    L = lib()
    L.pmp_oracle_vxx_rhs.argtypes = [C.POINTER(Problem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _check(L.pmp_oracle_vxx_rhs(C.byref(problem), x.ctypes.data, lam.ctypes.data, vxx.ctypes.data, jac.ctypes.data,
                                out.ctypes.data))
    return jac[0], jac[1], jac[2], jac[3], out


def vxx_rhs_fwd(problem: Problem, x, lam, vxx):
    """any system: (fx, flam, gx, glam, vxx_dot) by forward-mode differentiation of the oracle's pmp_rhs (dual numbers), the
    construction of the CUDA adaptor for generated functors (csrc/pmp_vxx_dual.cuh); fp64, one state per call."""
    x = np.ascontiguousarray(x, np.float64); lam = np.ascontiguousarray(lam, np.float64)
    vxx = np.ascontiguousarray(vxx, np.float64)
    nx = x.shape[0]
    jac = np.zeros((4, nx, nx)); out = np.zeros((nx, nx))
    L = lib()
    L.pmp_oracle_vxx_rhs_fwd.argtypes = [C.POINTER(Problem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _check(L.pmp_oracle_vxx_rhs_fwd(C.byref(problem), x.ctypes.data, lam.ctypes.data, vxx.ctypes.data, jac.ctypes.data,
                                    out.ctypes.data))
    return jac[0], jac[1], jac[2], jac[3], out


class NnCostate(C.Structure):
    _fields_ = [("n_nets", C.c_int32), ("n_hidden_layers", C.c_int32), ("width", C.c_int32), ("reserved", C.c_int32),
                ("weights", C.c_void_p), ("x_mean", C.c_double * 16), ("x_std", C.c_double * 16),
                ("v_mean", C.c_double), ("v_std", C.c_double), ("n_sigma", C.c_double), ("sigma_max", C.c_double)]


def nn_struct(weights: np.ndarray, nx, width, n_hidden, n_nets, x_mean=None, x_std=None, v_mean=0.0, v_std=1.0, n_sigma=2.0,
              sigma_max=0.5) -> NnCostate:
    """pmp_nn_costate_t over a HOST weight array (keep `weights` alive while the struct is in use)."""
    c = NnCostate(n_nets=n_nets, n_hidden_layers=n_hidden, width=width, reserved=0, weights=weights.ctypes.data,
                  v_mean=v_mean, v_std=v_std, n_sigma=n_sigma, sigma_max=sigma_max)
    for q in range(16):
        c.x_mean[q] = float(x_mean[q]) if x_mean is not None and q < nx else 0.0
        c.x_std[q] = float(x_std[q]) if x_std is not None and q < nx else 1.0
    return c


def nn_costate(nn: NnCostate, x: np.ndarray):
    """x [nx, n] (fp64 weights) -> lam [nx, n], v_mean [n], v_std [n]: the ensemble law of levelsets.py:789-852."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    nx, n = x.shape
    lam, vm, vs = np.empty_like(x), np.empty(n), np.empty(n)
    L = lib()
    L.pmp_oracle_nn_costate.argtypes = [C.POINTER(NnCostate), C.c_int32, C.c_int64] + [C.c_void_p] * 4
    _check(L.pmp_oracle_nn_costate(C.byref(nn), nx, n, x.ctypes.data, lam.ctypes.data, vm.ctypes.data, vs.ctypes.data))
    return lam, vm, vs


def forward_batch(problem: Problem, opts: Opts, x0: np.ndarray, P: np.ndarray = None, x_eq=None, nn: NnCostate = None) -> dict:
    """Forward closed-loop simulation under lambda(x) = P (x - x_eq): x0 [nx, n] -> dict(z_end [nx+1, n],
    n_accepted, n_steps, result [, ts, x, cost])   (levelsets.py:818-836, eval_utils.py:48-99)."""
    dt = _np_dtype(opts.dtype)
    x0 = np.ascontiguousarray(x0, dtype=dt)
    nx, n = x0.shape
    P = None if P is None else np.ascontiguousarray(P, dtype=np.float64)
    xe = None if x_eq is None else np.ascontiguousarray(x_eq, dtype=np.float64)
    out = {"z_end": np.empty((nx + 1, n), dt), "n_accepted": np.empty(n, np.int32), "n_steps": np.empty(n, np.int32),
           "result": np.empty(n, np.int32)}
    ptr = lambda a: None if a is None else a.ctypes.data
    if opts.save_dense:
        s1 = opts.max_steps + 1
        out.update(ts=np.empty((n, s1), dt), x=np.empty((n, s1, nx), dt), cost=np.empty((n, s1), dt))
    L = lib()
    L.pmp_oracle_forward_batch.argtypes = [C.POINTER(Problem), C.POINTER(Opts), C.c_int64] + [C.c_void_p] * 3 + \
        [C.POINTER(NnCostate)] + [C.c_void_p] * 7
    _check(L.pmp_oracle_forward_batch(C.byref(problem), C.byref(opts), n, ptr(x0), ptr(P), ptr(xe),
                                      None if nn is None else C.byref(nn), ptr(out["z_end"]),
                                      ptr(out.get("ts")), ptr(out.get("x")), ptr(out.get("cost")), ptr(out["n_accepted"]),
                                      ptr(out["n_steps"]), ptr(out["result"])))
    return out


def opcount(problem: Problem, opts: Opts, y0: np.ndarray) -> dict:
    """Algorithmic flops / specials per RHS and per step attempt."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    out = np.zeros(4, np.int64)
    _check(lib().pmp_oracle_opcount(C.byref(problem), C.byref(opts), y0.ctypes.data, out.ctypes.data))
    return {"rhs_flops": int(out[0]), "rhs_special": int(out[1]), "attempt_flops": int(out[2]),
            "attempt_special": int(out[3])}
