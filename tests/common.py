"""Shared helpers of the test-suite: seeded inputs that go, as the same arrays, to the CPU oracle
(oracle/) and to the CUDA library (through the C ABI)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu, systems, terminal  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def c_problem(name):
    pp = {"flatquad": systems.flatquad_problem_params, "orbits": systems.orbits_problem_params,
          "lqr": systems.lqr_statepenalty_problem_params}[name]()
    return systems.to_c_problem(pp), pp


def oracle_problem(name):
    return {"flatquad": O.flatquad_problem, "orbits": O.orbits_problem, "lqr": O.lqr_problem}[name]()


def as_oracle_opts(opts: _capi.Opts) -> O.Opts:
    """Byte-identical copy (the two ctypes classes mirror the same C struct)."""
    o = O.Opts()
    C.memmove(C.byref(o), C.byref(opts), C.sizeof(O.Opts))
    return o


_P = {}


def flatquad_P():
    if "P" not in _P:
        _P["K"], _P["P"] = pu.get_terminal_lqr(systems.flatquad_problem_params())
    return _P["P"]


def flatquad_terminal(n, seed=1, dtype=np.float64):
    """SoA [14, n] terminal states on the LQR level set V_f = 1e-3 (levelsets.py:551-599)."""
    st = terminal.sample_terminal_states(flatquad_P(), 1e-3, n, seed=seed, dtype=np.float64)
    y = np.concatenate([st["x"].T, st["t"][None], st["v"][None], st["vx"].T], 0)
    return np.ascontiguousarray(y, dtype=dtype)


def lqr_P():
    """CARE(A, B, Q=I, R=1): V = x'Px, lambda = 2Px (corrected convention)."""
    import scipy.linalg
    A = np.array([[0.0, 1.0], [0.0, 0.0]]); B = np.array([[0.0], [1.0]])
    return scipy.linalg.solve_continuous_are(A, B, np.eye(2), np.eye(1))


def _soa(st, dtype):
    return np.ascontiguousarray(np.concatenate([st["x"].T, st["t"][None], st["v"][None], st["vx"].T], 0), dtype=dtype)


def lqr_terminal(n=1024, radius=0.1, dtype=np.float64):
    """experiment_lqr_statepenalty.py:123-125: x_f = radius * sqrtm(P)^-1 (sin th, cos th)."""
    return _soa(terminal.ellipse_terminal_states(lqr_P(), n, radius), dtype)


def orbits_P0():
    """experiment_orbits.py:62-75: P0 = CARE(A,B,Q,R on the Hessians)/2 at x_eq = (0,1)."""
    import scipy.linalg
    A, B, Q, R = systems.linearise(systems.orbits_problem_params())
    return scipy.linalg.solve_continuous_are(A, B, Q, R) / 2


def orbits_terminal(n, radius=0.01, dtype=np.float64):
    """experiment_orbits.py:121-129: x_f = x_eq + radius * sqrtm(P0)^-1 (sin th, cos th)."""
    return _soa(terminal.ellipse_terminal_states(orbits_P0(), n, radius, x_eq=[0.0, 1.0]), dtype)


def scaled_err(a, b, tol, rows=None):
    """max over rows of |a-b| / (tol + tol*|b|): error in units of the solver tolerance."""
    if rows is not None:
        a, b = a[rows], b[rows]
    return (np.abs(a - b) / (tol + tol * np.abs(b))).max(0)


def cuda_solve(problem, opts, y_np):
    """numpy SoA -> CUDA library through the C ABI -> numpy dict (same keys as oracle.solve_batch)."""
    import torch
    dt = torch.float32 if opts.dtype == _capi.F32 else torch.float64
    y = torch.as_tensor(np.ascontiguousarray(y_np), dtype=dt, device="cuda").contiguous()
    out = pu.solve_batch_soa(problem, opts, y)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def flatquad_terminal_vxx(n, seed=1, dtype=np.float64, asym=0.0):
    """[14 + 36, n]: flatquad_terminal plus the terminal Hessian vxx = P_lqr on every trajectory
    (levelsets.py:595-597), row 14 + 6i + j = vxx[i][j].  asym > 0 adds a seeded non-symmetric perturbation
    (exercises the general 36-entry path)."""
    y = flatquad_terminal(n, seed=seed, dtype=np.float64)
    P = np.asarray(flatquad_P(), dtype=np.float64)
    P = 0.5 * (P + P.T)
    V = np.repeat(P.reshape(36, 1), n, axis=1)
    if asym > 0:
        V = V + asym * np.random.Generator(np.random.PCG64(seed + 7)).standard_normal((36, n))
    return np.ascontiguousarray(np.concatenate([y, V], 0), dtype=dtype)
