"""The three control systems of the reference's experiments, as `problem_params` / `algo_params` dicts
with the reference's own keys, plus their mapping to the compiled CUDA functors.

The reference defines f, l inside each experiment script with jax.numpy and differentiates them with
jax autodiff.  Here the same f, l are plain functions that accept numpy arrays (vectorised over
leading axes) or sympy symbols; the derivatives the hot path needs (dH/dx, H_u, H_uu) are
hand-derived inside the CUDA functors (csrc/pmp_systems.cuh), and the host-side linearisation for
the terminal LQR (get_terminal_lqr) is done symbolically with sympy.

  flatquad                               flatquad_landing_experiment.py:265-437
  orbits                                 experiment_orbits.py:24-104
  double_integrator_linear_statepenalty  experiment_lqr_statepenalty.py:23-106
"""
from __future__ import annotations

import numpy as np

from . import _capi

__all__ = ["flatquad_problem_params", "flatquad_algo_params", "orbits_problem_params",
           "orbits_algo_params", "lqr_statepenalty_problem_params", "lqr_statepenalty_algo_params",
           "to_c_problem", "linearise", "SYSTEM_IDS"]

SYSTEM_IDS = {
    "flatquad": _capi.SYS_FLATQUAD,
    "orbits": _capi.SYS_ORBITS,
    "double_integrator_linear_statepenalty": _capi.SYS_LQR_STATEPENALTY,
}


def _is_sym(*args) -> bool:
    for a in args:
        for e in np.ravel(np.asarray(a, dtype=object)):
            if type(e).__module__.startswith("sympy"):
                return True
    return False


def _m(*args):
    """math namespace matching the argument type: numpy, or symnp (the jax.numpy subset on sympy objects)."""
    if _is_sym(*args):
        from . import symnp
        return symnp
    return np


def _stack(items, like):
    if _is_sym(*items):
        return np.array(items, dtype=object)
    return np.stack(np.broadcast_arrays(*[np.asarray(i, dtype=float) for i in items]), axis=-1)


# =====================================================================================================
# flatquad
# =====================================================================================================
FLATQUAD_CONSTANTS = dict(m=20.0, g=9.81, r=0.5)  # flatquad_landing_experiment.py:269-273
FLATQUAD_CONSTANTS["I"] = FLATQUAD_CONSTANTS["m"] * (FLATQUAD_CONSTANTS["r"] / 2) ** 2
# state_length_scales on (px, py), (cos, sin), (vx, vy), omega            (:313)
FLATQUAD_SCALES = (0.3, float(np.deg2rad(30.0)), 0.5, float(np.deg2rad(120.0)))


def _flatquad_fl(constants, scales):
    m, g, r, I = (constants[k] for k in ("m", "g", "r", "I"))
    q_pos, q_ang, q_vel, q_om = (1.0 / s ** 2 for s in scales)

    def f(x, u):  # flatquad_landing_experiment.py:277-292
        M = _m(x, u)
        x = np.asarray(x, dtype=object if M is not np else float)
        u = np.asarray(u, dtype=object if M is not np else float)
        Fl, Fr = u[..., 0], u[..., 1]
        Phi, vx, vy, om = x[..., 2], x[..., 3], x[..., 4], x[..., 5]
        return _stack([vx, vy, om, -M.sin(Phi) * (Fl + Fr) / m, M.cos(Phi) * (Fl + Fr) / m - g,
                       (Fr - Fl) * r / I], x)

    def l(x, u):  # flatquad_landing_experiment.py:294-329
        M = _m(x, u)
        x = np.asarray(x, dtype=object if M is not np else float)
        u = np.asarray(u, dtype=object if M is not np else float)
        Fl, Fr = u[..., 0], u[..., 1]
        px, py, Phi, vx, vy, om = (x[..., i] for i in range(6))
        state_cost = (q_pos * (px ** 2 + py ** 2) + q_ang * ((M.cos(Phi) - 1) ** 2 + M.sin(Phi) ** 2)
                      + q_vel * (vx ** 2 + vy ** 2) + q_om * om ** 2)
        ax = -M.sin(Phi) * (Fl + Fr) / m
        ay = M.cos(Phi) * (Fl + Fr) / m - g
        aw = (Fr - Fl) * r / I
        return state_cost + ax ** 2 + ay ** 2 + aw ** 2

    return f, l


def flatquad_problem_params(constants=None, scales=None) -> dict:
    """problem_params of flatquad_landing_experiment.py:340-365 (same keys and values)."""
    constants = dict(FLATQUAD_CONSTANTS if constants is None else constants)
    scales = tuple(FLATQUAD_SCALES if scales is None else scales)
    f, l = _flatquad_fl(constants, scales)
    m, g = constants["m"], constants["g"]
    umax = m * g * 1.2 / 2
    return {
        "system_name": "flatquad",
        "f": f,
        "l": l,
        "h": lambda x: float(np.dot(x, x)),
        "nx": 6,
        "state_names": ("x", "y", "Phi", "vx", "vy", "omega"),
        "T": lambda x: np.concatenate([x[..., 0:2], np.cos(x[..., 2:3]), np.sin(x[..., 2:3]), x[..., 3:]], -1),
        "nu": 2,
        "U_interval": [np.zeros(2), umax * np.ones(2)],
        "V_f": 0.001,
        "V_max": 1000.0,
        "u_eq": np.ones(2) * m * g / 2,
        "x_eq": np.zeros(6),
        # extension: the closure constants of the experiment script, needed by the CUDA functor
        "constants": constants,
        "state_length_scales": scales,
    }


def flatquad_algo_params() -> dict:
    """The solver-related subset of algo_params, flatquad_landing_experiment.py:368-392."""
    return {
        "pontryagin_solver_vxx": False,
        "pontryagin_solver_atol": 1e-5,
        "pontryagin_solver_rtol": 1e-5,
        "pontryagin_solver_maxsteps": 128,
        "pontryagin_solver_T": 3.0,
        "vxx_max_norm": 1e4,
        "throw": False,
    }


# =====================================================================================================
# orbits
# =====================================================================================================
def _orbits_fl(w_dist=0.1, w_u=10.0):
    def f(t, x, u):  # experiment_orbits.py:24-30
        M = _m(x, u)
        x = np.asarray(x, dtype=object if M is not np else float)
        u = np.asarray(u, dtype=object if M is not np else float)
        uu = u[..., 0]
        rho = x[..., 0] ** 2 + x[..., 1] ** 2 - 1
        return _stack([uu * x[..., 0] - rho * x[..., 1], rho * x[..., 0] + uu * x[..., 1]], x)

    def l(t, x, u):  # experiment_orbits.py:33-40
        M = _m(x, u)
        x = np.asarray(x, dtype=object if M is not np else float)
        u = np.asarray(u, dtype=object if M is not np else float)
        rho = x[..., 0] ** 2 + x[..., 1] ** 2 - 1
        return rho ** 2 + w_dist * (x[..., 0] ** 2 + (x[..., 1] - 1) ** 2) + w_u * u[..., 0] ** 2

    return f, l


def orbits_problem_params() -> dict:
    """problem_params of experiment_orbits.py:83-95."""
    f, l = _orbits_fl()
    return {
        "system_name": "orbits",
        "f": f,
        "l": l,
        "T": np.inf,
        "nx": 2,
        "nu": 1,
        "U_interval": [-0.2, 0.2],
        "terminal_constraint": False,
        "V_max": 4.2,
        "x_eq": np.array([0.0, 1.0]),
        "u_eq": np.zeros(1),
        "constants": {"w_dist": 0.1, "w_u": 10.0},
    }


def orbits_algo_params() -> dict:
    """experiment_orbits.py:97-104."""
    return {
        "pontryagin_solver_dt": 1 / 16,
        "pontryagin_solver_adaptive": True,
        "pontryagin_solver_dense": False,
        "pontryagin_solver_rtol": 1e-4,
        "pontryagin_solver_atol": 1e-4,
        "pontryagin_solver_maxsteps": 1024,
    }


# =====================================================================================================
# double integrator with soft state constraints
# =====================================================================================================
def _lqr_fl(a=0.01, r_u=1.0):
    def f(t, x, u):  # experiment_lqr_statepenalty.py:23-28
        M = _m(x, u)
        x = np.asarray(x, dtype=object if M is not np else float)
        u = np.asarray(u, dtype=object if M is not np else float)
        return _stack([x[..., 1], u[..., 0]], x)

    def l(t, x, u):  # experiment_lqr_statepenalty.py:30-38
        M = _m(x, u)
        x = np.asarray(x, dtype=object if M is not np else float)
        u = np.asarray(u, dtype=object if M is not np else float)
        z1, z0 = M.maximum(0, x[..., 1] - 1), M.maximum(0, x[..., 0] - 1)
        return x[..., 0] ** 2 + x[..., 1] ** 2 + r_u * u[..., 0] ** 2 + z1 ** 2 / a + z0 ** 2 / a

    return f, l


def lqr_statepenalty_problem_params(U_interval=(-1.0, 1.0), a=0.01) -> dict:
    """problem_params of experiment_lqr_statepenalty.py:80-92."""
    f, l = _lqr_fl(a=a)
    return {
        "system_name": "double_integrator_linear_statepenalty",
        "f": f,
        "l": l,
        "h": lambda x: float(np.dot(x, x)),
        "T": 8,
        "nx": 2,
        "nu": 1,
        "U_interval": [U_interval[0], U_interval[1]],
        "terminal_constraint": False,
        "V_max": 30,
        "x_eq": np.zeros(2),
        "u_eq": np.zeros(1),
        "constants": {"a": a, "r_u": 1.0},
    }


def lqr_statepenalty_algo_params() -> dict:
    """experiment_lqr_statepenalty.py:99-106."""
    return {
        "pontryagin_solver_dt": 1 / 16,
        "pontryagin_solver_adaptive": True,
        "pontryagin_solver_dense": False,
        "pontryagin_solver_rtol": 1e-10,
        "pontryagin_solver_atol": 1e-10,
        "pontryagin_solver_maxsteps": 1024,
    }


# =====================================================================================================
# mapping to the C ABI
# =====================================================================================================
def to_c_problem(problem_params: dict) -> _capi.Problem:
    """problem_params -> pmp_problem_t.  problem_params['system_name'] selects one of the hand-written functors;
    any other system (or problem_params['codegen'] = True) is generated from problem_params['f'], ['l'] and
    compiled into a plug-in library (codegen.py) -- cached, so only the first call pays for sympy + nvcc."""
    name = problem_params.get("system_name")
    if problem_params.get("codegen", name not in SYSTEM_IDS):
        if problem_params.get("codegen") is False or "f" not in problem_params or "l" not in problem_params:
            raise NotImplementedError(
                f"system_name={name!r}: not one of the compiled systems {sorted(SYSTEM_IDS)}, and no f, l to "
                "generate a functor from (or problem_params['codegen'] is False)")
        from . import codegen
        return codegen.plugin_problem(problem_params)
    sid = SYSTEM_IDS[name]
    nx, nu = int(problem_params["nx"]), int(problem_params["nu"])
    p = _capi.Problem(system_id=sid, nx=nx, nu=nu)
    lo, hi = problem_params["U_interval"]
    lo = np.broadcast_to(np.asarray(lo, dtype=float), (nu,))
    hi = np.broadcast_to(np.asarray(hi, dtype=float), (nu,))
    for i in range(nu):
        p.u_lo[i], p.u_hi[i] = float(lo[i]), float(hi[i])
    if sid == _capi.SYS_FLATQUAD:
        c = problem_params.get("constants", FLATQUAD_CONSTANTS)
        sc = problem_params.get("state_length_scales", FLATQUAD_SCALES)
        vals = [c["m"], c["g"], c["r"], c["I"]] + [1.0 / s ** 2 for s in sc]
    elif sid == _capi.SYS_ORBITS:
        c = problem_params.get("constants", {"w_dist": 0.1, "w_u": 10.0})
        vals = [c["w_dist"], c["w_u"]]
    else:
        c = problem_params.get("constants", {"a": 0.01, "r_u": 1.0})
        vals = [c["a"], c["r_u"]]
    for i, v in enumerate(vals):
        p.params[i] = float(v)
    return p


def _call_f(fn, x, u):
    """The reference has 2-arg f(x,u) (flatquad) and stale 3-arg f(t,x,u) (orbits, lqr)."""
    import inspect
    try:
        nargs = len(inspect.signature(fn).parameters)
    except (TypeError, ValueError):
        nargs = 2
    return fn(x, u) if nargs == 2 else fn(0.0, x, u)


def linearise(problem_params: dict, x_eq=None, u_eq=None):
    """A = df/dx, B = df/du, Q = d2l/dx2, R = d2l/du2 at (x_eq, u_eq), as the reference obtains with
    jax.jacobian / jax.hessian (pontryagin_utils.py:616-619).  Symbolic (sympy) when f, l accept
    sympy symbols, central finite differences otherwise."""
    nx, nu = int(problem_params["nx"]), int(problem_params["nu"])
    x_eq = np.asarray(problem_params["x_eq"] if x_eq is None else x_eq, dtype=float).reshape(nx)
    u_eq = np.asarray(problem_params["u_eq"] if u_eq is None else u_eq, dtype=float).reshape(nu)
    f, l = problem_params["f"], problem_params["l"]
    try:
        import sympy
        xs = sympy.symbols(f"x0:{nx}", real=True)
        us = sympy.symbols(f"u0:{nu}", real=True)
        item = lambda e: sympy.sympify(e.item() if isinstance(e, np.ndarray) else e)
        fs = sympy.Matrix([item(e) for e in _call_f(f, np.array(xs, dtype=object), np.array(us, dtype=object))])
        ls = item(_call_f(l, np.array(xs, dtype=object), np.array(us, dtype=object)))
        sub = {**dict(zip(xs, x_eq)), **dict(zip(us, u_eq))}
        ev = lambda M: np.array(M.subs(sub).evalf(30), dtype=float)
        # Max(0, z) has a Heaviside derivative; away from the kink subs() resolves it
        A = ev(fs.jacobian(xs))
        B = ev(fs.jacobian(us))
        Q = ev(sympy.hessian(ls, xs))
        R = ev(sympy.hessian(ls, us))
        return A, B.reshape(nx, nu), Q, R
    except Exception:  # user callables that only take numeric arrays
        return _linearise_fd(f, l, x_eq, u_eq)


def _linearise_fd(f, l, x_eq, u_eq, h=1e-5):
    nx, nu = x_eq.size, u_eq.size
    ff = lambda x, u: np.asarray(_call_f(f, x, u), dtype=float)
    ll = lambda x, u: float(_call_f(l, x, u))
    A = np.zeros((nx, nx)); B = np.zeros((nx, nu))
    for j in range(nx):
        e = np.zeros(nx); e[j] = h
        A[:, j] = (ff(x_eq + e, u_eq) - ff(x_eq - e, u_eq)) / (2 * h)
    for j in range(nu):
        e = np.zeros(nu); e[j] = h
        B[:, j] = (ff(x_eq, u_eq + e) - ff(x_eq, u_eq - e)) / (2 * h)

    def hess(fun, z0, hh=1e-4):
        n = z0.size
        H = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                ei = np.zeros(n); ei[i] = hh
                ej = np.zeros(n); ej[j] = hh
                H[i, j] = H[j, i] = (fun(z0 + ei + ej) - fun(z0 + ei - ej) - fun(z0 - ei + ej)
                                     + fun(z0 - ei - ej)) / (4 * hh * hh)
        return H

    Q = hess(lambda x: ll(x, u_eq), x_eq)
    R = hess(lambda u: ll(x_eq, u), u_eq)
    return A, B, Q, R
