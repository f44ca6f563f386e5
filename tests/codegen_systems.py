"""Systems for the codegen tests that have NO hand-written functor, written the way the reference writes its
dynamics (jax.numpy style, here against symnp), plus a plain-numpy evaluation of the PMP right hand side from
the sympy-derived expressions (independent of the generated CUDA: lambdify + a brute-force box QP)."""
import os
import sys

import numpy as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from approximate_optimal_control_b200 import symnp as np  # noqa: E402  (stands in for jax.numpy)


def cart2d_problem_params():
    """4 states, 2 inputs; H_uu depends on the state and has an off-diagonal term."""
    def f(x, u):
        p1, p2, v1, v2 = x
        return np.array([v1, v2, u[0] * np.cos(p2) - 0.1 * v1, u[1] + np.sin(p1) - 0.1 * v2])

    def l(x, u):
        p1, p2, v1, v2 = x
        Q = np.diag(np.array([1.0, 2.0, 0.5, 0.5]))
        return x.T @ Q @ x + (0.2 + 0.5 * p1 ** 2) * u[0] ** 2 + 0.3 * u[1] ** 2 + 0.1 * u[0] * u[1]

    return {"system_name": "cart2d_test", "f": f, "l": l, "nx": 4, "nu": 2,
            "U_interval": [onp.array([-1.0, -0.5]), onp.array([1.5, 0.75])],
            "x_eq": onp.zeros(4), "u_eq": onp.zeros(2)}


def pendulum_problem_params():
    """damped pendulum with a torque input, nu = 1 (u_star_new form), exp / tanh in the running cost."""
    def f(x, u):
        th, om = x
        return np.array([om, -9.81 * np.sin(th) - 0.2 * om + 2.0 * u[0]])

    def l(x, u):
        th, om = x
        return 2.0 * (1 - np.cos(th)) + 0.1 * om ** 2 + 0.5 * np.tanh(om) ** 2 + 0.05 * np.exp(-th ** 2) + (0.5 + 0.1 * np.cos(th)) * u[0] ** 2

    return {"system_name": "pendulum_test", "f": f, "l": l, "nx": 2, "nu": 1, "U_interval": [-2.0, 2.0],
            "x_eq": onp.zeros(2), "u_eq": onp.zeros(1)}


def box_qp_2d(Hu, Huu, lo, hi):
    """argmin 1/2 u'Huu u + Hu'u over the box, by enumeration of the 9 active sets (KKT check)."""
    H = onp.array([[Huu[0], Huu[1]], [Huu[1], Huu[2]]], dtype=float)
    g = onp.asarray(Hu, dtype=float)
    best, bestv = None, onp.inf
    for a0 in (None, lo[0], hi[0]):
        for a1 in (None, lo[1], hi[1]):
            u = onp.zeros(2)
            if a0 is None and a1 is None:
                u = onp.linalg.solve(H, -g)
            elif a0 is None:
                u[1] = a1; u[0] = -(g[0] + H[0, 1] * a1) / H[0, 0]
            elif a1 is None:
                u[0] = a0; u[1] = -(g[1] + H[0, 1] * a0) / H[1, 1]
            else:
                u[:] = (a0, a1)
            if (u < lo - 1e-12).any() or (u > hi + 1e-12).any():
                continue
            v = 0.5 * u @ H @ u + g @ u
            if v < bestv:
                best, bestv = u, v
    return best


def numpy_field(pp):
    """forward-time f_extended(x, lam) -> (xdot, vdot, lamdot, u*) in numpy from the DERIVED sympy expressions."""
    from approximate_optimal_control_b200 import codegen
    d = codegen.derive(pp)
    model, rhs = codegen.lambdify_rhs(d)
    lo, hi = (onp.broadcast_to(onp.asarray(b, dtype=float), (d.nu,)) for b in pp["U_interval"])

    def field(x, lam):
        if d.nu == 2:
            Hu, Huu = model(x, lam)
            u = box_qp_2d(Hu, Huu, lo, hi)
        else:
            num, den = model(x, lam)
            u = onp.clip(onp.array([num / den]), lo, hi)
        f, l, hx = rhs(x, u, lam)
        return onp.asarray(f, dtype=float), -float(l), -onp.asarray(hx, dtype=float), u

    return field


# ---- caller-supplied feedback laws for eval_utils.closed_loop_eval_general (eval_utils.py:48-99) -------------------------
def lqr_feedback_law(P, r_u=1.0, lo=-1.0, hi=1.0):
    """the LQR law of the double integrator as a plain function of x: lambda = P x, u = clip(-lambda_1 / (2 r_u), U)
    (what closed_loop_eval_lqr + the hand-written functor's u_star_new compute)."""
    P = onp.asarray(P, dtype=float)

    def law(x):
        lam1 = P[1, 0] * x[0] + P[1, 1] * x[1]
        return np.array([np.clip(-lam1 / (2.0 * r_u), lo, hi)])
    return law


def pendulum_feedback_law():
    def law(x):
        return np.array([np.clip(-(3.0 * x[0] + 1.5 * x[1]) + 0.2 * np.sin(x[0]), -2.0, 2.0)])
    return law


def numpy_closed_loop(pp, law):
    """(f(x, u(x)), l(x, u(x)), u(x)) as numpy callables, by lambdifying the traced expressions."""
    import sympy as sp
    from approximate_optimal_control_b200 import codegen
    d = codegen.derive(dict(pp, ustar_fct=law))
    sub = dict(zip(d.u, d.policy))
    mods = [{"Max": onp.maximum, "Min": onp.minimum}, "numpy"]
    f = sp.lambdify([d.x], [e.subs(sub) for e in d.f], modules=mods)
    l = sp.lambdify([d.x], d.l.subs(sub), modules=mods)
    u = sp.lambdify([d.x], d.policy[0], modules=mods)
    return (lambda x: onp.asarray(f(x), dtype=float)), (lambda x: float(l(x))), (lambda x: float(u(x)))
