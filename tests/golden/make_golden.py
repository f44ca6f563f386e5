#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ by RUNNING THE REFERENCE'S OWN PYTHON CODE.

jax / diffrax cannot be installed in the build container (no network), so the reference modules are
imported from /root/reference with a small stand-in for the `jax` API implemented on torch
(torch.func.jacrev / hessian / vmap for autodiff, torch ops for jax.numpy).  What runs is the
reference's unmodified source:

  * pontryagin_utils.u_star_2d, u_star_new, define_backward_solver -> f_extended   (imported)
  * the f, l closures of the three experiment scripts                               (exec of the
    defining source lines, read from /root/reference at generation time -- nothing is copied)

diffrax itself (the integrator) is NOT available, so trajectory fixtures are produced by integrating
the reference's own f_extended with scipy.integrate.solve_ivp(DOP853, rtol=1e-12): they pin the
CONVERGED solution, not diffrax's step sequence.

Run once in the build container:   python tests/golden/make_golden.py
The .npz outputs are committed; the GPU box never needs /root/reference.
"""
import os
import sys
import textwrap
import types
import warnings

import numpy as onp
import scipy.integrate  # noqa: F401  (imported before the jax stand-in is installed)
import scipy.linalg  # noqa: F401
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


# ------------------------------------------------------------------------------------------------
# jax stand-in
# ------------------------------------------------------------------------------------------------
def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(x, dtype=torch.get_default_dtype())


def _array(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)):
        return torch.stack([_array(e) for e in x])
    return torch.as_tensor(x, dtype=torch.get_default_dtype())


def _make_jnp():
    np_ = types.ModuleType("jax.numpy")
    np_.array = _array
    np_.asarray = _array
    np_.zeros = lambda n, *a, **k: torch.zeros(n)
    np_.ones = lambda n, *a, **k: torch.ones(n)
    np_.eye = lambda n: torch.eye(n)
    np_.diag = lambda v: torch.diag(_t(v))
    np_.sin = lambda x: torch.sin(_t(x))
    np_.cos = lambda x: torch.cos(_t(x))
    np_.sqrt = lambda x: torch.sqrt(_t(x))
    np_.deg2rad = lambda x: _t(x) * (onp.pi / 180.0)
    np_.maximum = lambda a, b: torch.maximum(_t(a), _t(b))
    np_.minimum = lambda a, b: torch.minimum(_t(a), _t(b))
    np_.max = lambda a: torch.max(_t(a))
    np_.clip = lambda x, lo, hi: torch.minimum(torch.maximum(_t(x), _t(lo)), _t(hi))
    np_.argmin = lambda a: torch.argmin(a)
    np_.roll = lambda a, s, axis: torch.roll(a, s, axis)
    np_.row_stack = lambda xs: torch.vstack([_t(x) for x in xs])
    np_.vstack = np_.row_stack
    np_.concatenate = lambda xs, axis=0: torch.cat([_t(x).reshape(-1) if _t(x).dim() == 0 else _t(x) for x in xs], axis)
    np_.inf = float("inf")
    np_.nan = float("nan")
    np_.pi = onp.pi
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.solve = lambda A, b: torch.linalg.solve(A, b)
    linalg.inv = torch.linalg.inv
    linalg.norm = torch.linalg.norm
    np_.linalg = linalg
    return np_


def install_shim():
    # jax arrays accept x.reshape() (-> scalar) and x.reshape(other.shape)
    _reshape = torch.Tensor.reshape
    torch.Tensor.reshape = lambda self, *shape: _reshape(self, *(shape if shape else ((),)))
    jax = types.ModuleType("jax")
    jnp = _make_jnp()
    jax.numpy = jnp
    jax.jacobian = lambda f, argnums=0: torch.func.jacrev(f, argnums=argnums)
    # reverse-over-reverse: forward-mode (torch.func.hessian = jacfwd(jacrev)) mixes dtypes in float32 mode
    jax.hessian = lambda f, argnums=0: torch.func.jacrev(torch.func.jacrev(f, argnums=argnums), argnums=argnums)
    jax.vmap = lambda f, in_axes=0: torch.func.vmap(f, in_dims=in_axes)
    # u_star_2d(smooth=True): jax.scipy.special.logsumexp, jax.nn.softmax (pontryagin_utils.py:231-247)
    special = types.ModuleType("jax.scipy.special")
    special.logsumexp = lambda a: torch.logsumexp(_t(a), dim=0)
    jscipy = types.ModuleType("jax.scipy")
    jscipy.special = special
    jax.scipy = jscipy
    jax.nn = types.SimpleNamespace(softmax=lambda a: torch.softmax(_t(a), dim=0))
    sys.modules["jax.scipy"] = jscipy
    sys.modules["jax.scipy.special"] = special
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.Array = type("Array", (), {})  # scipy's array-API probe looks for it when 'jax' is in sys.modules
    sys.modules["jax"] = jax
    sys.modules["jax.numpy"] = jnp
    for name in ("ipdb", "diffrax"):
        sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, REF)
    return jnp


def exec_lines(path, first, last, ns):
    """exec source lines [first, last] (1-based, inclusive) of a reference file in namespace ns."""
    with open(os.path.join(REF, path)) as fh:
        src = "".join(fh.readlines()[first - 1:last])
    exec(compile(textwrap.dedent(src), path, "exec"), ns)
    return ns


# ------------------------------------------------------------------------------------------------
def flatquad_ref(jnp):
    ns = exec_lines("flatquad_landing_experiment.py", 269, 329, {"np": jnp})
    umax = ns["umax"]
    pp = {"system_name": "flatquad", "f": ns["f"], "l": ns["l"], "nx": 6, "nu": 2,
          "U_interval": [torch.zeros(2), umax * torch.ones(2)]}
    return pp


def flatquad_inputs(n, rng):
    """(x, lambda) pairs covering unconstrained, one- and two-active-constraint regimes."""
    xs = rng.standard_normal((n, 6)) * onp.array([1.0, 1.0, 0.6, 2.0, 2.0, 1.0])
    scale = 10.0 ** rng.uniform(-1, 2.5, size=(n, 1))
    lams = rng.standard_normal((n, 6)) * scale
    # the reference's own regression point (flatquad_landing_experiment.py:171-172)
    xs[0] = [4.538097, 6.19019, -0.10679064, -2.4105127, -3.6202462, 0.04513513]
    lams[0] = [17.425396, 22.878305, 0.73656803, -2.9136074, -6.116059, 0.28396177]
    # near equilibrium (terminal region): lambda = small
    xs[1:n // 8] *= 0.01
    lams[1:n // 8] = rng.standard_normal((n // 8 - 1, 6)) * 0.3
    return xs, lams


def gen_rhs_flatquad(jnp, dtype, n=512):
    import pontryagin_utils as ref
    torch.set_default_dtype(dtype)
    pp = flatquad_ref(jnp)
    algo = {"pontryagin_solver_vxx": False}
    _, f_extended = ref.define_backward_solver(pp, algo)
    rng = onp.random.Generator(onp.random.PCG64(7))
    xs, lams = flatquad_inputs(n, rng)
    us = onp.zeros((n, 2)); dx = onp.zeros((n, 6)); dlam = onp.zeros((n, 6)); dv = onp.zeros(n); dt = onp.zeros(n)
    for i in range(n):
        x = torch.as_tensor(xs[i], dtype=dtype); lam = torch.as_tensor(lams[i], dtype=dtype)
        us[i] = ref.u_star_2d(x, lam, pp).numpy()
        sd = f_extended(0.0, {"x": x, "v": torch.zeros(()), "vx": lam, "t": 0})
        dx[i] = sd["x"].numpy(); dlam[i] = sd["vx"].numpy(); dv[i] = float(sd["v"]); dt[i] = float(sd["t"])
    tag = "f64" if dtype == torch.float64 else "f32"
    onp.savez(os.path.join(OUT, f"rhs_flatquad_{tag}.npz"), x=xs, lam=lams, u=us, dx=dx, dlam=dlam, dv=dv, dt=dt)
    print("rhs_flatquad", tag, "inside-box fraction:",
          onp.mean((us > 1e-9).all(1) & (us < 117.72 - 1e-6).all(1)))


def gen_rhs_2d(jnp, which, n=256):
    """orbits / lqr: u_star_new from the reference; x' = dH/dlam, lam' = -dH/dx, v' = -l assembled as
    in pontryagin_utils.py:438-439,449 from the experiment's own f, l (3-arg signatures)."""
    import pontryagin_utils as ref
    torch.set_default_dtype(torch.float64)
    if which == "orbits":
        ns = exec_lines("experiment_orbits.py", 24, 40, {"np": jnp})
        U = [-0.2, 0.2]
    else:
        ns = exec_lines("experiment_lqr_statepenalty.py", 23, 38, {"np": jnp})
        U = [-1.0, 1.0]
    f, l = ns["f"], ns["l"]
    pp = {"f": f, "l": l, "nx": 2, "nu": 1, "U_interval": U}
    rng = onp.random.Generator(onp.random.PCG64(11))
    xs = rng.standard_normal((n, 2)) * 1.5 + (onp.array([0.0, 1.0]) if which == "orbits" else 0.0)
    lams = rng.standard_normal((n, 2)) * 10.0 ** rng.uniform(-1, 1.5, size=(n, 1))
    us = onp.zeros((n, 1)); dx = onp.zeros((n, 2)); dlam = onp.zeros((n, 2)); dv = onp.zeros(n)
    H = lambda x, u, lam: l(0.0, x, u) + lam @ f(0.0, x, u)
    for i in range(n):
        x = torch.as_tensor(xs[i]); lam = torch.as_tensor(lams[i])
        u = ref.u_star_new(x, lam, pp)
        us[i] = u.numpy()
        dx[i] = torch.func.jacrev(H, argnums=2)(x, u, lam).numpy()
        dlam[i] = -torch.func.jacrev(H, argnums=0)(x, u, lam).numpy()
        dv[i] = -float(l(0.0, x, u))
    onp.savez(os.path.join(OUT, f"rhs_{which}_f64.npz"), x=xs, lam=lams, u=us, dx=dx, dlam=dlam, dv=dv)
    print("rhs", which, "clipped fraction:", onp.mean((us <= U[0]) | (us >= U[1])))


def gen_traj_flatquad(jnp, n=8):
    """Converged backward trajectories of the reference's f_extended (fp64) from LQR terminal states."""
    import scipy.integrate
    import scipy.linalg
    import pontryagin_utils as ref
    torch.set_default_dtype(torch.float64)
    pp = flatquad_ref(jnp)
    _, f_extended = ref.define_backward_solver(pp, {"pontryagin_solver_vxx": False})
    # terminal LQR exactly as get_terminal_lqr (pontryagin_utils.py:609-628) with the shim autodiff
    x_eq = torch.zeros(6); u_eq = torch.ones(2) * 20 * 9.81 / 2
    A = torch.func.jacrev(pp["f"], argnums=0)(x_eq, u_eq).numpy()
    B = torch.func.jacrev(pp["f"], argnums=1)(x_eq, u_eq).numpy()
    Q = sys.modules["jax"].hessian(pp["l"], argnums=0)(x_eq, u_eq).numpy()
    R = sys.modules["jax"].hessian(pp["l"], argnums=1)(x_eq, u_eq).numpy()
    P = scipy.linalg.solve_continuous_are(A, B, Q, R)
    L = onp.linalg.cholesky(P)
    rng = onp.random.Generator(onp.random.PCG64(3))
    z = rng.standard_normal((n, 6)); s = z / onp.linalg.norm(z, axis=1)[:, None]
    xf = s @ onp.linalg.inv(L) * onp.sqrt(1e-3) * onp.sqrt(2)
    y0 = onp.concatenate([xf, onp.zeros((n, 1)), onp.full((n, 1), 1e-3), xf @ P.T], 1)  # x,t,v,vx

    def rhs(t, y):
        sd = f_extended(t, {"x": torch.as_tensor(y[:6]), "t": y[6], "v": torch.as_tensor(y[7]),
                            "vx": torch.as_tensor(y[8:])})
        return onp.concatenate([sd["x"].numpy(), [1.0], [float(sd["v"])], sd["vx"].numpy()])

    t_eval = onp.array([-1.0, -2.0, -3.0])
    ys = onp.zeros((n, 3, 14))
    for i in range(n):
        sol = scipy.integrate.solve_ivp(rhs, (0.0, -3.0), y0[i], method="DOP853", rtol=1e-12, atol=1e-14,
                                        t_eval=t_eval)
        assert sol.success
        ys[i] = sol.y.T
        print("traj", i, "nfev", sol.nfev, "v(-3) =", ys[i, -1, 7])
    onp.savez(os.path.join(OUT, "traj_flatquad_f64.npz"), P=P, A=A, B=B, Q=Q, R=R, y0=y0, t_eval=t_eval, ys=ys)


def gen_vxx_flatquad(jnp, n=192):
    """f_extended with pontryagin_solver_vxx=True (pontryagin_utils.py:460-467): vxx_dot and the four
    Jacobians of pmp_rhs THROUGH u_star_2d (args='debug' returns them), fp64."""
    import pontryagin_utils as ref
    torch.set_default_dtype(torch.float64)
    pp = flatquad_ref(jnp)
    _, f_extended = ref.define_backward_solver(pp, {"pontryagin_solver_vxx": True})
    rng = onp.random.Generator(onp.random.PCG64(21))
    xs, lams = flatquad_inputs(n, rng)
    A = rng.standard_normal((n, 6, 6))
    vxxs = onp.einsum("nij,nkj->nik", A, A) * 3.0 + 5.0 * onp.eye(6)[None]
    vxxs[n // 2:] += 0.3 * rng.standard_normal((n - n // 2, 6, 6))  # not exactly symmetric, as after integration
    out = {k: onp.zeros((n, 6, 6)) for k in ("vxx_dot", "fx", "flam", "gx", "glam")}
    us = onp.zeros((n, 2))
    for i in range(n):
        st = {"x": torch.as_tensor(xs[i]), "t": 0, "v": torch.zeros(()), "vx": torch.as_tensor(lams[i]),
              "vxx": torch.as_tensor(vxxs[i])}
        sd, aux = f_extended(0.0, st, "debug")
        out["vxx_dot"][i] = sd["vxx"].numpy()
        for k in ("fx", "flam", "gx", "glam"):
            out[k][i] = aux[k].numpy()
        us[i] = aux["u_star"].numpy()
    onp.savez(os.path.join(OUT, "vxx_flatquad_f64.npz"), x=xs, lam=lams, vxx=vxxs, u=us, **out)
    sat = ((us <= 1e-9) | (us >= 117.72 - 1e-6))
    print("vxx golden: both free", (~sat).all(1).mean(), "one clipped", (sat.sum(1) == 1).mean(), "corner", sat.all(1).mean())


def gen_ustar_smooth_flatquad(jnp, n=256):
    """u_star_2d(x, lambda, problem_params, smooth=True) (pontryagin_utils.py:217-254): the softmax blend of the five
    candidates, from the reference itself (fp64)."""
    import pontryagin_utils as ref
    torch.set_default_dtype(torch.float64)
    pp = flatquad_ref(jnp)
    rng = onp.random.Generator(onp.random.PCG64(31))
    xs, lams = flatquad_inputs(n, rng)
    # half of the sample within a few smoothing lengths of an active-set change (the unconstrained minimiser close to an
    # edge or a corner of the box), where the blend differs from the hard selection: pick that minimiser u_t, then the
    # costate entries (lambda_4, lambda_5) that produce it: H_u(0) = -H_uu u_t, H_u(0) = base -+ rot,
    # base = (lam4 cos - lam3 sin)/m - 2 g cos/m, rot = lam5 r/I  (flatquad_landing_experiment.py:269-329)
    m, g, r = 20.0, 9.81, 0.5
    I = m * (r / 2) ** 2
    d, o = 2 * (1 / m ** 2 + r ** 2 / I ** 2), 2 * (1 / m ** 2 - r ** 2 / I ** 2)
    umax = m * g * 1.2 / 2
    for i in range(n // 2, n):
        edge = rng.integers(0, 4)
        along = rng.uniform(-0.05, 1.05) * umax if rng.uniform() < 0.7 else rng.choice([0.0, umax]) + rng.normal(0, 0.15)
        off = rng.normal(0, 0.15)
        ut = {0: (along, off), 1: (off, along), 2: (along, umax + off), 3: (umax + off, along)}[int(edge)]
        hu0, hu1 = -(d * ut[0] + o * ut[1]), -(o * ut[0] + d * ut[1])
        base, rot = 0.5 * (hu0 + hu1), 0.5 * (hu1 - hu0)
        xs[i, 2] = rng.uniform(-1.0, 1.0)
        sn, cs = onp.sin(xs[i, 2]), onp.cos(xs[i, 2])
        lams[i, 3] = rng.normal(0, 5.0)
        lams[i, 4] = ((base + 2 * g / m * cs) * m + lams[i, 3] * sn) / cs
        lams[i, 5] = rot * I / r
    us = onp.zeros((n, 2)); uh = onp.zeros((n, 2))
    for i in range(n):
        x = torch.as_tensor(xs[i]); lam = torch.as_tensor(lams[i])
        us[i] = ref.u_star_2d(x, lam, pp, smooth=True).numpy()
        uh[i] = ref.u_star_2d(x, lam, pp).numpy()
    onp.savez(os.path.join(OUT, "ustar_smooth_flatquad_f64.npz"), x=xs, lam=lams, u_smooth=us, u=uh)
    d = onp.abs(us - uh).max(1)
    print("ustar_smooth: |smooth - hard| median %.3g, max %.3g, > 1e-3 at %d of %d points" % (onp.median(d), d.max(), (d > 1e-3).sum(), n))


if __name__ == "__main__":
    jnp = install_shim()
    only = sys.argv[1:]
    if not only or "rhs" in only:
        gen_rhs_flatquad(jnp, torch.float64)
        gen_rhs_flatquad(jnp, torch.float32)
        gen_rhs_2d(jnp, "orbits")
        gen_rhs_2d(jnp, "lqr")
    if not only or "traj" in only:
        gen_traj_flatquad(jnp)
    if not only or "vxx" in only:
        gen_vxx_flatquad(jnp)
    if not only or "smooth" in only:
        gen_ustar_smooth_flatquad(jnp)
