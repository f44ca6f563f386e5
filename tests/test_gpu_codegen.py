"""Plug-in dynamics on the GPU (codegen.py).

* The three systems of the reference, GENERATED from their f, l, must reproduce the reference goldens and the
  oracle exactly as the hand-written functors do (same bars as test_gpu_parity.py).
* Two systems that exist nowhere else (cart2d: nx=4, nu=2 with a state-dependent, non-diagonal H_uu; pendulum:
  nx=2, nu=1) are checked against an independent numpy evaluation of the sympy-derived right hand side
  (tests/codegen_systems.py: lambdify + brute-force box QP) integrated with scipy DOP853.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.integrate

from common import (GOLDEN, O, _capi, as_oracle_opts, cuda_solve, flatquad_terminal, lqr_terminal, oracle_problem,
                    orbits_terminal, scaled_err)
from approximate_optimal_control_b200 import pontryagin_utils as pu, systems, terminal
import codegen_systems as CS

pytestmark = pytest.mark.gpu


def generated(mk):
    pp = dict(mk())
    pp["codegen"] = True
    prob = systems.to_c_problem(pp)
    assert prob.system_id == _capi.SYS_GENERATED
    return prob, pp


def pointwise(prob, x, lam, dtype=_capi.F64, flags=None):
    """(u*, f_extended rows) of the library serving `prob` for points x, lam [n, nx]."""
    import torch
    n, nx = x.shape
    dt = torch.float64 if dtype == _capi.F64 else torch.float32
    xd = torch.as_tensor(x.T.copy(), dtype=dt, device="cuda").contiguous()
    ld = torch.as_tensor(lam.T.copy(), dtype=dt, device="cuda").contiguous()
    u = torch.empty((prob.nu, n), dtype=dt, device="cuda")
    L = _capi.lib_of(prob)
    if flags is None:
        _capi.check(L.pmp_ustar_batch_cuda(C.byref(prob), dtype, n, xd.data_ptr(), ld.data_ptr(), u.data_ptr(), None), L)
    else:
        _capi.check(L.pmp_ustar_batch_cuda_flags(C.byref(prob), dtype, flags, n, xd.data_ptr(), ld.data_ptr(), u.data_ptr(), None), L)
    y = torch.cat([xd, torch.zeros((2, n), dtype=dt, device="cuda"), ld]).contiguous()
    dy = torch.empty_like(y)
    _capi.check(L.pmp_rhs_batch_cuda(C.byref(prob), dtype, n, y.data_ptr(), dy.data_ptr(), None), L)
    torch.cuda.synchronize()
    return u.cpu().numpy().T.astype(np.float64), dy.cpu().numpy().astype(np.float64)


# ---- the reference's systems, generated ----------------------------------------------------------------------
def test_generated_flatquad_vs_reference_golden():
    g = np.load(f"{GOLDEN}/rhs_flatquad_f64.npz")
    prob, _ = generated(systems.flatquad_problem_params)
    u, dy = pointwise(prob, g["x"], g["lam"])
    assert np.abs(u - g["u"]).max() <= 1e-9
    n = g["x"].shape[0]
    ref = np.concatenate([g["dx"].T, np.ones((1, n)), g["dv"][None], g["dlam"].T], 0)
    assert (np.abs(dy - ref) / (1 + np.abs(ref))).max() <= 1e-11
    u2, _ = pointwise(prob, g["x"], g["lam"], flags=0)  # KKT selection
    assert np.abs(u2 - g["u"]).max() <= 1e-9


@pytest.mark.parametrize("name,mk", [("orbits", systems.orbits_problem_params), ("lqr", systems.lqr_statepenalty_problem_params)])
def test_generated_nu1_systems_vs_reference_golden(name, mk):
    g = np.load(f"{GOLDEN}/rhs_{name}_f64.npz")
    prob, _ = generated(mk)
    u, dy = pointwise(prob, g["x"], g["lam"])
    n = g["x"].shape[0]
    assert np.abs(u[:, 0] - g["u"].reshape(n)).max() <= 1e-14
    ref = np.concatenate([g["dx"].T, np.ones((1, n)), g["dv"][None], g["dlam"].T], 0)
    assert (np.abs(dy - ref) / (1 + np.abs(ref))).max() <= 1e-13


def test_generated_flatquad_solve_matches_oracle_and_handwritten_fp64():
    tol = 1e-5
    y = flatquad_terminal(1 << 12)
    prob, _ = generated(systems.flatquad_problem_params)
    opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, rtol=tol, atol=tol)
    got = cuda_solve(prob, opts, y)
    ref = O.solve_batch(oracle_problem("flatquad"), as_oracle_opts(opts), y)
    same = (got["n_steps"] == ref["n_steps"]) & (got["n_accepted"] == ref["n_accepted"])
    assert same.mean() >= 0.999 and (got["result"] == ref["result"]).all()
    err = scaled_err(got["y_end"], ref["y_end"], tol)
    assert err[same].max() <= 10.0 and np.percentile(err, 99) <= 0.1
    hand = cuda_solve(systems.to_c_problem(systems.flatquad_problem_params()), opts, y)
    both = same & (hand["n_steps"] == got["n_steps"])
    assert np.percentile(scaled_err(got["y_end"], hand["y_end"], tol)[both], 99) <= 0.1


def test_generated_lqr_rk4_fp64_c1():
    """BASELINE config 0 through the generated functor: <= 1e-10 relative to the oracle."""
    y = lqr_terminal(1024)
    prob, _ = generated(systems.lqr_statepenalty_problem_params)
    opts = _capi.make_opts(method="rk4", dtype=_capi.F64, adaptive=0, t0=0.0, t1=-8.0, dt0=-1 / 16, max_steps=128)
    got = cuda_solve(prob, opts, y)
    ref = O.solve_batch(oracle_problem("lqr"), as_oracle_opts(opts), y)
    assert (got["n_steps"] == 128).all() and (got["result"] == 0).all()
    rel = np.abs(got["y_end"] - ref["y_end"]) / np.maximum(np.abs(ref["y_end"]), 1e-300)
    assert rel.max() <= 1e-10, rel.max(1)


def test_generated_orbits_event_and_value_mode_fp64():
    y = orbits_terminal(1 << 12)
    prob, _ = generated(systems.orbits_problem_params)
    for kw in (dict(event=_capi.EVENT_VALUE_GE, event_level=4.2, t1=-50.0, dt0=-1 / 16),
               dict(indep=_capi.INDEP_VALUE, t1=4.2, dt0=1 / 16)):
        opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, rtol=1e-4, atol=1e-4, max_steps=1024, dtmin=0.0, dtmax=0.0, **kw)
        got = cuda_solve(prob, opts, y)
        ref = O.solve_batch(oracle_problem("orbits"), as_oracle_opts(opts), y)
        same = got["n_steps"] == ref["n_steps"]
        assert same.mean() >= 0.995 and (got["result"] == ref["result"]).mean() >= 0.999
        assert np.percentile(scaled_err(got["y_end"], ref["y_end"], 1e-4)[same], 99) <= 1.0


# ---- systems without a hand-written functor -------------------------------------------------------------------
@pytest.mark.parametrize("mk", [CS.cart2d_problem_params, CS.pendulum_problem_params])
def test_new_system_pointwise_vs_numpy(mk):
    pp = mk()
    prob = systems.to_c_problem(pp)
    field = CS.numpy_field(pp)
    rng = np.random.Generator(np.random.PCG64(11))
    n, nx = 256, pp["nx"]
    x = rng.standard_normal((n, nx)) * 1.5
    lam = rng.standard_normal((n, nx)) * np.array([4.0] * nx)
    u_lit, dy = pointwise(prob, x, lam)
    u_kkt, _ = pointwise(prob, x, lam, flags=0)
    sat = 0
    for i in range(n):
        xd, vd, ld, u = field(x[i], lam[i])
        assert np.allclose(u_lit[i], u, rtol=1e-10, atol=1e-10), (i, u_lit[i], u)
        assert np.allclose(u_kkt[i], u, rtol=1e-10, atol=1e-10), (i, u_kkt[i], u)
        assert np.allclose(dy[:nx, i], xd, rtol=1e-11, atol=1e-11)
        assert np.isclose(dy[nx + 1, i], vd, rtol=1e-11, atol=1e-11)
        assert np.allclose(dy[nx + 2:, i], ld, rtol=1e-11, atol=1e-10)
        lo, hi = (np.broadcast_to(np.asarray(b, dtype=float), u.shape) for b in pp["U_interval"])
        sat += int(((u <= lo + 1e-12) | (u >= hi - 1e-12)).any())
    assert 0.1 * n < sat < 0.95 * n  # the sample exercises both the interior and the constrained branches


@pytest.mark.parametrize("mk,T", [(CS.cart2d_problem_params, 1.5), (CS.pendulum_problem_params, 2.0)])
def test_new_system_trajectories_vs_scipy(mk, T):
    """define_backward_solver on a generated system, fp64 at a tight tolerance, against scipy DOP853 on the numpy
    field (u* from the brute-force box QP), from terminal states on the LQR level set."""
    import torch
    pp = mk()
    nx = pp["nx"]
    K, P = pu.get_terminal_lqr(pp)
    st = terminal.sample_terminal_states(P, 0.05, 8, seed=5)
    ap = {"pontryagin_solver_rtol": 1e-10, "pontryagin_solver_atol": 1e-10, "pontryagin_solver_T": T,
          "pontryagin_solver_maxsteps": 2048, "dtype": "float64"}
    solve, f_ext = pu.define_backward_solver(pp, ap)
    sol = solve({k: torch.as_tensor(v, device="cuda") for k, v in st.items()})
    assert bool((sol.result == 0).all())
    field = CS.numpy_field(pp)

    def rhs(t, z):
        xd, vd, ld, _ = field(z[:nx], z[nx + 1:])
        return np.concatenate([xd, [vd], ld])

    for i in range(8):
        z0 = np.concatenate([st["x"][i], [st["v"][i]], st["vx"][i]])
        ref = scipy.integrate.solve_ivp(rhs, (0.0, -T), z0, method="DOP853", rtol=1e-12, atol=1e-13).y[:, -1]
        got = np.concatenate([sol.y_end["x"][i].cpu().numpy(), [float(sol.y_end["v"][i])], sol.y_end["vx"][i].cpu().numpy()])
        assert np.allclose(got, ref, rtol=2e-7, atol=2e-8), (i, np.abs(got - ref).max())
    # dense output of the generated system through the same interpolation kernel
    mid = sol.evaluate(torch.full((8,), -0.5 * T, dtype=torch.float64, device="cuda"))
    assert mid["x"].shape == (8, nx) and bool(torch.isfinite(mid["x"]).all())


@pytest.mark.parametrize("mk,term,kw", [
    (systems.orbits_problem_params, orbits_terminal, dict(indep=_capi.INDEP_VALUE, t0=0.0, t1=1.0, dt0=1 / 64)),
    (CS.pendulum_problem_params, None, dict(t0=0.0, t1=-1.0, dt0=-0.05))])
def test_generated_systems_differentiate_like_handwritten_ones(mk, term, kw):
    """pmp_solve_jvp_cuda on GENERATED functors (dual numbers through the sympy-derived code, exp / tanh included): equal to
    the hand-written functor's tangents where one exists, and to central finite differences of the solve otherwise."""
    import torch
    prob, pp = generated(mk)
    n = 128
    rng = np.random.Generator(np.random.PCG64(5))
    if term is not None:
        y = term(n)
    else:
        y = np.concatenate([rng.uniform(-0.5, 0.5, (2, n)), np.zeros((1, n)), np.full((1, n), 0.1), rng.uniform(-0.5, 0.5, (2, n))], 0)
    dy = rng.standard_normal(y.shape)
    dy[prob.nx] = 0.0
    if kw.get("indep") == _capi.INDEP_VALUE:
        dy[prob.nx + 1] = 0.0
    opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, rtol=1e-9, atol=1e-9, max_steps=4096, dtmin=0.0, dtmax=0.0, **kw)
    yd, dyd = (torch.as_tensor(a, device="cuda").contiguous() for a in (y, dy))
    got = {k: v.cpu().numpy() for k, v in pu.solve_jvp_soa(prob, opts, yd, dyd).items()}
    assert (got["result"] == 0).all()
    if term is not None:  # same dynamics, hand-written
        from common import c_problem
        hand = {k: v.cpu().numpy() for k, v in pu.solve_jvp_soa(c_problem("orbits")[0], opts, yd, dyd).items()}
        same = hand["n_steps"] == got["n_steps"]
        assert same.mean() >= 0.95
        scale = np.abs(hand["dy_end"]).max(0) + 1e-300
        assert (np.abs(got["dy_end"] - hand["dy_end"]).max(0) / scale)[same].max() <= 1e-4
    eps = 1e-6
    a = cuda_solve(prob, opts, y + eps * dy)["y_end"]
    b = cuda_solve(prob, opts, y - eps * dy)["y_end"]
    fd = (a - b) / (2 * eps)
    rel = np.abs(got["dy_end"] - fd).max(0) / (np.abs(fd).max(0) + 1e-300)
    assert np.median(rel) <= 1e-4 and np.percentile(rel, 90) <= 2e-2, np.percentile(rel, [50, 90, 100])
