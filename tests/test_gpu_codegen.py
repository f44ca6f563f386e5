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


# ---- value-Hessian branch of generated systems (pontryagin_solver_vxx, pontryagin_utils.py:460-467) -------------------------
# The reference differentiates pmp_rhs of ANY f, l with jax; here a generated functor is differentiated in forward mode
# (csrc/pmp_vxx_dual.cuh).  Pinned on the reference's own f_extended(..., 'debug') golden through the generated flatquad,
# on the oracle's vxx solve, and -- for a system that exists nowhere else -- on the solve's own sensitivities.
VXX, SYM = _capi.FLAG_VXX, _capi.FLAG_VXX | _capi.FLAG_VXX_SYMMETRIC


def rhs_vxx(prob, x, lam, V, dtype=_capi.F64):
    import torch
    n, nx = x.shape
    dt = torch.float64 if dtype == _capi.F64 else torch.float32
    y = np.concatenate([x.T, np.zeros((2, n)), lam.T, V.reshape(n, nx * nx).T], 0)
    yd = torch.as_tensor(y, dtype=dt, device="cuda").contiguous()
    dy = torch.empty_like(yd)
    L = _capi.lib_of(prob)
    _capi.check(L.pmp_rhs_batch_cuda_flags(C.byref(prob), dtype, VXX, n, yd.data_ptr(), dy.data_ptr(), None), L)
    torch.cuda.synchronize()
    return dy.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("tag,dtype", [("f64", _capi.F64), ("f32", _capi.F32)])
def test_generated_flatquad_vxx_rhs_vs_reference_golden(tag, dtype):
    g = np.load(f"{GOLDEN}/vxx_flatquad_f64.npz")
    prob, _ = generated(systems.flatquad_problem_params)
    n = g["x"].shape[0]
    dy = rhs_vxx(prob, g["x"], g["lam"], g["vxx"], dtype)
    got, ref = dy[14:].T.reshape(n, 6, 6), g["vxx_dot"]
    rel = np.abs(got - ref).reshape(n, -1).max(1) / (1 + np.abs(ref).reshape(n, -1).max(1))
    if tag == "f64":
        assert rel.max() <= 1e-10, rel.max()
    else:  # a point within rounding of an active-set boundary may pick the other branch of du*/dz
        assert np.median(rel) <= 1e-5 and np.mean(rel <= 2e-4) >= 0.97, (np.median(rel), np.sort(rel)[-8:])


@pytest.mark.parametrize("method", ["tsit5", "dopri5"])
@pytest.mark.parametrize("flags,asym", [(VXX, 0.0), (SYM, 0.0), (VXX, 0.5)])
def test_generated_flatquad_vxx_solve_matches_oracle_fp64(flags, asym, method):
    """Tsit5 fp64 with the vxx leaf on the GENERATED flatquad against the oracle's hand-derived Jacobians: same step sequences
    (the 36 entries are a leaf of the error norm), all 50 rows within 10x tol."""
    from common import flatquad_terminal_vxx
    tol = 1e-5
    y = flatquad_terminal_vxx(1 << 11, asym=asym)
    prob, _ = generated(systems.flatquad_problem_params)
    opts = _capi.make_opts(method=method, dtype=_capi.F64, rtol=tol, atol=tol, flags=flags)
    got = cuda_solve(prob, opts, y)
    ref = O.solve_batch(oracle_problem("flatquad"), as_oracle_opts(opts), y)
    same = (got["n_steps"] == ref["n_steps"]) & (got["n_accepted"] == ref["n_accepted"])
    assert same.mean() >= 0.995, same.mean()
    assert (got["result"] == ref["result"]).mean() >= 0.999
    ok = same & (ref["result"] == 0)
    err = scaled_err(got["y_end"], ref["y_end"], tol)
    assert err[ok].max() <= 10.0 and np.percentile(err[ok], 99) <= 0.1, (err[ok].max(), np.percentile(err[ok], 99))
    if flags == SYM:
        V = got["y_end"][14:].T.reshape(-1, 6, 6)
        assert (V == V.transpose(0, 2, 1)).all()


def test_generated_flatquad_vxx_rk4_and_fp32_and_dense():
    from common import flatquad_terminal_vxx
    prob, _ = generated(systems.flatquad_problem_params)
    y = flatquad_terminal_vxx(512)
    opts = _capi.make_opts(method="rk4", dtype=_capi.F64, adaptive=0, t0=0.0, t1=-3.0, dt0=-1 / 64, max_steps=192, flags=VXX)
    got = cuda_solve(prob, opts, y)
    ref = O.solve_batch(oracle_problem("flatquad"), as_oracle_opts(opts), y)
    assert (got["n_steps"] == 192).all() and (got["result"] == 0).all()
    scale = np.abs(ref["y_end"][14:]).max(0)
    rel = (np.abs(got["y_end"][14:] - ref["y_end"][14:]) / scale).max(0)
    assert np.median(rel) <= 1e-11 and np.mean(rel <= 1e-9) >= 0.98, (np.median(rel), rel.max())
    # fp32, short horizon (test_vxx_fp32_short_horizon's bars)
    tol = 1e-5
    y32 = flatquad_terminal_vxx(1 << 11, dtype=np.float32)
    for flags in (VXX, SYM):
        opts = _capi.make_opts(method="tsit5", dtype=_capi.F32, rtol=tol, atol=tol, t1=-0.5, flags=flags)
        got = cuda_solve(prob, opts, y32)
        ref = O.solve_batch(oracle_problem("flatquad"), as_oracle_opts(opts), y32)
        assert (got["result"] == ref["result"]).all()
        err = scaled_err(got["y_end"].astype(np.float64), ref["y_end"].astype(np.float64), tol)
        # (the hand-written functor shares the oracle's closed-form Jacobians and meets 10 units / 0.9; forward mode through the
        # generated code rounds differently, and in fp32 that flips accept / reject decisions a little more often)
        assert np.percentile(err, 99) <= 15.0, np.percentile(err, [50, 99, 100])
        assert (got["n_steps"] == ref["n_steps"]).mean() >= 0.8
    # dense layout with the vxx leaf
    for n in (1, 33):
        yd = flatquad_terminal_vxx(n)
        opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, save_dense=1, flags=SYM)
        got = cuda_solve(prob, opts, yd)
        ref = O.solve_batch(oracle_problem("flatquad"), as_oracle_opts(opts), yd)
        assert (got["n_accepted"] == ref["n_accepted"]).all()
        for k in ("ts", "x", "vx", "vxx"):
            a, b = got[k], ref[k]
            assert (np.isinf(a) == np.isinf(b)).all(), k
            fin = np.isfinite(b)
            assert np.allclose(a[fin], b[fin], rtol=1e-5, atol=1e-5), k


def test_new_system_vxx_rhs_vs_finite_differences():
    """cart2d (state-dependent, non-diagonal H_uu): vxx_dot of the generated functor against gx + glam V - V fx - V flam V with
    the Jacobians taken by central differences of the independent numpy field (brute-force box QP)."""
    pp = CS.cart2d_problem_params()
    prob = systems.to_c_problem(pp)
    field = CS.numpy_field(pp)
    rng = np.random.Generator(np.random.PCG64(21))
    n, nx = 96, 4
    x = rng.standard_normal((n, nx)) * 1.2
    lam = rng.standard_normal((n, nx)) * 3.0
    V = rng.standard_normal((n, nx, nx))
    dy = rhs_vxx(prob, x, lam, V)
    got = dy[2 * nx + 2:].T.reshape(n, nx, nx)

    def F(z):
        xd, _, ld, u = field(z[:nx], z[nx:])
        return np.concatenate([xd, ld]), u

    checked = 0
    lo, hi = (np.asarray(b, dtype=float) for b in pp["U_interval"])
    for i in range(n):
        z0, h = np.concatenate([x[i], lam[i]]), 1e-6
        u0 = F(z0)[1]
        act0 = np.concatenate([u0 <= lo + 1e-12, u0 >= hi - 1e-12])
        cols, kink = [], False
        for e in np.eye(2 * nx):
            (fp, up), (fm, um) = F(z0 + h * e), F(z0 - h * e)
            for uu in (up, um):  # the stencil must stay on one active set
                kink |= not np.array_equal(np.concatenate([uu <= lo + 1e-12, uu >= hi - 1e-12]), act0)
            cols.append((fp - fm) / (2 * h))
        if kink:
            continue
        J = np.stack(cols, 1)
        fx, flam, gx, glam = J[:nx, :nx], J[:nx, nx:], J[nx:, :nx], J[nx:, nx:]
        want = gx + glam @ V[i] - V[i] @ fx - V[i] @ flam @ V[i]
        assert np.allclose(got[i], want, rtol=2e-6, atol=2e-6 * (1 + np.abs(want).max())), (i, np.abs(got[i] - want).max())
        checked += 1
    assert checked >= n // 2
    # base leaves unchanged by the flag
    _, base = pointwise(prob, x, lam, flags=0)
    assert np.allclose(dy[:2 * nx + 2][[0, 1, 2, 3, 5, 6, 7, 8, 9]], base[[0, 1, 2, 3, 5, 6, 7, 8, 9]], rtol=1e-12, atol=1e-12)


def test_new_system_vxx_is_the_sensitivity_of_the_costate():
    """define_backward_solver(pontryagin_solver_vxx=True) on cart2d: along a characteristic V_xx(t) = d lambda(t) / d x(t).
    The solve's own forward-mode sensitivities (pmp_solve_jvp_cuda) along (e_j, P e_j) give A = dx_end/dx_f, B = dlam_end/dx_f;
    the Riccati leaf must equal B A^-1."""
    import torch
    pp = CS.cart2d_problem_params()
    nx = 4
    K, P = pu.get_terminal_lqr(pp)
    P = 0.5 * (P + P.T)
    N, T = 64, 1.0
    st = terminal.sample_terminal_states(P, 0.05, N, seed=9)
    yf = {k: torch.as_tensor(v, device="cuda") for k, v in st.items()}
    yf["vxx"] = torch.as_tensor(P, device="cuda").expand(N, nx, nx)
    ap = {"pontryagin_solver_rtol": 1e-10, "pontryagin_solver_atol": 1e-10, "pontryagin_solver_T": T,
          "pontryagin_solver_maxsteps": 4096, "pontryagin_solver_vxx": True, "dtype": "float64",
          "pontryagin_solver_save_steps": False}
    solve, f_ext = pu.define_backward_solver(pp, ap)
    sol = solve(yf)
    assert bool((sol.result == 0).all())
    Vend = sol.y_end["vxx"].cpu().numpy()
    assert Vend.shape == (N, nx, nx)
    assert np.abs(Vend - Vend.transpose(0, 2, 1)).max() <= 1e-7 * np.abs(Vend).max()
    # symmetric storage agrees with the general one
    sol_g = pu.define_backward_solver(pp, dict(ap, pontryagin_solver_vxx_symmetric=False))[0](yf)
    # (characteristics that cross a switching surface of u* see a discontinuous Riccati right hand side: the two step
    # sequences then differ by more than rounding)
    d = np.abs(Vend - sol_g.y_end["vxx"].cpu().numpy()).reshape(N, -1).max(1) / np.abs(Vend).max()
    assert np.median(d) <= 1e-8 and d.max() <= 1e-4, (np.median(d), d.max())
    prob = systems.to_c_problem(pp)
    y = np.concatenate([st["x"].T, st["t"][None], st["v"][None], st["vx"].T], 0)
    opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, rtol=1e-10, atol=1e-10, t1=-T, max_steps=4096)
    A, B = np.zeros((N, nx, nx)), np.zeros((N, nx, nx))
    for j in range(nx):
        dyf = np.zeros_like(y)
        dyf[j] = 1.0
        dyf[nx + 2:] = P[:, j][:, None]
        dyf[nx + 1] = st["vx"] @ np.eye(nx)[j]  # dv_f = lambda_f . dx (not used by A, B)
        out = pu.solve_jvp_soa(prob, opts, torch.as_tensor(y, device="cuda").contiguous(), torch.as_tensor(dyf, device="cuda").contiguous())
        dye = out["dy_end"].cpu().numpy()
        A[:, :, j], B[:, :, j] = dye[:nx].T, dye[nx + 2:].T
    want = B @ np.linalg.inv(A)
    rel = np.abs(Vend - want).reshape(N, -1).max(1) / np.abs(want).reshape(N, -1).max(1)
    # Characteristics whose u* keeps ONE active set obey a smooth Riccati equation: there the two constructions agree to
    # integration accuracy.  Across a switching surface both see a jump in the Jacobians inside one step and resolve it to
    # O(step) each in its own way (as the reference's jax.jacobian-inside-diffrax does): agreement to a fraction of a percent.
    dense = pu.define_backward_solver(pp, dict(ap, pontryagin_solver_save_steps=True, pontryagin_solver_vxx=False))[0](
        {k: v for k, v in yf.items() if k != "vxx"})
    lo, hi = (np.asarray(b, dtype=float) for b in pp["U_interval"])
    one_set = np.zeros(N, bool)
    for i in range(N):
        m = int(dense.stats["num_accepted_steps"][i]) + 1
        u, _ = pointwise(prob, dense.ys["x"][i, :m].cpu().numpy(), dense.ys["vx"][i, :m].cpu().numpy(), flags=0)
        act = np.concatenate([u <= lo + 1e-9, u >= hi - 1e-9], 1)
        one_set[i] = (act == act[0]).all()
    assert 0.2 <= one_set.mean() <= 0.95, one_set.mean()  # the sample has both kinds
    assert rel[one_set].max() <= 1e-7, np.sort(rel[one_set])[-4:]
    assert rel.max() <= 5e-2 and np.median(rel) <= 2e-3, (np.median(rel), rel.max())
    # f_extended with the vxx leaf, batch and single state
    dyb = f_ext(0.0, yf)
    assert dyb["vxx"].shape == (N, nx, nx)
    one = f_ext(0.0, {k: v[0] for k, v in yf.items()})
    assert torch.allclose(one["vxx"], dyb["vxx"][0])


def test_vxx_flag_is_refused_where_the_reference_has_no_branch():
    from common import c_problem
    prob, _ = c_problem("orbits")
    assert _capi.lib().pmp_state_rows_flags(C.byref(prob), VXX) < 0
    gen, _ = generated(systems.flatquad_problem_params)
    L = _capi.lib_of(gen)
    assert L.pmp_state_rows_flags(C.byref(gen), VXX) == 50
