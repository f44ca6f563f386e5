"""Forward closed-loop simulation (pmp_forward_batch_cuda; levelsets.py:818-836 forward_sim_lqr and
eval_utils.py:48-99 closed_loop_eval_general with the LQR costate law) against the CPU oracle."""
import numpy as np
import pytest

from common import (O, _capi, as_oracle_opts, c_problem, flatquad_P, flatquad_terminal, lqr_P, oracle_problem, scaled_err)
from approximate_optimal_control_b200 import eval_utils, levelsets, pontryagin_utils as pu, systems

pytestmark = pytest.mark.gpu


def cuda_forward(prob, opts, x0, P, x_eq=None):
    import torch
    dt = torch.float32 if opts.dtype == _capi.F32 else torch.float64
    out = pu.forward_batch_soa(prob, opts, torch.as_tensor(np.ascontiguousarray(x0), dtype=dt, device="cuda").contiguous(), P, x_eq)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def x0_flatquad(n, scale=40.0, seed=1):
    return scale * flatquad_terminal(n, seed=seed)[:6]  # on the LQR level set V = 1e-3 * scale^2


@pytest.mark.parametrize("method", ["tsit5", "dopri5"])
def test_forward_flatquad_adaptive_fp64(method):
    tol = 1e-5
    x0 = x0_flatquad(1 << 12)
    prob, _ = c_problem("flatquad")
    opts = _capi.make_opts(method=method, dtype=_capi.F64, rtol=tol, atol=tol, t0=0.0, t1=10.0, dt0=0.1, dtmin=0.05, dtmax=0.0,
                           max_steps=256)
    got = cuda_forward(prob, opts, x0, flatquad_P())
    ref = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(opts), x0, flatquad_P())
    same = (got["n_steps"] == ref["n_steps"]) & (got["n_accepted"] == ref["n_accepted"])
    assert same.mean() >= 0.995 and (got["result"] == ref["result"]).all()
    err = scaled_err(got["z_end"], ref["z_end"], tol)
    assert err[same].max() <= 10.0 and np.percentile(err, 99) <= 0.1
    assert np.abs(got["z_end"][:6]).max() < 0.05 * np.abs(x0).max()  # the LQR feedback lands the quadrotor


def test_forward_rk4_constant_step_fp64_and_fp32():
    x0 = x0_flatquad(1024)
    prob, _ = c_problem("flatquad")
    for dtype, bar in ((_capi.F64, 1e-10), (_capi.F32, 2e-3)):
        opts = _capi.make_opts(method="rk4", dtype=dtype, adaptive=0, t0=0.0, t1=4.0, dt0=1 / 64, max_steps=256, dtmin=0.0, dtmax=0.0)
        got = cuda_forward(prob, opts, x0, flatquad_P())
        ref = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(opts), x0, flatquad_P())
        assert (got["n_steps"] == 256).all() and (got["result"] == 0).all()
        scale = np.abs(ref["z_end"]).max(0)
        rel = (np.abs(got["z_end"].astype(np.float64) - ref["z_end"]) / scale).max(0)
        assert np.percentile(rel, 98) <= bar, np.percentile(rel, [50, 98, 100])


def test_forward_event_dense_layout_and_lqr_system():
    x0 = x0_flatquad(333)
    prob, _ = c_problem("flatquad")
    opts = _capi.make_opts(dtype=_capi.F64, t0=0.0, t1=10.0, dt0=0.1, dtmin=0.05, dtmax=0.0, max_steps=200, save_dense=1,
                           event=_capi.EVENT_LQR_VALUE_LE, event_level=0.2)
    got = cuda_forward(prob, opts, x0, flatquad_P())
    ref = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(opts), x0, flatquad_P())
    assert (ref["result"] == 1).all() and (got["result"] == ref["result"]).all() and (got["n_accepted"] == ref["n_accepted"]).all()
    for k in ("ts", "x", "cost"):
        a, b = got[k], ref[k]
        assert (np.isinf(a) == np.isinf(b)).all(), k
        fin = np.isfinite(b)
        assert np.allclose(a[fin], b[fin], rtol=1e-6, atol=1e-7), k
    # a NaN initial state retires at once
    x0[:, 5] = np.nan
    got = cuda_forward(prob, opts, x0, flatquad_P())
    assert got["result"][5] == _capi.RESULT_NAN and got["n_steps"][5] == 0 and (got["result"][:5] == 1).all()
    # two-state system with x_eq != 0 (orbits, lambda = 2 P0 (x - x_eq))
    from common import orbits_P0, orbits_terminal
    po, _ = c_problem("orbits")
    xo = orbits_terminal(256, radius=0.3)[:2]
    oo = _capi.make_opts(dtype=_capi.F64, t0=0.0, t1=5.0, dt0=0.1, rtol=1e-7, atol=1e-7, dtmin=0.0, dtmax=0.0, max_steps=512)
    g2 = cuda_forward(po, oo, xo, 2 * orbits_P0(), [0.0, 1.0])
    r2 = O.forward_batch(oracle_problem("orbits"), as_oracle_opts(oo), xo, 2 * orbits_P0(), [0.0, 1.0])
    assert (g2["n_steps"] == r2["n_steps"]).mean() >= 0.99
    assert np.percentile(scaled_err(g2["z_end"], r2["z_end"], 1e-7), 99) <= 1.0


def test_forward_public_api_and_generated_functor():
    """levelsets.forward_sim_lqr / eval_utils.closed_loop_eval_lqr (the reference's call shapes); the functor generated
    from f, l runs the same kernel in its plug-in library."""
    import torch
    pp, ap = systems.flatquad_problem_params(), dict(systems.flatquad_algo_params(), dtype="float64")
    P = flatquad_P()
    x0 = x0_flatquad(128).T.copy()
    sol = levelsets.forward_sim_lqr(pp, ap, P)(x0)
    S1 = ap["pontryagin_solver_maxsteps"] + 1
    assert sol.ys.shape == (128, S1, 6) and sol.ts.shape == (128, S1) and bool((sol.result == 0).all())
    one = levelsets.forward_sim_lqr(pp, ap, P)(x0[3])
    assert one.ys.shape == (S1, 6) and torch.allclose(one.ys[: int(one.stats["num_accepted_steps"]) + 1],
                                                      sol.ys[3, : int(sol.stats["num_accepted_steps"][3]) + 1])
    stop = levelsets.forward_sim_lqr(pp, ap, P, v_k=0.2)(x0)
    assert bool((stop.result == 1).all()) and bool((stop.stats["num_steps"] < sol.stats["num_steps"]).all())
    # eval_utils: constant step, cost in the last column of ys, SaveAt(steps=True) -> T/dt slots
    ev = dict(ap, sim_T=4.0, sim_dt=1 / 32)
    all_sols = eval_utils.closed_loop_eval_lqr(pp, ev, P, x0)
    assert all_sols.ys.shape == (128, 128, 7)
    ref = O.forward_batch(O.flatquad_problem(), O.make_opts(dtype=O.F64, adaptive=0, t0=0.0, t1=4.0, dt0=1 / 32, max_steps=128,
                                                            dtmin=0.0, dtmax=0.0), x0.T, P)
    mean_cost = float(eval_utils.compute_controlcost(pp, all_sols))
    assert abs(mean_cost - ref["z_end"][6].mean()) <= 1e-8 * abs(mean_cost)
    # infinite-horizon cost-to-go from the LQR level set V = 1.6 under the LQR law: a little above V (u saturates)
    assert 1.5 < mean_cost < 2.5
    with pytest.raises(NotImplementedError):
        eval_utils.closed_loop_eval_general(pp, ev, lambda x: x, x0)
    gen = dict(pp, codegen=True)
    sol_g = levelsets.forward_sim_lqr(gen, ap, P)(x0)
    assert float((sol_g.stats["num_steps"] == sol.stats["num_steps"]).float().mean()) >= 0.98
    same = sol_g.stats["num_steps"] == sol.stats["num_steps"]
    assert float((sol_g.y_end["cost"] - sol.y_end["cost"]).abs()[same].max()) <= 1e-6
