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
    gen = dict(pp, codegen=True)
    sol_g = levelsets.forward_sim_lqr(gen, ap, P)(x0)
    assert float((sol_g.stats["num_steps"] == sol.stats["num_steps"]).float().mean()) >= 0.98
    same = sol_g.stats["num_steps"] == sol.stats["num_steps"]
    assert float((sol_g.y_end["cost"] - sol.y_end["cost"]).abs()[same].max()) <= 1e-6


def test_closed_loop_eval_general_with_a_user_feedback_law():
    """eval_utils.closed_loop_eval_general (eval_utils.py:48-99) with a caller-supplied u*(x): the law is traced and compiled
    into a plug-in (codegen.py).  (1) The LQR law of the double integrator written out as such a function reproduces
    closed_loop_eval_lqr on the hand-written functor; (2) a saturated linear feedback on the pendulum (a system that exists
    nowhere else) against scipy on the same closed loop."""
    import scipy.integrate
    import torch
    import codegen_systems as CS
    from common import lqr_P
    pp = systems.lqr_statepenalty_problem_params()
    P = np.round(lqr_P(), 12)  # (rounded: the law's constants are part of the plug-in's content hash, prebuilt by build())
    ev = {"sim_T": 4.0, "sim_dt": 1 / 32, "dtype": "float64"}
    rng = np.random.Generator(np.random.PCG64(2))
    x0 = rng.uniform(-1.5, 1.5, (64, 2))
    a = eval_utils.closed_loop_eval_lqr(pp, ev, P, x0)
    b = eval_utils.closed_loop_eval_general(pp, ev, CS.lqr_feedback_law(P), x0)
    assert b.ys.shape == a.ys.shape == (64, 128, 3)
    assert float((a.ys - b.ys).abs().max()) <= 1e-11 * float(a.ys.abs().max())
    assert abs(float(eval_utils.compute_controlcost(pp, a)) - float(eval_utils.compute_controlcost(pp, b))) <= 1e-10
    # pendulum under u = clip(-k.x): constant-step Tsit5 (dt = 1/64) against DOP853 at 1e-12
    pq = CS.pendulum_problem_params()
    law = CS.pendulum_feedback_law()
    x0 = rng.uniform(-1.0, 1.0, (8, 2))
    fnp, lnp, unp = CS.numpy_closed_loop(pq, law)
    refs = [scipy.integrate.solve_ivp(lambda t, z: np.concatenate([fnp(z[:2]), [lnp(z[:2])]]), (0.0, 2.0), np.concatenate([x0[i], [0.0]]),
                                      method="DOP853", rtol=1e-12, atol=1e-13).y[:, -1] for i in range(8)]
    errs = []
    for steps in (128, 512):  # the constant step crosses the saturation kink with an O(dt^2) error: it must shrink with dt
        ev = {"sim_T": 2.0, "sim_dt": 2.0 / steps, "dtype": "float64"}
        sol = eval_utils.closed_loop_eval_general(pq, ev, law, x0)
        assert sol.ys.shape == (8, steps, 3) and bool((sol.result == 0).all())
        got = sol.ys[:, -1].cpu().numpy()
        errs.append(max(np.abs(got[i] - refs[i]).max() / (1 + np.abs(refs[i]).max()) for i in range(8)))
    assert errs[0] <= 5e-4 and errs[1] <= 2e-5 and errs[1] <= errs[0] / 4, errs
    sat = np.abs(np.array([unp(x) for x in x0])) >= 2.0 - 1e-12
    assert sat.any() and not sat.all()  # the sample starts on both sides of the saturation


# ---- value-network ensemble as the costate law (levelsets.py:838-933, eval_utils.py:10-45) -------------------------
def _ensemble(nx, dims, E, dtype, seed=0, scale=1.0):
    """Random flax-layout parameters (trained weights come out of the reference's training loop, which is out of scope):
    parity of the simulation does not depend on what the networks represent -- as long as the closed loop is well
    conditioned: with 3x larger random weights u* chatters between the edges of the input box, a third of the attempts
    is rejected and two runs of ANY implementation disagree by hundreds of tolerance units."""
    import torch
    from approximate_optimal_control_b200 import nn_costate as NC
    p = NC.init_flax_params(seed, nx, dims, E, scale=scale)
    norm = NC.DataNormaliser(x_means=np.linspace(-0.1, 0.1, nx), x_stds=np.linspace(0.5, 1.5, nx), v_mean=2.0, v_std=5.0)
    ens = NC.NnEnsemble.from_flax(p, norm, "cuda", dtype)
    host = NC.NnEnsemble.from_flax(p, norm, "cpu", dtype).weights.numpy().copy()
    onn = O.nn_struct(host, nx, dims[0], len(dims), E, norm.x_means, norm.x_stds, norm.v_mean, norm.v_std)
    return ens, onn, host


@pytest.mark.parametrize("nx,dims,E", [(6, (64, 64, 64), 8), (2, (32, 32), 5)])
def test_nn_value_batch_vs_oracle(nx, dims, E):
    import torch
    prob, _ = c_problem("flatquad" if nx == 6 else "orbits")
    x = np.random.Generator(np.random.PCG64(4)).standard_normal((1000, nx))
    ens, onn, _w = _ensemble(nx, dims, E, torch.float64)
    lam, vm, vs = ens.meanstd(prob, x)
    rl, rm, rs = O.nn_costate(onn, x.T)
    assert np.allclose(lam.cpu().numpy(), rl.T, rtol=1e-10, atol=1e-11)
    assert np.allclose(vm.cpu().numpy(), rm, rtol=1e-11, atol=1e-11) and np.allclose(vs.cpu().numpy(), rs, rtol=1e-9, atol=1e-11)
    ens32, _, _w32 = _ensemble(nx, dims, E, torch.float32)
    l32, m32, s32 = ens32.meanstd(prob, x)
    assert np.allclose(l32.cpu().numpy(), rl.T, rtol=2e-4, atol=2e-4) and np.allclose(m32.cpu().numpy(), rm, rtol=1e-5, atol=1e-4)
    one = ens.meanstd(prob, x[7])
    assert one[0].shape == (nx,) and torch.allclose(one[0], lam[7])


@pytest.mark.parametrize("cluster", ["0", "1"])
def test_forward_nn_ensemble_fp64_vs_oracle(cluster, monkeypatch):
    """both kernels: lane groups (default) and thread-block clusters with distributed shared memory"""
    import torch
    monkeypatch.setenv("B200PMP_NN_CLUSTER", cluster)
    tol = 1e-6
    prob, _ = c_problem("flatquad")
    ens, onn, _w = _ensemble(6, (64, 64, 64), 8, torch.float64)
    x0 = 0.5 * np.random.Generator(np.random.PCG64(8)).standard_normal((6, 256))
    opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, rtol=tol, atol=tol, t0=0.0, t1=1.5, dt0=0.1, dtmin=0.0, dtmax=0.0,
                           max_steps=256, save_dense=1)
    out = pu.forward_batch_soa(prob, opts, torch.as_tensor(x0, device="cuda").contiguous(), nn=ens)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    ref = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(opts), x0, nn=onn)
    same = (got["n_steps"] == ref["n_steps"]) & (got["n_accepted"] == ref["n_accepted"])
    assert same.mean() >= 0.99 and (got["result"] == ref["result"]).all() and (ref["result"] == 0).all()
    assert scaled_err(got["z_end"], ref["z_end"], tol)[same].max() <= 10.0
    fin = np.isfinite(ref["x"]) & same[:, None, None]
    assert (np.isinf(got["x"]) == np.isinf(ref["x"]))[same].all() and np.allclose(got["x"][fin], ref["x"][fin], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dims,E", [((64, 64, 64), 8), ((64, 64), 5), ((64,), 3)])
def test_forward_nn_ensemble_fp32_tensor_core_kernels(dims, E, monkeypatch):
    """fp32 ensembles of width 64: the tcgen05 cluster kernel (default, pmp_forward_umma.cuh), the warp-MMA kernel
    (B200PMP_NN_UMMA=0, pmp_forward_tc.cuh) and the CUDA-core kernels (B200PMP_NN_TC=0) against the oracle.  In fp32 the closed
    loop at rtol = atol = 1e-5 amplifies rounding to tens of tolerance units and flips accept/reject decisions (the fp32 oracle
    itself is 5-8 units from the converged fp64 solution at the median, 100-150 at p99; scripts/diag_nn_kernels.py), so the gate
    is the one of the backward solve: every kernel is as close to the CONVERGED solution as the fp32 oracle is, percentile by
    percentile -- which is what split precision (3xTF32) has to preserve -- and the tensor-core kernels repeat the oracle's
    step sequence as often as the CUDA-core kernel with exact exp / log1p does."""
    import torch
    tol = 1e-5
    prob, _ = c_problem("flatquad")
    ens, onn, _w = _ensemble(6, dims, E, torch.float32, seed=3)
    _e64, onn64, _w64 = _ensemble(6, dims, E, torch.float64, seed=3)
    x0 = 0.5 * np.random.Generator(np.random.PCG64(11)).standard_normal((6, 1000))
    mk = lambda dtype, t: _capi.make_opts(method="tsit5", dtype=dtype, rtol=t, atol=t, t0=0.0, t1=1.5, dt0=0.1, dtmin=0.0,
                                           dtmax=0.0, max_steps=256)
    opts = mk(_capi.F32, tol)
    truth = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(mk(_capi.F64, 1e-11)), x0, nn=onn64)
    ref = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(opts), x0.astype(np.float32), nn=onn)
    floor = scaled_err(ref["z_end"], truth["z_end"], tol)  # what fp32 arithmetic alone does to the answer
    runs, same_frac = {}, {}
    for name, env in (("tcgen05", {}), ("warp-mma", {"B200PMP_NN_UMMA": "0"}), ("cuda-core", {"B200PMP_NN_TC": "0"})):
        for k in ("B200PMP_NN_UMMA", "B200PMP_NN_TC"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        out = pu.forward_batch_soa(prob, opts, torch.as_tensor(x0, dtype=torch.float32, device="cuda").contiguous(), nn=ens)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in out.items()}
        runs[name] = got
        assert (got["result"] == 0).all() and (ref["result"] == 0).all()
        same_frac[name] = ((got["n_steps"] == ref["n_steps"]) & (got["n_accepted"] == ref["n_accepted"])).mean()
        err = scaled_err(got["z_end"], truth["z_end"], tol)
        for pc, slack in ((50, 1.25), (90, 1.25), (99, 1.5)):
            assert np.percentile(err, pc) <= slack * np.percentile(floor, pc) + 0.5, (name, pc, np.percentile(err, pc), np.percentile(floor, pc))
        assert abs(int(got["n_steps"].sum()) - int(ref["n_steps"].sum())) <= 0.01 * ref["n_steps"].sum(), name
    assert same_frac["cuda-core"] >= 0.7 and min(same_frac["tcgen05"], same_frac["warp-mma"]) >= same_frac["cuda-core"] - 0.03, same_frac
    # run-to-run: the tcgen05 kernel is deterministic (sums in rank order, one issue order)
    for k in ("B200PMP_NN_UMMA", "B200PMP_NN_TC"):
        monkeypatch.delenv(k, raising=False)
    out2 = pu.forward_batch_soa(prob, opts, torch.as_tensor(x0, dtype=torch.float32, device="cuda").contiguous(), nn=ens)
    torch.cuda.synchronize()
    assert np.array_equal(out2["z_end"].cpu().numpy(), runs["tcgen05"]["z_end"])


@pytest.mark.parametrize("cluster", ["0", "1"])
def test_forward_nn_until_value_and_public_api(cluster, monkeypatch):
    import torch
    monkeypatch.setenv("B200PMP_NN_CLUSTER", cluster)
    from approximate_optimal_control_b200 import nn_costate as NC
    pp, ap = systems.flatquad_problem_params(), dict(systems.flatquad_algo_params(), forward_sim_T=1.5)
    ens, onn, _w = _ensemble(6, (64, 64, 64), 8, torch.float32, seed=2)
    x0 = 0.5 * np.random.Generator(np.random.PCG64(9)).standard_normal((200, 6))
    sol = levelsets.forward_sim_nn(pp, ap, ens)(x0)
    assert sol.ys.shape == (200, ap["pontryagin_solver_maxsteps"] + 1, 6) and bool((sol.result == 0).all())
    # the ensemble's upper confidence band along the trajectories decides where forward_sim_nn_until_value stops
    # (the event is evaluated at the step ends = the saved nodes: v_k = median over the trajectories of the smallest
    # band any of their nodes reaches, so about half of them stop)
    nodes = sol.ys.reshape(-1, 6)
    fin = torch.isfinite(nodes).all(1)
    lam, vm, vs = ens.meanstd(systems.to_c_problem(pp), nodes[fin])
    ens.sigma_max = float(vs.max()) * 2
    ucb = torch.full((nodes.shape[0],), float("inf"), device="cuda")
    ucb[fin] = vm + 2 * vs
    ucb[:: sol.ys.shape[1]] = float("inf")  # slot 0 (t0) is not an event point
    v_k = float(ucb.reshape(200, -1).min(1).values.median()) + 1e-3
    stop = levelsets.forward_sim_nn_until_value(pp, ap, ens, v_k)(x0)
    res = stop.result.cpu().numpy()
    assert (res == 1).any() and (res == 0).any()
    l2, m2, s2 = ens.meanstd(systems.to_c_problem(pp), stop.y_end["x"])
    assert bool(((m2 + 2 * s2)[stop.result == 1] <= v_k + 1e-4).all())
    onn.sigma_max = ens.sigma_max
    o = _capi.make_opts(dtype=_capi.F32, rtol=ap["pontryagin_solver_rtol"], atol=ap["pontryagin_solver_atol"], t0=0.0, t1=1.5,
                        dt0=0.1, dtmin=0.05, dtmax=0.0, max_steps=ap["pontryagin_solver_maxsteps"], event=_capi.EVENT_NN_VALUE_UCB_LE,
                        event_level=v_k)
    ref = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(o), x0.T.astype(np.float32), nn=onn)
    assert (res == ref["result"]).mean() >= 0.97
    # eval_utils.closed_loop_eval_nn_ensemble on a nu = 1 system: raw input gradients (apply_grad), constant step
    po = systems.orbits_problem_params()
    fl = NC.init_flax_params(5, 2, (32, 32), 4, scale=2.0)
    xo = np.array([0.0, 1.0]) + 0.2 * np.random.Generator(np.random.PCG64(3)).standard_normal((64, 2))
    ev = {"sim_T": 2.0, "sim_dt": 1 / 32, "dtype": "float64", "nn_ensemble_size": 4}
    all_sols = eval_utils.closed_loop_eval_nn_ensemble(po, ev, None, fl, xo)
    assert all_sols.ys.shape == (64, 64, 3)
    host = NC.NnEnsemble.from_flax(fl, NC.DataNormaliser.identity(2), "cpu", torch.float64).weights.numpy().copy()
    r2 = O.forward_batch(O.orbits_problem(), O.make_opts(dtype=O.F64, adaptive=0, t0=0.0, t1=2.0, dt0=1 / 32, max_steps=64, dtmin=0.0,
                                                         dtmax=0.0), xo.T, nn=O.nn_struct(host, 2, 32, 2, 4))
    assert abs(float(eval_utils.compute_controlcost(po, all_sols)) - r2["z_end"][2].mean()) <= 1e-9 * abs(r2["z_end"][2].mean())
