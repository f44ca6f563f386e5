"""CPU tests of the host-side mirror of the reference interface (no GPU compute)."""
import os
import sys

import numpy as np
import pytest
import torch

from common import ROOT, _capi, flatquad_P
from approximate_optimal_control_b200 import pontryagin_utils as pu, sharding, systems, terminal


def test_flatquad_linearisation_closed_form():
    pp = systems.flatquad_problem_params()
    A, B, Q, R = systems.linearise(pp)
    m, g, r = 20.0, 9.81, 0.5
    I = m * (r / 2) ** 2
    A0 = np.zeros((6, 6)); A0[0, 3] = A0[1, 4] = A0[2, 5] = 1; A0[3, 2] = -g
    B0 = np.zeros((6, 2)); B0[4, :] = 1 / m; B0[5, 0] = -r / I; B0[5, 1] = r / I
    q = [1 / 0.3 ** 2, 1 / 0.3 ** 2, 1 / np.deg2rad(30) ** 2, 4.0, 4.0, 1 / np.deg2rad(120) ** 2]
    Q0 = 2 * np.diag(q); Q0[2, 2] += 2 * g * g
    R0 = 2 * np.array([[1 / m ** 2 + r ** 2 / I ** 2, 1 / m ** 2 - r ** 2 / I ** 2]] * 2); R0[1] = R0[1][::-1]
    for got, want in ((A, A0), (B, B0), (Q, Q0), (R, R0)):
        assert np.allclose(got, want, atol=1e-12)
    # H_uu = [[.325, -.315], [-.315, .325]]
    assert np.allclose(R, [[0.325, -0.315], [-0.315, 0.325]])
    # finite-difference fallback agrees
    A2, B2, Q2, R2 = systems._linearise_fd(pp["f"], pp["l"], pp["x_eq"], pp["u_eq"])
    assert np.allclose(A2, A, atol=1e-6) and np.allclose(B2, B, atol=1e-6)
    assert np.allclose(Q2, Q, rtol=1e-4, atol=1e-4) and np.allclose(R2, R, atol=1e-5)


def test_get_terminal_lqr_known_values():
    """P11 = 31.336, P22 = 21.773, P33 = 297.19, cond(P) = 93.4; stable closed loop."""
    pp = systems.flatquad_problem_params()
    K, P = pu.get_terminal_lqr(pp)
    assert np.allclose(np.diag(P)[:3], [31.336, 21.773, 297.19], rtol=1e-4)
    assert abs(np.linalg.cond(P) - 93.4) < 0.1
    A, B, Q, R = systems.linearise(pp)
    poles = np.linalg.eigvals(A - B @ K)
    assert (poles.real < 0).all() and np.isclose(poles.real.min(), -2.04, atol=0.01)
    assert K.shape == (2, 6)
    bad = dict(pp, u_eq=np.zeros(2))
    with pytest.raises(AssertionError):
        pu.get_terminal_lqr(bad)
    with pytest.raises(ValueError):
        pu.lqr(np.eye(2), np.zeros((2, 1)), np.eye(2), np.eye(1))  # unstable, uncontrollable


def test_terminal_samples_on_level_set():
    P = flatquad_P()
    st = terminal.sample_terminal_states(P, 1e-3, 1000, seed=1)
    v = 0.5 * np.einsum("ni,ij,nj->n", st["x"], P, st["x"])
    assert np.allclose(v, 1e-3) and np.allclose(st["v"], 1e-3)           # levelsets.py:562
    assert np.allclose(st["vx"], st["x"] @ P.T) and (st["t"] == 0).all()  # levelsets.py:585-592
    st2 = terminal.sample_terminal_states(P, 1e-3, 1000, seed=1)
    assert np.array_equal(st["x"], st2["x"])


def test_to_c_problem_and_errors():
    p = systems.to_c_problem(systems.flatquad_problem_params())
    assert (p.system_id, p.nx, p.nu) == (0, 6, 2) and np.isclose(p.u_hi[0], 117.72) and p.params[0] == 20.0
    p = systems.to_c_problem(systems.orbits_problem_params())
    assert (p.system_id, p.nx, p.nu, p.u_lo[0], p.u_hi[0]) == (1, 2, 1, -0.2, 0.2)
    with pytest.raises(NotImplementedError):
        systems.to_c_problem({"system_name": "pendulum", "nx": 2, "nu": 1, "U_interval": [-1, 1]})
    pp = systems.flatquad_problem_params()
    pu.define_backward_solver(pp, dict(systems.flatquad_algo_params(), pontryagin_solver_vxx=True))  # flatquad: built
    with pytest.raises(NotImplementedError):  # the Jacobians through u* exist for the flatquad functor only
        pu.define_backward_solver(systems.orbits_problem_params(), {"pontryagin_solver_vxx": True})
    with pytest.raises(NotImplementedError):
        pu.u_star_2d(np.zeros(2), np.zeros(2), systems.orbits_problem_params())  # nu != 2, as the reference
    with pytest.raises(AssertionError):
        pu.u_star_new(np.zeros(6), np.zeros(6), pp)  # nu != 1, as the reference


def test_host_f_l_match_oracle_rhs_pieces():
    """The numpy f, l of systems.py (kept for callers that evaluate costs on results) agree with the
    oracle's f_extended at u = u*."""
    from common import O
    pp = systems.flatquad_problem_params()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 6)); lam = rng.standard_normal((50, 6)) * 5
    u = O.ustar_batch(O.flatquad_problem(), O.F64, x.T, lam.T).T
    dy = O.rhs_batch(O.flatquad_problem(), O.F64, np.concatenate([x.T, np.zeros((2, 50)), lam.T], 0))
    assert np.allclose(pp["f"](x, u), dy[:6].T, atol=1e-12)
    assert np.allclose(-pp["l"](x, u), dy[7], atol=1e-10)


def test_shard_indices_cover_everything():
    for n, w in ((10, 2), (11, 4), (1 << 10, 8), (5, 8)):
        for inter in (False, True):
            idx = torch.cat([sharding.shard_indices(n, r, w, inter) for r in range(w)])
            real = idx[idx >= 0]
            assert sorted(real.tolist()) == list(range(n))
            assert idx.numel() == w * sharding.shard_size(n, w)


def _worker(rank, world, n, inter, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(3 * n, dtype=torch.float64).reshape(3, n)
    local = sharding.take_shard(full, n, rank, world, inter)
    # "solve": every rank transforms its own shard, no communication
    local = torch.where(torch.isnan(local), local, local * 2 + 1)
    out = sharding.allgather_endpoints(local, n, inter)
    q.put((rank, torch.equal(out, full * 2 + 1), tuple(out.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,inter", [(64, False), (64, True), (37, False), (37, True)])
def test_sharded_allgather_world2_gloo(n, inter):
    """world_size-2 run on CPU (gloo) of the N>1 path: shard by index -> local work -> ONE all-gather
    restores the global endpoint array in the original order on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n + int(inter)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, n, inter, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for (_, ok, _) in res), res
    assert all(shape == (3, n) for (_, _, shape) in res)


# ---- host logic of the rows around the path ---------------------------------------------------------------------
def test_symnp_matches_numpy_on_numbers_and_traces_symbols():
    """symnp (the jax.numpy stand-in the dynamics are traced with) falls through to numpy on numbers and builds sympy
    expressions on symbols; both views of the same function agree."""
    import sympy as sp
    from approximate_optimal_control_b200 import symnp
    rng = np.random.Generator(np.random.PCG64(0))

    def g(x, u):  # jax.numpy style, as the reference writes its dynamics
        M = symnp.array([[symnp.cos(x[2]), -symnp.sin(x[2])], [symnp.sin(x[2]), symnp.cos(x[2])]])
        v = M @ symnp.array([x[0], x[1]])
        return symnp.concatenate([v, symnp.array([symnp.maximum(0, x[0] - 1) ** 2 + symnp.exp(-u[0] ** 2) + symnp.tanh(x[1])])])

    xs, us = sp.symbols("a0:3"), sp.symbols("b0:1")
    expr = g(symnp.SymArray(xs), symnp.SymArray(us))
    f = sp.lambdify([xs, us], [sp.sympify(e.item() if isinstance(e, np.ndarray) else e) for e in expr], modules="numpy")
    for _ in range(20):
        x, u = rng.standard_normal(3) * 2, rng.standard_normal(1)
        assert np.allclose(np.asarray(f(x, u), dtype=float), np.asarray(g(x, u), dtype=float), rtol=1e-13, atol=1e-13)
    # jax's 0-d array semantics on symbols: x[0] keeps .reshape() / .T, reshape() without arguments gives a scalar
    s = symnp.SymArray(xs)
    assert (s[0] ** 2).reshape((1,)).shape == (1,) and (s[:1] * 2).reshape().shape == ()


def test_nn_ensemble_packing_and_normaliser():
    """flax parameter pytrees -> the flat layout of pmp_nn_costate_t (K1 [nx][W], b1, hidden {K [W][W], b}, Kout [W], bout,
    3 pad reals per member); data_normaliser statistics and their inverses."""
    import torch
    from approximate_optimal_control_b200 import nn_costate as NC
    nx, W, E = 6, 32, 3
    p = NC.init_flax_params(1, nx, (W, W), E)
    rng = np.random.Generator(np.random.PCG64(1))
    data = {"x": rng.standard_normal((500, nx)) * np.arange(1, nx + 1), "v": rng.standard_normal(500) * 3 + 7}
    norm = NC.DataNormaliser(data)
    assert np.allclose(norm.normalise_x(data["x"]).mean(0), 0, atol=1e-12) and np.allclose(norm.normalise_x(data["x"]).std(0), 1)
    assert np.allclose(norm.unnormalise_v(norm.normalise_v(data["v"])), data["v"])
    vx = rng.standard_normal((5, nx))
    assert np.allclose(norm.unnormalise_vx(norm.normalise_vx(vx)), vx)
    ens = NC.NnEnsemble.from_flax(p, norm, "cpu", torch.float64)
    stride = nx * W + W + (W * W + W) + W + 4
    w = ens.weights.numpy().reshape(E, stride)
    L = p["params"]
    for e in range(E):
        assert np.array_equal(w[e, :nx * W].reshape(nx, W), L["Dense_0"]["kernel"][e])
        assert np.array_equal(w[e, nx * W:nx * W + W], L["Dense_0"]["bias"][e])
        o = nx * W + W
        assert np.array_equal(w[e, o:o + W * W].reshape(W, W), L["Dense_1"]["kernel"][e])
        assert np.array_equal(w[e, o + W * W + W:o + W * W + 2 * W], L["Dense_2"]["kernel"][e][:, 0])
        assert w[e, o + W * W + 2 * W] == L["Dense_2"]["bias"][e][0] and (w[e, -3:] == 0).all()
    c = ens.to_c()
    assert (c.n_nets, c.n_hidden_layers, c.width) == (E, 2, W) and np.isclose(c.x_std[2], norm.x_stds[2]) and c.x_std[10] == 1.0
    with pytest.raises(NotImplementedError):
        NC.NnEnsemble.from_flax(NC.init_flax_params(0, nx, (48,), 2), norm, "cpu")  # width not in {32, 64}
    with pytest.raises(NotImplementedError):
        NC.NnEnsemble.from_flax(NC.init_flax_params(0, nx, (32, 64), 2), norm, "cpu")  # ragged hidden widths


def test_throw_semantics_host_logic():
    """algo_params['throw'] (pontryagin_utils.py:517,524): True raises on result codes 2/3/4 naming the first failing
    trajectory; event terminations (1) and throw=False never raise."""
    import torch
    from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu
    r = torch.tensor([0, 1, 0, 2, 4], dtype=torch.int32)
    pu.raise_if_failed(r, {"throw": False})
    pu.raise_if_failed(r, {})
    pu.raise_if_failed(torch.tensor([0, 1, 1], dtype=torch.int32), {"throw": True})
    with pytest.raises(_capi.PmpError, match=r"trajectory 3 of 5 .* result 2.*max_steps.*2 trajectories"):
        pu.raise_if_failed(r, {"throw": True})


def test_dtype_inference_follows_jax_x64_switch():
    """jax computes float64 numpy inputs in fp32 unless jax_enable_x64 is set (flatquad_landing_experiment.py:26-27 vs
    experiment_orbits.py:18-19): same rule here; algo_params['dtype'] and float64 torch tensors are explicit requests."""
    import torch
    from approximate_optimal_control_b200 import pontryagin_utils as pu
    x64, x32 = np.zeros(3), np.zeros(3, np.float32)
    assert pu._infer_dtype({}, x64) == torch.float32 and pu._infer_dtype({}, x32) == torch.float32
    assert pu._infer_dtype({"dtype": "float64"}, x32) == torch.float64
    assert pu._infer_dtype({}, torch.zeros(3, dtype=torch.float64)) == torch.float64
    pu.enable_x64()
    try:
        assert pu._infer_dtype({}, x64) == torch.float64 and pu._infer_dtype({}, x32) == torch.float32
    finally:
        pu.enable_x64(False)


def test_pinned_outputs_layout_serves_the_pitched_copies(monkeypatch):
    """pontryagin_utils._pinned_outputs carves the endpoint leaves of a host solve out of ONE allocation so that csrc/pmp_host.cu
    can ship a piece as two pitched copies: x -> vx one pitch apart, and (fp32) t -> v -> n_accepted equally spaced, every leaf
    256-byte aligned, no leaf overlapping the next.  (Pinned memory needs a CUDA context: the allocator is stubbed here.)"""
    import torch
    from approximate_optimal_control_b200 import pontryagin_utils as pu
    real_empty = torch.empty

    def fake_empty(*a, **k):
        k.pop("pin_memory", None)
        return real_empty(*a, **k)
    monkeypatch.setattr(torch, "empty", fake_empty)
    for n, nx, dt in ((1000, 6, torch.float32), (100003, 6, torch.float32), (40001, 2, torch.float64), (1, 4, torch.float32)):
        ho = pu._pinned_outputs(n, nx, dt)
        es = 4 if dt == torch.float32 else 8
        assert ho["x"].shape == (n, nx) and ho["vx"].shape == (n, nx) and ho["t"].shape == (n,) and ho["result"].dtype == torch.int32
        base = ho["x"].data_ptr()
        order = ["x", "vx", "t", "v", "n_accepted", "n_steps", "result"]
        for a, b in zip(order, order[1:]):
            assert (ho[b].data_ptr() - base) % 256 == 0
            assert ho[b].data_ptr() >= ho[a].data_ptr() + ho[a].numel() * ho[a].element_size(), (a, b)
        assert ho["vx"].data_ptr() - ho["x"].data_ptr() >= n * nx * es
        pitch_tv = ho["v"].data_ptr() - ho["t"].data_ptr()
        assert pitch_tv >= n * es
        if es == 4:  # the packed counters travel as a third row into the n_accepted leaf
            assert ho["n_accepted"].data_ptr() - ho["v"].data_ptr() == pitch_tv
        for k in order:  # views of one block: writing one leaf leaves the others alone
            ho[k].zero_()
        ho["t"].fill_(1.0)
        assert float(ho["v"].sum()) == 0.0 and float(ho["vx"].abs().sum()) == 0.0 and int(ho["n_accepted"].sum()) == 0
