"""Parity of the value-Hessian branch (PMP_FLAG_VXX; algo_params['pontryagin_solver_vxx'],
pontryagin_utils.py:460-467, :508-519) of the CUDA path against the reference-generated goldens and the CPU
oracle.  Two device variants: the general 36-entry storage and the symmetric upper-triangle storage
(PMP_FLAG_VXX_SYMMETRIC); the oracle always integrates all 36 entries.

Bars as in test_gpu_parity.py: RK4 fp64 <= 1e-10 relative, adaptive <= 10x tol with identical integer outputs.
"""
import ctypes as C

import numpy as np
import pytest

from common import (GOLDEN, O, _capi, as_oracle_opts, c_problem, cuda_solve, flatquad_P, flatquad_terminal_vxx,
                    oracle_problem, scaled_err)

pytestmark = pytest.mark.gpu

VXX, SYM = _capi.FLAG_VXX, _capi.FLAG_VXX | _capi.FLAG_VXX_SYMMETRIC


def both(opts, y):
    prob, _ = c_problem("flatquad")
    got = cuda_solve(prob, opts, y)
    ref = O.solve_batch(oracle_problem("flatquad"), as_oracle_opts(opts), y)
    return got, ref


@pytest.mark.parametrize("tag,dtype,tol", [("f64", _capi.F64, 1e-11), ("f32", _capi.F32, 2e-4)])
def test_vxx_rhs_vs_reference_golden(tag, dtype, tol):
    """pmp_rhs_batch_cuda_flags(PMP_FLAG_VXX) against the reference's own f_extended with pontryagin_solver_vxx
    (golden made by tests/golden/make_golden.py::gen_vxx_flatquad; 192 points, all active sets)."""
    import torch
    g = np.load(f"{GOLDEN}/vxx_flatquad_f64.npz")
    prob, _ = c_problem("flatquad")
    dt = torch.float64 if dtype == _capi.F64 else torch.float32
    n = g["x"].shape[0]
    y = np.concatenate([g["x"].T, np.zeros((2, n)), g["lam"].T, g["vxx"].reshape(n, 36).T], 0)
    yd = torch.as_tensor(y, dtype=dt, device="cuda").contiguous()
    dy = torch.empty_like(yd)
    _capi.check(_capi.lib().pmp_rhs_batch_cuda_flags(C.byref(prob), dtype, VXX, n, yd.data_ptr(), dy.data_ptr(), None))
    torch.cuda.synchronize()
    got = dy.cpu().numpy().astype(np.float64)[14:].T.reshape(n, 6, 6)
    ref = g["vxx_dot"]
    rel = np.abs(got - ref).reshape(n, -1).max(1) / (1 + np.abs(ref).reshape(n, -1).max(1))
    if tag == "f64":
        assert rel.max() <= tol, rel.max()
    else:
        # fp32: a point within rounding of an active-set boundary may pick the other branch of du*/dz
        assert np.median(rel) <= 1e-5 and np.mean(rel <= tol) >= 0.97, (np.median(rel), np.sort(rel)[-8:])
    # the base leaves are those of the plain f_extended
    base = torch.empty((14, n), dtype=dt, device="cuda")
    _capi.check(_capi.lib().pmp_rhs_batch_cuda(C.byref(prob), dtype, n, yd[:14].contiguous().data_ptr(), base.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.allclose(dy[:14].cpu().numpy(), base.cpu().numpy(), rtol=1e-5 if tag == "f32" else 1e-12, atol=1e-4 if tag == "f32" else 1e-10)


@pytest.mark.parametrize("flags", [VXX, SYM])
def test_vxx_rk4_fp64(flags):
    """fixed-step RK4 fp64: <= 1e-10 relative (per-trajectory magnitude), as test_flatquad_rk4_fp64."""
    y = flatquad_terminal_vxx(2048)
    opts = _capi.make_opts(method="rk4", dtype=_capi.F64, adaptive=0, t0=0.0, t1=-3.0, dt0=-1 / 64, max_steps=192, flags=flags)
    got, ref = both(opts, y)
    assert (got["n_steps"] == 192).all() and (got["result"] == 0).all()
    for rows in (list(range(6)) + list(range(7, 14)), list(range(14, 50))):
        scale = np.abs(ref["y_end"][rows]).max(0)
        rel = (np.abs(got["y_end"][rows] - ref["y_end"][rows]) / scale).max(0)
        assert np.median(rel) <= 1e-12, np.median(rel)
        assert np.mean(rel <= 1e-10) >= 0.98, np.mean(rel <= 1e-10)
        assert rel.max() <= 1e-6, rel.max()


@pytest.mark.parametrize("method", ["tsit5", "dopri5"])
@pytest.mark.parametrize("flags,asym", [(VXX, 0.0), (SYM, 0.0), (VXX, 0.5)])
def test_vxx_adaptive_fp64(method, flags, asym):
    """fp64 adaptive: step sequences identical to the oracle (the 36 vxx entries are a leaf of the error norm),
    endpoints of all 50 rows within 10x tol.  asym: non-symmetric terminal Hessians on the general path."""
    tol = 1e-5
    y = flatquad_terminal_vxx(1 << 12, asym=asym)
    opts = _capi.make_opts(method=method, dtype=_capi.F64, rtol=tol, atol=tol, flags=flags)
    got, ref = both(opts, y)
    same = (got["n_steps"] == ref["n_steps"]) & (got["n_accepted"] == ref["n_accepted"])
    assert same.mean() >= 0.995, same.mean()
    assert (got["result"] == ref["result"]).mean() >= 0.999
    ok = same & (ref["result"] == 0)
    err = scaled_err(got["y_end"], ref["y_end"], tol)
    assert err[ok].max() <= 10.0, err[ok].max()
    assert np.percentile(err[ok], 99) <= 0.1
    if flags == SYM:  # mirrored on output
        V = got["y_end"][14:].T.reshape(-1, 6, 6)
        assert (V == V.transpose(0, 2, 1)).all()


def test_vxx_fp32_short_horizon():
    tol = 1e-5
    y = flatquad_terminal_vxx(1 << 12, dtype=np.float32)
    for flags in (VXX, SYM):
        opts = _capi.make_opts(method="tsit5", dtype=_capi.F32, rtol=tol, atol=tol, t1=-0.5, flags=flags)
        got, ref = both(opts, y)
        assert (got["result"] == ref["result"]).all()
        err = scaled_err(got["y_end"].astype(np.float64), ref["y_end"].astype(np.float64), tol)
        assert np.percentile(err, 99) <= 10.0, np.percentile(err, [50, 99, 100])
        assert (got["n_steps"] == ref["n_steps"]).mean() >= 0.9


@pytest.mark.parametrize("flags", [VXX, SYM])
def test_vxx_norm_event(flags):
    """DiscreteTerminatingEvent(||vxx||_F > vxx_max_norm) (pontryagin_utils.py:508-519): same result codes and
    step counts as the oracle."""
    y = flatquad_terminal_vxx(1 << 12)
    opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, flags=flags, event=_capi.EVENT_VXX_NORM, event_level=345.0)
    got, ref = both(opts, y)
    assert (ref["result"] == 1).any() and (ref["result"] == 0).any()
    assert (got["result"] == ref["result"]).mean() >= 0.999
    assert (got["n_steps"] == ref["n_steps"]).mean() >= 0.995
    fro = np.linalg.norm(got["y_end"][14:], axis=0)
    assert (fro[got["result"] == 1] > 345.0).all()


@pytest.mark.parametrize("flags", [VXX, SYM])
@pytest.mark.parametrize("n", [1, 33, 700])
def test_vxx_dense_layout_matches_oracle_fp64(flags, n):
    y = flatquad_terminal_vxx(n)
    opts = _capi.make_opts(method="tsit5", dtype=_capi.F64, save_dense=1, flags=flags)
    got, ref = both(opts, y)
    assert (got["n_accepted"] == ref["n_accepted"]).all()
    for k in ("ts", "x", "t", "v", "vx", "vxx"):
        a, b = got[k], ref[k]
        assert a.shape == b.shape, k
        assert (np.isinf(a) == np.isinf(b)).all(), k
        assert (np.sign(a[np.isinf(a)]) == np.sign(b[np.isinf(b)])).all(), k
        fin = np.isfinite(b)
        assert np.allclose(a[fin], b[fin], rtol=1e-5, atol=1e-5), (k, np.abs(a[fin] - b[fin]).max())
    # endpoint == ys[num_accepted_steps]
    i = np.arange(n)
    assert np.array_equal(got["vxx"][i, got["n_accepted"]].reshape(n, 36).T, got["y_end"][14:])


def test_vxx_writes_stay_inside_their_buffers():
    """guard bands around every output of the vxx kernels (ragged n, NaN lane, early max_steps)."""
    import torch
    from approximate_optimal_control_b200 import pontryagin_utils as pu
    prob, _ = c_problem("flatquad")
    G = 4096
    sentinel = -12345.0
    for dtype, tdt in ((_capi.F32, torch.float32), (_capi.F64, torch.float64)):
        for flags in (VXX, SYM):
            for n, max_steps in ((1, 128), (333, 128), (1031, 17)):
                y = flatquad_terminal_vxx(n)
                y[:, n // 2] = np.nan
                yd = torch.as_tensor(y, dtype=tdt, device="cuda").contiguous()
                s1 = max_steps + 1
                shapes = {"y_end": (50, n), "ts": (n, s1), "x": (n, s1, 6), "t": (n, s1), "v": (n, s1), "vx": (n, s1, 6),
                          "vxx": (n, s1, 6, 6)}
                big, out = {}, {}
                for k, shp in shapes.items():
                    m = int(np.prod(shp))
                    big[k] = torch.full((m + 2 * G,), sentinel, dtype=tdt, device="cuda")
                    out[k] = big[k][G:G + m].view(shp)
                ints = {}
                for k in ("n_accepted", "n_steps", "result"):
                    ints[k] = torch.full((n + 2 * G,), -777, dtype=torch.int32, device="cuda")
                    out[k] = ints[k][G:G + n]
                pu.solve_batch_soa(prob, _capi.make_opts(dtype=dtype, save_dense=1, max_steps=max_steps, flags=flags), yd, out)
                torch.cuda.synchronize()
                for k, b in big.items():
                    assert (b[:G] == sentinel).all() and (b[-G:] == sentinel).all(), (k, n, flags)
                    assert not (b[G:-G] == sentinel).any(), (k, n, flags)
                for k, b in ints.items():
                    assert (b[:G] == -777).all() and (b[-G:] == -777).all() and not (b[G:-G] == -777).any(), (k, n)
                assert int(out["result"][n // 2]) == _capi.RESULT_NAN


def test_vxx_public_api():
    """define_backward_solver with algo_params['pontryagin_solver_vxx'] = True: state dicts with a 'vxx' leaf in,
    Solution with ys['vxx'] (N, S+1, nx, nx) out; symmetric terminal Hessians select the triangular kernel;
    'vxx_max_norm' installs the terminating event; select_train_pts filters on ||vxx||_F."""
    import torch
    from approximate_optimal_control_b200 import levelsets, pontryagin_utils as pu, systems, terminal
    pp = systems.flatquad_problem_params()
    P = torch.as_tensor(0.5 * (flatquad_P() + flatquad_P().T))
    N = 512
    st = terminal.sample_terminal_states(flatquad_P(), 1e-3, N, seed=3, dtype=np.float64)
    yf = {k: torch.as_tensor(v, device="cuda") for k, v in st.items()}
    yf["vxx"] = P.to("cuda").expand(N, 6, 6)
    ap = {"pontryagin_solver_rtol": 1e-6, "pontryagin_solver_atol": 1e-6, "pontryagin_solver_T": 2.0,
          "pontryagin_solver_maxsteps": 128, "pontryagin_solver_vxx": True, "dtype": "float64"}
    solve, f_ext = pu.define_backward_solver(pp, ap)
    sol = solve(yf)
    assert sol.ys["vxx"].shape == (N, 129, 6, 6) and sol.y_end["vxx"].shape == (N, 6, 6)
    assert torch.equal(sol.ys["vxx"][:, 0], yf["vxx"])
    # general path on the same (symmetric) data: rounding-level agreement
    sol_g = pu.define_backward_solver(pp, dict(ap, pontryagin_solver_vxx_symmetric=False))[0](yf)
    same = (sol.stats["num_steps"] == sol_g.stats["num_steps"])
    assert same.float().mean() >= 0.99
    d = (sol.y_end["vxx"] - sol_g.y_end["vxx"]).abs().amax((1, 2)) / sol_g.y_end["vxx"].abs().amax((1, 2))
    assert float(d[same].max()) <= 1e-6
    # f_extended on a batch and on one state
    dy = f_ext(0.0, yf)
    assert dy["vxx"].shape == (N, 6, 6)
    one = f_ext(0.0, {k: v[0] for k, v in yf.items()})
    assert one["vxx"].shape == (6, 6) and torch.allclose(one["vxx"], dy["vxx"][0])
    # the event + the data selection after the path (levelsets.py:684-689)
    sol_e = pu.define_backward_solver(pp, dict(ap, vxx_max_norm=345.0))[0](yf)
    assert bool((sol_e.result == 1).any())
    sel = levelsets.select_train_pts((0.0, 50.0), sol, vxx_max_norm=345.0)
    assert sel["vxx"].shape[1:] == (6, 6) and sel["vxx"].shape[0] == sel["x"].shape[0] > 0
    assert float(torch.linalg.matrix_norm(sel["vxx"]).max()) < 345.0
