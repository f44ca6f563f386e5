"""Device-side mirrors of the data-selection helpers that sit right after the hot path in
levelsets.testbed (/root/reference/levelsets.py): they consume the dense `Solution` the backward
solve returns and would otherwise round-trip ~8 GB of saved nodes through the host.

    define_solve_backward_lqr(problem_params, algo_params, P)   levelsets.py:576-601  (the vmapped caller of the path)
    select_train_pts(value_interval, sols[, vxx_max_norm])      levelsets.py:667-709
    find_min_l(ys, v_lower, v_upper, problem_params)            levelsets.py:604-628
    forward_sim_lqr(problem_params, algo_params, P_lqr)(x0)     levelsets.py:818-836
    forward_sim_nn(problem_params, algo_params, ensemble)(x0)   levelsets.py:838-870   (the step BEFORE the backward solve
    forward_sim_nn_until_value(..., ensemble, v_k)(x0)          levelsets.py:873-933    in batched_oracle, :1201-1215)

Array work is torch on the GPU; the running cost l(x, u*(x, lambda)) comes from the CUDA f_extended
kernel (v' = -l, pontryagin_utils.py:449), i.e. from pmp_rhs_batch_cuda.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi, systems
from . import pontryagin_utils as pu


class data_normaliser(object):
    """nn_utils.py:112-178: per-component affine scaling of x and v to zero mean / unit variance and the induced scalings
    of vx and vxx.  The reference computes the statistics with four passes over the training set; here they come out of the
    selection kernel (sums accumulated while the nodes are being compacted): select_train_pts(..., return_normaliser=True)."""

    def __init__(self, train_ode_states=None, *, x_means=None, x_stds=None, v_mean=None, v_std=None):
        if train_ode_states is not None:  # the reference's constructor: statistics of a given data set (torch, on its device)
            x, v = train_ode_states["x"], train_ode_states["v"]
            x_means, x_stds = x.mean(0), x.std(0, unbiased=False)
            v_mean, v_std = v.mean(), v.std(unbiased=False)
        self.x_means, self.x_stds, self.v_mean, self.v_std = x_means, x_stds, v_mean, v_std

    def normalise_x(self, x): return (x - self.x_means) / self.x_stds
    def unnormalise_x(self, xn): return xn * self.x_stds + self.x_means
    def normalise_v(self, v): return (v - self.v_mean) / self.v_std
    def unnormalise_v(self, vn): return vn * self.v_std + self.v_mean
    def normalise_vx(self, vx): return vx * self.x_stds / self.v_std
    def unnormalise_vx(self, vx_n): return (vx_n / self.x_stds) * self.v_std
    def normalise_vxx(self, vxx): return self.x_stds[..., :, None] * vxx * self.x_stds[..., None, :] / self.v_std
    def unnormalise_vxx(self, vxx_n): return (vxx_n * self.v_std) / self.x_stds[..., :, None] / self.x_stds[..., None, :]

    def normalise_all_dict(self, train_ode_states):
        oup = {"x": self.normalise_x(train_ode_states["x"]), "v": self.normalise_v(train_ode_states["v"]),
               "vx": self.normalise_vx(train_ode_states["vx"])}
        if "vxx" in train_ode_states:
            oup["vxx"] = self.normalise_vxx(train_ode_states["vxx"])
        return oup


def select_train_pts(value_interval, sols, vxx_max_norm=None, return_normaliser=False, verbose=True):
    """levelsets.py:667-709: all saved nodes whose value lies in [v_lower, v_upper] (inf/NaN padding drops out of the
    comparison), as a dict of flat arrays {'x','t','v','vx'[,'vxx']} in the order of the reference's boolean-mask gather.
    With a 'vxx' leaf (pontryagin_solver_vxx) nodes with ||vxx||_F >= vxx_max_norm are dropped as well (:684-689; the
    reference reads algo_params['vxx_max_norm'] from its closure, here it is an argument).

    Two kernels over the device-resident Solution (csrc/pmp_select.cu): a per-trajectory count, then a stable compaction
    fused with the data_normaliser's sums (nn_utils.py:119-134) and the NaN check; only the n_accepted + 1 saved nodes of a
    trajectory are read.  return_normaliser=True also returns the data_normaliser of the selected set."""
    v_lower, v_upper = value_interval
    ys = sols.ys
    v = ys["v"]
    if v.dim() == 1:  # a single trajectory
        ys = {k: a[None] for k, a in ys.items()}
        v = ys["v"]
    n, s1 = v.shape
    nx = ys["x"].shape[-1]
    dev, dtype = v.device, v.dtype
    has_vxx = "vxx" in ys
    if has_vxx and vxx_max_norm is None:
        raise ValueError("sols carries a 'vxx' leaf: pass vxx_max_norm (algo_params['vxx_max_norm'])")
    L = _capi.lib()
    code = pu._code(dtype)
    leaves = {k: ys[k].contiguous() for k in (("x", "t", "v", "vx", "vxx") if has_vxx else ("x", "t", "v", "vx"))}
    nacc = sols.stats["num_accepted_steps"].reshape(-1).to(torch.int32).contiguous() if sols.stats is not None else None
    vxx_p = leaves["vxx"].data_ptr() if has_vxx else None
    nvxx = nx * nx if has_vxx else 0
    vmax = float(vxx_max_norm) if has_vxx else 0.0
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    stream = pu._stream_ptr(dev)
    with torch.cuda.device(dev):
        _capi.check(L.pmp_select_count_cuda(code, n, s1, leaves["v"].data_ptr(), vxx_p, nvxx, nacc.data_ptr() if nacc is not None else None,
                                            float(v_lower), float(v_upper), vmax, counts.data_ptr(), stream), L)
        ends = torch.cumsum(counts, 0, dtype=torch.int64)  # (the one library call of this path: a prefix sum over n counts)
        offsets = (ends - counts).contiguous()
        m = int(ends[-1]) if n else 0
        mm = max(m, 1)  # (an empty selection still hands valid pointers to the C entry)
        out = {"x": torch.empty((mm, nx), dtype=dtype, device=dev), "t": torch.empty(mm, dtype=dtype, device=dev),
               "v": torch.empty(mm, dtype=dtype, device=dev), "vx": torch.empty((mm, nx), dtype=dtype, device=dev)}
        if has_vxx:
            out["vxx"] = torch.empty((mm, nx, nx), dtype=dtype, device=dev)
        stats = torch.zeros(2 * nx + 2, dtype=torch.float64, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        _capi.check(L.pmp_select_write_cuda(code, n, s1, nx, leaves["x"].data_ptr(), leaves["t"].data_ptr(), leaves["v"].data_ptr(),
                                            leaves["vx"].data_ptr(), vxx_p, nvxx, nacc.data_ptr() if nacc is not None else None,
                                            offsets.data_ptr(), float(v_lower), float(v_upper), vmax, out["x"].data_ptr(),
                                            out["t"].data_ptr(), out["v"].data_ptr(), out["vx"].data_ptr(),
                                            out["vxx"].data_ptr() if has_vxx else None, stats.data_ptr(), flags.data_ptr(), stream), L)
    out = {k: a[:m] for k, a in out.items()}
    if verbose:
        n_fin = int((nacc.long() + 1).clamp(max=s1).sum()) if nacc is not None else int(torch.isfinite(v).sum())
        print(f"full (train+test) dataset size: {m} points (= {100.0 * m / max(n_fin, 1):.2f}% of valid points)")
    if int(flags) & 1:
        raise RuntimeError("There are still NaNs in training data.")
    if not return_normaliser:
        return out
    st = stats / max(m, 1)
    mean_x, mean_v = st[:nx], st[2 * nx]
    var_x = (st[nx:2 * nx] - mean_x * mean_x).clamp(min=0)
    var_v = (st[2 * nx + 1] - mean_v * mean_v).clamp(min=0)
    norm = data_normaliser(x_means=mean_x.to(dtype), x_stds=var_x.sqrt().to(dtype), v_mean=mean_v.to(dtype), v_std=var_v.sqrt().to(dtype))
    return out, norm


def running_cost(ys: dict, problem_params: dict) -> torch.Tensor:
    """l(x, u*(x, vx)) at every saved node, shape of ys['v']; padding slots give NaN."""
    problem = systems.to_c_problem(problem_params)
    nx = problem.nx
    x, vx = ys["x"], ys["vx"]
    shape = x.shape[:-1]
    dev, dtype = x.device, x.dtype
    n = x.numel() // nx
    y = torch.zeros((2 * nx + 2, n), dtype=dtype, device=dev)
    y[:nx] = x.reshape(n, nx).t()
    y[nx + 2:] = vx.reshape(n, nx).t()  # (a 'vxx' leaf does not enter the running cost)
    dy = torch.empty_like(y)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib_of(problem).pmp_rhs_batch_cuda(C.byref(problem), pu._code(dtype), n, y.data_ptr(), dy.data_ptr(),
                                                   pu._stream_ptr(dev)))
    l = -dy[nx + 1]
    return torch.where(torch.isfinite(l), l, torch.full_like(l, float("nan"))).reshape(shape)


def find_min_l(ys: dict, v_lower, v_upper, problem_params: dict):
    """levelsets.py:604-628: smallest running cost l(x, u*) over the saved nodes in the value band."""
    all_ls = running_cost(ys, problem_params)
    inside = (ys["v"] >= v_lower) & (ys["v"] <= v_upper) & ~torch.isnan(all_ls)
    if not bool(inside.any()):
        return torch.tensor(float("nan"), dtype=all_ls.dtype, device=all_ls.device)
    return all_ls[inside].min()


def define_solve_backward_lqr(problem_params: dict, algo_params: dict, P_lqr):
    """levelsets.py:576-601: `solve_backward_lqr(x_f, algo_params)` -- the function the reference vmaps over its terminal
    samples, `sols_orig = jax.vmap(solve_backward_lqr, in_axes=(0, None))(xfs, algo_params)`.  It builds the terminal state
    from x_f alone, v_f = 1/2 x_f' P_lqr x_f, vx_f = P_lqr x_f, t = 0 (:585-592, deviations from x_eq), and solves backward.

    Returns solve_backward_lqr(xfs): xfs (N, nx) [or (nx,)] -> Solution.  Host xfs (numpy / CPU tensor) of >= 2^16 samples
    with algo_params['pontryagin_solver_save_steps'] False: only xfs crosses the host link (pmp_solve_host_lqr forms v_f and
    vx_f on the device) and the endpoints come back in pinned host tensors.  Otherwise the terminal state is formed with
    torch on the device and handed to define_backward_solver's solve_backward (dense Solution)."""
    problem = systems.to_c_problem(problem_params)
    nx = problem.nx
    solve_backward, _ = pu.define_backward_solver(problem_params, algo_params)
    x_eq = problem_params.get("x_eq")

    def solve_backward_lqr(x_f, algo_params_=None):
        ap = algo_params if algo_params_ is None else algo_params_
        save_steps = bool(ap.get("pontryagin_solver_save_steps", True))
        with_vxx = bool(ap.get("pontryagin_solver_vxx", False))
        dev = pu._device(x_f)
        dtype = pu._infer_dtype(ap, x_f)
        if pu._host_leaf(x_f) and np.ndim(x_f) == 2 and not save_steps and not with_vxx and np.shape(x_f)[0] >= (1 << 16):
            T = float(ap["pontryagin_solver_T"])
            method = ap.get("pontryagin_solver_method", pu._DEFAULTS["method"])
            adaptive = bool(ap.get("pontryagin_solver_adaptive", True)) and str(method).lower() != "rk4"
            opts = pu._solver_opts(ap, dtype, t0=0.0, t1=-T, dt0=-abs(float(ap.get("pontryagin_solver_dt0", pu._DEFAULTS["dt0"]))),
                                   adaptive=int(adaptive), save_dense=0)
            sol = pu._solve_host_lqr(problem, opts, nx, x_f, P_lqr, x_eq, dtype, dev, int(ap.get("pontryagin_solver_host_chunks", 0)))
            pu.raise_if_failed(sol.result, ap)
            return sol
        x = pu._as_tensor(x_f, dtype, dev)
        Pt = pu._as_tensor(P_lqr, dtype, dev)
        dx = x if x_eq is None else x - pu._as_tensor(x_eq, dtype, dev)
        vx = dx @ Pt.T
        state_f = {"x": x, "t": torch.zeros(x.shape[:-1], dtype=dtype, device=dev), "v": 0.5 * (dx * vx).sum(-1), "vx": vx}
        if with_vxx:
            state_f["vxx"] = Pt
        return solve_backward(state_f)

    return solve_backward_lqr


def _forward_sim(problem_params, algo_params, *, P=None, nn=None, event=None, level=0.0):
    from .solution import Solution
    problem = systems.to_c_problem(problem_params)
    nx = problem.nx

    def forward_sim(x0):
        dev = pu._device(x0) if nn is None else nn.weights.device
        dtype = pu._infer_dtype(algo_params, x0) if nn is None else nn.weights.dtype
        xt = pu._as_tensor(x0, dtype, dev)
        batched = xt.dim() == 2
        xs = xt.reshape(-1, nx).t().contiguous()
        T = float(algo_params.get("forward_sim_T", 10.0))
        over = {} if event is None else dict(event=event, event_level=float(level))
        opts = pu._solver_opts(algo_params, dtype, t0=0.0, t1=T, dt0=0.1, adaptive=1, save_dense=1,
                               dtmin=float(algo_params.get("forward_sim_dtmin", 0.05)), dtmax=0.0, **over)
        out = pu.forward_batch_soa(problem, opts, xs, P, problem_params.get("x_eq"), nn=nn)
        pu.raise_if_failed(out["result"], algo_params, "forward_sim")  # levelsets.py:834,867,929: throw=algo_params['throw']
        stats = {"num_steps": out["n_steps"], "num_accepted_steps": out["n_accepted"],
                 "num_rejected_steps": out["n_steps"] - out["n_accepted"]}
        sol = Solution(t0=0.0, t1=T, ts=out["ts"], ys=out["x"], y_end={"x": out["z_end"][:nx].t(), "cost": out["z_end"][nx]},
                       stats=stats, result=out["result"], batched=True)
        sol.extra["cost"] = out["cost"]
        return sol if batched else sol[0]

    return forward_sim


def forward_sim_lqr(problem_params: dict, algo_params: dict, P_lqr, v_k=None):
    """levelsets.py:818-836: closed-loop forward simulation x' = f(x, u_star_2d(x, P_lqr (x - x_eq))) with
    Tsit5 + PIDController(rtol, atol, dtmin=.05), t: 0 -> 10, dt0 = .1, SaveAt(steps, t0), max_steps, throw=False.
    Returns forward_sim(x0): x0 (nx,) or (N, nx) [replaces jax.vmap] -> Solution with ys (N, S+1, nx), ts, stats, result,
    y_end = {'x', 'cost'}.  v_k: stop once 1/2 dx'P dx <= v_k (result == 1).  Optional algo_params: 'forward_sim_T' (10),
    'forward_sim_dtmin' (.05)."""
    return _forward_sim(problem_params, algo_params, P=P_lqr, event=None if v_k is None else _capi.EVENT_LQR_VALUE_LE,
                        level=0.0 if v_k is None else v_k)


def forward_sim_nn(problem_params: dict, algo_params: dict, ensemble):
    """levelsets.py:838-870 (vmap=True): as forward_sim_lqr with lambda(x) = gradient of the MEAN of the value-network
    ensemble (`ensemble`: nn_costate.NnEnsemble, built from the flax params and the data normaliser)."""
    return _forward_sim(problem_params, algo_params, nn=ensemble)


def forward_sim_nn_until_value(problem_params: dict, algo_params: dict, ensemble, v_k):
    """levelsets.py:873-933: additionally stops (result == 1) once the state is very likely inside the value level set v_k:
    v_mean + 2 v_std <= v_k and v_std <= 0.5 over the ensemble (ensemble.n_sigma, ensemble.sigma_max)."""
    return _forward_sim(problem_params, algo_params, nn=ensemble, event=_capi.EVENT_NN_VALUE_UCB_LE, level=v_k)
