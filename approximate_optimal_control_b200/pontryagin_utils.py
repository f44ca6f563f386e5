"""Host-side mirror of the reference's `pontryagin_utils` module for the PMP backward-shooting path.

Same names, argument meaning and error behaviour as /root/reference/pontryagin_utils.py:

    define_backward_solver(problem_params, algo_params) -> (solve_backward, f_extended)   :414-529
    u_star_2d(x, costate, problem_params, smooth=False, debug_oups=False)                 :11-254
    u_star_new(x, costate, problem_params)                                                :532-562
    lqr(A, B, Q, R) -> (K, X, eigVals)                                                    :569-599
    get_terminal_lqr(problem_params) -> (K_lqr, P_lqr)                                    :603-628
    make_pontryagin_solver_reparam(problem_params, algo_params) -> solver(y0, v0, v1)
        (absent from the snapshot; contract from experiment_orbits.py:117,181-184,
         experiment_lqr_statepenalty.py:119,147, rrt_sampler.py:62-77)

Differences, all forced by the move from "jax closure traced on CPU" to "compiled CUDA kernel":
  * the arithmetic runs in libb200pmp.so (hand-written sm_100a CUDA) through ctypes; arrays are
    torch CUDA tensors (numpy / CPU tensors are accepted and copied to the current CUDA device);
  * `jax.vmap(solve_backward)` is replaced by passing a leading batch axis: every function accepts
    either one state (the reference's shapes) or a batch;
  * dynamics are compiled functors: the three systems of the reference's experiments are hand-written
    (problem_params['system_name'] selects one); any other f, l -- or problem_params['codegen'] = True --
    goes through codegen.py (sympy derivatives in place of jax autodiff, nvcc, a plug-in library);
  * there is no CPU fallback: without the library or a CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import scipy.linalg
import torch

from . import _capi, systems
from .solution import Solution, StepStats

__all__ = ["define_backward_solver", "u_star_2d", "u_star_new", "lqr", "get_terminal_lqr",
           "make_pontryagin_solver_reparam", "solve_batch_soa", "forward_batch_soa", "pack_state", "unpack_state"]

# solver constants hard-coded in the reference (pontryagin_utils.py:500-501,515); overridable through
# optional algo_params keys of the same name prefixed with 'pontryagin_solver_'
_DEFAULTS = dict(dtmin=0.0005, dtmax=0.5, dt0=0.1, safety=0.9, factormin=0.2, factormax=10.0,
                 force_dtmin=1, method="tsit5")


# ---- tensor plumbing --------------------------------------------------------------------------------
def _device(*xs) -> torch.device:
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise _capi.PmpError("no CUDA device visible: the PMP path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_tensor(x, dtype, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


_X64 = False


def enable_x64(on: bool = True) -> None:
    """Mirror of jax.config.update("jax_enable_x64", True) (experiment_orbits.py:18-19): without it jax computes in fp32
    whatever the dtype of the numpy inputs (flatquad_landing_experiment.py:26-27 leaves it commented out)."""
    global _X64
    _X64 = bool(on)


def _infer_dtype(algo_params, *xs) -> torch.dtype:
    """algo_params['dtype'] ('float32'/'float64') wins.  Else, like jax: float64 NUMPY / python inputs are computed in
    fp32 unless enable_x64() was called (jax downcasts them; the reference's flatquad path runs fp32 although
    numpy defaults to float64), while an explicit float64 torch tensor -- which cannot arise by accident -- selects fp64."""
    d = (algo_params or {}).get("dtype")
    if d is not None:
        return {"float32": torch.float32, "float64": torch.float64, torch.float32: torch.float32,
                torch.float64: torch.float64, np.float32: torch.float32, np.float64: torch.float64}[d]
    for x in xs:
        if isinstance(x, torch.Tensor) and x.dtype == torch.float64:
            return torch.float64
        if _X64 and isinstance(x, np.ndarray) and x.dtype == np.float64:
            return torch.float64
    return torch.float32


_RESULT_TEXT = {2: "the maximum number of solver steps was reached (max_steps)",
                3: "the minimum step size was reached (dt_min)",
                4: "NaN / infinite state or time encountered"}


def raise_if_failed(result: torch.Tensor, algo_params: dict, what: str = "solve_backward") -> None:
    """diffrax `throw` semantics (pontryagin_utils.py:515-524, levelsets.py:834,867,929 pass throw=algo_params['throw']):
    with throw=True a solve that ends with max_steps_reached / dt_min_reached / nan raises; with throw=False (the value in
    flatquad_landing_experiment.py:392) failures are per-trajectory `result` codes and inf/NaN payloads.  A terminating
    event (result 1) is a success either way.  The check reads the result codes back (one sync), so it only runs for throw=True."""
    if not (algo_params or {}).get("throw", False):
        return
    r = result.reshape(-1)
    bad = torch.nonzero(r >= 2)
    if bad.numel():
        i = int(bad[0])
        code = int(r[i])
        raise _capi.PmpError(f"{what} (throw=True): trajectory {i} of {r.numel()} failed with result {code}: "
                             f"{_RESULT_TEXT.get(code, 'unknown')}; {int(bad.numel())} trajectories failed in all")


def _code(dtype: torch.dtype) -> int:
    return _capi.F64 if dtype == torch.float64 else _capi.F32


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(device, stream) scratch, grown geometrically and reused across calls."""
    key = (device.index or 0, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def pack_state(state: dict, nx: int, dtype: torch.dtype, device, vxx: bool = False) -> Tuple[torch.Tensor, bool]:
    """The reference's state dict {'x','t','v','vx'[,'vxx']} (pontryagin_utils.py:421-427, levelsets.py:588-597)
    -> struct-of-arrays [NS (+ nx*nx), n].  Returns (y, batched)."""
    x = _as_tensor(state["x"], dtype, device)
    batched = x.dim() == 2
    x = x.reshape(-1, nx)
    n = x.shape[0]
    ns = 2 * nx + 2
    y = torch.empty((ns + (nx * nx if vxx else 0), n), dtype=dtype, device=device)
    y[:nx] = x.t()
    y[nx] = _as_tensor(state.get("t", 0.0), dtype, device).reshape(-1).expand(n)
    y[nx + 1] = _as_tensor(state["v"], dtype, device).reshape(-1).expand(n)
    y[nx + 2:ns] = _as_tensor(state["vx"], dtype, device).reshape(-1, nx).t()
    if vxx:  # algo_params['pontryagin_solver_vxx']: rows ns + i*nx + j = vxx[i][j]
        y[ns:] = _as_tensor(state["vxx"], dtype, device).reshape(-1, nx * nx).expand(n, nx * nx).t()
    return y, batched


def unpack_state(y: torch.Tensor, nx: int, batched: bool = True) -> dict:
    """[NS, n] -> {'x': (n,nx), 't': (n,), 'v': (n,), 'vx': (n,nx)[, 'vxx': (n,nx,nx)]} (leading axis dropped
    if single)."""
    ns = 2 * nx + 2
    d = {"x": y[:nx].t(), "t": y[nx], "v": y[nx + 1], "vx": y[nx + 2:ns].t()}
    if y.shape[0] > ns:
        d["vxx"] = y[ns:].t().reshape(-1, nx, nx)
    if not batched:
        d = {k: v[0] for k, v in d.items()}
    return d


# ---- the batched solve on the C ABI -----------------------------------------------------------------
def solve_batch_soa(problem: _capi.Problem, opts: _capi.Opts, y_f: torch.Tensor, out=None, push=None, push_mc=None):
    """Thin wrapper of pmp_solve_batch_cuda on struct-of-arrays CUDA tensors.  Asynchronous on the
    current stream.  `out` may carry preallocated outputs (keys y_end, n_accepted, n_steps, result,
    and ts, x, t, v, vx (, vxx) when opts.save_dense).  push = peer_slots: pmp_solve_batch_push_cuda, the integrator with the
    multi-GPU gather fused in -- tensors (or device pointers) of this rank's [NS, n] block in every peer's gather buffer;
    push_mc: the address of that block through the buffer's NVSwitch multicast mapping (one store then reaches every GPU)."""
    L = _capi.lib_of(problem)
    assert y_f.is_cuda and y_f.is_contiguous()
    ns, n = y_f.shape
    nx = problem.nx
    dev = y_f.device
    out = {} if out is None else out
    if "y_end" not in out:
        out["y_end"] = torch.empty_like(y_f)
    for k in ("n_accepted", "n_steps", "result"):
        if k not in out:
            out[k] = torch.empty(n, dtype=torch.int32, device=dev)
    dense = None
    if opts.save_dense:
        s1 = opts.max_steps + 1
        shapes = {"ts": (n, s1), "x": (n, s1, nx), "t": (n, s1), "v": (n, s1), "vx": (n, s1, nx)}
        if opts.flags & _capi.FLAG_VXX:
            shapes["vxx"] = (n, s1, nx, nx)
        for k, shp in shapes.items():
            if k not in out:
                out[k] = torch.empty(shp, dtype=y_f.dtype, device=dev)
        dense = _capi.Dense(*(out[k].data_ptr() for k in ("ts", "x", "t", "v", "vx")),
                            out["vxx"].data_ptr() if "vxx" in shapes else None)
    nbytes = _capi.check(L.pmp_solve_workspace_bytes(C.byref(problem), C.byref(opts), n))
    ws = _workspace(dev, nbytes)
    with torch.cuda.device(dev):
        if push is not None:
            slots = list(push)
            ptrs = (C.c_void_p * max(1, len(slots)))(*[int(t.data_ptr() if isinstance(t, torch.Tensor) else t) for t in slots])
            _capi.check(L.pmp_solve_batch_push_cuda(
                C.byref(problem), C.byref(opts), n, y_f.data_ptr(), out["y_end"].data_ptr(), out["n_accepted"].data_ptr(),
                out["n_steps"].data_ptr(), out["result"].data_ptr(), ws.data_ptr(), ws.numel(), ptrs, len(slots),
                C.c_void_p(int(push_mc)) if push_mc else None, _stream_ptr(dev)), L)
            return out
        _capi.check(L.pmp_solve_batch_cuda(
            C.byref(problem), C.byref(opts), n, y_f.data_ptr(), out["y_end"].data_ptr(),
            C.byref(dense) if dense is not None else None, out["n_accepted"].data_ptr(),
            out["n_steps"].data_ptr(), out["result"].data_ptr(), ws.data_ptr(), ws.numel(),
            _stream_ptr(dev)))
    return out


def solve_jvp_soa(problem: _capi.Problem, opts: _capi.Opts, y_f: torch.Tensor, dy_f: torch.Tensor):
    """Thin wrapper of pmp_solve_jvp_cuda: y_f, dy_f [NS, n] CUDA tensors -> dict(y_end, dy_end [NS, n], n_accepted, n_steps,
    result): the endpoints and their directional derivative along dy_f (forward mode through the whole solve)."""
    L = _capi.lib_of(problem)
    assert y_f.is_cuda and y_f.is_contiguous() and dy_f.is_cuda and dy_f.is_contiguous() and y_f.shape == dy_f.shape
    n = y_f.shape[1]
    dev = y_f.device
    out = {"y_end": torch.empty_like(y_f), "dy_end": torch.empty_like(y_f)}
    for k in ("n_accepted", "n_steps", "result"):
        out[k] = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(L.pmp_solve_jvp_cuda(C.byref(problem), C.byref(opts), n, y_f.data_ptr(), dy_f.data_ptr(), out["y_end"].data_ptr(),
                                         out["dy_end"].data_ptr(), out["n_accepted"].data_ptr(), out["n_steps"].data_ptr(),
                                         out["result"].data_ptr(), _stream_ptr(dev)), L)
    return out


def forward_batch_soa(problem: _capi.Problem, opts: _capi.Opts, x0: torch.Tensor, P=None, x_eq=None, out=None, nn=None):
    """Thin wrapper of pmp_forward_batch_cuda / pmp_forward_nn_batch_cuda: closed-loop forward simulation under
    lambda(x) = P (x - x_eq), or under the costate law of a value-network ensemble (nn: nn_costate.NnEnsemble), with the
    running cost accumulated as an extra state.  x0 [nx, n] CUDA tensor -> dict(z_end [nx+1, n], n_accepted, n_steps,
    result [, ts (n, S+1), x (n, S+1, nx), cost (n, S+1)]).  Asynchronous on the current stream."""
    L = _capi.lib_of(problem)
    assert x0.is_cuda and x0.is_contiguous()
    nx, n = x0.shape
    dev = x0.device
    if nn is None:
        Ph = np.ascontiguousarray(np.asarray(P.detach().cpu() if isinstance(P, torch.Tensor) else P, dtype=np.float64).reshape(nx, nx))
        xe = None if x_eq is None else np.ascontiguousarray(np.asarray(x_eq, dtype=np.float64).reshape(nx))
    else:
        assert nn.weights.dtype == x0.dtype and nn.weights.device == dev and nn.nx == nx
    out = {} if out is None else out
    out.setdefault("z_end", torch.empty((nx + 1, n), dtype=x0.dtype, device=dev))
    for k in ("n_accepted", "n_steps", "result"):
        out.setdefault(k, torch.empty(n, dtype=torch.int32, device=dev))
    dp = [None, None, None]
    if opts.save_dense:
        s1 = opts.max_steps + 1
        out.setdefault("ts", torch.empty((n, s1), dtype=x0.dtype, device=dev))
        out.setdefault("x", torch.empty((n, s1, nx), dtype=x0.dtype, device=dev))
        out.setdefault("cost", torch.empty((n, s1), dtype=x0.dtype, device=dev))
        dp = [out["ts"].data_ptr(), out["x"].data_ptr(), out["cost"].data_ptr()]
    with torch.cuda.device(dev):
        if nn is None:
            _capi.check(L.pmp_forward_batch_cuda(
                C.byref(problem), C.byref(opts), n, x0.data_ptr(), Ph.ctypes.data, None if xe is None else xe.ctypes.data,
                out["z_end"].data_ptr(), *dp, out["n_accepted"].data_ptr(), out["n_steps"].data_ptr(),
                out["result"].data_ptr(), _stream_ptr(dev)), L)
        else:
            c = nn.to_c()
            _capi.check(L.pmp_forward_nn_batch_cuda(
                C.byref(problem), C.byref(opts), C.byref(c), n, x0.data_ptr(), out["z_end"].data_ptr(), *dp,
                out["n_accepted"].data_ptr(), out["n_steps"].data_ptr(), out["result"].data_ptr(), _stream_ptr(dev)), L)
    return out


def _interp(problem, opts, nx, sol: Solution, target, mode: int):
    """Solution.evaluate / state_at_value on pmp_interp_batch_cuda: locate the accepted step that contains
    each query (torch.searchsorted on the saved ts / v), gather its start node, let the kernel recompute
    the step's stage derivatives and evaluate the interpolant."""
    value_mode = opts.indep == _capi.INDEP_VALUE
    ts = sol.ts
    ysd = sol.extra.get("ys_dict", sol.ys)
    batched = ts.dim() == 2
    if not batched:
        ts = ts[None]
        ysd = {k: v[None] for k, v in ysd.items()}
    acc = sol.stats["num_accepted_steps"].reshape(-1).long()
    N, S1 = ts.shape
    dev, dtype = ts.device, ts.dtype
    tq = _as_tensor(target, dtype, dev)
    if batched:
        if tq.dim() == 0:
            tq = tq.expand(N)
        multi = tq.dim() == 2
        q = tq.reshape(N, -1)
    else:
        multi = tq.dim() == 1
        q = tq.reshape(1, -1)
    M = q.shape[1]
    # keys increase along the slots; unused slots are +inf
    if mode == 0:
        direction = 1.0 if value_mode or (opts.t0 < opts.t1) else -1.0
        keys, kq = ts * direction, q * direction
    else:
        keys, kq = ysd["v"], q
    idx = torch.searchsorted(keys.contiguous(), kq.contiguous(), right=True) - 1
    last = (acc - 1).clamp(min=0)[:, None]
    inside = (kq >= keys[:, :1]) & (kq <= keys.gather(1, acc[:, None])) & (acc[:, None] > 0)
    idx = idx.clamp(min=0).minimum(last)
    rows = torch.arange(N, device=dev)[:, None].expand(N, M)
    y0 = torch.empty((2 * nx + 2, N * M), dtype=dtype, device=dev)
    y0[:nx] = ysd["x"][rows, idx].reshape(N * M, nx).t()
    y0[nx] = ysd["t"][rows, idx].reshape(-1)
    y0[nx + 1] = ysd["v"][rows, idx].reshape(-1)
    y0[nx + 2:] = ysd["vx"][rows, idx].reshape(N * M, nx).t()
    t1 = ts[rows, idx + 1].reshape(-1).contiguous()
    tgt = q.reshape(-1).contiguous()
    out = torch.empty_like(y0)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib_of(problem).pmp_interp_batch_cuda(C.byref(problem), C.byref(opts), N * M, y0.data_ptr(),
                                                                t1.data_ptr(), tgt.data_ptr(), mode, out.data_ptr(),
                                                                _stream_ptr(dev)))
    out = torch.where(inside.reshape(1, -1), out, torch.full_like(out, float("nan")))
    d = unpack_state(out, nx)
    shape = (N, M) if (batched and multi) else ((N,) if batched else ((M,) if multi else ()))
    d = {k: v.reshape(*shape, *v.shape[1:]) for k, v in d.items()}
    if value_mode:  # flat [x, lambda, t] vector of the value-parameterised solver (rrt_sampler.py:62)
        return torch.cat([d["x"], d["vx"], d["t"][..., None]], dim=-1)
    return d


# ---- host in -> host out: pmp_solve_host ------------------------------------------------------------------------
def _host_leaf(a) -> bool:
    return isinstance(a, np.ndarray) or (isinstance(a, torch.Tensor) and not a.is_cuda)


def _pinned_outputs(n: int, nx: int, dtype: torch.dtype) -> dict:
    """The endpoint leaves and counters of one host solve as views of ONE pinned allocation, x | vx | t | v | n_accepted |
    n_steps | result: leaves at a constant distance from each other leave the device as one pitched copy per piece
    (csrc/pmp_host.cu), and a call makes one trip to the pinned allocator instead of seven."""
    es = torch.empty((), dtype=dtype).element_size()
    up = lambda b: (b + 255) // 256 * 256
    sv, ss, si = up(n * nx * es), up(n * es), up(n * 4)
    block = torch.empty(2 * sv + 2 * ss + 3 * si, dtype=torch.uint8, pin_memory=True)
    off = [0]

    def take(nbytes, shape, dt):
        v = block[off[0]:off[0] + int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()].view(dt).view(*shape)
        off[0] += nbytes
        return v
    ho = dict(x=take(sv, (n, nx), dtype), vx=take(sv, (n, nx), dtype), t=take(ss, (n,), dtype), v=take(ss, (n,), dtype),
              n_accepted=take(si, (n,), torch.int32), n_steps=take(si, (n,), torch.int32), result=take(si, (n,), torch.int32))
    return ho


def _solve_host_lqr(problem, opts, nx, x_f, P_lqr, x_eq, dtype, dev, n_pieces=0):
    """Host terminal samples x_f (n, nx) -> endpoints in pinned host tensors through pmp_solve_host_lqr: only x_f crosses the
    link, vx_f = P (x_f - x_eq) and v_f = 1/2 (x_f - x_eq)' vx_f are formed on the device (levelsets.py:585-586)."""
    x = (x_f if isinstance(x_f, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_f))).to(dtype).contiguous()
    n = x.shape[0]
    if x.shape != (n, nx):
        raise ValueError(f"x_f must be [n, {nx}]")
    Ph = np.ascontiguousarray(np.asarray(P_lqr.detach().cpu() if isinstance(P_lqr, torch.Tensor) else P_lqr, dtype=np.float64).reshape(nx, nx))
    xe = None if x_eq is None else np.ascontiguousarray(np.asarray(x_eq, dtype=np.float64).reshape(nx))
    ho = _pinned_outputs(n, nx, dtype)
    L = _capi.lib_of(problem)
    with torch.cuda.device(dev):
        _capi.check(L.pmp_solve_host_lqr(C.byref(problem), C.byref(opts), n, x.data_ptr(), None, Ph.ctypes.data,
                                         None if xe is None else xe.ctypes.data, ho["x"].data_ptr(), ho["t"].data_ptr(),
                                         ho["v"].data_ptr(), ho["vx"].data_ptr(), ho["n_accepted"].data_ptr(),
                                         ho["n_steps"].data_ptr(), ho["result"].data_ptr(), int(n_pieces)), L)
    stats = StepStats(ho["n_steps"], ho["n_accepted"])
    return Solution(t0=opts.t0, t1=opts.t1, ts=None, ys=None, y_end={k: ho[k] for k in ("x", "t", "v", "vx")}, stats=stats,
                    result=ho["result"], batched=True)


def _solve_host_pipelined(problem, opts, nx, y_f, dtype, dev, n_pieces=0):
    """solve_backward for HOST inputs (the reference's habitat: everything lives in host memory): one call of
    pmp_solve_host (csrc/pmp_host.cu), which cuts the batch into pieces and overlaps the host->device copy, the
    integration and the device->host copy of neighbouring pieces on its own streams.  Endpoints, step counts and result
    codes come back in pinned host tensors.  Inputs overlap only if they are pinned too (torch pin_memory); pageable
    numpy arrays work, the copies then serialise."""
    host = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dtype).contiguous()
    x, vx = host(y_f["x"]), host(y_f["vx"])
    n = x.shape[0]
    if x.shape != (n, nx) or vx.shape != (n, nx):
        raise ValueError(f"'x', 'vx' must be [n, {nx}]")

    def per_traj(a, name):  # [n] leaf; scalars are broadcast
        a = host(np.asarray(a) if not isinstance(a, torch.Tensor) else a).reshape(-1)
        if a.numel() == 1 and n != 1:
            a = a.expand(n).contiguous()
        if a.numel() != n:
            raise ValueError(f"'{name}' must be a scalar or [n]")
        return a

    v = per_traj(y_f["v"], "v")
    t = None
    if "t" in y_f:
        t_in = y_f["t"]
        if np.ndim(t_in) == 0 or np.size(t_in) == 1 and n != 1:
            if float(np.asarray(t_in).reshape(-1)[0]) != float(opts.t0):
                t = per_traj(t_in, "t")
        else:
            t = per_traj(t_in, "t")
    ho = _pinned_outputs(n, nx, dtype)
    L = _capi.lib_of(problem)
    with torch.cuda.device(dev):
        _capi.check(L.pmp_solve_host(C.byref(problem), C.byref(opts), n, x.data_ptr(), t.data_ptr() if t is not None else None,
                                     v.data_ptr(), vx.data_ptr(), ho["x"].data_ptr(), ho["t"].data_ptr(), ho["v"].data_ptr(),
                                     ho["vx"].data_ptr(), ho["n_accepted"].data_ptr(), ho["n_steps"].data_ptr(),
                                     ho["result"].data_ptr(), int(n_pieces)), L)
    stats = StepStats(ho["n_steps"], ho["n_accepted"])
    return Solution(t0=opts.t0, t1=opts.t1, ts=None, ys=None, y_end={k: ho[k] for k in ("x", "t", "v", "vx")}, stats=stats,
                    result=ho["result"], batched=True)


def _wrap_solution(out, nx, batched, t0, t1, opts, problem=None) -> Solution:
    ys = None
    ts = None
    if opts.save_dense:
        ts = out["ts"]
        ys = {"x": out["x"], "t": out["t"], "v": out["v"], "vx": out["vx"]}
        if "vxx" in out:
            ys["vxx"] = out["vxx"]
    stats = {"num_steps": out["n_steps"], "num_accepted_steps": out["n_accepted"],
             "num_rejected_steps": out["n_steps"] - out["n_accepted"]}
    sol = Solution(t0=t0, t1=t1, ts=ts, ys=ys, y_end=unpack_state(out["y_end"], nx), stats=stats,
                   result=out["result"], batched=True)
    if problem is not None:
        import functools
        sol.extra["interp"] = functools.partial(_interp, problem, opts, nx)
    return sol if batched else sol[0]


def _solver_opts(algo_params: dict, dtype: torch.dtype, **over) -> _capi.Opts:
    g = lambda k: algo_params.get("pontryagin_solver_" + k, _DEFAULTS[k])
    kw = dict(method=g("method"), dtype=_code(dtype), dtmin=g("dtmin"), dtmax=g("dtmax"),
              safety=g("safety"), factormin=g("factormin"), factormax=g("factormax"),
              force_dtmin=g("force_dtmin"),
              rtol=algo_params.get("pontryagin_solver_rtol", 1e-5),
              atol=algo_params.get("pontryagin_solver_atol", 1e-5),
              max_steps=int(algo_params["pontryagin_solver_maxsteps"]))
    kw.update(over)
    return _capi.make_opts(**kw)


def define_backward_solver(problem_params: dict, algo_params: dict):
    """pontryagin_utils.py:414-529.  Returns (solve_backward, f_extended).

    solve_backward(y_f): y_f = {'x','t','v','vx'} with the reference's shapes, or with a leading
    batch axis N on every leaf (replaces jax.vmap, levelsets.py:601,1254).  Integrates the extended
    characteristic system from t0=0 to t1=-T with Tsit5 + PIDController(rtol, atol, dtmin=5e-4,
    dtmax=.5), dt0=-0.1, at most `pontryagin_solver_maxsteps` step attempts, throw=False semantics
    (per-trajectory `result` codes, inf padding); algo_params['throw'] = True raises PmpError naming the first trajectory
    that ended with max_steps / dt_min / NaN instead (raise_if_failed).  Optional algo_params keys:
    'pontryagin_solver_method' ('tsit5' | 'dopri5' | 'rk4'), 'pontryagin_solver_adaptive' (default
    True; RK4 is fixed-step), 'pontryagin_solver_dt0', '..._dtmin', '..._dtmax',
    'pontryagin_solver_save_steps' (default True = SaveAt(steps=True)), 'dtype'.
    """
    problem = systems.to_c_problem(problem_params)
    nx = problem.nx
    # pontryagin_utils.py:425,460: also propagate the value Hessian ('vxx' leaf of the state)
    with_vxx = bool(algo_params.get("pontryagin_solver_vxx", False))
    # flatquad: closed-form Jacobians of pmp_rhs through u*; generated functors: forward-mode differentiation of the
    # generated code (csrc/pmp_vxx_dual.cuh), as the reference's jax.jacobian(pmp_rhs) handles any f, l.  orbits / LQR are
    # u_star_new systems: the reference has no vxx branch for them (its define_backward_solver goes through u_star_2d)
    if with_vxx and problem.system_id not in (_capi.SYS_FLATQUAD, _capi.SYS_GENERATED):
        raise NotImplementedError("pontryagin_solver_vxx=True: flatquad and generated systems (problem_params['codegen'] = "
                                  "True) only; the reference's vxx branch runs through u_star_2d")

    def f_extended(t, state, args=None):
        """RHS of the extended system in forward time (pontryagin_utils.py:418-482)."""
        dev = _device(state["x"])
        dtype = _infer_dtype(algo_params, state["x"])
        y, batched = pack_state(state, nx, dtype, dev, with_vxx)
        dy = torch.empty_like(y)
        with torch.cuda.device(dev):
            _capi.check(_capi.lib_of(problem).pmp_rhs_batch_cuda_flags(
                C.byref(problem), _code(dtype), _capi.FLAG_VXX if with_vxx else 0, y.shape[1], y.data_ptr(),
                dy.data_ptr(), _stream_ptr(dev)))
        return unpack_state(dy, nx, batched)

    def solve_backward(y_f, save_steps=None):
        dev = _device(y_f["x"])
        dtype = _infer_dtype(algo_params, y_f["x"])
        T = float(algo_params["pontryagin_solver_T"])
        method = algo_params.get("pontryagin_solver_method", _DEFAULTS["method"])
        adaptive = bool(algo_params.get("pontryagin_solver_adaptive", True)) and str(method).lower() != "rk4"
        if save_steps is None:
            save_steps = bool(algo_params.get("pontryagin_solver_save_steps", True))
        dt0 = float(algo_params.get("pontryagin_solver_dt0", _DEFAULTS["dt0"]))
        # host in -> host out: endpoint-only batches handed over in host memory go through pmp_solve_host (results in
        # pinned host tensors); 'pontryagin_solver_host_chunks': -1 disables (device tensors come back), 0 lets the
        # library choose the number of pieces, k forces k
        if (not with_vxx and not save_steps and _host_leaf(y_f["x"]) and _host_leaf(y_f["vx"]) and np.ndim(y_f["x"]) == 2):
            n_all = int(np.shape(y_f["x"])[0])
            chunks = int(algo_params.get("pontryagin_solver_host_chunks", 0 if n_all >= (1 << 18) else -1))
            if chunks >= 0:
                opts = _solver_opts(algo_params, dtype, t0=0.0, t1=-T, dt0=-abs(dt0), adaptive=int(adaptive), save_dense=0)
                sol = _solve_host_pipelined(problem, opts, nx, y_f, dtype, dev, chunks)
                raise_if_failed(sol.result, algo_params)
                return sol
        y, batched = pack_state(y_f, nx, dtype, dev, with_vxx)
        over = {}
        if with_vxx:
            flags = _capi.FLAG_VXX
            # a symmetric terminal Hessian (the reference starts from P_lqr, levelsets.py:595) stays symmetric:
            # the kernel then integrates the upper triangle only
            vf = y[2 * nx + 2:].reshape(nx, nx, -1)
            if bool(algo_params.get("pontryagin_solver_vxx_symmetric", torch.equal(vf, vf.transpose(0, 1)))):
                flags |= _capi.FLAG_VXX_SYMMETRIC
            over["flags"] = flags
            if "vxx_max_norm" in algo_params:  # pontryagin_utils.py:508-512
                over.update(event=_capi.EVENT_VXX_NORM, event_level=float(algo_params["vxx_max_norm"]))
        opts = _solver_opts(algo_params, dtype, t0=0.0, t1=-T, dt0=-abs(dt0), adaptive=int(adaptive),
                            save_dense=int(save_steps), **over)
        out = solve_batch_soa(problem, opts, y)
        raise_if_failed(out["result"], algo_params)
        return _wrap_solution(out, nx, batched, 0.0, -T, opts, problem)

    return solve_backward, f_extended


def _ustar(x, costate, problem_params, flags=None):
    problem = systems.to_c_problem(problem_params)
    nx, nu = problem.nx, problem.nu
    dev = _device(x, costate)
    dtype = _infer_dtype(problem_params, x, costate)
    xt = _as_tensor(x, dtype, dev)
    batched = xt.dim() == 2
    xs = xt.reshape(-1, nx).t().contiguous()
    ls = _as_tensor(costate, dtype, dev).reshape(-1, nx).t().contiguous()
    u = torch.empty((nu, xs.shape[1]), dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        L = _capi.lib_of(problem)
        if flags is None:
            _capi.check(L.pmp_ustar_batch_cuda(C.byref(problem), _code(dtype), xs.shape[1], xs.data_ptr(), ls.data_ptr(),
                                               u.data_ptr(), _stream_ptr(dev)), L)
        else:
            _capi.check(L.pmp_ustar_batch_cuda_flags(C.byref(problem), _code(dtype), int(flags), xs.shape[1], xs.data_ptr(),
                                                     ls.data_ptr(), u.data_ptr(), _stream_ptr(dev)), L)
    u = u.t()
    return u if batched else u[0]


def u_star_2d(x, costate, problem_params, smooth=False, debug_oups=False):
    """pontryagin_utils.py:11-254: u* = argmin_u l(x,u) + costate.f(x,u) over the 2-D input box.
    x, costate: (nx,) or (N, nx).  smooth=True: the reference's softmax blend of the candidates (:217-254, a debugging aid)."""
    if not problem_params["nu"] == 2:
        raise NotImplementedError("this ustar function is for 2d only!")
    return _ustar(x, costate, problem_params, flags=_capi.FLAG_USTAR_SMOOTH if smooth else None)


def u_star_new(x, costate, problem_params):
    """pontryagin_utils.py:532-562: clip(solve(R+R', -costate' df/du), U) for nu == 1."""
    assert problem_params["nu"] == 1
    return _ustar(x, costate, problem_params)


def lqr(A, B, Q, R):
    """pontryagin_utils.py:569-599: continuous-time LQR; X such that the LQR value is 0.5 x'Xx."""
    A, B, Q, R = (np.asarray(M, dtype=float) for M in (A, B, Q, R))
    X = scipy.linalg.solve_continuous_are(A, B, Q, R)
    K = np.linalg.inv(R) @ (B.T @ X)
    eigVals = np.linalg.eigvals(A - B @ K)
    if not (eigVals.real < 0).all():
        raise ValueError("LQR closed loop not stable...")
    return K, X, eigVals


def get_terminal_lqr(problem_params):
    """pontryagin_utils.py:603-628: linearise at (x_eq, u_eq), check equilibrium and
    controllability, solve the CARE.  Host-side setup (scipy), not part of the GPU path."""
    x_eq = np.asarray(problem_params["x_eq"], dtype=float)
    u_eq = np.asarray(problem_params["u_eq"], dtype=float)
    f0 = np.asarray(systems._call_f(problem_params["f"], x_eq, u_eq), dtype=float)
    assert np.allclose(f0, 0), "(x_eq, u_eq) does not seem to be an equilibrium"
    A, B, Q, R = systems.linearise(problem_params)
    nx = problem_params["nx"]
    ctrb = np.hstack([np.linalg.matrix_power(A, j) @ B for j in range(nx)])
    if np.linalg.matrix_rank(ctrb) < nx:
        raise ValueError("linearisation not controllable aaaaah what did you do idiot")
    K_lqr, P_lqr, _ = lqr(A, B, Q, R)
    return K_lqr, P_lqr


def make_pontryagin_solver_reparam(problem_params: dict, algo_params: dict):
    """The value-reparameterised characteristic solver the stale experiments call
    (experiment_orbits.py:117, experiment_lqr_statepenalty.py:119, rrt_sampler.py:64-70) and the
    snapshot no longer defines.  Contract rebuilt from the call sites:

        solver(y0, v0, v1) -> Solution
          y0 = [x (nx), lambda (nx), t (1)] flat, or (N, 2nx+1)       rrt_sampler.py:62
          v0 = value at y0 (scalar or (N,)), v1 = V_max               experiment_orbits.py:128-129
          integrates dy/dv = (f, -H_x, 1) / (dv/dt), dv/dt = -l(x,u*) experiment_orbits.py:109-111
          sol.ts = value levels, sol.ys = (S+1, 2nx+1) [batched: (N, S+1, 2nx+1)]  rrt_sampler.py:76-82
    algo_params: pontryagin_solver_{dt, adaptive, rtol, atol, maxsteps} (experiment_orbits.py:97-104).
    solver.jvp(y0, v0, v1, dy0) / solver.jacobian(y0, v0, v1): forward-mode sensitivities with respect to y0 (the
    reference differentiates the solver with jax.jacobian, experiment_orbits.py:186-197).
    No reference output exists for this function: parity is pinned only against the oracle.
    """
    problem = systems.to_c_problem(problem_params)
    nx = problem.nx

    def _soa(y0t, v0, dtype, dev):  # flat [x, lambda, t] rows -> [NS, n] (x, t, v, vx)
        n = y0t.shape[0]
        y = torch.empty((2 * nx + 2, n), dtype=dtype, device=dev)
        y[:nx] = y0t[:, :nx].t()
        y[nx] = y0t[:, 2 * nx]
        y[nx + 1] = _as_tensor(v0, dtype, dev).reshape(-1).expand(n)
        y[nx + 2:] = y0t[:, nx:2 * nx].t()
        return y

    def _flat(y):  # [NS, n] -> (n, 2nx+1) flat [x, lambda, t]
        return torch.cat([y[:nx].t(), y[nx + 2:].t(), y[nx][:, None]], dim=1)

    def _opts(dtype, v1, save_steps):
        adaptive = bool(algo_params.get("pontryagin_solver_adaptive", True))
        return _solver_opts(algo_params, dtype, indep=_capi.INDEP_VALUE, t0=0.0, t1=float(v1),
                            dt0=float(algo_params.get("pontryagin_solver_dt", 1 / 16)),
                            adaptive=int(adaptive), save_dense=int(save_steps),
                            dtmin=float(algo_params.get("pontryagin_solver_dtmin", 0.0)),
                            dtmax=float(algo_params.get("pontryagin_solver_dtmax", 0.0)))

    def jvp(y0, v0, v1, dy0):
        """(y(v1), dy(v1)): the state at value level v1 and its derivative along dy0, d y(v1) / d y0 . dy0 -- what
        jax.jvp / jax.jacobian through `pontryagin_solver(y0, v0, v1)` gives (experiment_orbits.py:186-197: dsol_final =
        jax.jacobian(get_final_x)(theta) is jvp with dy0 = d y0 / d theta).  y0, dy0: (2nx+1,) or (N, 2nx+1); v0, v1 fixed."""
        dev = _device(y0)
        dtype = _infer_dtype(algo_params, y0)
        y0t = _as_tensor(y0, dtype, dev)
        batched = y0t.dim() == 2
        y0t = y0t.reshape(-1, 2 * nx + 1)
        d0t = _as_tensor(dy0, dtype, dev).reshape(-1, 2 * nx + 1).expand_as(y0t)
        y = _soa(y0t, v0, dtype, dev)
        dy = _soa(d0t, 0.0, dtype, dev)
        out = solve_jvp_soa(problem, _opts(dtype, v1, False), y.contiguous(), dy.contiguous())
        raise_if_failed(out["result"], algo_params, "pontryagin_solver_reparam.jvp")
        y1, dy1 = _flat(out["y_end"]), _flat(out["dy_end"])
        return (y1, dy1) if batched else (y1[0], dy1[0])

    def jacobian(y0, v0, v1):
        """d y(v1) / d y0, shape (2nx+1, 2nx+1) [batched: (N, 2nx+1, 2nx+1)]: 2nx+1 forward-mode columns in one launch."""
        dev = _device(y0)
        dtype = _infer_dtype(algo_params, y0)
        y0t = _as_tensor(y0, dtype, dev)
        batched = y0t.dim() == 2
        y0t = y0t.reshape(-1, 2 * nx + 1)
        n, m = y0t.shape
        eye = torch.eye(m, dtype=dtype, device=dev)
        yy = y0t[:, None, :].expand(n, m, m).reshape(n * m, m)
        dd = eye[None].expand(n, m, m).reshape(n * m, m)
        v0t = _as_tensor(v0, dtype, dev).reshape(-1).expand(n)[:, None].expand(n, m).reshape(-1)
        _, dy1 = jvp(yy, v0t, v1, dd)
        J = dy1.reshape(n, m, m).transpose(1, 2)  # [i, out, in]
        return J if batched else J[0]

    def solver(y0, v0, v1, save_steps=True):
        dev = _device(y0)
        dtype = _infer_dtype(algo_params, y0)
        y0t = _as_tensor(y0, dtype, dev)
        batched = y0t.dim() == 2
        y0t = y0t.reshape(-1, 2 * nx + 1)
        y = _soa(y0t, v0, dtype, dev)
        opts = _opts(dtype, v1, save_steps)
        out = solve_batch_soa(problem, opts, y)
        raise_if_failed(out["result"], algo_params, "pontryagin_solver_reparam")
        sol = _wrap_solution(out, nx, True, None, float(v1), opts, problem)
        if sol.ys is not None:  # flat (.., 2nx+1) layout of the missing solver
            sol.extra["ys_dict"] = sol.ys
            sol.ys = torch.cat([sol.ys["x"], sol.ys["vx"], sol.ys["t"][..., None]], dim=-1)
        return sol if batched else sol[0]

    solver.jvp = jvp
    solver.jacobian = jacobian
    return solver
