"""Host-side mirror of the reference's `eval_utils` (closed-loop cost evaluation, /root/reference/eval_utils.py) for
costate laws the CUDA path can evaluate.

    closed_loop_eval_nn_ensemble(problem_params, algo_params, V_nn, nn_params, x0s)   eval_utils.py:10-45
    closed_loop_eval_lqr(problem_params, algo_params, P_lqr, x0s)     the same with lambda(x) = P (x - x_eq)
    compute_controlcost(problem_params, all_sols)                     eval_utils.py:101-107
    closed_loop_eval_general(problem_params, algo_params, ustar_fct, x0s)   eval_utils.py:48-99: the feedback law is traced
                                                                      with sympy and compiled into a plug-in (codegen.py)

z = (x, cost), z' = (f(x,u*), l(x,u*)), Tsit5 with the CONSTANT step algo_params['sim_dt'] from 0 to
algo_params['sim_T'], max_steps = T/dt, SaveAt(steps=True): all_sols.ys has shape (N, T/dt, nx+1).
"""
from __future__ import annotations

import torch

from . import _capi, systems
from . import pontryagin_utils as pu
from .solution import Solution


def closed_loop_eval_general(problem_params, algo_params, ustar_fct, x0s):
    """eval_utils.py:48-99: closed loop under an arbitrary feedback law `ustar_fct: X -> U` ("should be a jax function ... no
    time dependence"), one extra state recording the control cost.  The law is traced with sympy like f and l (write it against
    approximate_optimal_control_b200.symnp where the reference's script says jax.numpy) and compiled, together with f and l,
    into a plug-in whose u*(x, lambda) IS that law; the forward kernel then runs with a zero costate."""
    pp = dict(problem_params)
    pp["ustar_fct"] = ustar_fct
    pp["codegen"] = True
    pp["system_name"] = str(problem_params.get("system_name", "user")) + "_with_feedback_law"
    nx = int(pp["nx"])
    return _closed_loop_eval(pp, algo_params, x0s, P=torch.zeros((nx, nx), dtype=torch.float64))


def closed_loop_eval_nn_ensemble(problem_params, algo_params, V_nn, nn_params, x0s):
    """eval_utils.py:10-45: u* from the mean over the ensemble of the networks' input gradients (V_nn.apply_grad: no
    normaliser).  V_nn is unused (the architecture is read off nn_params); nn_params: flax params with a leading
    ensemble axis, or a ready nn_costate.NnEnsemble."""
    from .nn_costate import DataNormaliser, NnEnsemble
    nx = int(problem_params["nx"])
    dev = pu._device(x0s)
    dtype = pu._infer_dtype(algo_params, x0s)
    ens = nn_params if isinstance(nn_params, NnEnsemble) else NnEnsemble.from_flax(nn_params, DataNormaliser.identity(nx), dev, dtype)
    return _closed_loop_eval(problem_params, algo_params, x0s, nn=ens)


def closed_loop_eval_lqr(problem_params, algo_params, P_lqr, x0s):
    return _closed_loop_eval(problem_params, algo_params, x0s, P=P_lqr)


def _closed_loop_eval(problem_params, algo_params, x0s, P=None, nn=None):
    problem = systems.to_c_problem(problem_params)
    nx = problem.nx
    T, dt = float(algo_params["sim_T"]), float(algo_params["sim_dt"])
    max_steps = int(T / dt)
    dev = pu._device(x0s) if nn is None else nn.weights.device
    dtype = pu._infer_dtype(algo_params, x0s) if nn is None else nn.weights.dtype
    xs = pu._as_tensor(x0s, dtype, dev).reshape(-1, nx).t().contiguous()
    opts = _capi.make_opts(method=algo_params.get("pontryagin_solver_method", "tsit5"), dtype=pu._code(dtype), adaptive=0,
                           t0=0.0, t1=T, dt0=dt, max_steps=max_steps, save_dense=1, dtmin=0.0, dtmax=0.0)
    out = pu.forward_batch_soa(problem, opts, xs, P, problem_params.get("x_eq"), nn=nn)
    pu.raise_if_failed(out["result"], {"throw": True}, "closed_loop_eval")  # eval_utils.py:86-89: diffeqsolve's default throw=True
    ys = torch.cat([out["x"], out["cost"][..., None]], dim=-1)[:, 1:]  # SaveAt(steps=True): no t0 slot
    stats = {"num_steps": out["n_steps"], "num_accepted_steps": out["n_accepted"],
             "num_rejected_steps": out["n_steps"] - out["n_accepted"]}
    return Solution(t0=0.0, t1=T, ts=out["ts"][:, 1:], ys=ys, y_end={"x": out["z_end"][:nx].t(), "cost": out["z_end"][nx]},
                    stats=stats, result=out["result"], batched=True)


def compute_controlcost(problem_params, all_sols):
    """eval_utils.py:101-107: mean over the initial states of the cost recorded at the last saved step."""
    control_costs = all_sols.ys[:, -1, problem_params["nx"]]
    return control_costs.mean()
