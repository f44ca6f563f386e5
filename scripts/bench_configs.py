#!/usr/bin/env python
"""Kernel-time table for the five BASELINE.json configurations (C1-C5 of SURVEY.md 8(d)) on one GPU.
bench.py is the driver-facing harness (C4/C5); this script covers the rest for DESIGN.md / profiles/.
usage: python scripts/bench_configs.py [> profiles/rNN_configs.jsonl]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.linalg
import torch

from approximate_optimal_control_b200 import _capi, flops, pontryagin_utils as pu, systems, terminal


def soa(st, dtype):
    return np.ascontiguousarray(np.concatenate([st["x"].T, st["t"][None], st["v"][None], st["vx"].T], 0), dtype=dtype)


def timed(problem, opts, y, reps=5):
    out = {}
    ts = []
    for i in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); pu.solve_batch_soa(problem, opts, y, out); b.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out


def line(name, problem, opts, y_np, note=""):
    tdt = torch.float32 if opts.dtype == _capi.F32 else torch.float64
    y = torch.as_tensor(y_np, dtype=tdt, device="cuda").contiguous()
    ms, out = timed(problem, opts, y)
    n = y.shape[1]
    attempts = int(out["n_steps"].sum().item())
    res = np.bincount(out["result"].cpu().numpy(), minlength=5).tolist()
    d = {"config": name, "n": n, "dtype": "f32" if opts.dtype == _capi.F32 else "f64", "ms": ms,
         "trajectories_per_s": n / ms * 1e3, "attempts_per_s": attempts / ms * 1e3,
         "mean_attempts": attempts / n, "result_counts[ok,event,max_steps,dt_min,nan]": res, "note": note}
    key = (problem.system_id, opts.method)
    if key in flops.ATTEMPT_FLOPS:
        d["algorithmic_tflops"] = flops.solve_flops(problem.system_id, opts.method, attempts, n) / ms / 1e9
    print(json.dumps(d), flush=True)


def main():
    # C1: LQR state penalty, 1024 samples, RK4 dt=1/16, T=8, fp64 (experiment_lqr_statepenalty.py)
    A = np.array([[0.0, 1.0], [0.0, 0.0]]); B = np.array([[0.0], [1.0]])
    P = scipy.linalg.solve_continuous_are(A, B, np.eye(2), np.eye(1))
    pl = systems.to_c_problem(systems.lqr_statepenalty_problem_params())
    line("C1 lqr_statepenalty RK4 dt=1/16 T=8", pl,
         _capi.make_opts(method="rk4", dtype=_capi.F64, adaptive=0, t1=-8.0, dt0=-1 / 16, max_steps=128),
         soa(terminal.ellipse_terminal_states(P, 1024, 0.1), np.float64), "launch-latency bound at n=1024")
    # C2: flatquad 2^16 adaptive, Tsit5 and Dopri5, fp32 and fp64
    pp = systems.flatquad_problem_params()
    _, Pq = pu.get_terminal_lqr(pp)
    pf = systems.to_c_problem(pp)
    y16 = soa(terminal.sample_terminal_states(Pq, pp["V_f"], 1 << 16, seed=1), np.float64)
    for m in ("tsit5", "dopri5"):
        for dt in (_capi.F32, _capi.F64):
            line(f"C2 flatquad 2^16 {m}", pf, _capi.make_opts(method=m, dtype=dt), y16)
    # C3: orbits 2^18, Tsit5 rtol=atol=1e-4, <=1024 attempts, stop at v >= 4.2; and value-parameterised
    po = systems.orbits_problem_params()
    Ao, Bo, Qo, Ro = systems.linearise(po)
    P0 = scipy.linalg.solve_continuous_are(Ao, Bo, Qo, Ro) / 2
    pc = systems.to_c_problem(po)
    y18 = soa(terminal.ellipse_terminal_states(P0, 1 << 18, 0.01, x_eq=[0.0, 1.0]), np.float64)
    line("C3 orbits 2^18 tsit5 event v>=4.2", pc,
         _capi.make_opts(method="tsit5", dtype=_capi.F64, rtol=1e-4, atol=1e-4, max_steps=1024, t1=-50.0, dt0=-1 / 16,
                         event=_capi.EVENT_VALUE_GE, event_level=4.2, dtmin=0.0, dtmax=0.0), y18,
         "theta-ordered samples: neighbouring lanes have similar lengths")
    perm = np.random.Generator(np.random.PCG64(0)).permutation(1 << 18)
    line("C3 orbits 2^18 tsit5 event v>=4.2, shuffled", pc,
         _capi.make_opts(method="tsit5", dtype=_capi.F64, rtol=1e-4, atol=1e-4, max_steps=1024, t1=-50.0, dt0=-1 / 16,
                         event=_capi.EVENT_VALUE_GE, event_level=4.2, dtmin=0.0, dtmax=0.0), y18[:, perm])
    line("C3 orbits 2^18 tsit5 value-parameterised to V_max=4.2", pc,
         _capi.make_opts(method="tsit5", dtype=_capi.F64, indep=_capi.INDEP_VALUE, rtol=1e-4, atol=1e-4, max_steps=1024,
                         t1=4.2, dt0=1 / 16, dtmin=0.0, dtmax=0.0), y18)
    # C4 / C5: flatquad 2^20 endpoints and dense (bench.py's workloads), for completeness
    y20 = soa(terminal.sample_terminal_states(Pq, pp["V_f"], 1 << 20, seed=1), np.float64)
    for dt in (_capi.F32, _capi.F64):
        line("C4 flatquad 2^20 tsit5 endpoints", pf, _capi.make_opts(dtype=dt), y20)
        line("C4 flatquad 2^20 tsit5 endpoints, literal u_star_2d", pf, _capi.make_opts(dtype=dt, flags=_capi.FLAG_USTAR_LITERAL), y20)
        line("C5 flatquad 2^20 tsit5 dense", pf, _capi.make_opts(dtype=dt, save_dense=1), y20)
    line("flatquad 2^20 rk4 dt=1/32 (96 steps)", pf, _capi.make_opts(method="rk4", adaptive=0, dt0=-1 / 32, dtype=_capi.F32), y20)


if __name__ == "__main__":
    main()
