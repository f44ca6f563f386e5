#!/usr/bin/env python
"""e2e wall time of solve_backward on pinned host tensors (bench.py's e2e workload: 2^20 flatquad rollouts, T = 3) for a
list of pmp_solve_host plans (B200PMP_HOST_PLAN / B200PMP_HOST_GROUP tuning knobs).
usage: python scripts/sweep_host_pipeline.py 'plan|group' ..."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from approximate_optimal_control_b200 import levelsets, pontryagin_utils as pu, systems, terminal

n = 1 << 20
pp, ap = systems.flatquad_problem_params(), systems.flatquad_algo_params()
_, P = pu.get_terminal_lqr(pp)
st = terminal.sample_terminal_states(P, pp["V_f"], n, seed=1, dtype=np.float32)
host_in = torch.as_tensor(st["x"]).pin_memory()  # the reference's call: xfs only (levelsets.py:601)
solve = levelsets.define_solve_backward_lqr(pp, dict(ap, pontryagin_solver_save_steps=False, dtype="float32"), P)
for cfg in sys.argv[1:] or ["|"]:
    plan, group = (cfg.split("|") + [""])[:2]
    for k, v in (("B200PMP_HOST_PLAN", plan), ("B200PMP_HOST_GROUP", group)):
        os.environ.pop(k, None)
        if v:
            os.environ[k] = v
    for _ in range(3):
        solve(host_in)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(20):
            solve(host_in)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) / 20 * 1e3)
    print(json.dumps({"plan": plan or "auto", "group_log2": group or "18", "ms_per_call_best": round(min(ts), 4),
                      "ms_per_call_median": round(float(np.median(ts)), 4), "traj_per_s": n / min(ts) * 1e3}), flush=True)
