#!/usr/bin/env python
"""One configuration of the integrator, a few launches: the command ncu wraps (B200_PROFILING.md).
usage: python scripts/prof_case.py [f32|f64] [endpoints|dense] [tsit5|dopri5|rk4] [log2n] [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import c_problem, flatquad_terminal
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu

dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
workload = sys.argv[2] if len(sys.argv) > 2 else "endpoints"
method = sys.argv[3] if len(sys.argv) > 3 else "tsit5"
log2n = int(sys.argv[4]) if len(sys.argv) > 4 else 20
launches = int(sys.argv[5]) if len(sys.argv) > 5 else 3
prob, _ = c_problem("flatquad")
tdt = torch.float32 if dtype == "f32" else torch.float64
y = torch.as_tensor(flatquad_terminal(1 << log2n), dtype=tdt, device="cuda").contiguous()
kw = dict(adaptive=0, dt0=-1 / 32) if method == "rk4" else {}
opts = _capi.make_opts(method=method, dtype=_capi.F32 if dtype == "f32" else _capi.F64, save_dense=int(workload == "dense"), **kw)
out = {}
for _ in range(launches):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pu.solve_batch_soa(prob, opts, y, out); b.record(); torch.cuda.synchronize()
    print(f"{dtype} {workload} {method} 2^{log2n}: {a.elapsed_time(b):.3f} ms, {int(out['n_steps'].sum())} attempts")
