#!/usr/bin/env python
"""One GPU: what the PUSH instantiation of the integrator (pmp_solve_batch_push_cuda, the multi-GPU gather fused in) costs by
itself -- 0 peers (the detour code only) and k "peers" that are plain local buffers (the stores, without NVLink) -- against the
plain kernel.  2^20 flatquad rollouts, Tsit5 fp32.  usage: python scripts/bench_push_kernel.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from common import c_problem, flatquad_terminal
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu

n = 1 << 20
prob, _ = c_problem("flatquad")
y = torch.as_tensor(flatquad_terminal(n, dtype=np.float32), device="cuda").contiguous()
opts = _capi.make_opts(method="tsit5", dtype=_capi.F32)
out = {}
bufs = [torch.empty_like(y) for _ in range(7)]


def timed(push, reps=10):
    ts = []
    for i in range(reps + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); pu.solve_batch_soa(prob, opts, y, out, push=push); b.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


ref = pu.solve_batch_soa(prob, opts, y)["y_end"].clone()
print(json.dumps({"kernel": "plain", "ms": timed(None)}))
for k in (0, 1, 3, 7):
    ms = timed(bufs[:k])
    same = all(bool(torch.equal(b, ref)) for b in bufs[:k]) and bool(torch.equal(out["y_end"], ref))
    print(json.dumps({"kernel": f"PUSH, {k} local 'peers'", "ms": ms, "bit_identical": same}), flush=True)
