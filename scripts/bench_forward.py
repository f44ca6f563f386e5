#!/usr/bin/env python
"""Kernel time of the forward closed-loop simulations (row N3): LQR costate law and value-network ensemble (8 x 6-64-64-64-1),
flatquad, Tsit5 + PIDController(dtmin=.05), t: 0 -> 10 (levelsets.py:818-933).   usage: python scripts/bench_forward.py [log2_n]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from common import c_problem, flatquad_P, flatquad_terminal
from approximate_optimal_control_b200 import _capi, nn_costate as NC, pontryagin_utils as pu

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 14
n = 1 << lg
prob, _ = c_problem("flatquad")
x0 = 40.0 * flatquad_terminal(n)[:6]
for dtype, tdt, name in ((_capi.F32, torch.float32, "f32"), (_capi.F64, torch.float64, "f64")):
    x = torch.as_tensor(x0, dtype=tdt, device="cuda").contiguous()
    ens = NC.NnEnsemble.from_flax(NC.init_flax_params(0, 6, (64, 64, 64), 8), NC.DataNormaliser.identity(6), "cuda", tdt)
    for law, kw in (("lqr", dict(P=flatquad_P())), ("nn-ensemble 8x(6-64-64-64-1)", dict(nn=ens))):
        opts = _capi.make_opts(dtype=dtype, t0=0.0, t1=10.0, dt0=0.1, dtmin=0.05, dtmax=0.0, max_steps=256)
        out = {}; ts = []
        for i in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); pu.forward_batch_soa(prob, opts, x, out=out, **kw); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        att = int(out["n_steps"].sum())
        d = {"config": f"flatquad forward closed loop 2^{lg}, {law}", "dtype": name, "ms": float(np.median(ts[1:])),
             "trajectories_per_s": n / np.median(ts[1:]) * 1e3, "attempts_per_s": att / np.median(ts[1:]) * 1e3, "mean_attempts": att / n}
        if "nn" in kw:
            d["mlp_gflops"] = att * 6 * 8 * 2 * 2 * (6 * 64 + 64 * 64 * 2 + 64) / np.median(ts[1:]) / 1e6
        print(json.dumps(d), flush=True)
