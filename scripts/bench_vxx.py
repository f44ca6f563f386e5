#!/usr/bin/env python
"""Kernel time of the value-Hessian branch (PMP_FLAG_VXX, row N1 of SURVEY.md 8(f)) on one GPU: general
36-entry storage and symmetric upper-triangle storage, fp32 / fp64, endpoints and dense.
usage: python scripts/bench_vxx.py [log2_n] [--generated] [> profiles/rNN_vxx.jsonl]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from common import c_problem, flatquad_terminal_vxx
from approximate_optimal_control_b200 import _capi, flops, pontryagin_utils as pu


def timed(problem, opts, y, reps=5):
    out = {}
    ts = []
    for i in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); pu.solve_batch_soa(problem, opts, y, out); b.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts)), out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    lg = int(args[0]) if args else 18
    n = 1 << lg
    prob, _ = c_problem("flatquad")
    generated = "--generated" in sys.argv  # the functor generated from f, l: vxx by forward mode (csrc/pmp_vxx_dual.cuh)
    if generated:
        from approximate_optimal_control_b200 import systems
        prob = systems.to_c_problem(dict(systems.flatquad_problem_params(), codegen=True))
    y64 = flatquad_terminal_vxx(n)
    V, S = _capi.FLAG_VXX, _capi.FLAG_VXX | _capi.FLAG_VXX_SYMMETRIC
    cases = [("tsit5", 0, 0), ("tsit5", 0, V), ("tsit5", 0, S), ("dopri5", 0, S), ("rk4", 0, S), ("tsit5", 1, V), ("tsit5", 1, S)]
    for dtype, tdt, name in ((_capi.F32, torch.float32, "f32"), (_capi.F64, torch.float64, "f64")):
        for method, dense, flags in cases:
            if dense and n > (1 << 18):
                continue
            rows = 50 if flags else 14
            y = torch.as_tensor(y64[:rows], dtype=tdt, device="cuda").contiguous()
            kw = dict(method=method, dtype=dtype, save_dense=dense, flags=flags)
            if method == "rk4":
                kw.update(adaptive=0, dt0=-1 / 32, max_steps=128)
            opts = _capi.make_opts(**kw)
            med, best, out = timed(prob, opts, y)
            attempts = int(out["n_steps"].sum().item())
            d = {"config": f"{'GENERATED ' if generated else ''}flatquad 2^{lg} {method} {'dense' if dense else 'endpoints'} "
                           f"{'vxx-symmetric' if flags == S else ('vxx' if flags == V else 'no vxx')}",
                 "dtype": name, "ms": med, "ms_best": best, "trajectories_per_s": n / med * 1e3,
                 "attempts_per_s": attempts / med * 1e3, "mean_attempts": attempts / n,
                 "result_counts": np.bincount(out["result"].cpu().numpy(), minlength=5).tolist()}
            if flags:
                d["reference_algorithm_tflops"] = flops.solve_flops_vxx(opts.method, attempts, n) / med / 1e9
            if dense:
                d["dense_GBps"] = n * 129 * (15 + (36 if flags else 0)) * (4 if dtype == _capi.F32 else 8) / med / 1e6
            print(json.dumps(d), flush=True)
            del out, y


if __name__ == "__main__":
    main()
