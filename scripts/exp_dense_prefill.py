#!/usr/bin/env python
"""Experiment behind DESIGN.md 3.4: does the +inf padding of the dense output overlap the integration better when it streams
sequentially from a separate kernel (here: torch fill_ on a side stream) than when the retiring warps write it row by row?
Run once with the regular library and once with a variant built with -DPMP_EXP_NOPAD (real nodes only):
    python scripts/build_variant.py nopad pmp_sys_flatquad.cu -DPMP_EXP_NOPAD
    B200PMP_LIB=approximate_optimal_control_b200/build/variants/lib_nopad.so python scripts/exp_dense_prefill.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from common import c_problem, flatquad_terminal
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu

n = 1 << 20
prob, _ = c_problem("flatquad")
y = torch.as_tensor(flatquad_terminal(n), dtype=torch.float32, device="cuda").contiguous()
opts = _capi.make_opts(method="tsit5", dtype=_capi.F32, save_dense=1)
out = {}
pu.solve_batch_soa(prob, opts, y, out)
torch.cuda.synchronize()
dense = [v for k, v in out.items() if isinstance(v, torch.Tensor) and v.numel() >= n * 129]
side = torch.cuda.Stream()

L = _capi.lib()
pat = torch.full((4,), float("inf"), dtype=torch.float32, device="cuda")
FB = int(os.environ.get("FILL_BLOCKS", "148"))


def fill_all(stream):
    for t in dense:
        nb = t.numel() * 4 // 16 * 16
        _capi.check(L.pmp_host_probe_copy(t.data_ptr(), pat.data_ptr(), nb, -FB, stream.cuda_stream), L)


def timed(with_fill):
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        if with_fill:
            side.wait_stream(torch.cuda.current_stream())
            fill_all(side)
        pu.solve_batch_soa(prob, opts, y, out)
        if with_fill:
            torch.cuda.current_stream().wait_stream(side)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[1:]))

fill_only = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    fill_all(torch.cuda.current_stream())
    b.record(); torch.cuda.synchronize(); fill_only.append(a.elapsed_time(b))
print(json.dumps({"lib": os.environ.get("B200PMP_LIB", "regular"), "dense_tensors": len(dense), "bytes": sum(t.numel() * 4 for t in dense),
                  "solve_ms": timed(False), "solve_with_concurrent_fill_ms": timed(True), "fill_only_ms": float(np.median(fill_only[1:])), "fill_blocks": FB}))
