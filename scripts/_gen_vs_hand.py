"""generated (codegen) flatquad functor vs the hand-written one: kernel time on 2^20 trajectories"""
import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from common import flatquad_terminal
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu, systems
y64 = flatquad_terminal(1 << 20)
pg = dict(systems.flatquad_problem_params()); pg["codegen"] = True
probs = {"hand-written": systems.to_c_problem(systems.flatquad_problem_params()), "generated": systems.to_c_problem(pg)}
for dtype, tdt, name in ((_capi.F32, torch.float32, "f32"), (_capi.F64, torch.float64, "f64")):
    y = torch.as_tensor(y64, dtype=tdt, device="cuda").contiguous()
    for k, prob in probs.items():
        for dense in (0, 1):
            opts = _capi.make_opts(dtype=dtype, save_dense=dense)
            out = {}; ts = []
            for i in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record(); pu.solve_batch_soa(prob, opts, y, out); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            print(json.dumps({"config": f"flatquad 2^20 tsit5 {'dense' if dense else 'endpoints'}", "functor": k, "dtype": name,
                              "ms": float(np.median(ts[1:])), "attempts": int(out["n_steps"].sum())}), flush=True)
            del out
