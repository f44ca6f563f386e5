import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from common import flatquad_terminal_vxx, c_problem
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu
prob, _ = c_problem("flatquad")
S = _capi.FLAG_VXX | _capi.FLAG_VXX_SYMMETRIC
y64 = flatquad_terminal_vxx(1 << 20)
def run(name, lg, tdt, **kw):
    n = 1 << lg
    opts = _capi.make_opts(**kw)
    rows = 50 if kw.get("flags") else 14
    yy = torch.as_tensor(y64[:rows, :n], dtype=tdt, device="cuda").contiguous()
    res = []
    for rm in (1, 2, 3, 4, 6, 8, 12, 16):
        os.environ["B200PMP_REFILL_MIN"] = str(rm)
        out = {}
        ts = []
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); pu.solve_batch_soa(prob, opts, yy, out); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res.append(f"{rm}:{min(ts[1:]):.3f}")
    print(name, " ".join(res), flush=True)
run("f32 tsit5 2^20 endpoints", 20, torch.float32, dtype=_capi.F32)
run("f32 dopri5 2^20 endpoints", 20, torch.float32, dtype=_capi.F32, method="dopri5")
run("f32 tsit5 2^20 dense", 20, torch.float32, dtype=_capi.F32, save_dense=1)
run("f64 tsit5 2^20 endpoints", 20, torch.float64, dtype=_capi.F64)
run("f32 tsit5 2^16 endpoints", 16, torch.float32, dtype=_capi.F32)
run("f32 tsit5 2^18 vxx sym", 18, torch.float32, dtype=_capi.F32, flags=S)
run("f32 tsit5 2^18 vxx gen", 18, torch.float32, dtype=_capi.F32, flags=_capi.FLAG_VXX)
run("f64 tsit5 2^17 vxx sym", 17, torch.float64, dtype=_capi.F64, flags=S)
