"""One vxx solve for ncu: python scripts/vxx_case.py [f32|f64] [sym|gen] [log2 n] [method] [dense]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from common import flatquad_terminal_vxx, c_problem
from approximate_optimal_control_b200 import _capi, pontryagin_utils as pu
a = sys.argv[1:] + ["f32", "sym", "17", "tsit5", "0"][len(sys.argv) - 1:]
dt, tdt = (_capi.F32, torch.float32) if a[0] == "f32" else (_capi.F64, torch.float64)
flags = _capi.FLAG_VXX | (_capi.FLAG_VXX_SYMMETRIC if a[1] == "sym" else 0)
prob, _ = c_problem("flatquad")
y = torch.as_tensor(flatquad_terminal_vxx(1 << int(a[2])), dtype=tdt, device="cuda").contiguous()
kw = dict(method=a[3], dtype=dt, flags=flags, save_dense=int(a[4]))
if a[3] == "rk4":
    kw.update(adaptive=0, dt0=-1 / 32)
out = {}
for _ in range(3):
    pu.solve_batch_soa(prob, _capi.make_opts(**kw), y, out)
torch.cuda.synchronize()
print("ok", int(out["n_steps"].sum()))
