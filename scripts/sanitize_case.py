"""Small dense + endpoints solves for compute-sanitizer (memcheck / racecheck on the shared-memory staging)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from common import flatquad_terminal, orbits_terminal, c_problem, cuda_solve
from approximate_optimal_control_b200 import _capi
prob, _ = c_problem("flatquad")
for n in (1000, 4099):
    y = flatquad_terminal(n)
    y[:, 7] = np.nan
    for dt, nd in ((_capi.F32, np.float32), (_capi.F64, np.float64)):
        for dense in (0, 1):
            out = cuda_solve(prob, _capi.make_opts(dtype=dt, save_dense=dense, max_steps=40), y.astype(nd))
            assert out["result"][7] == 4
po, _ = c_problem("orbits")
out = cuda_solve(po, _capi.make_opts(dtype=_capi.F64, save_dense=1, rtol=1e-4, atol=1e-4, max_steps=64, t1=-50.0, dt0=-1 / 16,
                                     event=1, event_level=4.2, dtmin=0.0, dtmax=0.0), orbits_terminal(3000))
print("ok", np.bincount(out["result"]))
