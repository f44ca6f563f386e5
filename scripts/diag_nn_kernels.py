import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from common import (O, _capi, as_oracle_opts, c_problem, oracle_problem, scaled_err)
from approximate_optimal_control_b200 import pontryagin_utils as pu
from test_gpu_forward import _ensemble
tol = 1e-5
prob, _ = c_problem("flatquad")
for dims, E in (((64, 64, 64), 8), ((64, 64), 5), ((64,), 3)):
    ens, onn, _w = _ensemble(6, dims, E, torch.float32, seed=3)
    _e64, onn64, _w64 = _ensemble(6, dims, E, torch.float64, seed=3)
    x0 = 0.5 * np.random.Generator(np.random.PCG64(11)).standard_normal((6, 1000))
    mk = lambda dtype, t: _capi.make_opts(method="tsit5", dtype=dtype, rtol=t, atol=t, t0=0.0, t1=1.5, dt0=0.1, dtmin=0.0, dtmax=0.0, max_steps=256)
    opts = mk(_capi.F32, tol)
    truth = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(mk(_capi.F64, 1e-11)), x0, nn=onn64)
    ref = O.forward_batch(oracle_problem("flatquad"), as_oracle_opts(opts), x0.astype(np.float32), nn=onn)
    floor = scaled_err(ref["z_end"], truth["z_end"], tol)
    print(dims, E, "oracle fp32 vs truth p50/p90/p99/max", [round(float(np.percentile(floor, p)), 2) for p in (50, 90, 99, 100)], "rejected frac", 1 - ref["n_accepted"].sum() / ref["n_steps"].sum())
    for name, env in (("tcgen05", {}), ("warp-mma", {"B200PMP_NN_UMMA": "0"}), ("cuda-core", {"B200PMP_NN_TC": "0"})):
        for k in ("B200PMP_NN_UMMA", "B200PMP_NN_TC"): os.environ.pop(k, None)
        os.environ.update(env)
        out = pu.forward_batch_soa(prob, opts, torch.as_tensor(x0, dtype=torch.float32, device="cuda").contiguous(), nn=ens)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in out.items()}
        same = (got["n_steps"] == ref["n_steps"]) & (got["n_accepted"] == ref["n_accepted"])
        err = scaled_err(got["z_end"], truth["z_end"], tol)
        print("  ", name, "same steps", round(float(same.mean()), 3), "vs truth p50/p90/p99/max", [round(float(np.percentile(err, p)), 2) for p in (50, 90, 99, 100)],
              "vs oracle32 (same) max", round(float(scaled_err(got["z_end"], ref["z_end"], tol)[same].max()), 2))
