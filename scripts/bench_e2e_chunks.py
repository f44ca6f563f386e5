"""Host in -> host out (define_backward_solver -> pmp_solve_host): wall time per call against the number of pieces the
batch is cut into (-1: copy in, one launch, copy out with torch).  python scripts/bench_e2e_chunks.py [log2n]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from approximate_optimal_control_b200 import pontryagin_utils as pu, systems, terminal  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log2n
pp = systems.flatquad_problem_params()
K, P = pu.get_terminal_lqr(pp)
st = terminal.sample_terminal_states(P, 0.05, n, seed=0)
host_in = {k: torch.as_tensor(v, dtype=torch.float32).pin_memory() for k, v in st.items()}
ap = {"pontryagin_solver_rtol": 1e-5, "pontryagin_solver_atol": 1e-5, "pontryagin_solver_T": 2.0,
      "pontryagin_solver_maxsteps": 256, "pontryagin_solver_save_steps": False, "dtype": "float32"}
for chunks in (-1, 1, 0, 2, 4, 6, 8, 12, 16):
    solve = pu.define_backward_solver(pp, dict(ap, pontryagin_solver_host_chunks=chunks))[0]

    def step():
        if chunks < 0:
            dev_in = {k: v.to("cuda", non_blocking=True) for k, v in host_in.items()}
            ye = solve(dev_in).y_end
            out = {k: v.to("cpu", non_blocking=True) for k, v in ye.items()}
            torch.cuda.synchronize()
            return out
        return solve(host_in)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(json.dumps({"log2n": log2n, "pieces": {-1: "device tensors, one launch", 0: "auto"}.get(chunks, chunks),
                      "ms_per_call": round(ms, 3), "traj_per_s": n / ms * 1e3}))
