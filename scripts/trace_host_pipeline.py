"""Stage timestamps of pmp_solve_host (B200PMP_HOST_TRACE=1) for 2^20 flatquad rollouts from pinned host tensors."""
import os, sys, torch
sys.path.insert(0, ".")
from approximate_optimal_control_b200 import pontryagin_utils as pu, systems, terminal
n = 1 << 20
pp = systems.flatquad_problem_params()
K, P = pu.get_terminal_lqr(pp)
st = terminal.sample_terminal_states(P, 0.05, n, seed=0)
host_in = {k: torch.as_tensor(v, dtype=torch.float32).pin_memory() for k, v in st.items()}
ap = {"pontryagin_solver_rtol": 1e-5, "pontryagin_solver_atol": 1e-5, "pontryagin_solver_T": 2.0,
      "pontryagin_solver_maxsteps": 256, "pontryagin_solver_save_steps": False, "dtype": "float32"}
for pieces in (0, 4, 8):
    solve = pu.define_backward_solver(pp, dict(ap, pontryagin_solver_host_chunks=pieces))[0]
    os.environ.pop("B200PMP_HOST_TRACE", None)
    for _ in range(3):
        solve(host_in)
    os.environ["B200PMP_HOST_TRACE"] = "1"
    print(f"--- pieces {pieces}", file=sys.stderr, flush=True)
    solve(host_in)
