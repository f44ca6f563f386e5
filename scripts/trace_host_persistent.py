import os, sys, torch, numpy as np
sys.path.insert(0, ".")
from approximate_optimal_control_b200 import levelsets, pontryagin_utils as pu, systems, terminal
n = 1 << 20
pp, ap = systems.flatquad_problem_params(), systems.flatquad_algo_params()
_, P = pu.get_terminal_lqr(pp)
st = terminal.sample_terminal_states(P, pp["V_f"], n, seed=1, dtype=np.float32)
host_in = torch.as_tensor(st["x"]).pin_memory()  # the reference's call: xfs only (levelsets.py:601)
solve = levelsets.define_solve_backward_lqr(pp, dict(ap, pontryagin_solver_save_steps=False, dtype="float32"), P)
for _ in range(3):
    solve(host_in)
os.environ["B200PMP_HOST_TRACE"] = "1"
solve(host_in)
import time
os.environ.pop("B200PMP_HOST_TRACE")
for _ in range(3):
    solve(host_in)
t = time.perf_counter()
for _ in range(20):
    solve(host_in)
print(f"python call: {(time.perf_counter() - t) / 20 * 1e3:.3f} ms per 2^20", file=sys.stderr)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    solve(host_in)
pr.disable()
pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(12)
