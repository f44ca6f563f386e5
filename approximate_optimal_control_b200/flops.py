"""Algorithmic work per unit of the PMP path -- the numerator of the FMA roofline.

Counting rule: add, sub, mul = 1 flop (an FMA = 2); div, sqrt = 1; sin, cos, pow are "special" and
are NOT counted as flops; min / max / abs / negate / compare / select = 0.  The integers below are
what the CPU oracle reports when its RHS / stepper templates are instantiated on an op-counting
scalar (oracle/opcount.hpp; u* computed once per RHS with the quadratic model, problem constants
folded); tests/test_oracle.py::test_flop_constants_match_oracle keeps them in sync.  They follow the
reference's ALGORITHM (all five u* candidates scored twice, every stage combination of the tableau),
not the instructions the CUDA kernel happens to execute.
"""
from . import _capi

# (system_id) -> flops / specials of one f_extended evaluation
RHS_FLOPS = {_capi.SYS_FLATQUAD: 204, _capi.SYS_ORBITS: 64, _capi.SYS_LQR_STATEPENALTY: 27}
RHS_SPECIAL = {_capi.SYS_FLATQUAD: 2, _capi.SYS_ORBITS: 0, _capi.SYS_LQR_STATEPENALTY: 0}

# (system_id, method) -> flops of one step attempt (stage combinations + RHS evaluations + error
# estimate + RMS norm + controller); adaptive for Tsit5/Dopri5, fixed step for RK4
ATTEMPT_FLOPS = {
    (_capi.SYS_FLATQUAD, _capi.METHOD_TSIT5): 2175,
    (_capi.SYS_FLATQUAD, _capi.METHOD_DOPRI5): 2175,
    (_capi.SYS_FLATQUAD, _capi.METHOD_RK4): 1160,
    (_capi.SYS_ORBITS, _capi.METHOD_TSIT5): 799,
    (_capi.SYS_LQR_STATEPENALTY, _capi.METHOD_RK4): 260,
}
ATTEMPT_SPECIAL = {
    (_capi.SYS_FLATQUAD, _capi.METHOD_TSIT5): 13,
    (_capi.SYS_FLATQUAD, _capi.METHOD_DOPRI5): 13,
    (_capi.SYS_FLATQUAD, _capi.METHOD_RK4): 8,
}


# value-Hessian branch (PMP_FLAG_VXX), flatquad: one f_extended with the vxx leaf as the ORACLE evaluates it (closed-form
# Jacobians through u*, then gx + glam vxx - vxx fx - vxx flam vxx with four dense 6x6 products) and one step attempt on
# the 50-scalar state.  The CUDA kernel exploits the sparsity of the Jacobians (Flatquad::vxx_dot, ~1/4 of these flops),
# so a rate quoted on these constants is "reference-algorithm flops", not FMA-pipe utilisation.
RHS_FLOPS_VXX = 2205
ATTEMPT_FLOPS_VXX = {_capi.METHOD_TSIT5: 16593, _capi.METHOD_DOPRI5: 16593, _capi.METHOD_RK4: 10028}


def solve_flops_vxx(method: int, total_attempts: int, n_traj: int) -> int:
    return total_attempts * ATTEMPT_FLOPS_VXX[method] + n_traj * RHS_FLOPS_VXX


def solve_flops(system_id: int, method: int, total_attempts: int, n_traj: int) -> int:
    """flops_total = sum_traj (n_steps_traj * F_attempt + F_rhs)   (the F_rhs is the FSAL start-up)."""
    return total_attempts * ATTEMPT_FLOPS[(system_id, method)] + n_traj * RHS_FLOPS[system_id]


def dense_bytes(n_traj: int, nx: int, max_steps: int, itemsize: int) -> int:
    """Dense (SaveAt(steps=True)) output: every slot of ts, t, v, x[nx], vx[nx] written exactly once."""
    return n_traj * (max_steps + 1) * (2 * nx + 3) * itemsize
