"""approximate_optimal_control_b200 -- B200-native batched Pontryagin backward-shooting rollouts.

One hot path of mbjd/approximate_optimal_control (pontryagin_utils.define_backward_solver ->
f_extended / solve_backward, vmapped by levelsets.testbed) as hand-written sm_100a CUDA behind the
reference's own Python call signatures.  See DESIGN.md.

    from approximate_optimal_control_b200 import pontryagin_utils, systems
    problem_params, algo_params = systems.flatquad_problem_params(), systems.flatquad_algo_params()
    solve_backward, f_extended = pontryagin_utils.define_backward_solver(problem_params, algo_params)
"""
from . import _capi, build, sharding, systems  # noqa: F401
from . import pontryagin_utils  # noqa: F401
from . import levelsets  # noqa: F401
from .solution import Solution  # noqa: F401
from .terminal import sample_terminal_states  # noqa: F401

__version__ = "0.1.0"
