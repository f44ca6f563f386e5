// Instantiates the integrator with value-Hessian propagation (PMP_FLAG_VXX, pontryagin_utils.py:460-467)
// for the flat quadrotor, the only system of the reference that runs with pontryagin_solver_vxx.
#include "pmp_solve.cuh"
namespace pmp {
int solve_flatquad_vxx(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
                       const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    return launch_solve_vxx<Flatquad>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);
}
}  // namespace pmp
