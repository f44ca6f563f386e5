// Plug-in dynamics with the value-Hessian leaf (PMP_FLAG_VXX, pontryagin_utils.py:460-467): the generated functor wrapped in
// the forward-mode adaptor of pmp_vxx_dual.cuh, so that define_backward_solver(..., pontryagin_solver_vxx=True) works for any
// f(x,u), l(x,u) as the reference's jax.jacobian(pmp_rhs) does.
#ifndef PMP_PLUGIN
#error "pmp_sys_generated_vxx.cu is a compile unit of codegen plug-ins only"
#endif
#include "pmp_solve.cuh"
#include "pmp_vxx_dual.cuh"
namespace pmp {
template <class R> using GeneratedVxx = DualVxx<Generated, R>;
int solve_generated_vxx(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
                        const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    return launch_solve_vxx<GeneratedVxx, 7, 1>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);
}
}  // namespace pmp
