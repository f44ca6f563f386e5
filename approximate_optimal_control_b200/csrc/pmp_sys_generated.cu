// Plug-in dynamics: instantiates the integrator, the pointwise kernels and the dense-output
// kernels for ONE system whose functor was generated from the user's f(x,u), l(x,u) by codegen.py (sympy takes
// the place of jax autodiff, pontryagin_utils.py:35-38,438-439).  Built with -DPMP_PLUGIN
// -DPMP_GENERATED_HEADER="<file>" into a library of its own that exports the same C ABI (include/b200pmp.h) and
// answers to system_id PMP_SYS_GENERATED only.
#ifndef PMP_PLUGIN
#error "pmp_sys_generated.cu is a compile unit of codegen plug-ins only"
#endif
#include "pmp_solve.cuh"
namespace pmp {
int generated_nx() { return Generated<float>::NX; }
int generated_nu() { return Generated<float>::NU; }
int solve_generated(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
                    const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    return launch_solve<Generated, PMP_GEN_MINB32, PMP_GEN_MINB64>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);
}
int solve_generated_hostq(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st) {
    return launch_hostq<Generated, PMP_GEN_MINB32, PMP_GEN_MINB64>(p, o, n, q, st);
}
int solve_generated_push(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end, int32_t* a, int32_t* s,
                   int32_t* r, void* ws, const PushArgs& push, cudaStream_t st) {
    return launch_solve_push<Generated, PMP_GEN_MINB32, PMP_GEN_MINB64>(p, o, n, y_f, y_end, a, s, r, ws, push, st);
}
}  // namespace pmp
