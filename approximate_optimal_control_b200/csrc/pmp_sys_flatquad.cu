// Instantiates the integrator for the flat quadrotor (flatquad_landing_experiment.py:269-365).
#include "pmp_solve.cuh"
namespace pmp {
int solve_flatquad(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
                   const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    #ifndef PMP_MINB32
#define PMP_MINB32 3
#endif
    return launch_solve<Flatquad, PMP_MINB32, 1>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);
}
int solve_flatquad_hostq(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st) {
    return launch_hostq<Flatquad, PMP_MINB32, 1>(p, o, n, q, st);
}
int solve_flatquad_push(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end, int32_t* a, int32_t* s,
                   int32_t* r, void* ws, const PushArgs& push, cudaStream_t st) {
    return launch_solve_push<Flatquad, PMP_MINB32, 1>(p, o, n, y_f, y_end, a, s, r, ws, push, st);
}
}  // namespace pmp
