// Instantiates the integrator for the state-penalised double integrator
// (experiment_lqr_statepenalty.py:23-38).
#include "pmp_solve.cuh"
namespace pmp {
int solve_lqr(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
              const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    return launch_solve<LqrStatePenalty, 4, 3>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);
}
int solve_lqr_hostq(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st) {
    return launch_hostq<LqrStatePenalty, 4, 3>(p, o, n, q, st);
}
int solve_lqr_push(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end, int32_t* a, int32_t* s,
                   int32_t* r, void* ws, const PushArgs& push, cudaStream_t st) {
    return launch_solve_push<LqrStatePenalty, 4, 3>(p, o, n, y_f, y_end, a, s, r, ws, push, st);
}
}  // namespace pmp
