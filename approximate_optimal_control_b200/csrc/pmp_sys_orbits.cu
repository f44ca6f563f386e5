// Instantiates the integrator for the orbits system (experiment_orbits.py:24-40).
#include "pmp_solve.cuh"
namespace pmp {
int solve_orbits(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
                 const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    return launch_solve<Orbits, 4, 3>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);
}
int solve_orbits_hostq(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st) {
    return launch_hostq<Orbits, 4, 3>(p, o, n, q, st);
}
int solve_orbits_push(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end, int32_t* a, int32_t* s,
                   int32_t* r, void* ws, const PushArgs& push, cudaStream_t st) {
    return launch_solve_push<Orbits, 4, 3>(p, o, n, y_f, y_end, a, s, r, ws, push, st);
}
}  // namespace pmp
