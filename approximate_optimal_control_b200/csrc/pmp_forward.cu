// Instantiates the forward closed-loop simulation kernels (pmp_forward.cuh) for the compiled systems, or for the
// generated one in a codegen plug-in.
#include "pmp_forward.cuh"
namespace pmp {
int forward_nn_batch(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const pmp_nn_costate_t* nn,
                     void* z_end, void* ts, void* xs, void* cost, int32_t* a, int32_t* s, int32_t* r, cudaStream_t st);

int forward_batch(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const double* P,
                  const double* x_eq, const pmp_nn_costate_t* nn, void* z_end, void* ts, void* xs, void* cost, int32_t* a,
                  int32_t* s, int32_t* r, cudaStream_t st) {
    if (nn) return forward_nn_batch(p, o, n, x0, nn, z_end, ts, xs, cost, a, s, r, st);
#ifdef PMP_PLUGIN
    return launch_forward<Generated>(p, o, n, x0, P, x_eq, z_end, ts, xs, cost, a, s, r, st);
#else
    switch (p.system_id) {
        case PMP_SYS_FLATQUAD: return launch_forward<Flatquad>(p, o, n, x0, P, x_eq, z_end, ts, xs, cost, a, s, r, st);
        case PMP_SYS_ORBITS: return launch_forward<Orbits>(p, o, n, x0, P, x_eq, z_end, ts, xs, cost, a, s, r, st);
        case PMP_SYS_LQR_STATEPENALTY: return launch_forward<LqrStatePenalty>(p, o, n, x0, P, x_eq, z_end, ts, xs, cost, a, s, r, st);
    }
    return -1000;
#endif
}
}  // namespace pmp
