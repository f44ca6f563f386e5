// Instantiates the forward closed-loop simulation kernels with the value-network-ensemble costate law (NnCostate,
// pmp_forward.cuh; levelsets.py:838-933, eval_utils.py:10-45) -- a compile unit of its own (36 kernels).
#include "pmp_forward.cuh"
namespace pmp {
int forward_nn_batch(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const pmp_nn_costate_t* nn,
                     void* z_end, void* ts, void* xs, void* cost, int32_t* a, int32_t* s, int32_t* r, cudaStream_t st) {
#ifdef PMP_PLUGIN
    return launch_forward_nn<Generated>(p, o, n, x0, nn, z_end, ts, xs, cost, a, s, r, st);
#else
    switch (p.system_id) {
        case PMP_SYS_FLATQUAD: return launch_forward_nn<Flatquad>(p, o, n, x0, nn, z_end, ts, xs, cost, a, s, r, st);
        case PMP_SYS_ORBITS: return launch_forward_nn<Orbits>(p, o, n, x0, nn, z_end, ts, xs, cost, a, s, r, st);
        case PMP_SYS_LQR_STATEPENALTY: return launch_forward_nn<LqrStatePenalty>(p, o, n, x0, nn, z_end, ts, xs, cost, a, s, r, st);
    }
    return -1000;
#endif
}

int nn_value_batch(const pmp_problem_t& p, const pmp_nn_costate_t* nn, int dtype, long long n, const void* x, void* lam,
                   void* vm, void* vs, cudaStream_t st) {
#ifdef PMP_PLUGIN
    return launch_nn_value<Generated<float>::NX>(nn, dtype, n, x, lam, vm, vs, st);
#else
    if (p.nx == 6) return launch_nn_value<6>(nn, dtype, n, x, lam, vm, vs, st);
    if (p.nx == 2) return launch_nn_value<2>(nn, dtype, n, x, lam, vm, vs, st);
    return -1000;
#endif
}
}  // namespace pmp
