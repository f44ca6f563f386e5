// Instantiates the dense-output (Solution.evaluate) kernels for the three systems.
#include "pmp_interp.cuh"
namespace pmp {
int interp_batch(const pmp_problem_t& p, const pmp_opts_t& o, long long nq, const void* y0, const void* t1,
                 const void* target, void* y_out, int mode, cudaStream_t st) {
#ifdef PMP_PLUGIN
    return launch_interp<Generated>(p, o, nq, y0, t1, target, y_out, mode, st);
#else
    switch (p.system_id) {
        case PMP_SYS_FLATQUAD: return launch_interp<Flatquad>(p, o, nq, y0, t1, target, y_out, mode, st);
        case PMP_SYS_ORBITS: return launch_interp<Orbits>(p, o, nq, y0, t1, target, y_out, mode, st);
        case PMP_SYS_LQR_STATEPENALTY: return launch_interp<LqrStatePenalty>(p, o, nq, y0, t1, target, y_out, mode, st);
    }
    return -1000;
#endif
}
}  // namespace pmp
