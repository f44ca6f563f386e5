// Instantiates the forward-mode sensitivity kernels (pmp_sens.cuh) for every compiled system (or, in a codegen plug-in,
// for the generated one).
#include "pmp_sens.cuh"
namespace pmp {
int solve_jvp(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, const void* dy_f, void* y_end, void* dy_end,
              int32_t* a, int32_t* s, int32_t* r, cudaStream_t st) {
#ifdef PMP_PLUGIN
    return launch_sens<Generated>(p, o, n, y_f, dy_f, y_end, dy_end, a, s, r, st);
#else
    switch (p.system_id) {
        case PMP_SYS_FLATQUAD: return launch_sens<Flatquad>(p, o, n, y_f, dy_f, y_end, dy_end, a, s, r, st);
        case PMP_SYS_ORBITS: return launch_sens<Orbits>(p, o, n, y_f, dy_f, y_end, dy_end, a, s, r, st);
        case PMP_SYS_LQR_STATEPENALTY: return launch_sens<LqrStatePenalty>(p, o, n, y_f, dy_f, y_end, dy_end, a, s, r, st);
    }
    return -1000;
#endif
}
}  // namespace pmp
