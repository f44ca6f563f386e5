// Instantiates the tcgen05 ensemble kernels (pmp_forward_umma.cuh) for the compiled systems, or for the generated one in a
// codegen plug-in.
#include "pmp_forward_umma.cuh"
namespace pmp {
template <class Sys>
int launch_forward_umma_entry(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<float>& io,
                              const NnCostate<float, Sys::NX, 64>& pol, cudaStream_t st) {
    return launch_forward_umma<Sys>(p, o, n, io, pol, st);
}
#define PMP_INST(SYS)                                                                                                               \
    template int launch_forward_umma_entry<SYS<float>>(const pmp_problem_t&, const pmp_opts_t&, long long, const ForwardIO<float>&, \
                                                       const NnCostate<float, SYS<float>::NX, 64>&, cudaStream_t);
#ifdef PMP_PLUGIN
#if PMP_GEN_NX <= 8
PMP_INST(Generated)
#endif
#else
PMP_INST(Flatquad)
PMP_INST(Orbits)
PMP_INST(LqrStatePenalty)
#endif
}  // namespace pmp
