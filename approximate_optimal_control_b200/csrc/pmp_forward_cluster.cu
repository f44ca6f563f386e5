// Instantiates the cluster kernels of the value-network-ensemble costate law (pmp_forward_cluster.cuh) for the compiled
// systems, or for the generated one in a codegen plug-in.
#include "pmp_forward_cluster.cuh"
namespace pmp {
template <class Sys, class R, int W>
int launch_forward_cluster(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<R>& io,
                           const NnCostate<R, Sys::NX, W>& pol, cudaStream_t st) {
    return launch_forward_cluster_w<Sys, R, W>(p, o, n, io, pol, st);
}
#define PMP_INST(SYS, R, W)                                                                                        \
    template int launch_forward_cluster<SYS<R>, R, W>(const pmp_problem_t&, const pmp_opts_t&, long long, const ForwardIO<R>&, \
                                                       const NnCostate<R, SYS<R>::NX, W>&, cudaStream_t);
#define PMP_INST_SYS(SYS) PMP_INST(SYS, float, 64) PMP_INST(SYS, float, 32) PMP_INST(SYS, double, 64) PMP_INST(SYS, double, 32)
#ifdef PMP_PLUGIN
PMP_INST_SYS(Generated)
#else
PMP_INST_SYS(Flatquad)
PMP_INST_SYS(Orbits)
PMP_INST_SYS(LqrStatePenalty)
#endif
}  // namespace pmp
