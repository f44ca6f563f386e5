// pmp_eval.cu -- batched pointwise evaluations of the plug-in dynamics and the FMA-peak probe.
//   rhs_batch   <- jax.vmap(f_extended)             pontryagin_utils.py:418-482
//   ustar_batch <- jax.vmap(u_star_2d / u_star_new) pontryagin_utils.py:11-216, :532-562
//                  (batched over all saved nodes at levelsets.py:604-628)
#include "pmp_common.cuh"
#include "pmp_systems.cuh"
#ifdef PMP_PLUGIN
#include "pmp_vxx_dual.cuh"
#endif

namespace pmp {

template <class Sys, class R>
__global__ void __launch_bounds__(256) rhs_kernel(const typename Sys::Consts sc, const R* __restrict__ y,
                                                  R* __restrict__ dy, long long n) {
    constexpr int NX = Sys::NX;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R x[NX], lam[NX], xd[NX], ld[NX], vd;
#pragma unroll
    for (int q = 0; q < NX; q++) { x[q] = y[q * n + i]; lam[q] = y[(NX + 2 + q) * n + i]; }
    Sys::template rhs<0>(sc, x, lam, xd, vd, ld);  // literal u*
#pragma unroll
    for (int q = 0; q < NX; q++) { dy[q * n + i] = xd[q]; dy[(NX + 2 + q) * n + i] = ld[q]; }
    dy[NX * n + i] = R(1);  // state_dot['t'] = 1.   (pontryagin_utils.py:456)
    dy[(NX + 1) * n + i] = vd;
}

// f_extended with algo_params['pontryagin_solver_vxx'] (pontryagin_utils.py:460-467): y, dy [NS + nx*nx][n]
template <class Sys, class R>
__global__ void __launch_bounds__(128) rhs_vxx_kernel(const typename Sys::Consts sc, const R* __restrict__ y,
                                                      R* __restrict__ dy, long long n) {
    constexpr int NX = Sys::NX, NS = 2 * NX + 2, NV = NX * NX;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R x[NX], lam[NX], xd[NX], ld[NX], vd, V[NV];
#pragma unroll
    for (int q = 0; q < NX; q++) { x[q] = y[q * n + i]; lam[q] = y[(NX + 2 + q) * n + i]; }
#pragma unroll
    for (int q = 0; q < NV; q++) V[q] = y[(NS + q) * n + i];
    typename Sys::JacCoef jc;
    Sys::rhs_jac(sc, x, lam, xd, vd, ld, jc);
#pragma unroll
    for (int q = 0; q < NX; q++) { dy[q * n + i] = xd[q]; dy[(NX + 2 + q) * n + i] = ld[q]; }
    dy[NX * n + i] = R(1);
    dy[(NX + 1) * n + i] = vd;
    Sys::template vxx_dot<1>(sc, jc, V, [&](int q, R val) { dy[(NS + q) * n + i] = val; });
}

template <class Sys, class R, int UMODE>
__global__ void __launch_bounds__(256) ustar_kernel(const typename Sys::Consts sc, const R* __restrict__ x_,
                                                    const R* __restrict__ lam_, R* __restrict__ u_, long long n) {
    constexpr int NX = Sys::NX, NU = Sys::NU;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R x[NX], lam[NX], u[NU];
#pragma unroll
    for (int q = 0; q < NX; q++) { x[q] = x_[q * n + i]; lam[q] = lam_[q * n + i]; }
    if (UMODE == 0) Sys::ustar(sc, x, lam, u);
    else Sys::template ustar_mode<UMODE>(sc, x, lam, u);
#pragma unroll
    for (int q = 0; q < NU; q++) u_[q * n + i] = u[q];
}

template <template <class> class SysT>
int eval_dispatch(int what, const pmp_problem_t& p, int dtype, long long n, const void* a, const void* b, void* out,
                  cudaStream_t st) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == PMP_F32) {
        using S = SysT<float>;
        if (what == 0) rhs_kernel<S, float><<<grid, 256, 0, st>>>(S::make(p), (const float*)a, (float*)out, n);
        else if (what == 1) ustar_kernel<S, float, 0><<<grid, 256, 0, st>>>(S::make(p), (const float*)a, (const float*)b, (float*)out, n);
        else if (what == 4) ustar_kernel<S, float, 4><<<grid, 256, 0, st>>>(S::make(p), (const float*)a, (const float*)b, (float*)out, n);
        else ustar_kernel<S, float, 2><<<grid, 256, 0, st>>>(S::make(p), (const float*)a, (const float*)b, (float*)out, n);
    } else {
        using S = SysT<double>;
        if (what == 0) rhs_kernel<S, double><<<grid, 256, 0, st>>>(S::make(p), (const double*)a, (double*)out, n);
        else if (what == 1) ustar_kernel<S, double, 0><<<grid, 256, 0, st>>>(S::make(p), (const double*)a, (const double*)b, (double*)out, n);
        else if (what == 4) ustar_kernel<S, double, 4><<<grid, 256, 0, st>>>(S::make(p), (const double*)a, (const double*)b, (double*)out, n);
        else ustar_kernel<S, double, 2><<<grid, 256, 0, st>>>(S::make(p), (const double*)a, (const double*)b, (double*)out, n);
    }
    return (int)cudaGetLastError();
}

int eval_batch(int what, const pmp_problem_t& p, int dtype, long long n, const void* a, const void* b, void* out,
               cudaStream_t st) {
    if (what == 3) {  // f_extended with the vxx leaf: flatquad (closed-form Jacobians) or the generated functor (forward mode)
        if (n == 0) return 0;
        const unsigned grid = (unsigned)((n + 127) / 128);
#ifdef PMP_PLUGIN
        using SF = DualVxx<Generated, float>;
        using SD = DualVxx<Generated, double>;
#else
        using SF = Flatquad<float>;
        using SD = Flatquad<double>;
#endif
        if (dtype == PMP_F32) rhs_vxx_kernel<SF, float><<<grid, 128, 0, st>>>(SF::make(p), (const float*)a, (float*)out, n);
        else rhs_vxx_kernel<SD, double><<<grid, 128, 0, st>>>(SD::make(p), (const double*)a, (double*)out, n);
        return (int)cudaGetLastError();
    }
#ifdef PMP_PLUGIN
    return eval_dispatch<Generated>(what, p, dtype, n, a, b, out, st);
#else
    switch (p.system_id) {
        case PMP_SYS_FLATQUAD: return eval_dispatch<Flatquad>(what, p, dtype, n, a, b, out, st);
        case PMP_SYS_ORBITS: return eval_dispatch<Orbits>(what, p, dtype, n, a, b, out, st);
        case PMP_SYS_LQR_STATEPENALTY: return eval_dispatch<LqrStatePenalty>(what, p, dtype, n, a, b, out, st);
    }
    return -1000;
#endif
}

// ---- FMA peak probe: 16 independent dependent-FMA chains per thread, all in registers ------------
constexpr int kPeakChains = 16, kPeakBlock = 256, kPeakBlocksPerSM = 4;

template <class R, int VARIANT>
__global__ void __launch_bounds__(kPeakBlock) fma_peak_kernel(int iters, R b, R c, R* sink) {
    R a[kPeakChains];
#pragma unroll
    for (int j = 0; j < kPeakChains; j++) a[j] = R(threadIdx.x + j) * R(1e-3);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
#pragma unroll
            for (int j = 0; j < kPeakChains; j++) {
                if (VARIANT == 0) a[j] = Math<R>::fma(a[j], b, c);
                else a[j] = Math<R>::fma(a[j], R(0.999999), R(1e-7));
            }
        }
    }
    R s = R(0);
#pragma unroll
    for (int j = 0; j < kPeakChains; j++) s += a[j];
    if (s == R(-12345.678)) *sink = s;  // never true; keeps the chains alive
}

// VARIANT 2: the same 16 chains as 8 packed pairs (FFMA2, fma.rn.f32x2): same flops, half the issue slots.
// Answers "does FFMA2 raise the FMA peak or only free issue slots?" -- see DESIGN.md 3.5.
__global__ void __launch_bounds__(kPeakBlock) fma2_peak_kernel(int iters, float b, float c, float* sink) {
    float2 a[kPeakChains / 2];
#pragma unroll
    for (int j = 0; j < kPeakChains / 2; j++) a[j] = make_float2((threadIdx.x + 2 * j) * 1e-3f, (threadIdx.x + 2 * j + 1) * 1e-3f);
    const float2 bb = make_float2(b, b), cc = make_float2(c, c);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
#pragma unroll
            for (int j = 0; j < kPeakChains / 2; j++) a[j] = __ffma2_rn(a[j], bb, cc);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kPeakChains / 2; j++) s += a[j].x + a[j].y;
    if (s == -12345.678f) *sink = s;
}

int fma_peak(int dtype, int variant, int iters, void* scratch, double* flops_out, cudaStream_t st) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * kPeakBlocksPerSM;
    if (flops_out) *flops_out = 2.0 * kPeakChains * 8.0 * (double)iters * (double)grid * kPeakBlock;
    if (dtype == PMP_F32) {
        if (variant == 0) fma_peak_kernel<float, 0><<<grid, kPeakBlock, 0, st>>>(iters, 0.999999f, 1e-7f, (float*)scratch);
        else if (variant == 2) fma2_peak_kernel<<<grid, kPeakBlock, 0, st>>>(iters, 0.999999f, 1e-7f, (float*)scratch);
        else fma_peak_kernel<float, 1><<<grid, kPeakBlock, 0, st>>>(iters, 0.f, 0.f, (float*)scratch);
    } else {
        if (variant == 0) fma_peak_kernel<double, 0><<<grid, kPeakBlock, 0, st>>>(iters, 0.999999, 1e-7, (double*)scratch);
        else fma_peak_kernel<double, 1><<<grid, kPeakBlock, 0, st>>>(iters, 0., 0., (double*)scratch);
    }
    return (int)cudaGetLastError();
}

}  // namespace pmp
