// pmp_common.cuh -- shared device helpers of libb200pmp (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200pmp.h"

namespace pmp {

// ---- scalar maths, one overload set per precision ------------------------------------------------
// fp32 uses FFMA-friendly forms and MUFU-based reciprocal / pow where the value only steers the
// step size (never the state); the state path uses IEEE ops.
template <class R> struct Math;

// Branch-free sin & cos for fp32, ~21 instructions for the pair (CUDA's sincosf is ~2x that plus a
// Payne-Hanek slow path behind a branch).  Cody-Waite reduction by pi/2 with the quotient obtained
// by the 1.5*2^23 magic-number trick (valid for |x| < ~1e5 rad; the pitch angle Phi stays within a
// few turns), then the Cephes minimax polynomials on [-pi/4, pi/4] (~1 ulp, exact-in-the-limit
// relative accuracy near 0, which the MUFU.SIN/COS intrinsics do not give).
// (A reduction by pi -- one sign flip, no swap, one more term per polynomial: 5 fewer ALU-pipe instructions -- was
//  measured in round 2 and made no difference: 1.190 vs 1.187 ms per 2^20 rollouts.)
__device__ __forceinline__ void pmp_sincosf(float x, float* sn_out, float* cs_out) {
    const float magic = 12582912.0f;  // 1.5 * 2^23
    const float kf_biased = fmaf(x, 0.636619772367581f, magic);
    const int k = __float_as_int(kf_biased);  // low mantissa bits hold round(x * 2/pi)
    const float kf = kf_biased - magic;
    float r = fmaf(kf, -1.57079625129699707031f, x);  // pi/2 = hi + mid + lo (fma: products are exact)
    r = fmaf(kf, -7.54978941586159635335e-08f, r);
    r = fmaf(kf, -5.39030285815811905290e-15f, r);
    const float r2 = r * r;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    const float sn = fmaf(r2 * r, sp, r);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    cp = fmaf(cp, r2, -0.5f);
    const float cs = fmaf(cp, r2, 1.0f);
    const bool swap = (k & 1) != 0;
    const float s0 = swap ? cs : sn, c0 = swap ? sn : cs;
    // quadrant signs: sin negated for k mod 4 in {2,3}, cos for k mod 4 in {1,2}
    *sn_out = __int_as_float(__float_as_int(s0) ^ ((k & 2) << 30));
    *cs_out = __int_as_float(__float_as_int(c0) ^ (((k + 1) & 2) << 30));
}

template <> struct Math<float> {
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { pmp_sincosf(x, s, c); }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float sat(float x) { return __saturatef(x); }  // clip to [0, 1]; NaN -> 0
    // register copy through the FMA pipe: x * one with `one` = 1.0 read from the parameter bank (exact for every x).  A MOV
    // runs on the half-rate ALU pipe and costs two issue cycles, an FMUL one; the accept path copies 2 (2nx+1) registers.
    static __device__ __forceinline__ float copy(float x, float one) { return x * one; }
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    // error / scale in the step-size controller: MUFU.RCP + FMUL (2 ulp).  Only the accept decision
    // at |scaled error - 1| < 1e-7 can differ from an IEEE division; the state never sees it.
    // (MUFU.RCP + FMUL: __fdividef adds range checks and rescaling around them, 5 instructions instead of 2; the divisor is
    //  the error scale atol + rtol |y| >= atol, never denormal or near overflow)
    static __device__ __forceinline__ float ctl_div(float a, float b) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
        return a * r;
    }
    // x^(-1/order) for x > 0 through log2/exp2 (MUFU.LG2 / MUFU.EX2): relative error ~1e-6, used
    // only for the step-size factor.
    static __device__ __forceinline__ float pow_neg_inv(float x, float inv_order) {
        return exp2f(-inv_order * __log2f(x));
    }
    // Packed pairs: Blackwell's FFMA2 / FMUL2 (fma.rn.f32x2) do two IEEE FMAs per lane and ISSUE SLOT, each
    // component rounded exactly like fmaf; the scalar coefficient is a broadcast operand (R.F32), so it costs no
    // extra register.  The integrator is issue-bound, not FMA-pipe-bound: its stage combinations and error
    // estimate (one coefficient times a vector of 2nx+1 components) run on these.
    typedef float2 V2;
    static __device__ __forceinline__ V2 mk2(float a, float b) { return make_float2(a, b); }
    static __device__ __forceinline__ V2 fma2(float c, V2 a, V2 b) { return __ffma2_rn(make_float2(c, c), a, b); }
    static __device__ __forceinline__ V2 fma2v(V2 c, V2 a, V2 b) { return __ffma2_rn(c, a, b); }
    static __device__ __forceinline__ V2 mul2(float c, V2 a) { return __fmul2_rn(make_float2(c, c), a); }
    static __device__ __forceinline__ V2 mul2v(V2 c, V2 a) { return __fmul2_rn(c, a); }
    static __device__ __forceinline__ V2 add2(float c, V2 a) { return __fadd2_rn(make_float2(c, c), a); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fffffff); }
    static __device__ __forceinline__ float clip_tol() { return 1e-6f; }  // diffrax _clip_to_end
};

// Branch-free sin & cos for fp64 (same scheme as pmp_sincosf): Cody-Waite reduction by pi/2 with the
// 1.5*2^52 magic-number quotient (|x| < ~1e9 rad), fdlibm's __kernel_sin / __kernel_cos minimax
// polynomials on [-pi/4, pi/4] (< 1 ulp).  CUDA's sincos(double) costs ~2.5x the instructions plus a
// Payne-Hanek slow path behind a branch, 6 times per step attempt.
__device__ __forceinline__ void pmp_sincos(double x, double* sn_out, double* cs_out) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    const double kf_biased = ::fma(x, 0.63661977236758138, magic);
    const int k = __double2loint(kf_biased);
    const double kf = kf_biased - magic;
    double r = ::fma(kf, -1.5707963267948966, x);     // pi/2 = hi + mid + lo; the fma products are exact
    r = ::fma(kf, -6.123233995736766e-17, r);
    r = ::fma(kf, 1.4973849048591698e-33, r);
    const double z = r * r;
    double sp = ::fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    sp = ::fma(sp, z, 2.75573137070700676789e-06);
    sp = ::fma(sp, z, -1.98412698298579493134e-04);
    sp = ::fma(sp, z, 8.33333333332248946124e-03);
    sp = ::fma(sp, z, -1.66666666666666324348e-01);
    const double sn = ::fma(z * r, sp, r);
    double cp = ::fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    cp = ::fma(cp, z, -2.75573143513906633035e-07);
    cp = ::fma(cp, z, 2.48015872894767294178e-05);
    cp = ::fma(cp, z, -1.38888888888741095749e-03);
    cp = ::fma(cp, z, 4.16666666666666019037e-02);
    const double cs = ::fma(z * z, cp, ::fma(z, -0.5, 1.0));
    const bool swap = (k & 1) != 0;
    const double s0 = swap ? cs : sn, c0 = swap ? sn : cs;
    *sn_out = __hiloint2double(__double2hiint(s0) ^ ((k & 2) << 30), __double2loint(s0));
    *cs_out = __hiloint2double(__double2hiint(c0) ^ (((k + 1) & 2) << 30), __double2loint(c0));
}

template <> struct Math<double> {
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { pmp_sincos(x, s, c); }
    static __device__ __forceinline__ double abs(double x) { return fabs(x); }
    static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double min(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double sat(double x) { return fmin(fmax(x, 0.0), 1.0); }
    static __device__ __forceinline__ double copy(double x, double) { return x; }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
    static __device__ __forceinline__ double ctl_div(double a, double b) { return a / b; }
    static __device__ __forceinline__ double pow_neg_inv(double x, double inv_order) {
        return ::pow(x, -inv_order);
    }
    // pairs are two scalar operations in fp64 (no packed DFMA exists); same interface as Math<float>
    typedef double2 V2;
    static __device__ __forceinline__ V2 mk2(double a, double b) { return make_double2(a, b); }
    static __device__ __forceinline__ V2 fma2(double c, V2 a, V2 b) { return make_double2(::fma(c, a.x, b.x), ::fma(c, a.y, b.y)); }
    static __device__ __forceinline__ V2 fma2v(V2 c, V2 a, V2 b) { return make_double2(::fma(c.x, a.x, b.x), ::fma(c.y, a.y, b.y)); }
    static __device__ __forceinline__ V2 mul2(double c, V2 a) { return make_double2(c * a.x, c * a.y); }
    static __device__ __forceinline__ V2 mul2v(V2 c, V2 a) { return make_double2(c.x * a.x, c.y * a.y); }
    static __device__ __forceinline__ V2 add2(double c, V2 a) { return make_double2(c + a.x, c + a.y); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7fffffffffffffffLL); }
    static __device__ __forceinline__ double clip_tol() { return 1e-10; }
};

// jnp.maximum / jnp.minimum / jnp.clip propagate NaN; fmaxf/fminf do not.
template <class R> __device__ __forceinline__ R nan_max(R a, R b) {
    R m = Math<R>::max(a, b);
    return (a != a) ? a : ((b != b) ? b : m);
}
template <class R> __device__ __forceinline__ R nan_min(R a, R b) {
    R m = Math<R>::min(a, b);
    return (a != a) ? a : ((b != b) ? b : m);
}
// clip(x, lo, hi) with finite lo/hi: NaN x stays NaN
template <class R> __device__ __forceinline__ R clipf(R x, R lo, R hi) {
    R c = Math<R>::min(Math<R>::max(x, lo), hi);
    return (x != x) ? x : c;
}

// ---- solver options in the kernel's precision (lives in the kernel parameter / constant bank) ----
template <class R> struct SolveOpts {
    R T0, T1;       // internal (direction-normalised) start / end of the independent variable
    R dt0;          // initial step, > 0
    R dir;          // +1 / -1
    R rtol, atol, dtmin, dtmax, safety, factormin, factormax, inv_order;
    R event_level;
    R one;          // 1.0 (see Math<R>::copy)
    int max_steps, adaptive, force_dtmin, event, use_dtmin, use_dtmax;
    int refill_min;  // lanes of a warp that must be waiting before the retire + refill detour is taken
};

// Butcher tableau in the kernel's precision, used by the fp64 instantiations only: 64-bit immediates do
// not exist, so a compile-time double coefficient costs two register moves per use, while a
// parameter-bank entry is a c[][] operand of the DMUL that consumes it.  fp32 keeps the compile-time
// immediates (measured 4 % faster than the parameter bank), so its record is empty.
template <class R> struct TabParams {
    R ta[7][7], tb[7], te[7];
    template <class TB> static TabParams make() {
        TabParams t;
        for (int i = 0; i < 7; i++) {
            t.tb[i] = i < TB::S ? R(TB::b(i)) : R(0);
            t.te[i] = i < TB::S ? R(TB::berr(i)) : R(0);
            for (int j = 0; j < 7; j++) t.ta[i][j] = (i < TB::S && j < TB::S) ? R(TB::a(i, j)) : R(0);
        }
        return t;
    }
    template <class TB> __device__ __forceinline__ R a(int i, int j) const { return ta[i][j]; }
    template <class TB> __device__ __forceinline__ R b(int j) const { return tb[j]; }
    template <class TB> __device__ __forceinline__ R e(int j) const { return te[j]; }
};
template <> struct TabParams<float> {
    template <class TB> static TabParams make() { return TabParams(); }
    template <class TB> __device__ __forceinline__ float a(int i, int j) const { return float(TB::a(i, j)); }
    template <class TB> __device__ __forceinline__ float b(int j) const { return float(TB::b(j)); }
    template <class TB> __device__ __forceinline__ float e(int j) const { return float(TB::berr(j)); }
};

template <class R> struct DensePtrs {
    R* p[6];
};
enum { kDenseTs = 0, kDenseT = 1, kDenseV = 2, kDenseX = 3, kDenseVx = 4, kDenseVxx = 5 };

constexpr int kMaxPeers = 7;  // one node: up to 8 GPUs
// store through a multicast (multimem) address: the NVSwitch replicates the write into every GPU of the multicast group
__device__ __forceinline__ void mc_store(float* p, float v) { asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void mc_store(double* p, double v) { asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
template <class R> struct SolveIO {
    const R* y_f;   // [NS][n]
    const R* k1;    // [ND][n]  field at the terminal state (prep kernel)
    R* y_end;       // [NS][n]
    DensePtrs<R> dense;
    int32_t *n_accepted, *n_steps, *result;
    unsigned long long* counter;  // next unclaimed trajectory index
    long long n;
    // PUSH instantiation (multi-GPU gather fused into the integrator, pmp_solve_batch_push_cuda): this rank's [NS][n] block in
    // every peer's gather buffer (peer memory over NVLink)
    R* peer[kMaxPeers];
    int n_peers;
    R* mc;  // non-NULL: the same block through the NVSwitch MULTICAST mapping of the gather buffer (one multimem.st reaches every GPU)
};

// ---- elementary functions the generated functors (codegen.py) may call, one overload set per scalar type (pmp_sens.cuh adds
// the dual-number overloads, so that a generated system differentiates like a hand-written one)
namespace fx {
__device__ __forceinline__ float exp(float x) { return ::expf(x); }
__device__ __forceinline__ double exp(double x) { return ::exp(x); }
__device__ __forceinline__ float log(float x) { return ::logf(x); }
__device__ __forceinline__ double log(double x) { return ::log(x); }
__device__ __forceinline__ float tanh(float x) { return ::tanhf(x); }
__device__ __forceinline__ double tanh(double x) { return ::tanh(x); }
__device__ __forceinline__ float atan2(float a, float b) { return ::atan2f(a, b); }
__device__ __forceinline__ double atan2(double a, double b) { return ::atan2(a, b); }
__device__ __forceinline__ float pow(float a, float b) { return ::powf(a, b); }
__device__ __forceinline__ double pow(double a, double b) { return ::pow(a, b); }
}  // namespace fx

}  // namespace pmp
#include "pmp_dual.cuh"
namespace pmp {

// ---- pmp_solve_host's persistent launch (csrc/pmp_host.cu) -------------------------------------------------------
// ONE launch of the integrator serves the whole host batch while the copy engines are still delivering it: the
// kernel reads the reference's state-dict leaves (array of structs, as they lie in host memory) straight from the
// whole-batch device arrays the host->device stream fills, may only start trajectories below *ready (an 8-byte
// copy the same stream issues behind every piece), evaluates k1 = F(y_f) itself, writes the endpoints in the
// state-dict layout and counts retired trajectories per piece, so that the device->host stream can wait for "piece
// complete" with a stream memory operation.  No other kernel runs meanwhile: nothing the
// integrator waits for needs an SM.
template <class R> struct HostQueue {
    const R *in_x, *in_vx, *in_t, *in_v;   // [n][nx], [n][nx], [n] or NULL (then t_fill), [n]
    const R* term;                         // non-NULL: [nx*nx + nx] = P row-major, x_eq; then in_vx / in_v are unused and the
                                           // terminal costate and value are formed here, vx_f = P (x_f - x_eq),
                                           // v_f = 1/2 (x_f - x_eq)' vx_f, as levelsets.py:585-586 does inside its vmap
    R *out_x, *out_vx, *out_t, *out_v;     // endpoints, same layouts; any may be NULL
    uint32_t* out_pk;                      // accepted | attempts << 14 | result << 28, or NULL
    int32_t *out_na, *out_ns, *out_res;    // (used when the counters do not fit the packed word)
    const unsigned long long* ready;       // trajectories [0, *ready) have landed
    unsigned int* done;                    // done[piece]: retired trajectories of piece [bounds[piece], bounds[piece+1])
    unsigned int* abort;                   // != 0: the watchdog gave up waiting for *ready (host reports an error)
    int n_pieces;
    long long bounds[33];
    R t_fill;
};

}  // namespace pmp
