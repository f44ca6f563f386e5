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

template <> struct Math<float> {
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    // x^(-1/order) for x > 0 through log2/exp2 (MUFU.LG2 / MUFU.EX2): relative error ~1e-6, used
    // only for the step-size factor.
    static __device__ __forceinline__ float pow_neg_inv(float x, float inv_order) {
        return exp2f(-inv_order * __log2f(x));
    }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fffffff); }
    static __device__ __forceinline__ float clip_tol() { return 1e-6f; }  // diffrax _clip_to_end
};

template <> struct Math<double> {
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double abs(double x) { return fabs(x); }
    static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double min(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
    static __device__ __forceinline__ double pow_neg_inv(double x, double inv_order) {
        return ::pow(x, -inv_order);
    }
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
    int max_steps, adaptive, force_dtmin, event, use_dtmin, use_dtmax;
};

template <class R> struct DensePtrs {
    R *ts, *x, *t, *v, *vx;
};

template <class R> struct SolveIO {
    const R* y_f;   // [NS][n]
    const R* k1;    // [ND][n]  field at the terminal state (prep kernel)
    R* y_end;       // [NS][n]
    DensePtrs<R> dense;
    int32_t *n_accepted, *n_steps, *result;
    unsigned long long* counter;  // next unclaimed trajectory index
    long long n;
};

}  // namespace pmp
