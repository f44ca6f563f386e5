// pmp_dual_math.cuh -- Math<Dual<R>>: the integrator's arithmetic vocabulary (pmp_common.cuh) on dual numbers, so that the
// dynamics functors and the Runge-Kutta code instantiate on (value, tangent) pairs unchanged.  Decisions (min / max / clip,
// step control) look at values; the tangent follows the operand that was chosen -- what jax's autodiff does to
// jnp.clip / argmin-indexing in u_star_2d (pontryagin_utils.py:11-216).  Used by pmp_sens.cuh (sensitivities of a solve)
// and pmp_vxx_dual.cuh (value-Hessian branch of any functor).
#pragma once
#include "pmp_common.cuh"

namespace pmp {

template <class R> struct Math<Dual<R>> {
    using D = Dual<R>;
    using MR = Math<R>;
    static __device__ __forceinline__ D fma(D a, D b, D c) { return D(MR::fma(a.v, b.v, c.v), MR::fma(a.d, b.v, MR::fma(a.v, b.d, c.d))); }
    static __device__ __forceinline__ void sincos(D x, D* s, D* c) {
        R sv, cv;
        MR::sincos(x.v, &sv, &cv);
        *s = D(sv, cv * x.d);
        *c = D(cv, -sv * x.d);
    }
    static __device__ __forceinline__ D abs(D x) { return x.v < R(0) ? -x : x; }
    // fmax / fmin semantics on the value (a NaN operand is dropped); the tangent follows the operand that was chosen
    static __device__ __forceinline__ D max(D a, D b) { return (a.v != a.v) ? b : ((b.v != b.v) ? a : (a.v >= b.v ? a : b)); }
    static __device__ __forceinline__ D min(D a, D b) { return (a.v != a.v) ? b : ((b.v != b.v) ? a : (a.v <= b.v ? a : b)); }
    static __device__ __forceinline__ D sat(D x) { return (x.v > R(0) && x.v < R(1)) ? x : D(MR::sat(x.v)); }
    static __device__ __forceinline__ D copy(D x, D) { return x; }
    static __device__ __forceinline__ D sqrt(D x) { const R s = MR::sqrt(x.v); return D(s, x.d / (s + s)); }
    static __device__ __forceinline__ D div(D a, D b) { const R q = MR::div(a.v, b.v); return D(q, MR::div(a.d - q * b.d, b.v)); }
    static __device__ __forceinline__ D rcp(D x) { return div(D(R(1)), x); }
    static __device__ __forceinline__ D ctl_div(D a, D b) { return D(MR::ctl_div(a.v, b.v)); }            // step control: value only
    static __device__ __forceinline__ D pow_neg_inv(D x, D e) { return D(MR::pow_neg_inv(x.v, e.v)); }   // step control: value only
    static __device__ __forceinline__ D inf() { return D(MR::inf()); }
    static __device__ __forceinline__ D nan() { return D(MR::nan()); }
    static __device__ __forceinline__ D clip_tol() { return D(MR::clip_tol()); }
    struct V2 { D x, y; };
    static __device__ __forceinline__ V2 mk2(D a, D b) { V2 r; r.x = a; r.y = b; return r; }
    static __device__ __forceinline__ V2 fma2(D c, V2 a, V2 b) { return mk2(fma(c, a.x, b.x), fma(c, a.y, b.y)); }
    static __device__ __forceinline__ V2 fma2v(V2 c, V2 a, V2 b) { return mk2(fma(c.x, a.x, b.x), fma(c.y, a.y, b.y)); }
    static __device__ __forceinline__ V2 mul2(D c, V2 a) { return mk2(c * a.x, c * a.y); }
    static __device__ __forceinline__ V2 mul2v(V2 c, V2 a) { return mk2(c.x * a.x, c.y * a.y); }
    static __device__ __forceinline__ V2 add2(D c, V2 a) { return mk2(c + a.x, c + a.y); }
};

}  // namespace pmp
