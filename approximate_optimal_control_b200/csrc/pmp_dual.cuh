// pmp_dual.cuh -- forward-mode dual numbers as a scalar type (value, tangent); see pmp_sens.cuh for what they are used for.
// Included by pmp_common.cuh so that the elementary-function overloads are visible where the functors are defined.
#pragma once

namespace pmp {

template <class R> struct Dual {
    R v, d;
    __host__ __device__ Dual() {}
    template <class T> __host__ __device__ Dual(T x) : v(R(x)), d(R(0)) {}
    __host__ __device__ Dual(R x, R dx) : v(x), d(dx) {}
};
template <class R> __host__ __device__ __forceinline__ Dual<R> operator+(Dual<R> a, Dual<R> b) { return Dual<R>(a.v + b.v, a.d + b.d); }
template <class R> __host__ __device__ __forceinline__ Dual<R> operator-(Dual<R> a, Dual<R> b) { return Dual<R>(a.v - b.v, a.d - b.d); }
template <class R> __host__ __device__ __forceinline__ Dual<R> operator*(Dual<R> a, Dual<R> b) { return Dual<R>(a.v * b.v, a.d * b.v + a.v * b.d); }
template <class R> __host__ __device__ __forceinline__ Dual<R> operator-(Dual<R> a) { return Dual<R>(-a.v, -a.d); }
template <class R> __host__ __device__ __forceinline__ Dual<R>& operator+=(Dual<R>& a, Dual<R> b) { a = a + b; return a; }
template <class R> __host__ __device__ __forceinline__ Dual<R>& operator-=(Dual<R>& a, Dual<R> b) { a = a - b; return a; }
template <class R> __host__ __device__ __forceinline__ Dual<R>& operator*=(Dual<R>& a, Dual<R> b) { a = a * b; return a; }
#define PMP_DUAL_MIXED(op)                                                                                                   \
    template <class R, class T> __host__ __device__ __forceinline__ auto operator op(Dual<R> a, T b)->decltype(R(b), Dual<R>()) { return a op Dual<R>(b); } \
    template <class R, class T> __host__ __device__ __forceinline__ auto operator op(T a, Dual<R> b)->decltype(R(a), Dual<R>()) { return Dual<R>(a) op b; }
PMP_DUAL_MIXED(+)
PMP_DUAL_MIXED(-)
PMP_DUAL_MIXED(*)
#undef PMP_DUAL_MIXED
// decisions look at values
#define PMP_DUAL_CMP(op)                                                                                                      \
    template <class R> __host__ __device__ __forceinline__ bool operator op(Dual<R> a, Dual<R> b) { return a.v op b.v; }       \
    template <class R, class T> __host__ __device__ __forceinline__ auto operator op(Dual<R> a, T b)->decltype(R(b), bool()) { return a.v op R(b); } \
    template <class R, class T> __host__ __device__ __forceinline__ auto operator op(T a, Dual<R> b)->decltype(R(a), bool()) { return R(a) op b.v; }
PMP_DUAL_CMP(<)
PMP_DUAL_CMP(>)
PMP_DUAL_CMP(<=)
PMP_DUAL_CMP(>=)
PMP_DUAL_CMP(==)
PMP_DUAL_CMP(!=)
#undef PMP_DUAL_CMP

// elementary functions of the generated functors (namespace fx, pmp_common.cuh) on dual numbers
namespace fx {
template <class R> __device__ __forceinline__ Dual<R> exp(Dual<R> x) { const R e = fx::exp(x.v); return Dual<R>(e, e * x.d); }
template <class R> __device__ __forceinline__ Dual<R> log(Dual<R> x) { return Dual<R>(fx::log(x.v), x.d / x.v); }
template <class R> __device__ __forceinline__ Dual<R> tanh(Dual<R> x) { const R t = fx::tanh(x.v); return Dual<R>(t, (R(1) - t * t) * x.d); }
template <class R> __device__ __forceinline__ Dual<R> atan2(Dual<R> a, Dual<R> b) {
    return Dual<R>(fx::atan2(a.v, b.v), (b.v * a.d - a.v * b.d) / (a.v * a.v + b.v * b.v));
}
template <class R> __device__ __forceinline__ Dual<R> pow(Dual<R> a, Dual<R> b) {
    const R p = fx::pow(a.v, b.v);
    return Dual<R>(p, p * (b.v * a.d / a.v + (b.d != R(0) ? fx::log(a.v) * b.d : R(0))));
}
}  // namespace fx

}  // namespace pmp
