// pmp_interp.cuh -- dense output: diffrax `Solution.evaluate(t)` for SaveAt(dense=True).
//
// The reference keeps, for every accepted step, the stage derivatives k1..k7 ("dense info",
// pontryagin_utils.py:504 SaveAt(dense=True)) and evaluates the solver's interpolant on demand
// (experiment_orbits.py:183-184 jax.vmap(sol.evaluate)(vs); misc.py:40-95 bisection on sol.evaluate).
// Storing 7 x 13 reals per step would multiply the dense output by 7; instead the stage derivatives
// of the ONE step that contains a query are recomputed from the saved node (deterministic: same
// tableau, same field, same step) and the interpolant is evaluated on them:
//     y(t0 + theta h) = y0 + h sum_i b_i(theta) k_i
// Tsit5: the free 4th-order interpolant of Tsitouras (2011); Dopri5: Dormand-Prince/Shampine dense
// output; RK4 (not a diffrax solver): cubic Hermite on (y0, f0, y1, f1).
// mode 1 additionally solves v(theta) = level for theta in [0,1] by bisection (the value-level
// root finding of misc.py:47-91 on one step).
#pragma once
#include "pmp_solve.cuh"

namespace pmp {

template <int METHOD> struct Interp;

template <> struct Interp<PMP_METHOD_TSIT5> {
    template <class R> static __host__ __device__ void weights(R th, R* w) {
        const R t2 = th * th;
        w[0] = R(-1.0530884977290216) * th * (th - R(1.3299890189751412)) *
               (t2 - R(1.4364028541716351) * th + R(0.7139816917074209));
        w[1] = R(0.1017) * t2 * (t2 - R(2.1966568338249754) * th + R(1.2949852507374631));
        w[2] = R(2.490627285651252793) * t2 * (t2 - R(2.38535645472061657) * th + R(1.57803468208092486));
        w[3] = R(-16.54810288924490272) * (th - R(1.21712927295533244)) * (th - R(0.61620406037800089)) * t2;
        w[4] = R(47.37952196281928122) * (th - R(1.203071208372362603)) * (th - R(0.658047292653547382)) * t2;
        w[5] = R(-34.87065786149660974) * (th - R(1.2)) * (th - R(0.666666666666666667)) * t2;
        w[6] = R(2.5) * (th - R(1.0)) * (th - R(0.6)) * t2;
    }
};

template <> struct Interp<PMP_METHOD_DOPRI5> {
    // Hairer, Norsett, Wanner, "Solving ODEs I", dense output of DOPRI5 (contd5)
    template <class R> static __host__ __device__ void weights(R th, R* w) {
        const R t2 = th * th, om = th - R(1), a = t2 * (R(3) - R(2) * th), c = t2 * om * om;
        w[0] = a * R(35.0 / 384) + th * om * om - c * R(5) * (R(2558722523.0) - R(31403016.0) * th) / R(11282082432.0);
        w[1] = R(0);
        w[2] = a * R(500.0 / 1113) + c * R(100) * (R(882725551.0) - R(15701508.0) * th) / R(32700410799.0);
        w[3] = a * R(125.0 / 192) - c * R(25) * (R(443332067.0) - R(31403016.0) * th) / R(1880347072.0);
        w[4] = a * R(-2187.0 / 6784) + c * R(32805) * (R(23143187.0) - R(3489224.0) * th) / R(199316789632.0);
        w[5] = a * R(11.0 / 84) - c * R(55) * (R(29972135.0) - R(7076736.0) * th) / R(822651844.0);
        w[6] = t2 * om + c * R(10) * (R(7414447.0) - R(829305.0) * th) / R(29380423.0);
    }
};

// query q: node y0[:, q] (public [NS][nq] rows x, t, v, vx), step end t1[q] (time, or value in value
// mode), target[q] = query time/value (mode 0) or value level (mode 1).  Output y_out [NS][nq].
template <class Sys, class R, int METHOD, bool VALUE, bool ULIT>
__global__ void __launch_bounds__(128) interp_kernel(const typename Sys::Consts sc, const R* __restrict__ y0_,
                                                     const R* __restrict__ t1_, const R* __restrict__ target_,
                                                     R* __restrict__ y_out, long long nq, int mode) {
    using M = Math<R>;
    using TB = Tab<METHOD>;
    constexpr int NX = Sys::NX, ND = 2 * NX + 1, S = TB::S, AUX = 2 * NX;
    constexpr int UM = ULIT ? 0 : 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    R y[ND], clock0, aux_fixed;
    load_state<R, NX, VALUE>(y0_, nq, i, y, clock0, aux_fixed);
    if (!VALUE) { clock0 = aux_fixed; }           // time mode: the clock is the t row of the node
    const R h = t1_[i] - clock0;                  // signed step
    R k[S + 1][ND];
    field<Sys, R, VALUE, UM>(sc, y, k[0]);
    R y1[ND];
#pragma unroll
    for (int s = 1; s < S; s++) {
        R ys[ND];
#pragma unroll
        for (int q = 0; q < ND; q++) {
            R acc = y[q];
#pragma unroll
            for (int j = 0; j < s; j++)
                if (TB::a(s, j) != 0.0) acc = M::fma(h * R(TB::a(s, j)), k[j][q], acc);
            ys[q] = acc;
        }
        field<Sys, R, VALUE, UM>(sc, ys, k[s]);
        if (TB::FSAL && s == S - 1) {
#pragma unroll
            for (int q = 0; q < ND; q++) y1[q] = ys[q];
        }
    }
    if (!TB::FSAL) {  // RK4: y1 and f(y1) for the Hermite interpolant
#pragma unroll
        for (int q = 0; q < ND; q++) {
            R acc = y[q];
#pragma unroll
            for (int j = 0; j < S; j++) acc = M::fma(h * R(TB::b(j)), k[j][q], acc);
            y1[q] = acc;
        }
        field<Sys, R, VALUE, UM>(sc, y1, k[S]);
    }
    auto eval = [&](R th, int q) -> R {
        if (TB::FSAL) {
            R w[7];
            Interp<METHOD == PMP_METHOD_RK4 ? PMP_METHOD_TSIT5 : METHOD>::weights(th, w);
            R acc = R(0);
#pragma unroll
            for (int j = 0; j < S; j++) acc = M::fma(w[j], k[j][q], acc);
            return M::fma(h, acc, y[q]);
        } else {  // cubic Hermite
            const R d = y1[q] - y[q];
            return (R(1) - th) * y[q] + th * y1[q] +
                   th * (th - R(1)) * ((R(1) - R(2) * th) * d + (th - R(1)) * h * k[0][q] + th * h * k[S][q]);
        }
    };
    R th;
    if (mode == 0) {
        th = M::div(target_[i] - clock0, h);
    } else {
        // time mode: v is the aux component; find theta in [0,1] with v(theta) = level (v monotone on a step)
        const R level = target_[i];
        R lo = R(0), hi = R(1);
        const R f0 = y[AUX] - level, f1 = eval(R(1), AUX) - level;
        const bool ok = (f0 <= R(0) && f1 >= R(0)) || (f0 >= R(0) && f1 <= R(0));
        for (int it = 0; it < (sizeof(R) == 4 ? 26 : 54); it++) {
            const R mid = R(0.5) * (lo + hi);
            const R fm = eval(mid, AUX) - level;
            const bool same = (fm >= R(0)) == (f0 >= R(0));
            lo = same ? mid : lo;
            hi = same ? hi : mid;
        }
        th = ok ? R(0.5) * (lo + hi) : M::nan();
    }
#pragma unroll
    for (int q = 0; q < NX; q++) {
        y_out[q * nq + i] = eval(th, q);
        y_out[(NX + 2 + q) * nq + i] = eval(th, NX + q);
    }
    const R clock = M::fma(th, h, clock0);
    if (VALUE) { y_out[NX * nq + i] = eval(th, AUX); y_out[(NX + 1) * nq + i] = clock; }
    else { y_out[NX * nq + i] = clock; y_out[(NX + 1) * nq + i] = eval(th, AUX); }
}

template <template <class> class SysT>
int launch_interp(const pmp_problem_t& p, const pmp_opts_t& o, long long nq, const void* y0, const void* t1,
                  const void* target, void* y_out, int mode, cudaStream_t st) {
    if (nq == 0) return 0;
    const unsigned grid = (unsigned)((nq + 127) / 128);
    const bool val = o.indep == PMP_INDEP_VALUE;
    const bool ulit = (o.flags & PMP_FLAG_USTAR_LITERAL) != 0 && SysT<float>::NU == 2;
#define PMP_IGO(R, METHOD, VALUE, ULIT)                                                                            \
    do {                                                                                                           \
        interp_kernel<SysT<R>, R, METHOD, VALUE, ULIT><<<grid, 128, 0, st>>>(SysT<R>::make(p), (const R*)y0, (const R*)t1, \
                                                                              (const R*)target, (R*)y_out, nq, mode); \
        return (int)cudaGetLastError();                                                                            \
    } while (0)
#define PMP_IBY(R, METHOD)                                                   \
    do {                                                                     \
        if (val) PMP_IGO(R, METHOD, true, false);                            \
        if (SysT<float>::NU == 2 && ulit) PMP_IGO(R, METHOD, false, (SysT<float>::NU == 2)); \
        PMP_IGO(R, METHOD, false, false);                                    \
    } while (0)
#define PMP_IMETH(R)                                                         \
    do {                                                                     \
        if (o.method == PMP_METHOD_RK4) PMP_IBY(R, PMP_METHOD_RK4);          \
        if (o.method == PMP_METHOD_TSIT5) PMP_IBY(R, PMP_METHOD_TSIT5);      \
        if (o.method == PMP_METHOD_DOPRI5) PMP_IBY(R, PMP_METHOD_DOPRI5);    \
    } while (0)
    if (o.dtype == PMP_F32) PMP_IMETH(float);
    if (o.dtype == PMP_F64) PMP_IMETH(double);
#undef PMP_IMETH
#undef PMP_IBY
#undef PMP_IGO
    return -1000;
}

}  // namespace pmp
