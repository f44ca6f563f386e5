// pmp_sens.cuh -- directional derivative of a whole solve with respect to its terminal state (forward mode).
//
// experiment_orbits.py:166-197 differentiates the value-parameterised characteristic solve with jax:
//     dsol_final = jax.jacobian(get_final_x)(theta),  get_final_x = theta -> y0(theta) -> pontryagin_solver(y0, v0, v1) -> x(v)
// i.e. a jvp through diffrax.diffeqsolve: every arithmetic operation of the Runge-Kutta steps carries a tangent, every
// decision (step-size control, clipping of u*, argmin over the candidates of u_star_2d) is taken on primal values.
// Here the same thing is done with dual numbers: Dual<R> = (value, tangent) is a scalar type like float / double, so the
// plug-in dynamics functors (hand-written or generated), the stage combinations and the error estimate of the
// integrator are INSTANTIATED on it unchanged -- no hand-derived variational equations, and any system works.  The step
// controller sees values only, so the primal part of the result is the solve the integrator itself produces.
// One thread per trajectory (this is not the throughput path: the caller differentiates a few thousand characteristics).
#pragma once
#include "pmp_solve.cuh"      // (Dual<R> itself lives in pmp_dual.cuh, included by pmp_common.cuh)
#include "pmp_dual_math.cuh"  // Math<Dual<R>>

namespace pmp {

// PairMap of the underlying system (the pairing is a property of the dynamics, not of the scalar)
template <class R> struct PairMap<Flatquad<Dual<R>>> : PairMap<Flatquad<R>> {};

// y_f, dy_f [NS][n] (public rows x, t, v, vx) -> y_end, dy_end [NS][n]; step control as in solve_kernel,
// decided on values.  The independent variable (t, or v in value mode) and its end points carry no tangent.
template <template <class> class SysT, class R, int METHOD, bool VALUE>
__global__ void __launch_bounds__(64) sens_kernel(const typename SysT<Dual<R>>::Consts sc, const SolveOpts<R> o, const R* __restrict__ y_f,
                                                  const R* __restrict__ dy_f, R* __restrict__ y_end, R* __restrict__ dy_end,
                                                  int32_t* __restrict__ n_accepted, int32_t* __restrict__ n_steps,
                                                  int32_t* __restrict__ result_out, long long n) {
    using D = Dual<R>;
    using Sys = SysT<D>;
    using M = Math<D>;
    using MR = Math<R>;
    using TB = Tab<METHOD>;
    constexpr int NX = Sys::NX, ND = 2 * NX + 1, S = TB::S, AUX = 2 * NX;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D y[ND], k[S][ND];
    for (int q = 0; q < NX; q++) {
        y[q] = D(y_f[q * n + i], dy_f[q * n + i]);
        y[NX + q] = D(y_f[(NX + 2 + q) * n + i], dy_f[(NX + 2 + q) * n + i]);
    }
    const D t_in(y_f[NX * n + i], dy_f[NX * n + i]), v_in(y_f[(NX + 1) * n + i], dy_f[(NX + 1) * n + i]);
    y[AUX] = VALUE ? t_in : v_in;
    const R T0 = VALUE ? v_in.v * o.dir : o.T0;
    const R clock0 = T0 * o.dir;
    R tprev = T0, tnext = MR::min(T0 + o.dt0, o.T1);
    int nsteps = 0, nacc = 0, result = 0;
    bool at_dtmin = false;
    bool bad = false;
    for (int q = 0; q < ND; q++) bad = bad || (y[q].v != y[q].v);
    if (bad) result = PMP_RESULT_NAN;
    if (TB::FSAL && !bad) field<Sys, D, VALUE, 2>(sc, y, k[0]);
    const R inv_nleaf = VALUE ? R(1.0 / (2 * NX + 1)) : R(1.0 / (2 * NX + 2));
    while (result == 0 && tprev < o.T1) {
        const R dt = tnext - tprev;
        const D h(dt * o.dir);
        D y1[ND];
        if (!TB::FSAL) field<Sys, D, VALUE, 2>(sc, y, k[0]);
        for (int s = 1; s < S; s++) {
            D ys[ND], ha[S];
            const bool last_fsal = TB::FSAL && s == S - 1;
            for (int j = 0; j < s; j++) ha[j] = h * D(R(TB::a(s, j)));
            stage_state<Sys, D, TB, ND>(s, last_fsal, ha, y, k, ys);
            field<Sys, D, VALUE, 2>(sc, ys, k[s]);
            if (last_fsal)
                for (int q = 0; q < ND; q++) y1[q] = ys[q];
        }
        if (!TB::FSAL) {
            D hb[S];
            for (int j = 0; j < S; j++) hb[j] = h * D(R(TB::b(j)));
            for (int q = 0; q < ND; q++) {
                D acc = y[q];
                for (int j = 0; j < S; j++)
                    if (TB::b(j) != 0.0) acc = M::fma(hb[j], k[j][q], acc);
                y1[q] = acc;
            }
        }
        bool keep = true, step_nonfinite = false;
        R next_t0, next_t1;
        if (TB::EMBEDDED && o.adaptive) {
            D he[S];
            for (int j = 0; j < S; j++) he[j] = h * D(R(TB::berr(j)));
            R sumsq = error_sumsq<Sys, D, TB, ND>(he, y, y1, k, D(o.rtol), D(o.atol)).v;
            sumsq = (sumsq != sumsq) ? MR::inf() : sumsq;
            const R msq = sumsq * inv_nleaf;
            step_nonfinite = !(msq < MR::inf());
            keep = (msq < R(1)) || at_dtmin;
            R factor = o.safety * MR::pow_neg_inv(msq, R(0.5) * o.inv_order);
            factor = MR::min(MR::max(factor, keep ? R(1) : o.factormin), o.factormax);
            R dtn = dt * factor;
            if (o.use_dtmax) dtn = MR::min(dtn, o.dtmax);
            if (o.use_dtmin) {
                if (!o.force_dtmin && dtn < o.dtmin) result = PMP_RESULT_DT_MIN;
                at_dtmin = dtn <= o.dtmin;
                dtn = MR::max(dtn, o.dtmin);
            } else {
                at_dtmin = false;
            }
            next_t0 = keep ? tnext : tprev;
            next_t1 = next_t0 + dtn;
        } else {
            next_t0 = tnext;
            next_t1 = tnext + dt;
        }
        tprev = MR::min(next_t0, o.T1);
        if (next_t1 > o.T1 - MR::clip_tol()) next_t1 = keep ? o.T1 : MR::fma(R(0.5), o.T1 - tprev, tprev);
        tnext = next_t1;
        nsteps += 1;
        bool nan_state = false;
        if (keep) {
            nacc += 1;
            bool b2 = step_nonfinite;
            for (int q = 0; q < ND; q++) {
                y[q] = y1[q];
                if (!TB::EMBEDDED) b2 = b2 || (y1[q].v != y1[q].v);
                if (TB::FSAL) k[0][q] = k[S - 1][q];
            }
            nan_state = b2;
        }
        if (!VALUE && result == 0 && o.event == PMP_EVENT_VALUE_GE && y[AUX].v >= o.event_level) result = PMP_RESULT_EVENT;
        if (result == 0 && nan_state) result = PMP_RESULT_NAN;
        if (result == 0 && tprev < o.T1 && nsteps >= o.max_steps) result = PMP_RESULT_MAX_STEPS;
    }
    const R clock = tprev * o.dir;
    for (int q = 0; q < NX; q++) {
        y_end[q * n + i] = y[q].v; dy_end[q * n + i] = y[q].d;
        y_end[(NX + 2 + q) * n + i] = y[NX + q].v; dy_end[(NX + 2 + q) * n + i] = y[NX + q].d;
    }
    if (VALUE) {
        y_end[NX * n + i] = y[AUX].v; dy_end[NX * n + i] = y[AUX].d;
        y_end[(NX + 1) * n + i] = clock; dy_end[(NX + 1) * n + i] = R(0);
    } else {
        y_end[NX * n + i] = t_in.v + (clock - clock0); dy_end[NX * n + i] = t_in.d;
        y_end[(NX + 1) * n + i] = y[AUX].v; dy_end[(NX + 1) * n + i] = y[AUX].d;
    }
    if (n_accepted) n_accepted[i] = nacc;
    if (n_steps) n_steps[i] = nsteps;
    if (result_out) result_out[i] = result;
}

template <template <class> class SysT>
int launch_sens(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, const void* dy_f, void* y_end, void* dy_end,
                int32_t* a, int32_t* s, int32_t* r, cudaStream_t st) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 63) / 64);
    const bool val = o.indep == PMP_INDEP_VALUE;
#define PMP_SENS(R, METHOD)                                                                                                          \
    do {                                                                                                                             \
        const SolveOpts<R> so = make_solve_opts<R, Tab<METHOD>>(o, Tab<METHOD>::ORDER, SysT<R>::NX);                                   \
        if (val) sens_kernel<SysT, R, METHOD, true><<<grid, 64, 0, st>>>(SysT<Dual<R>>::make(p), so, (const R*)y_f, (const R*)dy_f, (R*)y_end, (R*)dy_end, a, s, r, n); \
        else sens_kernel<SysT, R, METHOD, false><<<grid, 64, 0, st>>>(SysT<Dual<R>>::make(p), so, (const R*)y_f, (const R*)dy_f, (R*)y_end, (R*)dy_end, a, s, r, n);   \
        return (int)cudaGetLastError();                                                                                              \
    } while (0)
#define PMP_SENS_M(R)                                                \
    do {                                                             \
        if (o.method == PMP_METHOD_RK4) PMP_SENS(R, PMP_METHOD_RK4);     \
        if (o.method == PMP_METHOD_TSIT5) PMP_SENS(R, PMP_METHOD_TSIT5); \
        if (o.method == PMP_METHOD_DOPRI5) PMP_SENS(R, PMP_METHOD_DOPRI5); \
    } while (0)
    if (o.dtype == PMP_F32) PMP_SENS_M(float);
    if (o.dtype == PMP_F64) PMP_SENS_M(double);
#undef PMP_SENS_M
#undef PMP_SENS
    return -1000;
}

}  // namespace pmp
