// pmp_forward.cuh -- batched FORWARD closed-loop simulation under a costate feedback law (SURVEY.md 8(f) N3,
// first slice).
//
// The reference simulates the closed loop x' = f(x, u*(x, lambda(x))) forward in time in two places:
//   * levelsets.py:818-836  forward_sim_lqr: lambda(x) = P_lqr x, Tsit5 + PIDController(rtol, atol, dtmin=.05),
//     t: 0 -> 10, SaveAt(steps, t0), max_steps, throw=False  (the NN variants :838-933 differ in lambda(x) only);
//   * eval_utils.py:48-99   closed_loop_eval_general: one extra state records the control cost,
//     z = (x, cost), z' = (f(x,u*), l(x,u*)), Tsit5 with a constant step dt, max_steps = T/dt.
// This kernel covers both for a LINEAR costate law lambda(x) = P (x - x_eq) (the LQR value function): the state
// is z = (x, cost); the error norm runs over the x leaf only (forward_sim_lqr has no cost state, and the
// constant-step evaluation has no error control), so either call site sees its own step sequence.  The neural
// costate law (gradient of an MLP ensemble) is a different kernel (a batched small-GEMM) and is not built.
// Optional terminating event: 1/2 (x-x_eq)'P(x-x_eq) <= level, the LQR analogue of the reference's "inside the
// value level set" stop (levelsets.py:897-915).
//
// One thread per trajectory, state and stage derivatives in registers, the same diffrax step-control
// semantics as pmp_solve.cuh.  Forward simulations run by the hundred, not by the million
// (levelsets.py:1201-1215): no work queue, no staging of the dense output.
#pragma once
#include "pmp_solve.cuh"

namespace pmp {

template <class R, int NX> struct LinearCostate {
    R P[NX * NX];  // row-major
    R x_eq[NX];
};

template <class R> struct ForwardIO {
    const R* x0;  // [NX][n]
    R* z_end;     // [NX + 1][n]: x, cost
    R *ts, *xs, *cost;  // dense: [n][S1], [n][S1][NX], [n][S1] (NULL: not saved)
    int32_t *n_accepted, *n_steps, *result;
    long long n;
};

template <class Sys, class R, int METHOD>
__global__ void __launch_bounds__(128)
forward_kernel(const typename Sys::Consts sc, const SolveOpts<R> o, const ForwardIO<R> io,
               const LinearCostate<R, Sys::NX> pol, const TabParams<R> tp) {
    using M = Math<R>;
    using TB = Tab<METHOD>;
    constexpr int NX = Sys::NX, NZ = NX + 1, S = TB::S;
    const long long n = io.n, i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int S1 = o.max_steps + 1;

    // z' = (f(x, u*(x, P (x - x_eq))), l(x, u*)); also the LQR value 1/2 dx'P dx of the state
    auto field = [&](const R* z, R* k, R& vlqr) {
        R dx[NX], lam[NX], ld[NX], vd;
#pragma unroll
        for (int q = 0; q < NX; q++) dx[q] = z[q] - pol.x_eq[q];
        R v2 = R(0);
#pragma unroll
        for (int r = 0; r < NX; r++) {
            R a = R(0);
#pragma unroll
            for (int q = 0; q < NX; q++) a = M::fma(pol.P[r * NX + q], dx[q], a);
            lam[r] = a;
            v2 = M::fma(a, dx[r], v2);
        }
        vlqr = R(0.5) * v2;
        Sys::template rhs<2>(sc, z, lam, k, vd, ld);
        k[NX] = -vd;
    };

    R z[NZ], k[S][NZ];
#pragma unroll
    for (int q = 0; q < NX; q++) z[q] = io.x0[q * n + i];
    z[NX] = R(0);
    R tprev = o.T0, tnext = M::min(o.T0 + o.dt0, o.T1), vlqr = R(0);
    int nsteps = 0, nacc = 0, result = 0;
    bool at_dtmin = false;
    const R inf = M::inf();
    auto save = [&](int slot) {
        if (!io.ts || slot >= S1) return;
        io.ts[i * S1 + slot] = tprev * o.dir;
#pragma unroll
        for (int q = 0; q < NX; q++) io.xs[(i * S1 + slot) * NX + q] = z[q];
        io.cost[i * S1 + slot] = z[NX];
    };
    bool bad = false;
#pragma unroll
    for (int q = 0; q < NX; q++) bad = bad || (z[q] != z[q]);
    if (bad) result = PMP_RESULT_NAN;
    save(0);
    if (TB::FSAL) field(z, k[0], vlqr);

    while (tprev < o.T1 && result == 0 && nsteps < o.max_steps) {
        const R dt = tnext - tprev, h = dt * o.dir;
        if (!TB::FSAL) field(z, k[0], vlqr);
        R z1[NZ], v1 = R(0);
#pragma unroll
        for (int s = 1; s < S; s++) {
            R zs[NZ];
#pragma unroll
            for (int q = 0; q < NZ; q++) {
                R acc = z[q];
#pragma unroll
                for (int j = 0; j < s; j++)
                    if (TB::a(s, j) != 0.0) acc = M::fma(h * tp.template a<TB>(s, j), k[j][q], acc);
                zs[q] = acc;
            }
            field(zs, k[s], v1);
            if (TB::FSAL && s == S - 1) {
#pragma unroll
                for (int q = 0; q < NZ; q++) z1[q] = zs[q];
            }
        }
        if (!TB::FSAL) {
#pragma unroll
            for (int q = 0; q < NZ; q++) {
                R acc = z[q];
#pragma unroll
                for (int j = 0; j < S; j++)
                    if (TB::b(j) != 0.0) acc = M::fma(h * tp.template b<TB>(j), k[j][q], acc);
                z1[q] = acc;
            }
        }
        bool keep = true, step_nonfinite = false;
        R next_t0, next_t1;
        if (TB::EMBEDDED && o.adaptive) {
            R sumsq = R(0);
#pragma unroll
            for (int q = 0; q < NX; q++) {  // rms norm over the x leaf (see the header)
                R e = (h * tp.template e<TB>(0)) * k[0][q];
#pragma unroll
                for (int j = 1; j < S; j++)
                    if (TB::berr(j) != 0.0) e = M::fma(h * tp.template e<TB>(j), k[j][q], e);
                const R scale = M::fma(o.rtol, M::max(M::abs(z[q]), M::abs(z1[q])), o.atol);
                const R ratio = M::ctl_div(e, scale);
                sumsq = M::fma(ratio, ratio, sumsq);
            }
            sumsq = (sumsq != sumsq) ? inf : sumsq;
            const R msq = sumsq * R(1.0 / NX);
            step_nonfinite = !(msq < inf);
            keep = (msq < R(1)) || at_dtmin;
            R factor = o.safety * M::pow_neg_inv(msq, R(0.5) * o.inv_order);
            factor = M::min(M::max(factor, keep ? R(1) : o.factormin), o.factormax);
            R dtn = dt * factor;
            if (o.use_dtmax) dtn = M::min(dtn, o.dtmax);
            if (o.use_dtmin) {
                if (!o.force_dtmin && dtn < o.dtmin) result = PMP_RESULT_DT_MIN;
                at_dtmin = dtn <= o.dtmin;
                dtn = M::max(dtn, o.dtmin);
            } else {
                at_dtmin = false;
            }
            next_t0 = keep ? tnext : tprev;
            next_t1 = next_t0 + dtn;
        } else {
            next_t0 = tnext;
            next_t1 = tnext + dt;
        }
        tprev = M::min(next_t0, o.T1);
        if (next_t1 > o.T1 - M::clip_tol()) next_t1 = keep ? o.T1 : M::fma(R(0.5), o.T1 - tprev, tprev);
        tnext = next_t1;
        nsteps += 1;
        bool nan_state = false;
        if (keep) {
            nacc += 1;
            nan_state = step_nonfinite;
#pragma unroll
            for (int q = 0; q < NZ; q++) {
                z[q] = z1[q];
                nan_state = nan_state || (z1[q] != z1[q]);
                if (TB::FSAL) k[0][q] = k[S - 1][q];
            }
            if (TB::FSAL) vlqr = v1;
            save(nacc);
        }
        if (!TB::FSAL && o.event == PMP_EVENT_LQR_VALUE_LE) {  // RK4: the value of the post-step state
            R kk[NZ];
            field(z, kk, vlqr);
        }
        if (result == 0 && o.event == PMP_EVENT_LQR_VALUE_LE && vlqr <= o.event_level) result = PMP_RESULT_EVENT;
        if (result == 0 && nan_state) result = PMP_RESULT_NAN;
    }
    if (result == 0 && tprev < o.T1) result = PMP_RESULT_MAX_STEPS;
#pragma unroll
    for (int q = 0; q < NZ; q++) io.z_end[q * n + i] = z[q];
    if (io.n_accepted) io.n_accepted[i] = nacc;
    if (io.n_steps) io.n_steps[i] = nsteps;
    if (io.result) io.result[i] = result;
    if (io.ts) {
        for (int slot = nacc + 1; slot < S1; slot++) {
            io.ts[i * S1 + slot] = inf * o.dir;
#pragma unroll
            for (int q = 0; q < NX; q++) io.xs[(i * S1 + slot) * NX + q] = inf;
            io.cost[i * S1 + slot] = inf;
        }
    }
}

template <template <class> class SysT>
int launch_forward(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const double* P,
                   const double* x_eq, void* z_end, void* ts, void* xs, void* cost, int32_t* a, int32_t* s, int32_t* r,
                   cudaStream_t st) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 127) / 128);
#define PMP_FGO(R, METHOD)                                                                                        \
    do {                                                                                                          \
        using Sys = SysT<R>;                                                                                      \
        using TB = Tab<METHOD>;                                                                                   \
        LinearCostate<R, Sys::NX> pol;                                                                            \
        for (int q = 0; q < Sys::NX * Sys::NX; q++) pol.P[q] = R(P[q]);                                           \
        for (int q = 0; q < Sys::NX; q++) pol.x_eq[q] = R(x_eq ? x_eq[q] : 0.0);                                  \
        ForwardIO<R> io{(const R*)x0, (R*)z_end, (R*)ts, (R*)xs, (R*)cost, a, s, r, n};                          \
        forward_kernel<Sys, R, METHOD><<<grid, 128, 0, st>>>(Sys::make(p), make_solve_opts<R, TB>(o, TB::ORDER, Sys::NX), \
                                                             io, pol, TabParams<R>::template make<TB>());        \
        return (int)cudaGetLastError();                                                                           \
    } while (0)
#define PMP_FMETH(R)                                                  \
    do {                                                              \
        if (o.method == PMP_METHOD_RK4) PMP_FGO(R, PMP_METHOD_RK4);   \
        if (o.method == PMP_METHOD_TSIT5) PMP_FGO(R, PMP_METHOD_TSIT5); \
        if (o.method == PMP_METHOD_DOPRI5) PMP_FGO(R, PMP_METHOD_DOPRI5); \
    } while (0)
    if (o.dtype == PMP_F32) PMP_FMETH(float);
    if (o.dtype == PMP_F64) PMP_FMETH(double);
#undef PMP_FMETH
#undef PMP_FGO
    return -1000;
}

}  // namespace pmp
