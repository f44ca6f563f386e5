// pmp_oracle.cpp -- C entry points of the CPU ORACLE (test infrastructure, NOT product code).
// See pmp_oracle.hpp for what is restated and the parity status.  Host pointers everywhere; the
// batch layout is the product's (include/b200pmp.h): struct-of-arrays y[row*n + i].
#include "pmp_oracle.hpp"

#include <cstring>
#include <string>
#include <vector>

#include "dual.hpp"
#include "opcount.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace pmpo;

namespace {
thread_local std::string g_err;
int fail(int code, const char* msg) {
    g_err = msg;
    return code;
}

template <class R, class Sys, bool VXX = false>
int solve_batch_t(const pmp_problem_t& p, const pmp_opts_t& o, int64_t n, const R* y_f, R* y_end,
                  const pmp_dense_t* dense, int32_t* n_acc, int32_t* n_steps, int32_t* result,
                  int n_threads, int model_pass1) {
    constexpr int NX = Sys::NX, NS = 2 * NX + 2 + (VXX ? NX * NX : 0);
    Sys sys(p);
    if constexpr (Sys::NU == 2) sys.model_pass1 = model_pass1 != 0;
    const int64_t S1 = o.max_steps + 1;
    (void)n_threads;
#ifdef _OPENMP
    if (n_threads < 1) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
#endif
    for (int64_t i = 0; i < n; i++) {
        R y0[NS], y1[NS];
        for (int q = 0; q < NS; q++) y0[q] = y_f[q * n + i];
        DenseSink<R> sink;
        if (dense && o.save_dense) {
            if (dense->ts) sink.ts = (R*)dense->ts + i * S1;
            if (dense->x) sink.x = (R*)dense->x + i * S1 * NX;
            if (dense->t) sink.t = (R*)dense->t + i * S1;
            if (dense->v) sink.v = (R*)dense->v + i * S1;
            if (dense->vx) sink.vx = (R*)dense->vx + i * S1 * NX;
            if (VXX && dense->vxx) sink.vxx = (R*)dense->vxx + i * S1 * NX * NX;
        }
        TrajStats st = solve_one<R, Sys, VXX>(sys, o, y0, y1, sink);
        for (int q = 0; q < NS; q++) y_end[q * n + i] = y1[q];
        if (n_acc) n_acc[i] = st.n_accepted;
        if (n_steps) n_steps[i] = st.n_steps;
        if (result) result[i] = st.result;
    }
    return 0;
}

template <class Sys32, class Sys64>
int solve_dispatch(const pmp_problem_t& p, const pmp_opts_t& o, int64_t n, const void* y_f,
                   void* y_end, const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r,
                   int nt, int mp1) {
    if (o.dtype == PMP_F32)
        return solve_batch_t<float, Sys32>(p, o, n, (const float*)y_f, (float*)y_end, dense, a, s, r, nt, mp1);
    return solve_batch_t<double, Sys64>(p, o, n, (const double*)y_f, (double*)y_end, dense, a, s, r, nt, mp1);
}

int check(const pmp_problem_t* p, const pmp_opts_t* o) {
    if (!p) return fail(PMP_ERR_BAD_ARG, "problem is NULL");
    if (p->system_id < 0 || p->system_id > 2) return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
    if (o) {
        if (!tableau_for(o->method)) return fail(PMP_ERR_UNSUPPORTED, "unknown method");
        if (o->dtype != PMP_F32 && o->dtype != PMP_F64) return fail(PMP_ERR_BAD_ARG, "bad dtype");
        if (o->max_steps < 1) return fail(PMP_ERR_BAD_ARG, "max_steps < 1");
    }
    return 0;
}

template <class R, class Sys> void rhs_batch_t(const pmp_problem_t& p, int64_t n, const R* y, R* dy, int mp1) {
    constexpr int NS = 2 * Sys::NX + 2;
    Sys sys(p);
    if constexpr (Sys::NU == 2) sys.model_pass1 = mp1 != 0;
    for (int64_t i = 0; i < n; i++) {
        R a[NS], b[NS];
        for (int q = 0; q < NS; q++) a[q] = y[q * n + i];
        sys.rhs(a, b);
        for (int q = 0; q < NS; q++) dy[q * n + i] = b[q];
    }
}
template <class R, class Sys>
void ustar_batch_t(const pmp_problem_t& p, int64_t n, const R* x, const R* lam, R* u, int mp1) {
    constexpr int NX = Sys::NX, NU = Sys::NU;
    Sys sys(p);
    if constexpr (Sys::NU == 2) sys.model_pass1 = mp1 != 0;
    for (int64_t i = 0; i < n; i++) {
        R xx[NX], ll[NX], uu[NU];
        for (int q = 0; q < NX; q++) { xx[q] = x[q * n + i]; ll[q] = lam[q * n + i]; }
        sys.ustar(xx, ll, uu);
        for (int q = 0; q < NU; q++) u[q * n + i] = uu[q];
    }
}
// forward closed-loop simulation (levelsets.py:818-836, eval_utils.py:48-99); same contract as pmp_forward_batch_cuda
// with host pointers: x0 [nx][n], z_end [nx+1][n], dense ts [n][S1], x [n][S1][nx], cost [n][S1] (all or none)
template <class R, class Sys>
int forward_batch_t(const pmp_problem_t& p, const pmp_opts_t& o, int64_t n, const R* x0, const double* P, const double* x_eq,
                    const pmp_nn_costate_t* nn, R* z_end, R* ts, R* xs, R* cost, int32_t* n_acc, int32_t* n_steps,
                    int32_t* result) {
    constexpr int NX = Sys::NX;
    Sys sys(p);
    const int64_t S1 = o.max_steps + 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int64_t i = 0; i < n; i++) {
        R a[NX], e[NX + 1];
        for (int q = 0; q < NX; q++) a[q] = x0[q * n + i];
        ForwardSink<R> sink;
        if (ts) { sink.ts = ts + i * S1; sink.x = xs + i * S1 * NX; sink.cost = cost + i * S1; }
        TrajStats st = nn ? forward_one<R, Sys>(sys, o, NnLaw<R, NX>{nn}, a, e, sink)
                          : forward_one<R, Sys>(sys, o, LinearLaw<R, NX>{P, x_eq}, a, e, sink);
        for (int q = 0; q <= NX; q++) z_end[q * n + i] = e[q];
        if (n_acc) n_acc[i] = st.n_accepted;
        if (n_steps) n_steps[i] = st.n_steps;
        if (result) result[i] = st.result;
    }
    return 0;
}

}  // namespace

extern "C" {

const char* pmp_oracle_last_error(void) { return g_err.c_str(); }

int pmp_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// same contract as pmp_solve_batch_cuda (include/b200pmp.h) with host pointers.
// n_threads: OpenMP threads (<1 = all).  model_pass1: score u* pass 1 with the quadratic model.
int pmp_oracle_solve_batch(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* y_f,
                           void* y_end, const pmp_dense_t* dense, int32_t* n_accepted,
                           int32_t* n_steps, int32_t* result, int n_threads, int model_pass1) {
    if (int e = check(p, o)) return e;
    if (!o) return fail(PMP_ERR_BAD_ARG, "opts is NULL");
    if (n < 0 || (n > 0 && (!y_f || !y_end))) return fail(PMP_ERR_BAD_ARG, "bad n / NULL buffers");
    if (o->flags & PMP_FLAG_VXX) {
        if (p->system_id != PMP_SYS_FLATQUAD) return fail(PMP_ERR_UNSUPPORTED, "vxx branch: flatquad only");
        if (o->dtype == PMP_F32)
            return solve_batch_t<float, Flatquad<float>, true>(*p, *o, n, (const float*)y_f, (float*)y_end, dense, n_accepted, n_steps, result, n_threads, model_pass1);
        return solve_batch_t<double, Flatquad<double>, true>(*p, *o, n, (const double*)y_f, (double*)y_end, dense, n_accepted, n_steps, result, n_threads, model_pass1);
    }
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD:
            return solve_dispatch<Flatquad<float>, Flatquad<double>>(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, n_threads, model_pass1);
        case PMP_SYS_ORBITS:
            return solve_dispatch<Orbits<float>, Orbits<double>>(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, n_threads, model_pass1);
        case PMP_SYS_LQR_STATEPENALTY:
            return solve_dispatch<LqrStatePenalty<float>, LqrStatePenalty<double>>(*p, *o, n, y_f, y_end, dense, n_accepted, n_steps, result, n_threads, model_pass1);
    }
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

int pmp_oracle_rhs_batch(const pmp_problem_t* p, int32_t dtype, int64_t n, const void* y, void* dy,
                         int model_pass1) {
    if (int e = check(p, nullptr)) return e;
#define RHS_CASE(SYS)                                                                         \
    if (dtype == PMP_F32) rhs_batch_t<float, SYS<float>>(*p, n, (const float*)y, (float*)dy, model_pass1); \
    else rhs_batch_t<double, SYS<double>>(*p, n, (const double*)y, (double*)dy, model_pass1);  \
    return 0;
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD: RHS_CASE(Flatquad)
        case PMP_SYS_ORBITS: RHS_CASE(Orbits)
        case PMP_SYS_LQR_STATEPENALTY: RHS_CASE(LqrStatePenalty)
    }
#undef RHS_CASE
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

int pmp_oracle_ustar_batch(const pmp_problem_t* p, int32_t dtype, int64_t n, const void* x,
                           const void* lam, void* u, int model_pass1) {
    if (int e = check(p, nullptr)) return e;
#define US_CASE(SYS)                                                                                  \
    if (dtype == PMP_F32) ustar_batch_t<float, SYS<float>>(*p, n, (const float*)x, (const float*)lam, (float*)u, model_pass1); \
    else ustar_batch_t<double, SYS<double>>(*p, n, (const double*)x, (const double*)lam, (double*)u, model_pass1); \
    return 0;
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD: US_CASE(Flatquad)
        case PMP_SYS_ORBITS: US_CASE(Orbits)
        case PMP_SYS_LQR_STATEPENALTY: US_CASE(LqrStatePenalty)
    }
#undef US_CASE
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

// flatquad only, fp64: u_star_2d(..., smooth=True) (pontryagin_utils.py:217-254) at n points, x, lam [6][n] -> u [2][n]
int pmp_oracle_ustar_smooth(const pmp_problem_t* p, int64_t n, const double* x, const double* lam, double* u) {
    if (int e = check(p, nullptr)) return e;
    if (p->system_id != PMP_SYS_FLATQUAD) return fail(PMP_ERR_UNSUPPORTED, "ustar_smooth: flatquad only");
    Flatquad<double> sys(*p);
    for (int64_t i = 0; i < n; i++) {
        double xi[6], li[6], ui[2];
        for (int q = 0; q < 6; q++) { xi[q] = x[q * n + i]; li[q] = lam[q * n + i]; }
        sys.ustar_smooth(xi, li, ui);
        u[i] = ui[0]; u[n + i] = ui[1];
    }
    return 0;
}

// flatquad only: Jacobians of pmp_rhs through u* and vxx_dot (pontryagin_utils.py:460-467), fp64, one state per call.
// jac_out = fx, flam, gx, glam (4 x 36, row-major), vxx_dot_out 36.
int pmp_oracle_vxx_rhs(const pmp_problem_t* p, const double* x, const double* lam, const double* vxx, double* jac_out,
                       double* vxx_dot_out) {
    if (int e = check(p, nullptr)) return e;
    if (p->system_id != PMP_SYS_FLATQUAD) return fail(PMP_ERR_UNSUPPORTED, "vxx branch: flatquad only");
    Flatquad<double> sys(*p);
    double fx[6][6], flam[6][6], gx[6][6], glam[6][6];
    sys.pmp_jacobians(x, lam, fx, flam, gx, glam);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            jac_out[i * 6 + j] = fx[i][j]; jac_out[36 + i * 6 + j] = flam[i][j];
            jac_out[72 + i * 6 + j] = gx[i][j]; jac_out[108 + i * 6 + j] = glam[i][j];
        }
    sys.vxx_dot(x, lam, vxx, vxx_dot_out);
    return 0;
}

// Any system: the same quantities by FORWARD-MODE differentiation of the oracle's own pmp_rhs (dual numbers through u*, decisions
// on values -- what jax.jacobian(pmp_rhs, argnums=(0, 1)) does at pontryagin_utils.py:462), the way the CUDA adaptor for generated
// functors forms them (csrc/pmp_vxx_dual.cuh).  jac_out = fx, flam, gx, glam (4 x nx*nx) from 2 nx passes; vxx_dot_out from nx
// passes along (e_j, vxx e_j):  vxx'[:, j] = d lamdot - vxx d xdot.  fp64, one state per call.
int pmp_oracle_vxx_rhs_fwd(const pmp_problem_t* p, const double* x, const double* lam, const double* vxx, double* jac_out,
                           double* vxx_dot_out) {
    if (int e = check(p, nullptr)) return e;
    auto run = [&](auto sys) {
        using Sys = decltype(sys);
        constexpr int NX = Sys::NX, NS = 2 * NX + 2, NV = NX * NX;
        auto pass = [&](const double* dx, const double* dlam, double* dxd, double* dld) {
            Dual y[NS], dy[NS];
            for (int q = 0; q < NX; q++) { y[q] = Dual(x[q], dx[q]); y[NX + 2 + q] = Dual(lam[q], dlam[q]); }
            y[NX] = Dual(0.0); y[NX + 1] = Dual(0.0);
            sys.rhs(y, dy);
            for (int q = 0; q < NX; q++) { dxd[q] = dy[q].d; dld[q] = dy[NX + 2 + q].d; }
        };
        double zero[NX], unit[NX], a[NX], b[NX];
        for (int q = 0; q < NX; q++) zero[q] = 0.0;
        for (int j = 0; j < NX; j++) {
            for (int q = 0; q < NX; q++) unit[q] = q == j ? 1.0 : 0.0;
            pass(unit, zero, a, b);  // d/dx_j
            for (int i = 0; i < NX; i++) { jac_out[i * NX + j] = a[i]; jac_out[2 * NV + i * NX + j] = b[i]; }
            pass(zero, unit, a, b);  // d/dlam_j
            for (int i = 0; i < NX; i++) { jac_out[NV + i * NX + j] = a[i]; jac_out[3 * NV + i * NX + j] = b[i]; }
            double col[NX];
            for (int q = 0; q < NX; q++) col[q] = vxx[q * NX + j];
            pass(unit, col, a, b);
            for (int i = 0; i < NX; i++) {
                double acc = b[i];
                for (int k = 0; k < NX; k++) acc -= vxx[i * NX + k] * a[k];
                vxx_dot_out[i * NX + j] = acc;
            }
        }
    };
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD: run(Flatquad<Dual>(*p)); return 0;
        case PMP_SYS_ORBITS: run(Orbits<Dual>(*p)); return 0;
        case PMP_SYS_LQR_STATEPENALTY: run(LqrStatePenalty<Dual>(*p)); return 0;
    }
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

// costate law of the ensemble at n points: x [nx][n] -> lam [nx][n], vmean [n], vstd [n]   (fp64 weights)
int pmp_oracle_nn_costate(const pmp_nn_costate_t* nn, int32_t nx, int64_t n, const double* x, double* lam, double* vmean,
                          double* vstd) {
    if (!nn || !x || !lam || !vmean || !vstd) return fail(PMP_ERR_BAD_ARG, "NULL argument");
    for (int64_t i = 0; i < n; i++) {
        double xi[16], li[16];
        for (int q = 0; q < nx; q++) xi[q] = x[q * n + i];
        if (nx == 6) NnLaw<double, 6>{nn}.eval(xi, li, vmean[i], vstd[i]);
        else if (nx == 2) NnLaw<double, 2>{nn}.eval(xi, li, vmean[i], vstd[i]);
        else return fail(PMP_ERR_UNSUPPORTED, "nx must be 2 or 6");
        for (int q = 0; q < nx; q++) lam[q * n + i] = li[q];
    }
    return 0;
}

// nn != NULL: the costate law of the value-network ensemble (weights in HOST memory, dtype of opts) replaces P (x - x_eq)
int pmp_oracle_forward_batch(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const void* x0, const double* P,
                             const double* x_eq, const pmp_nn_costate_t* nn, void* z_end, void* ts, void* xs, void* cost,
                             int32_t* n_accepted, int32_t* n_steps, int32_t* result) {
    if (int e = check(p, o)) return e;
    if (!o || !x0 || (!P && !nn) || !z_end || n < 0) return fail(PMP_ERR_BAD_ARG, "bad arguments");
#define FWD(R, SYS) return forward_batch_t<R, SYS<R>>(*p, *o, n, (const R*)x0, P, x_eq, nn, (R*)z_end, (R*)ts, (R*)xs, (R*)cost, n_accepted, n_steps, result)
    if (o->dtype == PMP_F32) {
        switch (p->system_id) {
            case PMP_SYS_FLATQUAD: FWD(float, Flatquad);
            case PMP_SYS_ORBITS: FWD(float, Orbits);
            case PMP_SYS_LQR_STATEPENALTY: FWD(float, LqrStatePenalty);
        }
    } else {
        switch (p->system_id) {
            case PMP_SYS_FLATQUAD: FWD(double, Flatquad);
            case PMP_SYS_ORBITS: FWD(double, Orbits);
            case PMP_SYS_LQR_STATEPENALTY: FWD(double, LqrStatePenalty);
        }
    }
#undef FWD
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

// dense output (diffrax Solution.evaluate) on one step per query; same contract as pmp_interp_batch_cuda
int pmp_oracle_interp_batch(const pmp_problem_t* p, const pmp_opts_t* o, int64_t nq, const void* y0, const void* t1,
                            const void* target, int32_t mode, void* y_out) {
    if (int e = check(p, o)) return e;
    auto run = [&](auto sys, auto zero) {
        using Sys = decltype(sys);
        using R = decltype(zero);
        constexpr int NS = 2 * Sys::NX + 2;
        for (int64_t i = 0; i < nq; i++) {
            R a[NS], b[NS];
            for (int q = 0; q < NS; q++) a[q] = ((const R*)y0)[q * nq + i];
            interp_one<R, Sys>(sys, *o, a, ((const R*)t1)[i], ((const R*)target)[i], mode, b);
            for (int q = 0; q < NS; q++) ((R*)y_out)[q * nq + i] = b[q];
        }
    };
    const bool f32 = o->dtype == PMP_F32;
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD: if (f32) run(Flatquad<float>(*p), 0.0f); else run(Flatquad<double>(*p), 0.0); return 0;
        case PMP_SYS_ORBITS: if (f32) run(Orbits<float>(*p), 0.0f); else run(Orbits<double>(*p), 0.0); return 0;
        case PMP_SYS_LQR_STATEPENALTY: if (f32) run(LqrStatePenalty<float>(*p), 0.0f); else run(LqrStatePenalty<double>(*p), 0.0); return 0;
    }
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

// Algorithmic work.  out[0..1] = flops, specials of ONE f_extended evaluation
// (u* scored with the quadratic model, i.e. the cheaper reading of the reference's algorithm);
// out[2..3] = flops, specials of ONE step attempt of opts->method excluding nothing (stages, RHS
// evaluations, error estimate, norm, controller); measured by running one attempt on the counting
// scalar from the state y0 [NS] (doubles).
// Directional derivative of the endpoint with respect to the terminal state (fp64): y_f, dy_f [NS][n] -> y_end, dy_end [NS][n].
// The test-side reference of pmp_solve_jvp_cuda (experiment_orbits.py:186-197: jax.jacobian through the value-parameterised solve).
int pmp_oracle_solve_jvp(const pmp_problem_t* p, const pmp_opts_t* o, int64_t n, const double* y_f, const double* dy_f,
                         double* y_end, double* dy_end, int32_t* n_accepted, int32_t* n_steps, int32_t* result) {
    if (int e = check(p, o)) return e;
    if (o->flags & PMP_FLAG_VXX) return fail(PMP_ERR_UNSUPPORTED, "jvp: no vxx leaf");
    auto run = [&](auto sys) {
        using Sys = decltype(sys);
        constexpr int NS = 2 * Sys::NX + 2;
        pmp_opts_t oo = *o;
        oo.save_dense = 0;
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t i = 0; i < n; i++) {
            Dual y[NS], ye[NS];
            for (int q = 0; q < NS; q++) y[q] = Dual(y_f[q * n + i], dy_f[q * n + i]);
            TrajStats st = solve_one<Dual, Sys>(sys, oo, y, ye, DenseSink<Dual>{});
            for (int q = 0; q < NS; q++) { y_end[q * n + i] = ye[q].v; dy_end[q * n + i] = ye[q].d; }
            if (n_accepted) n_accepted[i] = st.n_accepted;
            if (n_steps) n_steps[i] = st.n_steps;
            if (result) result[i] = st.result;
        }
    };
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD: run(Flatquad<Dual>(*p)); return 0;
        case PMP_SYS_ORBITS: run(Orbits<Dual>(*p)); return 0;
        case PMP_SYS_LQR_STATEPENALTY: run(LqrStatePenalty<Dual>(*p)); return 0;
    }
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

int pmp_oracle_opcount(const pmp_problem_t* p, const pmp_opts_t* o, const double* y0, int64_t* out) {
    if (int e = check(p, o)) return e;
    auto run = [&](auto sys) {
        using Sys = decltype(sys);
        constexpr int NS = 2 * Sys::NX + 2;
        if constexpr (Sys::NU == 2) sys.model_pass1 = true;
        Counted y[NS], dy[NS], ye[NS];
        for (int q = 0; q < NS; q++) y[q] = Counted(y0[q]);
        op_counts() = OpCounts{};
        sys.rhs(y, dy);
        out[0] = op_counts().flops;
        out[1] = op_counts().special;
        pmp_opts_t one = *o;
        one.max_steps = 1;
        one.save_dense = 0;
        op_counts() = OpCounts{};
        solve_one<Counted, Sys>(sys, one, y, ye, DenseSink<Counted>{});
        // solve_one evaluates the FSAL stage once up front: that RHS is per trajectory, not per attempt
        out[2] = op_counts().flops - out[0];
        out[3] = op_counts().special - out[1];
    };
    if (o->flags & PMP_FLAG_VXX) {
        // value-Hessian branch: y0 holds 50 doubles; out[0..1] = f_extended INCLUDING vxx_dot as the oracle
        // evaluates it (closed-form Jacobians, four dense 6x6 products, pontryagin_utils.py:465)
        if (p->system_id != PMP_SYS_FLATQUAD) return fail(PMP_ERR_UNSUPPORTED, "vxx branch: flatquad only");
        Flatquad<Counted> sys(*p);
        sys.model_pass1 = true;
        constexpr int NS = 50;
        Counted y[NS], dy[NS], ye[NS];
        for (int q = 0; q < NS; q++) y[q] = Counted(y0[q]);
        op_counts() = OpCounts{};
        sys.rhs(y, dy);
        sys.vxx_dot(y, y + 8, y + 14, dy + 14);
        out[0] = op_counts().flops;
        out[1] = op_counts().special;
        pmp_opts_t one = *o;
        one.max_steps = 1;
        one.save_dense = 0;
        one.event = PMP_EVENT_NONE;
        op_counts() = OpCounts{};
        solve_one<Counted, Flatquad<Counted>, true>(sys, one, y, ye, DenseSink<Counted>{});
        out[2] = op_counts().flops - out[0];
        out[3] = op_counts().special - out[1];
        return 0;
    }
    switch (p->system_id) {
        case PMP_SYS_FLATQUAD: run(Flatquad<Counted>(*p)); return 0;
        case PMP_SYS_ORBITS: run(Orbits<Counted>(*p)); return 0;
        case PMP_SYS_LQR_STATEPENALTY: run(LqrStatePenalty<Counted>(*p)); return 0;
    }
    return fail(PMP_ERR_UNSUPPORTED, "unknown system_id");
}

}  // extern "C"
