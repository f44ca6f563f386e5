// pmp_oracle.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A plain, scalar, readable restatement of the reference's PMP backward-shooting path.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it;
// nothing under approximate_optimal_control_b200/ includes, links or calls this directory.
//
// PARITY STATUS (see DESIGN.md "Oracle"):
//   * RHS maths (u_star_2d, u_star_new, f_extended): PINNED against the reference's own Python code
//     executed in the build container through a torch-backed stand-in for jax
//     (tests/golden/make_golden.py -> tests/golden/rhs_*.npz).
//   * Converged trajectories: PINNED against scipy.integrate.solve_ivp(DOP853, rtol=1e-12) driven by
//     the reference's own f_extended (tests/golden/traj_*.npz) and against the analytic LQR
//     expm known answer.
//   * Adaptive step control (accept/reject sequence): diffrax 0.4.1 is an un-vendored dependency
//     (requirements.txt:10) that cannot be installed offline and the reference holds no golden
//     vectors for it => the control-flow restatement below is PARITY UNPINNED.
//
// Everything is templated on the scalar type R so that the same code runs in float, double and on
// the op-counting scalar of opcount.hpp (the algorithmic-flop contract of SURVEY.md section 8(d)).
#pragma once
#include <cmath>
#include <vector>
#include <cstdint>
#include <limits>
#include <type_traits>

#include "../include/b200pmp.h"

namespace pmpo {

// ---- scalar helpers (overloaded for the counting scalar in opcount.hpp) -------------------------
inline float r_sin(float x) { return std::sin(x); }
inline double r_sin(double x) { return std::sin(x); }
inline float r_cos(float x) { return std::cos(x); }
inline double r_cos(double x) { return std::cos(x); }
inline float r_sqrt(float x) { return std::sqrt(x); }
inline double r_sqrt(double x) { return std::sqrt(x); }
inline float r_pow(float x, float y) { return std::pow(x, y); }
inline double r_pow(double x, double y) { return std::pow(x, y); }
inline float r_abs(float x) { return std::fabs(x); }
inline double r_abs(double x) { return std::fabs(x); }
inline bool r_isnan(float x) { return std::isnan(x); }
inline bool r_isnan(double x) { return std::isnan(x); }
inline double r_val(float x) { return x; }
inline double r_val(double x) { return x; }
// NaN-propagating max/min like jnp.maximum / jnp.minimum
template <class R> inline R r_max(R a, R b) {
    if (r_isnan(a)) return a;
    if (r_isnan(b)) return b;
    return r_val(a) > r_val(b) ? a : b;
}
template <class R> inline R r_min(R a, R b) {
    if (r_isnan(a)) return a;
    if (r_isnan(b)) return b;
    return r_val(a) < r_val(b) ? a : b;
}
template <class R> inline R r_clip(R x, R lo, R hi) { return r_min(r_max(x, lo), hi); }
template <class R> inline R r_inf() { return R(std::numeric_limits<double>::infinity()); }
template <class R> inline R r_nan() { return R(std::numeric_limits<double>::quiet_NaN()); }

// =================================================================================================
// Systems.  Each supplies f, l, dH/dx (u held fixed), and u* = argmin_u H.
// =================================================================================================

// ---- flat (planar) quadrotor: flatquad_landing_experiment.py:269-329 ----------------------------
// Quantities that do not depend on (x, lambda) -- 1/m, r/I, the constant Hessian H_uu, its inverse,
// the box vertices and per-edge constants -- are folded in the constructor, as a compiler would do
// for the reference's jitted code, so that the op-counting instantiation reports the per-call work
// only ("H_uu and vertex constants folded").
template <class R> struct Flatquad {
    static constexpr int NX = 6, NU = 2;
    R m, g, r, I, q_pos, q_ang, q_vel, q_om;
    R ulo[2], uhi[2];
    bool model_pass1 = false;  // score pass 1 with the quadratic model instead of the true H
    // folded constants
    R inv_m, r_I, two_g_m, two_g, two_qpos, two_qang, two_qvel, two_qom;
    R Huu[2][2], Hinv[2][2], half_h00, half_h11;
    R hull[4][2];
    int edge_axis[4];   // the coordinate that moves along edge k (a-b has one non-zero entry)
    R edge_g[4];        // (b' H_uu)[axis]
    R edge_d[4];        // (a-b)[axis]
    R edge_inv_den[4];  // 1 / ((a-b)' H_uu (a-b))

    explicit Flatquad(const pmp_problem_t& p) {
        m = R(p.params[0]); g = R(p.params[1]); r = R(p.params[2]); I = R(p.params[3]);
        q_pos = R(p.params[4]); q_ang = R(p.params[5]); q_vel = R(p.params[6]); q_om = R(p.params[7]);
        for (int i = 0; i < 2; i++) { ulo[i] = R(p.u_lo[i]); uhi[i] = R(p.u_hi[i]); }
        inv_m = R(1) / m; r_I = r / I; two_g = R(2) * g; two_g_m = two_g / m;
        two_qpos = R(2) * q_pos; two_qang = R(2) * q_ang; two_qvel = R(2) * q_vel; two_qom = R(2) * q_om;
        // H_uu = d^2/du^2 (ax^2+ay^2+aw^2) = 2 [[1/m^2+r^2/I^2, 1/m^2-r^2/I^2],[.,.]]  (constant)
        R d = R(2) * (inv_m * inv_m + r_I * r_I), o = R(2) * (inv_m * inv_m - r_I * r_I);
        Huu[0][0] = d; Huu[0][1] = o; Huu[1][0] = o; Huu[1][1] = d;
        R det = d * d - o * o;
        Hinv[0][0] = d / det; Hinv[0][1] = -o / det; Hinv[1][0] = -o / det; Hinv[1][1] = d / det;
        half_h00 = R(0.5) * d; half_h11 = R(0.5) * d;
        // pontryagin_utils.py:78-86 box vertices; edge k = (a = vertex k, b = vertex k-1)
        R h[4][2] = {{ulo[0], ulo[1]}, {ulo[0], uhi[1]}, {uhi[0], uhi[1]}, {uhi[0], ulo[1]}};
        for (int k = 0; k < 4; k++) { hull[k][0] = h[k][0]; hull[k][1] = h[k][1]; }
        for (int k = 0; k < 4; k++) {
            const R* a = hull[k];
            const R* b = hull[(k + 3) % 4];
            int ax = (r_val(a[0]) != r_val(b[0])) ? 0 : 1;
            edge_axis[k] = ax;
            edge_d[k] = a[ax] - b[ax];
            edge_g[k] = b[0] * Huu[0][ax] + b[1] * Huu[1][ax];
            edge_inv_den[k] = R(1) / (edge_d[k] * Huu[ax][ax] * edge_d[k]);
        }
    }

    // xdot = f(x,u)                                  flatquad_landing_experiment.py:277-292
    void f(const R* x, const R* u, R* xd) const {
        R s = r_sin(x[2]), c = r_cos(x[2]);
        f_sc(x, u, s, c, xd);
    }
    void f_sc(const R* x, const R* u, R s, R c, R* xd) const {
        R Sm = (u[0] + u[1]) * inv_m;
        xd[0] = x[3];
        xd[1] = x[4];
        xd[2] = x[5];
        xd[3] = -(s * Sm);
        xd[4] = c * Sm - g;
        xd[5] = (u[1] - u[0]) * r_I;
    }
    // running cost l(x,u)                            flatquad_landing_experiment.py:294-329
    // state cost on (px,py,cos,sin,vx,vy,omega) minus its value at 0, diagonal Q = 1/scale^2;
    // input cost = |accelerations|^2.
    R l(const R* x, const R* u) const {
        R s = r_sin(x[2]), c = r_cos(x[2]);
        R xd[6];
        f_sc(x, u, s, c, xd);
        return l_acc(x, s, c, xd[3], xd[4], xd[5]);
    }
    R l_acc(const R* x, R s, R c, R ax, R ay, R aw) const {
        R cm1 = c - R(1);
        R state_cost = q_pos * (x[0] * x[0] + x[1] * x[1]) + q_ang * (cm1 * cm1 + s * s) +
                       q_vel * (x[3] * x[3] + x[4] * x[4]) + q_om * (x[5] * x[5]);
        return state_cost + (ax * ax + ay * ay + aw * aw);
    }
    // dH/dx at fixed u, H = l + lam.f                pontryagin_utils.py:430,439 (hand-derived;
    // checked against the reference's autodiff by tests/golden/rhs_flatquad_*.npz)
    void dHdx_sm(const R* x, const R* lam, R s, R c, R Sm, R* hx) const {
        hx[0] = two_qpos * x[0];
        hx[1] = two_qpos * x[1];
        hx[2] = s * (two_qang + two_g * Sm) - Sm * (lam[3] * c + lam[4] * s);
        hx[3] = two_qvel * x[3] + lam[0];
        hx[4] = two_qvel * x[4] + lam[1];
        hx[5] = two_qom * x[5] + lam[2];
    }
    // gradient of H in u at u=0                      pontryagin_utils.py:37
    void hu0(const R* lam, R s, R c, R* Hu) const {
        R base = (lam[4] * c - lam[3] * s) * inv_m - two_g_m * c;
        R rot = lam[5] * r_I;
        Hu[0] = base - rot;
        Hu[1] = base + rot;
    }
    // 1/2 w'H_uu w + gvec'w in 9 flops
    R quad(R w0, R w1, R g0, R g1) const {
        return w0 * (half_h00 * w0 + Huu[0][1] * w1 + g0) + w1 * (half_h11 * w1 + g1);
    }

    // u_star_2d(x, costate, problem_params, smooth=False)     pontryagin_utils.py:11-216
    void ustar(const R* x, const R* lam, R* u) const {
        R s = r_sin(x[2]), c = r_cos(x[2]);
        ustar_sc(x, lam, s, c, u);
    }
    void ustar_sc(const R* x, const R* lam, R s, R c, R* u) const {
        int best;
        bool moving_free;
        ustar_active(x, lam, s, c, u, best, moving_free);
    }
    // u_star_2d(x, costate, problem_params, smooth=True)      pontryagin_utils.py:217-254 (fp64 restatement; a debugging aid of
    // the reference): softmax blend of the five candidates, scored with the TRUE H(u) = l + lambda.f as the reference does
    void ustar_smooth(const R* x, const R* lam, R* u) const {
        const R s = r_sin(x[2]), c = r_cos(x[2]);
        R Hu[2], cand[5][2], Hs[5];
        hu0(lam, s, c, Hu);
        cand[0][0] = -(Hinv[0][0] * Hu[0] + Hinv[0][1] * Hu[1]);
        cand[0][1] = -(Hinv[1][0] * Hu[0] + Hinv[1][1] * Hu[1]);
        for (int k = 0; k < 4; k++) {
            const R* a = hull[k];
            const R* b = hull[(k + 3) % 4];
            const int ax = edge_axis[k];
            R sc = r_clip(-((Hu[ax] + edge_g[k]) * edge_d[k]) * edge_inv_den[k], R(0), R(1));
            cand[k + 1][ax] = sc * a[ax] + (R(1) - sc) * b[ax];
            cand[k + 1][1 - ax] = a[1 - ax];
        }
        for (int k = 0; k < 5; k++) {  // :90 all_Hs = vmap(H_fct)(all_candidates)
            R xd[6];
            f_sc(x, cand[k], s, c, xd);
            R h = l_acc(x, s, c, xd[3], xd[4], xd[5]);
            for (int i = 0; i < 6; i++) h = h + lam[i] * xd[i];
            Hs[k] = h;
        }
        const double ss = 0.01;  // :232 smoothing_size
        auto lse = [](const double* a, int n) {
            double m = a[0];
            for (int i = 1; i < n; i++) m = a[i] > m ? a[i] : m;
            double acc = 0;
            for (int i = 0; i < n; i++) acc += std::exp(a[i] - m);
            return m + std::log(acc);
        };
        // :231-238 violations of the unconstrained solution, smooth max, softplus-clipped at 0, times 1000
        const double cv[4] = {(r_val(ulo[0]) - r_val(cand[0][0])) / ss, (r_val(ulo[1]) - r_val(cand[0][1])) / ss,
                              (r_val(cand[0][0]) - r_val(uhi[0])) / ss, (r_val(cand[0][1]) - r_val(uhi[1])) / ss};
        const double softmax_violation = lse(cv, 4) * ss;
        const double z[2] = {0.0, softmax_violation / ss};
        const double penalty = 1000.0 * lse(z, 2) * ss;
        // :241-251 weights = softmax(-all_Hs_adjusted / smoothing_size); u = weights' candidates
        double sc5[5], m = 0, wsum = 0, a0 = 0, a1 = 0;
        for (int k = 0; k < 5; k++) {
            sc5[k] = -(r_val(Hs[k]) + (k == 0 ? penalty : 0.0)) / ss;
            m = (k == 0 || sc5[k] > m) ? sc5[k] : m;
        }
        for (int k = 0; k < 5; k++) {
            const double w = std::exp(sc5[k] - m);
            wsum += w; a0 += w * r_val(cand[k][0]); a1 += w * r_val(cand[k][1]);
        }
        u[0] = R(a0 / wsum);
        u[1] = R(a1 / wsum);
    }
    // same, also reporting which candidate won (0 = unconstrained, k = edge k) and, for an edge, whether its
    // line-search parameter is strictly inside (0,1) (else the point sits in a corner)
    void ustar_active(const R* x, const R* lam, R s, R c, R* u, int& best_out, bool& moving_free) const {
        R Hu[2];
        bool edge_free[5] = {false, false, false, false, false};
        hu0(lam, s, c, Hu);
        R cand[5][2];
        // :41  unconstrained minimiser solve(H_uu, -H_u)
        cand[0][0] = -(Hinv[0][0] * Hu[0] + Hinv[0][1] * Hu[1]);
        cand[0][1] = -(Hinv[1][0] * Hu[0] + Hinv[1][1] * Hu[1]);
        // :53-75,86  minimiser along each box edge {s a + (1-s) b}:
        //   s = -(H_u + b'H_uu)(a-b) / ((a-b)'H_uu(a-b)), clipped to [0,1]
        for (int k = 0; k < 4; k++) {
            const R* a = hull[k];
            const R* b = hull[(k + 3) % 4];
            const int ax = edge_axis[k];
            R sc = r_clip(-((Hu[ax] + edge_g[k]) * edge_d[k]) * edge_inv_den[k], R(0), R(1));
            edge_free[k + 1] = r_val(sc) > 0.0 && r_val(sc) < 1.0;
            cand[k + 1][ax] = sc * a[ax] + (R(1) - sc) * b[ax];
            cand[k + 1][1 - ax] = a[1 - ax];
        }
        // :129  is the unconstrained candidate outside the box?  (NaN compares false => "inside")
        R viol = r_max(r_max(ulo[0] - cand[0][0], ulo[1] - cand[0][1]),
                       r_max(cand[0][0] - uhi[0], cand[0][1] - uhi[1]));
        bool outside = r_val(viol) > 0;
        // :88-90  score ALL five candidates (the reference vmaps H over them, no branching)
        R Hs[5];
        for (int k = 0; k < 5; k++) {
            if (model_pass1) {  // quadratic model, constant term dropped (equal by assumption, :34)
                Hs[k] = quad(cand[k][0], cand[k][1], Hu[0], Hu[1]);
            } else {  // :35,90 the true H(u) = l(x,u) + lam.f(x,u)
                R xd[6];
                f_sc(x, cand[k], s, c, xd);
                R h = l_acc(x, s, c, xd[3], xd[4], xd[5]);
                for (int i = 0; i < 6; i++) h = h + lam[i] * xd[i];
                Hs[k] = h;
            }
        }
        // :131-135  penalty = outside*inf is +inf when outside and NaN (False*inf) when inside;
        // argmin returns the NaN slot => inside: candidate 0 wins unconditionally; outside: best edge
        const R penalty = outside ? r_inf<R>() : r_nan<R>();
        Hs[0] = Hs[0] + penalty;
        int best = argmin5(Hs);
        // :151-158 second pass: re-expand about the winner, compare suboptimalities
        R us0 = cand[best][0], us1 = cand[best][1];
        R hn0 = Huu[0][0] * us0 + Huu[0][1] * us1 + Hu[0];
        R hn1 = Huu[1][0] * us0 + Huu[1][1] * us1 + Hu[1];
        for (int k = 0; k < 5; k++) Hs[k] = quad(cand[k][0] - us0, cand[k][1] - us1, hn0, hn1);
        Hs[0] = Hs[0] + penalty;
        best = argmin5(Hs);
        u[0] = cand[best][0];
        u[1] = cand[best][1];
        best_out = best;
        moving_free = edge_free[best];
    }

    // Jacobians of pmp_rhs(x, lam) = (f(x, u*), -H_x(x, u*, lam)) with respect to (x, lam), differentiated
    // THROUGH u_star_2d as jax does at pontryagin_utils.py:462-463: the argmin index is piecewise constant, the
    // winning candidate is differentiable (unconstrained: -H_uu^-1 H_u; edge: d/dz of the clipped line search,
    // zero when clipped).  (fx, flam), (gx, glam) in the reference's naming.
    void pmp_jacobians(const R* x, const R* lam, R fx[6][6], R flam[6][6], R gx[6][6], R glam[6][6]) const {
        R s = r_sin(x[2]), c = r_cos(x[2]);
        R u[2];
        int best;
        bool moving_free;
        ustar_active(x, lam, s, c, u, best, moving_free);
        // H_u(u=0) depends on z in {Phi, lam3, lam4, lam5} only: dHu[z][component]
        const R dbase[4] = {(-(lam[4] * s) - lam[3] * c) * inv_m + two_g_m * s, -(s * inv_m), c * inv_m, R(0)};
        const R drot[4] = {R(0), R(0), R(0), r_I};
        R du[4][2];  // du*/dz
        for (int z = 0; z < 4; z++) {
            const R d0 = dbase[z] - drot[z], d1 = dbase[z] + drot[z];
            if (best == 0) {
                du[z][0] = -(Hinv[0][0] * d0 + Hinv[0][1] * d1);
                du[z][1] = -(Hinv[1][0] * d0 + Hinv[1][1] * d1);
            } else {
                const int ax = edge_axis[best - 1];
                du[z][0] = R(0); du[z][1] = R(0);
                if (moving_free) du[z][ax] = -((ax == 0 ? d0 : d1) / Huu[ax][ax]);
            }
        }
        const R Sm = (u[0] + u[1]) * inv_m;
        const R fu[6][2] = {{R(0), R(0)}, {R(0), R(0)}, {R(0), R(0)}, {-(s * inv_m), -(s * inv_m)},
                            {c * inv_m, c * inv_m}, {-r_I, r_I}};
        const R hxu2 = (two_g * s - (lam[3] * c + lam[4] * s)) * inv_m;  // d(H_x[2])/dFl = d/dFr
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) { fx[i][j] = R(0); flam[i][j] = R(0); gx[i][j] = R(0); glam[i][j] = R(0); }
        // f_x
        fx[0][3] = R(1); fx[1][4] = R(1); fx[2][5] = R(1);
        fx[3][2] = -(c * Sm); fx[4][2] = -(s * Sm);
        // + f_u u_x (u depends on x through Phi only), f_u u_lam
        const int lcol[4] = {-1, 3, 4, 5};
        for (int i = 3; i < 6; i++) {
            fx[i][2] = fx[i][2] + (fu[i][0] * du[0][0] + fu[i][1] * du[0][1]);
            for (int z = 1; z < 4; z++) flam[i][lcol[z]] = fu[i][0] * du[z][0] + fu[i][1] * du[z][1];
        }
        // gx = -(H_xx + H_xu u_x)
        gx[0][0] = -two_qpos; gx[1][1] = -two_qpos; gx[3][3] = -two_qvel; gx[4][4] = -two_qvel; gx[5][5] = -two_qom;
        const R hxx22 = two_qang * c + two_g * c * Sm - Sm * (lam[4] * c - lam[3] * s);
        gx[2][2] = -(hxx22 + hxu2 * (du[0][0] + du[0][1]));
        // glam = -(f_x' + H_xu u_lam)
        glam[3][0] = R(-1); glam[4][1] = R(-1); glam[5][2] = R(-1);
        glam[2][3] = c * Sm; glam[2][4] = s * Sm;
        for (int z = 1; z < 4; z++) glam[2][lcol[z]] = glam[2][lcol[z]] - hxu2 * (du[z][0] + du[z][1]);
    }
    // vxx_dot = gx + glam vxx - vxx fx - vxx flam vxx        pontryagin_utils.py:465 (forward time)
    void vxx_dot(const R* x, const R* lam, const R* vxx /*[6][6] row-major*/, R* out) const {
        R fx[6][6], flam[6][6], gx[6][6], glam[6][6], t1[6][6];
        pmp_jacobians(x, lam, fx, flam, gx, glam);
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) {
                R a = R(0);
                for (int k = 0; k < 6; k++) a = a + flam[i][k] * vxx[k * 6 + j];
                t1[i][j] = a;  // flam vxx
            }
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) {
                R a = gx[i][j];
                for (int k = 0; k < 6; k++) a = a + glam[i][k] * vxx[k * 6 + j];
                for (int k = 0; k < 6; k++) a = a - vxx[i * 6 + k] * fx[k][j];
                for (int k = 0; k < 6; k++) a = a - vxx[i * 6 + k] * t1[k][j];
                out[i * 6 + j] = a;
            }
    }
    // numpy/jax argmin: the first NaN wins, else the first minimum
    static int argmin5(const R* Hs) {
        for (int k = 0; k < 5; k++)
            if (r_isnan(Hs[k])) return k;
        int best = 0;
        for (int k = 1; k < 5; k++)
            if (r_val(Hs[k]) < r_val(Hs[best])) best = k;
        return best;
    }

    // f_extended, forward time                       pontryagin_utils.py:418-458
    // (u* is computed once; the reference traces it twice, :435 and :446, with identical inputs)
    void rhs(const R* y, R* dy) const {
        const R* x = y;
        const R* lam = y + NX + 2;
        R s = r_sin(x[2]), c = r_cos(x[2]);
        R u[2];
        ustar_sc(x, lam, s, c, u);
        R Sm = (u[0] + u[1]) * inv_m;
        R ax = -(s * Sm), ay = c * Sm - g, aw = (u[1] - u[0]) * r_I;
        dy[0] = x[3]; dy[1] = x[4]; dy[2] = x[5]; dy[3] = ax; dy[4] = ay; dy[5] = aw;
        dy[NX] = R(1);
        dy[NX + 1] = -l_acc(x, s, c, ax, ay, aw);
        R hx[6];
        dHdx_sm(x, lam, s, c, Sm, hx);
        for (int i = 0; i < NX; i++) dy[NX + 2 + i] = -hx[i];
    }
    R running_cost(const R* y) const {
        const R* x = y;
        const R* lam = y + NX + 2;
        R u[2];
        ustar(x, lam, u);
        return l(x, u);
    }
};

// ---- "orbits" toy system: experiment_orbits.py:24-40 --------------------------------------------
template <class R> struct Orbits {
    static constexpr int NX = 2, NU = 1;
    R w_dist, w_u, ulo, uhi;
    explicit Orbits(const pmp_problem_t& p) {
        w_dist = R(p.params[0]); w_u = R(p.params[1]);
        ulo = R(p.u_lo[0]); uhi = R(p.u_hi[0]);
    }
    // f = [[u,-rho],[rho,u]] x, rho = |x|^2-1       experiment_orbits.py:24-30
    void f(const R* x, const R* u, R* xd) const {
        R rho = x[0] * x[0] + x[1] * x[1] - R(1);
        xd[0] = u[0] * x[0] - rho * x[1];
        xd[1] = rho * x[0] + u[0] * x[1];
    }
    // l = rho^2 + 0.1 |x-(0,1)|^2 + 10 u^2          experiment_orbits.py:33-40
    R l(const R* x, const R* u) const {
        R rho = x[0] * x[0] + x[1] * x[1] - R(1);
        R e1 = x[1] - R(1);
        return rho * rho + w_dist * (x[0] * x[0] + e1 * e1) + w_u * (u[0] * u[0]);
    }
    void dHdx(const R* x, const R* u, const R* lam, R* hx) const {
        R rho = x[0] * x[0] + x[1] * x[1] - R(1);
        R e1 = x[1] - R(1);
        // d rho / dx = 2x
        hx[0] = R(4) * rho * x[0] + R(2) * w_dist * x[0] + lam[0] * (u[0] - R(2) * x[0] * x[1]) +
                lam[1] * (rho + R(2) * x[0] * x[0]);
        hx[1] = R(4) * rho * x[1] + R(2) * w_dist * e1 + lam[0] * (-rho - R(2) * x[1] * x[1]) +
                lam[1] * (R(2) * x[0] * x[1] + u[0]);
    }
    // u_star_new: clip(solve(R+R', -lam' df/du), U)  pontryagin_utils.py:532-562, df/du = x
    void ustar(const R* x, const R* lam, R* u) const {
        R un = -(lam[0] * x[0] + lam[1] * x[1]) / (R(2) * w_u);
        u[0] = r_clip(un, ulo, uhi);
    }
    void rhs(const R* y, R* dy) const {
        const R* x = y;
        const R* lam = y + NX + 2;
        R u[1];
        ustar(x, lam, u);
        f(x, u, dy);
        dy[NX] = R(1);
        dy[NX + 1] = -l(x, u);
        R hx[2];
        dHdx(x, u, lam, hx);
        for (int i = 0; i < NX; i++) dy[NX + 2 + i] = -hx[i];
    }
    R running_cost(const R* y) const {
        R u[1];
        ustar(y, y + NX + 2, u);
        return l(y, u);
    }
};

// ---- double integrator with soft state constraints: experiment_lqr_statepenalty.py:23-38 --------
template <class R> struct LqrStatePenalty {
    static constexpr int NX = 2, NU = 1;
    R a, r_u, ulo, uhi;
    explicit LqrStatePenalty(const pmp_problem_t& p) {
        a = R(p.params[0]); r_u = R(p.params[1]);
        ulo = R(p.u_lo[0]); uhi = R(p.u_hi[0]);
    }
    void f(const R* x, const R* u, R* xd) const {  // :23-28
        xd[0] = x[1];
        xd[1] = u[0];
    }
    R l(const R* x, const R* u) const {  // :30-38
        R z1 = r_max(R(0), x[1] - R(1));
        R z0 = r_max(R(0), x[0] - R(1));
        return (x[0] * x[0] + x[1] * x[1]) + r_u * (u[0] * u[0]) + (z1 * z1 / a + z0 * z0 / a);
    }
    void dHdx(const R* x, const R* u, const R* lam, R* hx) const {
        (void)u;
        R z1 = r_max(R(0), x[1] - R(1));
        R z0 = r_max(R(0), x[0] - R(1));
        hx[0] = R(2) * x[0] + R(2) * z0 / a;
        hx[1] = R(2) * x[1] + R(2) * z1 / a + lam[0];
    }
    void ustar(const R* x, const R* lam, R* u) const {  // df/du = (0,1)
        (void)x;
        u[0] = r_clip(-lam[1] / (R(2) * r_u), ulo, uhi);
    }
    void rhs(const R* y, R* dy) const {
        const R* x = y;
        const R* lam = y + NX + 2;
        R u[1];
        ustar(x, lam, u);
        f(x, u, dy);
        dy[NX] = R(1);
        dy[NX + 1] = -l(x, u);
        R hx[2];
        dHdx(x, u, lam, hx);
        for (int i = 0; i < NX; i++) dy[NX + 2 + i] = -hx[i];
    }
    R running_cost(const R* y) const {
        R u[1];
        ustar(y, y + NX + 2, u);
        return l(y, u);
    }
};

// =================================================================================================
// Butcher tableaus (explicit RK, optional embedded error estimate, optional FSAL)
// =================================================================================================
struct Tableau {
    int stages;
    int order;
    bool fsal;      // last stage is evaluated at y1 and reused as the first stage of the next step
    bool embedded;  // has an error estimate
    double c[7];
    double a[7][7];
    double b[7];
    double berr[7];  // b - bhat
};

// diffrax.Tsit5 (Tsitouras 2011); constants as listed in SURVEY.md section 8(c)
inline const Tableau& tsit5() {
    static const Tableau t = [] {
        Tableau q{};
        q.stages = 7; q.order = 5; q.fsal = true; q.embedded = true;
        const double c[7] = {0, 0.161, 0.327, 0.9, 0.9800255409045097, 1, 1};
        const double b[7] = {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
                             -3.290069515436081, 2.324710524099774, 0.0};
        const double bh[7] = {0.09468075576583945, 0.009183565540343254, 0.4877705284247616,
                              1.234297566930479, -2.707712349983526, 1.866628418170587, 1.0 / 66.0};
        const double a[7][7] = {
            {0},
            {0.161},
            {-0.008480655492356992, 0.3354806554923570},
            {2.897153057105494, -6.359448489975075, 4.362295432869581},
            {5.325864828439259, -11.74888356406283, 7.495539342889836, -0.09249506636175525},
            {5.861455442946420, -12.92096931784711, 8.159367898576159, -0.07158497328140100,
             -0.02826905039406838},
            {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
             2.324710524099774}};
        for (int i = 0; i < 7; i++) {
            q.c[i] = c[i]; q.b[i] = b[i]; q.berr[i] = b[i] - bh[i];
            for (int j = 0; j < 7; j++) q.a[i][j] = a[i][j];
        }
        return q;
    }();
    return t;
}

// Dormand-Prince 5(4)
inline const Tableau& dopri5() {
    static const Tableau t = [] {
        Tableau q{};
        q.stages = 7; q.order = 5; q.fsal = true; q.embedded = true;
        const double c[7] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1, 1};
        const double b[7] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0};
        const double bh[7] = {5179.0 / 57600, 0, 7571.0 / 16695, 393.0 / 640, -92097.0 / 339200,
                              187.0 / 2100, 1.0 / 40};
        const double a[7][7] = {
            {0},
            {1.0 / 5},
            {3.0 / 40, 9.0 / 40},
            {44.0 / 45, -56.0 / 15, 32.0 / 9},
            {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729},
            {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656},
            {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
        for (int i = 0; i < 7; i++) {
            q.c[i] = c[i]; q.b[i] = b[i]; q.berr[i] = b[i] - bh[i];
            for (int j = 0; j < 7; j++) q.a[i][j] = a[i][j];
        }
        return q;
    }();
    return t;
}

// classic RK4
inline const Tableau& rk4() {
    static const Tableau t = [] {
        Tableau q{};
        q.stages = 4; q.order = 4; q.fsal = false; q.embedded = false;
        q.c[1] = 0.5; q.c[2] = 0.5; q.c[3] = 1.0;
        q.a[1][0] = 0.5; q.a[2][1] = 0.5; q.a[3][2] = 1.0;
        q.b[0] = 1.0 / 6; q.b[1] = 1.0 / 3; q.b[2] = 1.0 / 3; q.b[3] = 1.0 / 6;
        return q;
    }();
    return t;
}

inline const Tableau* tableau_for(int method) {
    switch (method) {
        case PMP_METHOD_RK4: return &rk4();
        case PMP_METHOD_TSIT5: return &tsit5();
        case PMP_METHOD_DOPRI5: return &dopri5();
    }
    return nullptr;
}

// =================================================================================================
// One trajectory.  Restates diffrax.diffeqsolve(ODETerm(f_extended), Tsit5(), t0, t1, dt0, y0,
// PIDController(rtol, atol, dtmin, dtmax), SaveAt(steps=True, t0=True), max_steps, throw=False)
// as called at pontryagin_utils.py:492-525, with diffrax 0.4.1 semantics.
// =================================================================================================
template <class R> struct DenseSink {  // one trajectory's rows of the dense buffers (may be null)
    R* ts = nullptr; R* x = nullptr; R* t = nullptr; R* v = nullptr; R* vx = nullptr; R* vxx = nullptr;
};

struct TrajStats {
    int32_t n_accepted = 0, n_steps = 0, result = 0;
};

// VXX: the state carries the value Hessian as well (pontryagin_utils.py:460-467): NV = NX*NX more scalars
// after vx, one more leaf of the error norm, optional ||vxx||_F event.  The code below calls the total
// state length NS.
template <class R, class Sys, bool VXX = false>
TrajStats solve_one(const Sys& sys, const pmp_opts_t& o, const R* y_f, R* y_end, DenseSink<R> sink) {
    constexpr int NX = Sys::NX, NB = 2 * NX + 2, NV = VXX ? NX * NX : 0, NS = NB + NV, IT = NX, IV = NX + 1;
    const Tableau& tb = *tableau_for(o.method);
    const bool value_mode = o.indep == PMP_INDEP_VALUE;
    const int S1 = o.max_steps + 1;
    const R inf = r_inf<R>();
    TrajStats st;

    R y[NS];
    for (int i = 0; i < NS; i++) y[i] = y_f[i];

    // time reversal: integrate tau = dir*t forward, vector field multiplied by dir
    const double t_start = value_mode ? double(r_val(y[IV])) : o.t0;
    const R dir = (t_start < o.t1) ? R(1) : R(-1);
    const R T0 = R(t_start) * dir, T1 = R(o.t1) * dir;
    R tprev = T0;
    R tnext = r_min(T0 + R(o.dt0) * dir, T1);
    const R clip_tol = std::is_same<R, float>::value ? R(1e-6) : R(1e-10);  // integrate.py _clip_to_end

    auto save = [&](int slot, R tnow) {
        if (slot >= S1) return;
        if (sink.ts) sink.ts[slot] = tnow * dir;
        if (sink.x) for (int i = 0; i < NX; i++) sink.x[slot * NX + i] = y[i];
        if (sink.t) sink.t[slot] = y[IT];
        if (sink.v) sink.v[slot] = y[IV];
        if (sink.vx) for (int i = 0; i < NX; i++) sink.vx[slot * NX + i] = y[NX + 2 + i];
        if (VXX && sink.vxx) for (int i = 0; i < NV; i++) sink.vxx[slot * NV + i] = y[NB + i];
    };
    auto pad = [&](int from) {
        for (int slot = from; slot < S1; slot++) {
            if (sink.ts) sink.ts[slot] = inf * dir;
            if (sink.x) for (int i = 0; i < NX; i++) sink.x[slot * NX + i] = inf;
            if (sink.t) sink.t[slot] = inf;
            if (sink.v) sink.v[slot] = inf;
            if (sink.vx) for (int i = 0; i < NX; i++) sink.vx[slot * NX + i] = inf;
            if (VXX && sink.vxx) for (int i = 0; i < NV; i++) sink.vxx[slot * NV + i] = inf;
        }
    };
    // internal-time vector field.  time mode: dir * f_extended.  value mode: the same field divided
    // by dv/dtau so that v advances at unit rate (experiment_orbits.py:109-111).
    const bool backward = r_val(dir) < 0;  // multiplying by dir = -1 is a sign flip, not a flop
    auto field = [&](const R* yy, R* k) {
        sys.rhs(yy, k);
        if constexpr (VXX) sys.vxx_dot(yy, yy + NX + 2, yy + NB, k + NB);
        if (value_mode) {
            R vdot = k[IV];  // dv/dt = -l: dy/dv = (dy/dt) / (dv/dt)
            for (int i = 0; i < NS; i++) k[i] = k[i] / vdot;
            k[IV] = R(1);
        }
        if (backward)
            for (int i = 0; i < NS; i++) k[i] = -k[i];
    };

    save(0, tprev);
    bool bad0 = false;
    for (int i = 0; i < NS; i++) bad0 = bad0 || r_isnan(y[i]);
    if (bad0) {  // NaN terminal state (levelsets.py:1233): retire at once (DESIGN.md "NaN lanes")
        st.result = PMP_RESULT_NAN;
        for (int i = 0; i < NS; i++) y_end[i] = y[i];
        pad(1);
        return st;
    }

    R k[7][NS];
    field(y, k[0]);
    bool at_dtmin = false;

    while (r_val(tprev) < r_val(T1) && st.result == 0 && st.n_steps < o.max_steps) {
        const R dt = tnext - tprev;
        R ys[NS], y1[NS], yerr[NS];
        const int last = tb.stages - 1;
        for (int i = 1; i < tb.stages; i++) {
            for (int q = 0; q < NS; q++) {
                R acc = R(tb.a[i][0]) * k[0][q];
                for (int j = 1; j < i; j++) acc = acc + R(tb.a[i][j]) * k[j][q];
                ys[q] = y[q] + dt * acc;
            }
            field(ys, k[i]);
            if (tb.fsal && i == last)
                for (int q = 0; q < NS; q++) y1[q] = ys[q];
        }
        if (!tb.fsal) {
            for (int q = 0; q < NS; q++) {
                R acc = R(tb.b[0]) * k[0][q];
                for (int j = 1; j < tb.stages; j++) acc = acc + R(tb.b[j]) * k[j][q];
                y1[q] = y[q] + dt * acc;
            }
        }
        bool keep = true;
        bool err_nonfinite = false;
        R next_t0, next_t1;
        if (o.adaptive && tb.embedded) {
            for (int q = 0; q < NS; q++) {
                R acc = R(tb.berr[0]) * k[0][q];
                for (int j = 1; j < tb.stages; j++) acc = acc + R(tb.berr[j]) * k[j][q];
                yerr[q] = dt * acc;
                if (r_isnan(yerr[q])) yerr[q] = inf;  // integrate.py: nan error -> inf
            }
            // per-leaf NaN guard of the scale (leaves: x, t, v, vx)
            auto leaf_of = [&](int q) { return q < NX ? 0 : (q == IT ? 1 : (q == IV ? 2 : (q < NB ? 3 : 4))); };
            bool leaf_nan[5] = {false, false, false, false, false};
            for (int q = 0; q < NS; q++) leaf_nan[leaf_of(q)] = leaf_nan[leaf_of(q)] || r_isnan(y1[q]);
            R sumsq = R(0);
            // value mode: the missing solver's state was y=[x, lambda, t] (rrt_sampler.py:62), v is
            // the independent variable and not a leaf of the error norm
            const int n_norm = value_mode ? NS - 1 : NS;
            for (int q = 0; q < NS; q++) {
                if (value_mode && q == IV) continue;
                R y1c = leaf_nan[leaf_of(q)] ? y[q] : y1[q];
                R sc = R(o.atol) + r_max(r_abs(y[q]), r_abs(y1c)) * R(o.rtol);
                R e = yerr[q] / sc;
                sumsq = sumsq + e * e;
            }
            R scaled = r_sqrt(sumsq / R(n_norm));  // rms_norm over the whole pytree (incl. the t leaf)
            keep = (r_val(scaled) < 1.0) || at_dtmin;
            err_nonfinite = !(r_val(scaled) < std::numeric_limits<double>::infinity());
            R inv = R(1) / scaled;
            R factor = R(o.safety) * r_pow(inv, R(1.0 / tb.order));
            R fmin = keep ? R(1) : R(o.factormin);
            factor = r_clip(factor, fmin, R(o.factormax));
            R dtn = (tnext - tprev) * factor;
            if (o.dtmax > 0) dtn = r_min(dtn, R(o.dtmax));
            if (o.dtmin > 0) {
                if (!o.force_dtmin && r_val(dtn) < o.dtmin) st.result = PMP_RESULT_DT_MIN;
                at_dtmin = r_val(dtn) <= o.dtmin;
                dtn = r_max(dtn, R(o.dtmin));
            } else {
                at_dtmin = false;
            }
            next_t0 = keep ? tnext : tprev;
            next_t1 = next_t0 + dtn;
        } else {  // ConstantStepSize: next interval has the same length as the last one
            next_t0 = tnext;
            next_t1 = tnext + (tnext - tprev);
        }
        // integrate.py: tprev = min(tprev, t1); tnext = _clip_to_end(tprev, tnext, t1, keep_step)
        tprev = r_min(next_t0, T1);
        if (r_val(next_t1) > r_val(T1 - clip_tol))
            next_t1 = keep ? T1 : tprev + R(0.5) * (T1 - tprev);
        tnext = next_t1;

        if (keep) {
            for (int q = 0; q < NS; q++) y[q] = y1[q];
            if (tb.fsal) for (int q = 0; q < NS; q++) k[0][q] = k[last][q];
            else field(y, k[0]);
        }
        st.n_steps += 1;
        if (keep) {
            st.n_accepted += 1;
            save(st.n_accepted, tprev);
        }
        if (st.result == 0 && o.event == PMP_EVENT_VALUE_GE && r_val(y[IV]) >= o.event_level)
            st.result = PMP_RESULT_EVENT;
        if (VXX && st.result == 0 && o.event == PMP_EVENT_VXX_NORM) {  // np.linalg.norm(state.y['vxx']) > vxx_max_norm
            double fro = 0;
            for (int q = NB; q < NS; q++) fro += r_val(y[q]) * r_val(y[q]);
            if (std::sqrt(fro) > o.event_level) st.result = PMP_RESULT_EVENT;
        }
        if (keep && st.result == 0) {
            // a force-accepted step (at dtmin) whose scaled error is not finite, or a NaN state:
            // diffrax would idle (every later attempt rejected) until max_steps; we retire at once
            bool bad = r_isnan(tnext) || err_nonfinite;
            for (int q = 0; q < NS; q++) bad = bad || r_isnan(y[q]);
            if (bad) st.result = PMP_RESULT_NAN;
        }
    }
    if (st.result == 0 && r_val(tprev) < r_val(T1)) st.result = PMP_RESULT_MAX_STEPS;
    for (int i = 0; i < NS; i++) y_end[i] = y[i];
    pad(st.n_accepted + 1);
    return st;
}


// =================================================================================================
// Forward closed-loop simulation under the linear costate law lambda(x) = P (x - x_eq):
//   levelsets.py:818-836 forward_sim_lqr   (x' = f(x, u_star_2d(x, P x)), Tsit5 + PIDController, t: 0 -> T)
//   eval_utils.py:48-99  closed_loop_eval_general  (one more state records the cost, cost' = l(x, u*))
// State z = (x, cost); the rms error norm runs over the x leaf only (forward_sim_lqr has no cost state; the
// constant-step evaluation has no error control).  Same diffrax step-control flow as solve_one.
// Optional event: 1/2 dx'P dx <= event_level (PMP_EVENT_LQR_VALUE_LE).
// =================================================================================================
template <class R> struct ForwardSink {
    R* ts = nullptr; R* x = nullptr; R* cost = nullptr;
};

// costate law of the LQR value function: lambda = P (x - x_eq), value 1/2 dx'P dx, no spread
template <class R, int NX> struct LinearLaw {
    const double* P; const double* x_eq;
    void eval(const R* x, R* lam, R& vmean, R& vstd) const {
        R v2 = R(0);
        for (int r = 0; r < NX; r++) {
            R a = R(0);
            for (int q = 0; q < NX; q++) a = a + R(P[r * NX + q]) * (x[q] - R(x_eq ? x_eq[q] : 0.0));
            lam[r] = a;
            v2 = v2 + a * (x[r] - R(x_eq ? x_eq[r] : 0.0));
        }
        vmean = R(0.5) * v2;
        vstd = R(0);
    }
    bool inside(R vmean, R, double level) const { return r_val(vmean) <= level; }
};

// costate law of a value-network ensemble (levelsets.py:726-778,838-933; nn_utils.py:119-199), weights in HOST memory with
// the layout of pmp_nn_costate_t: V_e(x) = v_std MLP_e((x - x_mean)/x_std) + v_mean, MLP = Dense softplus ... Dense(1);
// lambda = grad of the ensemble mean (forward + backward sweep per member), vmean / vstd = mean / population std of V_e.
template <class R, int NX> struct NnLaw {
    const pmp_nn_costate_t* nn;
    static R softplus(R a) { return r_max(a, R(0)) + R(std::log1p(std::exp(-std::fabs(r_val(a))))); }  // logaddexp(a, 0)
    static R sigmoid(R a) { return R(1) / (R(1) + R(std::exp(-r_val(a)))); }
    void eval(const R* x, R* lam, R& vmean, R& vstd) const {
        const int W = nn->width, L = nn->n_hidden_layers, E = nn->n_nets;
        const int stride = NX * W + W + (L - 1) * (W * W + W) + W + 4;
        const R* w = (const R*)nn->weights;
        std::vector<R> h(W), a(W), g(W), gn(W), sg((size_t)L * W);
        R xn[NX], gsum[NX], ve[16];
        for (int q = 0; q < NX; q++) { xn[q] = (x[q] - R(nn->x_mean[q])) / R(nn->x_std[q]); gsum[q] = R(0); }
        for (int e = 0; e < E; e++) {
            const R* p = w + (size_t)e * stride;
            for (int j = 0; j < W; j++) {
                R acc = p[NX * W + j];
                for (int q = 0; q < NX; q++) acc = acc + xn[q] * p[q * W + j];
                h[j] = softplus(acc); sg[j] = sigmoid(acc);
            }
            const R* pl = p + NX * W + W;
            for (int l = 1; l < L; l++, pl += W * W + W) {
                for (int j = 0; j < W; j++) {
                    R acc = pl[W * W + j];
                    for (int i = 0; i < W; i++) acc = acc + h[i] * pl[i * W + j];
                    a[j] = acc;
                }
                for (int j = 0; j < W; j++) { h[j] = softplus(a[j]); sg[(size_t)l * W + j] = sigmoid(a[j]); }
            }
            R v = pl[W];
            for (int j = 0; j < W; j++) { v = v + h[j] * pl[j]; g[j] = pl[j] * sg[(size_t)(L - 1) * W + j]; }
            for (int l = L - 1; l >= 1; l--) {
                pl -= W * W + W;
                for (int i = 0; i < W; i++) {
                    R d = R(0);
                    for (int j = 0; j < W; j++) d = d + pl[i * W + j] * g[j];
                    gn[i] = d * sg[(size_t)(l - 1) * W + i];
                }
                g = gn;
            }
            for (int q = 0; q < NX; q++) {
                R d = R(0);
                for (int j = 0; j < W; j++) d = d + p[q * W + j] * g[j];
                gsum[q] = gsum[q] + d * (R(nn->v_std) / R(nn->x_std[q]));
            }
            ve[e] = R(nn->v_std) * v + R(nn->v_mean);
        }
        R m = R(0);
        for (int e = 0; e < E; e++) m = m + ve[e];
        m = m / R(E);
        R var = R(0);
        for (int e = 0; e < E; e++) var = var + (ve[e] - m) * (ve[e] - m);
        vmean = m;
        vstd = r_sqrt(var / R(E));
        for (int q = 0; q < NX; q++) lam[q] = gsum[q] / R(E);
    }
    // levelsets.py:897-913
    bool inside(R vmean, R vstd, double level) const {
        return r_val(vmean) + nn->n_sigma * r_val(vstd) <= level && r_val(vstd) <= nn->sigma_max;
    }
};

template <class R, class Sys, class Law>
TrajStats forward_one(const Sys& sys, const pmp_opts_t& o, const Law& law, const R* x0, R* z_end, ForwardSink<R> sink) {
    constexpr int NX = Sys::NX, NZ = NX + 1, NS = 2 * NX + 2, IV = NX + 1;
    const Tableau& tb = *tableau_for(o.method);
    const int S1 = o.max_steps + 1;
    const R inf = r_inf<R>();
    TrajStats st;
    R z[NZ];
    for (int i = 0; i < NX; i++) z[i] = x0[i];
    z[NX] = R(0);
    const R dir = (o.t0 < o.t1) ? R(1) : R(-1);
    const R T0 = R(o.t0) * dir, T1 = R(o.t1) * dir;
    R tprev = T0, tnext = r_min(T0 + R(o.dt0) * dir, T1);
    const R clip_tol = std::is_same<R, float>::value ? R(1e-6) : R(1e-10);
    R vlqr = R(0), vsig = R(0);
    auto field = [&](const R* zz, R* k, R& v, R& sd) {
        R y[NS], dy[NS];
        for (int i = 0; i < NS; i++) y[i] = R(0);
        for (int r = 0; r < NX; r++) y[r] = zz[r];
        law.eval(zz, y + NX + 2, v, sd);
        sys.rhs(y, dy);
        for (int i = 0; i < NX; i++) k[i] = dy[i] * dir;
        k[NX] = -dy[IV] * dir;
    };
    auto save = [&](int slot) {
        if (slot >= S1) return;
        if (sink.ts) sink.ts[slot] = tprev * dir;
        if (sink.x) for (int i = 0; i < NX; i++) sink.x[slot * NX + i] = z[i];
        if (sink.cost) sink.cost[slot] = z[NX];
    };
    save(0);
    bool bad0 = false;
    for (int i = 0; i < NX; i++) bad0 = bad0 || r_isnan(z[i]);
    if (bad0) st.result = PMP_RESULT_NAN;
    R k[7][NZ];
    field(z, k[0], vlqr, vsig);
    bool at_dtmin = false;
    while (r_val(tprev) < r_val(T1) && st.result == 0 && st.n_steps < o.max_steps) {
        const R dt = tnext - tprev;
        R zs[NZ], z1[NZ], v1 = R(0), s1 = R(0);
        const int last = tb.stages - 1;
        for (int i = 1; i < tb.stages; i++) {
            for (int q = 0; q < NZ; q++) {
                R acc = R(tb.a[i][0]) * k[0][q];
                for (int j = 1; j < i; j++) acc = acc + R(tb.a[i][j]) * k[j][q];
                zs[q] = z[q] + dt * acc;
            }
            field(zs, k[i], v1, s1);
            if (tb.fsal && i == last)
                for (int q = 0; q < NZ; q++) z1[q] = zs[q];
        }
        if (!tb.fsal) {
            for (int q = 0; q < NZ; q++) {
                R acc = R(tb.b[0]) * k[0][q];
                for (int j = 1; j < tb.stages; j++) acc = acc + R(tb.b[j]) * k[j][q];
                z1[q] = z[q] + dt * acc;
            }
        }
        bool keep = true, err_nonfinite = false;
        R next_t0, next_t1;
        if (o.adaptive && tb.embedded) {
            R sumsq = R(0);
            bool nan1 = false;
            for (int q = 0; q < NX; q++) nan1 = nan1 || r_isnan(z1[q]);
            for (int q = 0; q < NX; q++) {
                R acc = R(tb.berr[0]) * k[0][q];
                for (int j = 1; j < tb.stages; j++) acc = acc + R(tb.berr[j]) * k[j][q];
                R e = dt * acc;
                if (r_isnan(e)) e = inf;
                R sc = R(o.atol) + r_max(r_abs(z[q]), r_abs(nan1 ? z[q] : z1[q])) * R(o.rtol);
                R r = e / sc;
                sumsq = sumsq + r * r;
            }
            R scaled = r_sqrt(sumsq / R(NX));
            keep = (r_val(scaled) < 1.0) || at_dtmin;
            err_nonfinite = !(r_val(scaled) < std::numeric_limits<double>::infinity());
            R factor = R(o.safety) * r_pow(R(1) / scaled, R(1.0 / tb.order));
            factor = r_clip(factor, keep ? R(1) : R(o.factormin), R(o.factormax));
            R dtn = dt * factor;
            if (o.dtmax > 0) dtn = r_min(dtn, R(o.dtmax));
            if (o.dtmin > 0) {
                if (!o.force_dtmin && r_val(dtn) < o.dtmin) st.result = PMP_RESULT_DT_MIN;
                at_dtmin = r_val(dtn) <= o.dtmin;
                dtn = r_max(dtn, R(o.dtmin));
            } else {
                at_dtmin = false;
            }
            next_t0 = keep ? tnext : tprev;
            next_t1 = next_t0 + dtn;
        } else {
            next_t0 = tnext;
            next_t1 = tnext + dt;
        }
        tprev = r_min(next_t0, T1);
        if (r_val(next_t1) > r_val(T1 - clip_tol)) next_t1 = keep ? T1 : tprev + R(0.5) * (T1 - tprev);
        tnext = next_t1;
        st.n_steps += 1;
        if (keep) {
            for (int q = 0; q < NZ; q++) z[q] = z1[q];
            if (tb.fsal) {
                for (int q = 0; q < NZ; q++) k[0][q] = k[last][q];
                vlqr = v1; vsig = s1;
            } else {
                field(z, k[0], vlqr, vsig);
            }
            st.n_accepted += 1;
            save(st.n_accepted);
        }
        if (st.result == 0 && o.event != PMP_EVENT_NONE && law.inside(vlqr, vsig, o.event_level)) st.result = PMP_RESULT_EVENT;
        if (keep && st.result == 0) {
            bool bad = err_nonfinite;
            for (int q = 0; q < NZ; q++) bad = bad || r_isnan(z[q]);
            if (bad) st.result = PMP_RESULT_NAN;
        }
    }
    if (st.result == 0 && r_val(tprev) < r_val(T1)) st.result = PMP_RESULT_MAX_STEPS;
    for (int i = 0; i < NZ; i++) z_end[i] = z[i];
    for (int slot = st.n_accepted + 1; slot < S1; slot++) {
        if (sink.ts) sink.ts[slot] = inf * dir;
        if (sink.x) for (int i = 0; i < NX; i++) sink.x[slot * NX + i] = inf;
        if (sink.cost) sink.cost[slot] = inf;
    }
    return st;
}

// =================================================================================================
// Dense output: diffrax Solution.evaluate on one step.  Recomputes the stage derivatives of the
// step [t0, t1] that starts at the saved node y0 and evaluates the interpolant (Tsit5: Tsitouras'
// free interpolant, constants as in SURVEY.md 8(c); Dopri5: Hairer's contd5; RK4: cubic Hermite).
// mode 0: evaluate at the independent variable `target`; mode 1: at the point where v == target.
// =================================================================================================
template <class R> void interp_weights(int method, R th, R* w) {
    const R t2 = th * th;
    if (method == PMP_METHOD_TSIT5) {
        w[0] = R(-1.0530884977290216) * th * (th - R(1.3299890189751412)) * (t2 - R(1.4364028541716351) * th + R(0.7139816917074209));
        w[1] = R(0.1017) * t2 * (t2 - R(2.1966568338249754) * th + R(1.2949852507374631));
        w[2] = R(2.490627285651252793) * t2 * (t2 - R(2.38535645472061657) * th + R(1.57803468208092486));
        w[3] = R(-16.54810288924490272) * (th - R(1.21712927295533244)) * (th - R(0.61620406037800089)) * t2;
        w[4] = R(47.37952196281928122) * (th - R(1.203071208372362603)) * (th - R(0.658047292653547382)) * t2;
        w[5] = R(-34.87065786149660974) * (th - R(1.2)) * (th - R(2.0 / 3.0)) * t2;
        w[6] = R(2.5) * (th - R(1.0)) * (th - R(0.6)) * t2;
    } else {
        const R om = th - R(1), a = t2 * (R(3) - R(2) * th), c = t2 * om * om;
        w[0] = a * R(35.0 / 384) + th * om * om - c * R(5) * (R(2558722523.0) - R(31403016.0) * th) / R(11282082432.0);
        w[1] = R(0);
        w[2] = a * R(500.0 / 1113) + c * R(100) * (R(882725551.0) - R(15701508.0) * th) / R(32700410799.0);
        w[3] = a * R(125.0 / 192) - c * R(25) * (R(443332067.0) - R(31403016.0) * th) / R(1880347072.0);
        w[4] = a * R(-2187.0 / 6784) + c * R(32805) * (R(23143187.0) - R(3489224.0) * th) / R(199316789632.0);
        w[5] = a * R(11.0 / 84) - c * R(55) * (R(29972135.0) - R(7076736.0) * th) / R(822651844.0);
        w[6] = t2 * om + c * R(10) * (R(7414447.0) - R(829305.0) * th) / R(29380423.0);
    }
}

template <class R, class Sys>
void interp_one(const Sys& sys, const pmp_opts_t& o, const R* y0, R t1, R target, int mode, R* out) {
    constexpr int NX = Sys::NX, NS = 2 * NX + 2, IT = NX, IV = NX + 1;
    const Tableau& tb = *tableau_for(o.method);
    const bool value_mode = o.indep == PMP_INDEP_VALUE;
    const int IC = value_mode ? IV : IT;  // the row that holds the independent variable
    const R h = t1 - y0[IC];
    auto field = [&](const R* yy, R* k) {  // forward-time field (signed h carries the direction)
        sys.rhs(yy, k);
        if (value_mode) {
            R vdot = k[IV];
            for (int i = 0; i < NS; i++) k[i] = k[i] / vdot;
            k[IV] = R(1);
        }
    };
    R k[8][NS], ys[NS], y1[NS];
    field(y0, k[0]);
    for (int i = 1; i < tb.stages; i++) {
        for (int q = 0; q < NS; q++) {
            R acc = y0[q];
            for (int j = 0; j < i; j++) acc = acc + (h * R(tb.a[i][j])) * k[j][q];
            ys[q] = acc;
        }
        field(ys, k[i]);
    }
    for (int q = 0; q < NS; q++) {
        R acc = y0[q];
        for (int j = 0; j < tb.stages; j++) acc = acc + (h * R(tb.b[j])) * k[j][q];
        y1[q] = acc;
    }
    if (!tb.fsal) field(y1, k[tb.stages]);
    auto eval = [&](R th, int q) -> R {
        if (tb.fsal) {
            R w[7];
            interp_weights<R>(o.method, th, w);
            R acc = R(0);
            for (int j = 0; j < 7; j++) acc = acc + w[j] * k[j][q];
            return y0[q] + h * acc;
        }
        const R d = y1[q] - y0[q];
        return (R(1) - th) * y0[q] + th * y1[q] +
               th * (th - R(1)) * ((R(1) - R(2) * th) * d + (th - R(1)) * h * k[0][q] + th * h * k[tb.stages][q]);
    };
    R th;
    if (mode == 0) {
        th = (target - y0[IC]) / h;
    } else {
        R lo = R(0), hi = R(1);
        const R f0 = y0[IV] - target, f1 = eval(R(1), IV) - target;
        const bool ok = (r_val(f0) <= 0 && r_val(f1) >= 0) || (r_val(f0) >= 0 && r_val(f1) <= 0);
        for (int it = 0; it < (std::is_same<R, float>::value ? 26 : 54); it++) {
            const R mid = R(0.5) * (lo + hi);
            const R fm = eval(mid, IV) - target;
            const bool same = (r_val(fm) >= 0) == (r_val(f0) >= 0);
            if (same) lo = mid; else hi = mid;
        }
        th = ok ? R(0.5) * (lo + hi) : r_nan<R>();
    }
    for (int q = 0; q < NS; q++) out[q] = eval(th, q);
    out[IC] = y0[IC] + th * h;
}

}  // namespace pmpo
