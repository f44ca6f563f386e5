// pmp_systems.cuh -- plug-in dynamics as compile-time device functors.
//
// Each system supplies the three pieces the reference obtains from jax autodiff of the user's
// f(x,u), l(x,u) (pontryagin_utils.py:35-38, :438-439):
//     ustar : u* = argmin_{u in box} H(x,u,lambda)           (u_star_2d :11-216 / u_star_new :532-562)
//     rhs   : forward-time f_extended = (f(x,u*), -l(x,u*), -dH/dx|_{u*})      (:418-458)
// hand-derived in closed form.  Constants that do not depend on (x, lambda) are folded on the host
// into a Consts record that travels in the kernel parameter (constant) bank, so the FFMAs read
// them as c[][] operands instead of occupying registers.
//
// Vector layout used by the integrator: x[NX], lam[NX] and one scalar (v) are separate arguments.
#pragma once
#include "pmp_common.cuh"

namespace pmp {

// =================================================================================================
// flat (planar) quadrotor -- flatquad_landing_experiment.py:269-329
//   x = (px, py, Phi, vx, vy, omega), u = (Fl, Fr) in [0, umax]^2
//   f = (vx, vy, omega, -sin(Phi) S/m, cos(Phi) S/m - g, D r/I),  S = Fl+Fr, D = Fr-Fl
//   l = q.(px^2+py^2, (cos-1)^2+sin^2, vx^2+vy^2, omega^2) + |(ax, ay, aw)|^2
//   H(u) is exactly quadratic in u with the CONSTANT Hessian
//       H_uu = 2 [[1/m^2 + r^2/I^2, 1/m^2 - r^2/I^2], [., .]]
// =================================================================================================
template <class R> struct Flatquad {
    static constexpr int NX = 6, NU = 2;
    struct Consts {
        R inv_m, g, r_I, two_g, two_g_m;
        R q_pos, q_ang, q_vel, q_om, two_qpos, two_qang, two_qvel, two_qom;
        R h00, h01, hh;    // H_uu diagonal, off-diagonal, half diagonal
        R hi00, hi01;      // inverse of H_uu
        R lo0, lo1, hi0, hi1;
        // edge k of the box = segment {s a_k + (1-s) b_k}, a_k = vertex k, b_k = vertex k-1 of
        // (lo,lo),(lo,hi),(hi,hi),(hi,lo) (pontryagin_utils.py:78-86).  Along the one moving axis:
        //   s = clip(eA*H_u[axis] + eB, 0, 1),   u[axis] = eb + s*ed
        R eA[4], eB[4], ed[4], eb[4];
        // KKT form: edge minimiser coordinate = clip(kN*H_u[axis] + kQ[k], lo, hi); its multiplier
        // = +-(h01*coordinate + H_u[other axis] + kM[k])
        R kN, kQ[4], kM[4];
    };

    static __host__ Consts make(const pmp_problem_t& p) {
        Consts c;
        const double m = p.params[0], g = p.params[1], r = p.params[2], I = p.params[3];
        c.inv_m = R(1.0 / m); c.g = R(g); c.r_I = R(r / I); c.two_g = R(2 * g); c.two_g_m = R(2 * g / m);
        c.q_pos = R(p.params[4]); c.q_ang = R(p.params[5]); c.q_vel = R(p.params[6]); c.q_om = R(p.params[7]);
        c.two_qpos = R(2 * p.params[4]); c.two_qang = R(2 * p.params[5]);
        c.two_qvel = R(2 * p.params[6]); c.two_qom = R(2 * p.params[7]);
        const double d = 2 * (1 / (m * m) + (r * r) / (I * I)), o = 2 * (1 / (m * m) - (r * r) / (I * I));
        const double det = d * d - o * o;
        c.h00 = R(d); c.h01 = R(o); c.hh = R(0.5 * d); c.hi00 = R(d / det); c.hi01 = R(-o / det);
        c.lo0 = R(p.u_lo[0]); c.lo1 = R(p.u_lo[1]); c.hi0 = R(p.u_hi[0]); c.hi1 = R(p.u_hi[1]);
        const double hull[4][2] = {{p.u_lo[0], p.u_lo[1]}, {p.u_lo[0], p.u_hi[1]},
                                   {p.u_hi[0], p.u_hi[1]}, {p.u_hi[0], p.u_lo[1]}};
        const double H[2][2] = {{d, o}, {o, d}};
        for (int k = 0; k < 4; k++) {
            const double* a = hull[k];
            const double* b = hull[(k + 3) % 4];
            const int ax = (a[0] != b[0]) ? 0 : 1;
            const double dd = a[ax] - b[ax];
            const double gk = b[0] * H[0][ax] + b[1] * H[1][ax];  // (b' H_uu)[axis]
            const double inv_den = 1.0 / (dd * H[ax][ax] * dd);
            c.eA[k] = R(-dd * inv_den);
            c.eB[k] = R(-gk * dd * inv_den);
            c.ed[k] = R(dd);
            c.eb[k] = R(b[ax]);
            // the fixed coordinate of edge k is a[1-ax]: minimiser of H along the edge in the moving
            // coordinate w: d*w + o*fixed + Hu[ax] = 0; multiplier: d*fixed + o*w + Hu[1-ax]
            const double fixed = a[1 - ax];
            c.kQ[k] = R(-o * fixed / d);
            c.kM[k] = R(d * fixed);
        }
        c.kN = R(-1.0 / d);
        return c;
    }

    // 1/2 w'H_uu w + g'w
    static __device__ __forceinline__ R quad(const Consts& c, R w0, R w1, R g0, R g1) {
        using M = Math<R>;
        R t = M::fma(c.hh, w0, g0);
        t = M::fma(c.h01, w1, t);
        R r1 = M::fma(c.hh, w1, g1);
        return M::fma(w0, t, w1 * r1);
    }

    // u_star_2d with sin/cos of Phi already known.
    // MODE & 1 (literal only): the caller guarantees all 32 lanes are converged, so the edge search
    //          is skipped when no lane leaves the box (warp-uniform branch).
    // MODE & 2: KKT selection among the edge minimisers (PMP_FLAG_USTAR_LITERAL clear), branch-free.
    template <int MODE>
    static __device__ __forceinline__ void ustar_sc(const Consts& c, const R* lam, R s, R co, R& u0, R& u1) {
        constexpr bool WARP_VOTE = (MODE & 1) != 0;
        using M = Math<R>;
        // H_u at u=0 (pontryagin_utils.py:37)
        const R base = M::fma(M::fma(lam[4], co, -(lam[3] * s)), c.inv_m, -(c.two_g_m * co));
        const R rot = lam[5] * c.r_I;
        const R Hu0 = base - rot, Hu1 = base + rot;
        // unconstrained minimiser solve(H_uu, -H_u)  (:41)
        u0 = -M::fma(c.hi00, Hu0, c.hi01 * Hu1);
        u1 = -M::fma(c.hi01, Hu0, c.hi00 * Hu1);
        // :129  outside the box?  NaN compares false => treated as inside => NaN u propagates
        const R viol = M::max(M::max(c.lo0 - u0, c.lo1 - u1), M::max(u0 - c.hi0, u1 - c.hi1));
        const bool outside = viol > R(0);
        if (MODE & 2) {
            // minimisers of H along the four edges (bottom u1=lo1, left u0=lo0, top u1=hi1, right
            // u0=hi0 -- the reference's candidate order, pontryagin_utils.py:78-86), each the clipped
            // root of the tangential derivative
            const R xb = M::min(M::max(M::fma(c.kN, Hu0, c.kQ[0]), c.lo0), c.hi0);
            const R yl = M::min(M::max(M::fma(c.kN, Hu1, c.kQ[1]), c.lo1), c.hi1);
            const R xt = M::min(M::max(M::fma(c.kN, Hu0, c.kQ[2]), c.lo0), c.hi0);
            const R yr = M::min(M::max(M::fma(c.kN, Hu1, c.kQ[3]), c.lo1), c.hi1);
            // KKT multiplier of the active bound = outward-pointing component of -grad H; the edge
            // minimiser that solves the box problem is the one whose multiplier is >= 0 (the others
            // are negative), so take the largest: first maximum wins, like the reference's argmin
            const R mb = M::fma(c.h01, xb, Hu1 + c.kM[0]);
            const R ml = M::fma(c.h01, yl, Hu0 + c.kM[1]);
            const R mt = -M::fma(c.h01, xt, Hu1 + c.kM[2]);
            const R mr = -M::fma(c.h01, yr, Hu0 + c.kM[3]);
            R bm = mb, bx = xb, by = c.lo1;
            bool gt = ml > bm;
            bm = gt ? ml : bm; bx = gt ? c.lo0 : bx; by = gt ? yl : by;
            gt = mt > bm;
            bm = gt ? mt : bm; bx = gt ? xt : bx; by = gt ? c.hi1 : by;
            gt = mr > bm;
            bx = gt ? c.hi0 : bx; by = gt ? yr : by;
            u0 = outside ? bx : u0;
            u1 = outside ? by : u1;
            return;
        }
        bool go = outside;
        if (WARP_VOTE) go = __any_sync(0xffffffffu, outside);
        if (go) {
            // :53-75 minimisers along the four edges; candidate k has one clipped moving coordinate
            const R s0 = M::min(M::max(M::fma(c.eA[0], Hu0, c.eB[0]), R(0)), R(1));
            const R s1 = M::min(M::max(M::fma(c.eA[1], Hu1, c.eB[1]), R(0)), R(1));
            const R s2 = M::min(M::max(M::fma(c.eA[2], Hu0, c.eB[2]), R(0)), R(1));
            const R s3 = M::min(M::max(M::fma(c.eA[3], Hu1, c.eB[3]), R(0)), R(1));
            const R cx[4] = {M::fma(s0, c.ed[0], c.eb[0]), c.lo0, M::fma(s2, c.ed[2], c.eb[2]), c.hi0};
            const R cy[4] = {c.lo1, M::fma(s1, c.ed[1], c.eb[1]), c.hi1, M::fma(s3, c.ed[3], c.eb[3])};
            // :88-135 pass 1: score with the quadratic model about u=0; first minimum wins
            R bx = cx[0], by = cy[0];
            R bh = quad(c, cx[0], cy[0], Hu0, Hu1);
#pragma unroll
            for (int k = 1; k < 4; k++) {
                const R h = quad(c, cx[k], cy[k], Hu0, Hu1);
                const bool lt = h < bh;
                bh = lt ? h : bh; bx = lt ? cx[k] : bx; by = lt ? cy[k] : by;
            }
            // :151-158 pass 2: re-expand about the winner and compare suboptimalities
            const R hn0 = M::fma(c.h00, bx, M::fma(c.h01, by, Hu0));
            const R hn1 = M::fma(c.h01, bx, M::fma(c.h00, by, Hu1));
            R fx = cx[0], fy = cy[0];
            R fh = quad(c, cx[0] - bx, cy[0] - by, hn0, hn1);
#pragma unroll
            for (int k = 1; k < 4; k++) {
                const R h = quad(c, cx[k] - bx, cy[k] - by, hn0, hn1);
                const bool lt = h < fh;
                fh = lt ? h : fh; fx = lt ? cx[k] : fx; fy = lt ? cy[k] : fy;
            }
            u0 = outside ? fx : u0;
            u1 = outside ? fy : u1;
        }
    }

    static __device__ __forceinline__ void ustar(const Consts& c, const R* x, const R* lam, R* u) {
        R s, co;
        Math<R>::sincos(x[2], &s, &co);
        ustar_sc<0>(c, lam, s, co, u[0], u[1]);  // literal two-pass form, no warp vote
    }
    template <int MODE>
    static __device__ __forceinline__ void ustar_mode(const Consts& c, const R* x, const R* lam, R* u) {
        R s, co;
        Math<R>::sincos(x[2], &s, &co);
        ustar_sc<MODE>(c, lam, s, co, u[0], u[1]);
    }

    // forward-time f_extended: xd = f(x,u*), vd = -l(x,u*), ld = -dH/dx
    template <int MODE>
    static __device__ __forceinline__ void rhs(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld) {
        using M = Math<R>;
        R s, co;
        M::sincos(x[2], &s, &co);
        R u0, u1;
        ustar_sc<MODE>(c, lam, s, co, u0, u1);
        const R Sm = (u0 + u1) * c.inv_m;
        const R ax = -(s * Sm);
        const R ay = M::fma(co, Sm, -c.g);
        const R aw = (u1 - u0) * c.r_I;
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = x[5]; xd[3] = ax; xd[4] = ay; xd[5] = aw;
        const R cm1 = co - R(1);
        R l = c.q_pos * M::fma(x[0], x[0], x[1] * x[1]);
        l = M::fma(c.q_ang, M::fma(cm1, cm1, s * s), l);
        l = M::fma(c.q_vel, M::fma(x[3], x[3], x[4] * x[4]), l);
        l = M::fma(c.q_om * x[5], x[5], l);
        l = M::fma(ax, ax, l);
        l = M::fma(ay, ay, l);
        l = M::fma(aw, aw, l);
        vd = -l;
        ld[0] = -(c.two_qpos * x[0]);
        ld[1] = -(c.two_qpos * x[1]);
        ld[2] = M::fma(Sm, M::fma(lam[3], co, lam[4] * s), -(s * M::fma(c.two_g, Sm, c.two_qang)));
        ld[3] = -M::fma(c.two_qvel, x[3], lam[0]);
        ld[4] = -M::fma(c.two_qvel, x[4], lam[1]);
        ld[5] = -M::fma(c.two_qom, x[5], lam[2]);
    }
};

// =================================================================================================
// "orbits" -- experiment_orbits.py:24-40.  f = [[u,-rho],[rho,u]] x, rho = |x|^2 - 1,
// l = rho^2 + w_dist |x-(0,1)|^2 + w_u u^2,  u* = clip(-(lam.x)/(2 w_u), U)  (u_star_new)
// =================================================================================================
template <class R> struct Orbits {
    static constexpr int NX = 2, NU = 1;
    struct Consts {
        R w_dist, w_u, two_w_dist, neg_inv_2wu, lo, hi;
    };
    static __host__ Consts make(const pmp_problem_t& p) {
        Consts c;
        c.w_dist = R(p.params[0]); c.w_u = R(p.params[1]); c.two_w_dist = R(2 * p.params[0]);
        c.neg_inv_2wu = R(-1.0 / (2 * p.params[1]));
        c.lo = R(p.u_lo[0]); c.hi = R(p.u_hi[0]);
        return c;
    }
    static __device__ __forceinline__ void ustar(const Consts& c, const R* x, const R* lam, R* u) {
        using M = Math<R>;
        u[0] = clipf(M::fma(lam[0], x[0], lam[1] * x[1]) * c.neg_inv_2wu, c.lo, c.hi);
    }
    template <int MODE>
    static __device__ __forceinline__ void ustar_mode(const Consts& c, const R* x, const R* lam, R* u) { ustar(c, x, lam, u); }
    template <int MODE>
    static __device__ __forceinline__ void rhs(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld) {
        using M = Math<R>;
        R u[1];
        ustar(c, x, lam, u);
        const R rho = M::fma(x[0], x[0], M::fma(x[1], x[1], R(-1)));
        const R e1 = x[1] - R(1);
        xd[0] = M::fma(u[0], x[0], -(rho * x[1]));
        xd[1] = M::fma(rho, x[0], u[0] * x[1]);
        R l = rho * rho;
        l = M::fma(c.w_dist, M::fma(x[0], x[0], e1 * e1), l);
        l = M::fma(c.w_u * u[0], u[0], l);
        vd = -l;
        const R x01 = R(2) * x[0] * x[1];
        R h0 = R(4) * rho * x[0];
        h0 = M::fma(c.two_w_dist, x[0], h0);
        h0 = M::fma(lam[0], u[0] - x01, h0);
        h0 = M::fma(lam[1], M::fma(R(2) * x[0], x[0], rho), h0);
        R h1 = R(4) * rho * x[1];
        h1 = M::fma(c.two_w_dist, e1, h1);
        h1 = M::fma(lam[0], -M::fma(R(2) * x[1], x[1], rho), h1);
        h1 = M::fma(lam[1], x01 + u[0], h1);
        ld[0] = -h0;
        ld[1] = -h1;
    }
};

// =================================================================================================
// double integrator with soft state constraints -- experiment_lqr_statepenalty.py:23-38
//   f = (x1, u),  l = |x|^2 + r_u u^2 + (max(0,x1-1)^2 + max(0,x0-1)^2)/a,  u* = clip(-lam1/(2 r_u), U)
// =================================================================================================
template <class R> struct LqrStatePenalty {
    static constexpr int NX = 2, NU = 1;
    struct Consts {
        R inv_a, two_inv_a, r_u, neg_inv_2ru, lo, hi;
    };
    static __host__ Consts make(const pmp_problem_t& p) {
        Consts c;
        c.inv_a = R(1.0 / p.params[0]); c.two_inv_a = R(2.0 / p.params[0]); c.r_u = R(p.params[1]);
        c.neg_inv_2ru = R(-1.0 / (2 * p.params[1]));
        c.lo = R(p.u_lo[0]); c.hi = R(p.u_hi[0]);
        return c;
    }
    static __device__ __forceinline__ void ustar(const Consts& c, const R* x, const R* lam, R* u) {
        u[0] = clipf(lam[1] * c.neg_inv_2ru, c.lo, c.hi);
    }
    template <int MODE>
    static __device__ __forceinline__ void ustar_mode(const Consts& c, const R* x, const R* lam, R* u) { ustar(c, x, lam, u); }
    template <int MODE>
    static __device__ __forceinline__ void rhs(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld) {
        using M = Math<R>;
        R u[1];
        ustar(c, x, lam, u);
        const R z0 = nan_max(R(0), x[0] - R(1));
        const R z1 = nan_max(R(0), x[1] - R(1));
        xd[0] = x[1];
        xd[1] = u[0];
        R l = M::fma(x[0], x[0], x[1] * x[1]);
        l = M::fma(c.r_u * u[0], u[0], l);
        l = M::fma(c.inv_a, M::fma(z1, z1, z0 * z0), l);
        vd = -l;
        ld[0] = -M::fma(c.two_inv_a, z0, R(2) * x[0]);
        ld[1] = -(M::fma(c.two_inv_a, z1, R(2) * x[1]) + lam[0]);
    }
};

}  // namespace pmp
