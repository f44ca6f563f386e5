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
#include "pmp_ustar2d.cuh"
#ifdef PMP_PLUGIN
#include PMP_GENERATED_HEADER  // template <class R> struct pmp::Generated, PMP_GEN_MINB32 / 64 (codegen.py)
#endif

namespace pmp {

// ---- storage of the value Hessian V_xx (PMP_FLAG_VXX) -------------------------------------------------
// VXX = 1: all nx*nx entries, row-major (the reference's layout).  VXX = 2 (PMP_FLAG_VXX_SYMMETRIC): the
// caller states that the terminal V_xx is symmetric; V_xx^T obeys the same Riccati equation (gx and flam
// are symmetric, glam = -fx^T), so the solution stays symmetric and only the upper triangle is stored and
// integrated (21 of 36 entries for nx = 6); off-diagonal entries count twice in the norms.
template <int NX, int VXX> struct VxxMap {
    static constexpr int NV = VXX == 1 ? NX * NX : (VXX == 2 ? NX * (NX + 1) / 2 : 0);
    // storage entry of V[i][j]
    __host__ __device__ static constexpr int ent(int i, int j) {
        if (VXX != 2) return i * NX + j;
        const int a = i < j ? i : j, b = i < j ? j : i;
        return a * NX - a * (a - 1) / 2 + (b - a);
    }
    // (i, j), i <= j for VXX = 2, of storage entry e, as the public row-major index i*NX + j
    __host__ __device__ static constexpr int pub(int e) {
        if (VXX != 2) return e;
        int i = 0, rem = e;
        while (rem >= NX - i) { rem -= NX - i; i++; }
        return i * NX + i + rem;
    }
    __host__ __device__ static constexpr bool offdiag(int e) { return VXX == 2 && (pub(e) / NX != pub(e) % NX); }
};

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
        R kNo;  // -h01/h00: slope of an edge minimiser in the fixed coordinate
        R w0, w1, kNow0, kNow1;                   // box widths hi-lo; kNo / width
        alignas(2 * sizeof(R)) R pNw[2], pLw[2];  // (kN/w0, kN/w1), (-lo0/w0, -lo1/w1)
        // vxx branch (rhs_jac): d(u0+u1)/dH_u and d(u1-u0)/d(lam5) of the unconstrained minimiser, kN r/I
        R jMS, jD3, kN_rI;
        // the same constants as PAIRS for the packed (FFMA2) form of the KKT selection: the two edges that move along
        // u0 (bottom, top) and the two that move along u1 (left, right) are evaluated together; pairs sit on 8-byte
        // boundaries so that they reach the FFMA2 as one 64-bit uniform-register operand
        alignas(2 * sizeof(R)) R pN[2];           // (kN, kN)
        alignas(2 * sizeof(R)) R pQ0[2], pQ1[2];  // (kQ[0], kQ[2]), (kQ[1], kQ[3])
        alignas(2 * sizeof(R)) R pM0[2], pM1[2];  // (kM[0], kM[2]), (kM[1], kM[3])
        alignas(2 * sizeof(R)) R pUa[2], pUb[2];  // (-hi00, -hi01), (-hi01, -hi00): (u0, u1) = Hu0 pUa + Hu1 pUb
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
        c.kNo = R(-o / d);
        {
            const double w0 = p.u_hi[0] - p.u_lo[0], w1 = p.u_hi[1] - p.u_lo[1];
            c.w0 = R(w0); c.w1 = R(w1); c.kNow0 = R(-o / d / w0); c.kNow1 = R(-o / d / w1);
            c.pNw[0] = R(-1.0 / d / w0); c.pNw[1] = R(-1.0 / d / w1);
            c.pLw[0] = R(-p.u_lo[0] / w0); c.pLw[1] = R(-p.u_lo[1] / w1);
        }
        c.jMS = R(-2.0 * (d - o) / det);
        c.jD3 = R(-2.0 * (r / I) * (d + o) / det);
        c.kN_rI = R(-(r / I) / d);
        c.pN[0] = c.pN[1] = c.kN;
        c.pQ0[0] = c.kQ[0]; c.pQ0[1] = c.kQ[2]; c.pQ1[0] = c.kQ[1]; c.pQ1[1] = c.kQ[3];
        c.pM0[0] = c.kM[0]; c.pM0[1] = c.kM[2]; c.pM1[0] = c.kM[1]; c.pM1[1] = c.kM[3];
        c.pUa[0] = -c.hi00; c.pUa[1] = -c.hi01; c.pUb[0] = -c.hi01; c.pUb[1] = -c.hi00;
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
        if (MODE & 4) {  // u_star_2d(smooth=True): pointwise evaluation only, through the run-time-H_uu routine
            ustar2d_box<R, 4>(Hu0, Hu1, c.h00, c.h01, c.h00, c.lo0, c.lo1, c.hi0, c.hi1, u0, u1);
            return;
        }
        if (MODE & 2) {
            // ---- KKT form (integrator default).  Cost model measured on B200: compares, selects and
            // min/max run on the half-rate ALU pipe and cost two issue cycles each, FMA-pipe instructions one, so this
            // path minimises ALU instructions.
            using V2 = typename M::V2;
            // unconstrained minimiser solve(H_uu, -H_u) (:41): (u0, u1) = Hu0 (-hi00, -hi01) + Hu1 (-hi01, -hi00)
            const V2 uu = M::fma2(Hu0, M::mk2(c.pUa[0], c.pUa[1]), M::mul2(Hu1, M::mk2(c.pUb[0], c.pUb[1])));
            // If uu leaves the box, the constrained minimiser lies on an edge whose bound uu violates: at the minimiser
            // u* the direction towards uu is a descent direction, so it must leave the box through an active bound,
            // i.e. uu is beyond that bound.  b = clip(uu) is that bound on every violated axis (and uu itself on the
            // others), so there are at most two candidates instead of the reference's four (pontryagin_utils.py:78-86;
            // the other two can never win):
            //   A: u1 = b1 fixed, u0 = minimiser along that edge = clip(-(Hu0 + h01 b1)/h00)   (only if axis 1 is violated)
            //   B: u0 = b0 fixed, u1 = clip(-(Hu1 + h01 b0)/h00)                               (only if axis 0 is violated)
            // With both axes violated the one whose KKT multiplier has the right sign wins: dH/du1 at A must push
            // outwards, i.e. have the sign of (uu1 - b1) negated.  Inside the box b == uu and neither test fires.
            const R b0 = M::min(M::max(uu.x, c.lo0), c.hi0), b1 = M::min(M::max(uu.y, c.lo1), c.hi1);
            const bool viol0 = b0 != uu.x, viol1 = b1 != uu.y;
            // clip(w, lo, hi) = lo + (hi-lo) sat((w-lo)/(hi-lo)): the saturation is a free modifier of the FMA that forms its
            // argument (fp32), against two ALU-pipe min/max for the direct form; the reference writes its edge minimisers
            // the same way (s = clip(., 0, 1), u = s a + (1-s) b, pontryagin_utils.py:73-75)
            const V2 t = M::fma2v(M::mk2(c.pNw[0], c.pNw[1]), M::mk2(Hu0, Hu1), M::mk2(c.pLw[0], c.pLw[1]));  // (-Hu/h00 - lo)/w
            const R u0A = M::fma(M::sat(M::fma(c.kNow0, b1, t.x)), c.w0, c.lo0);
            const R u1B = M::fma(M::sat(M::fma(c.kNow1, b0, t.y)), c.w1, c.lo1);
            const R gA = M::fma(c.h00, b1, M::fma(c.h01, u0A, Hu1));      // dH/du1 at candidate A
            const bool okA = gA * (b1 - uu.y) >= R(0);
            const bool useA = viol1 && (okA || !viol0);
            const bool useB = viol0 && !useA;
            u0 = useA ? u0A : b0;
            u1 = useB ? u1B : b1;
            return;
        }
        // unconstrained minimiser solve(H_uu, -H_u)  (:41)
        u0 = -M::fma(c.hi00, Hu0, c.hi01 * Hu1);
        u1 = -M::fma(c.hi01, Hu0, c.hi00 * Hu1);
        // :129  outside the box?  NaN compares false => treated as inside => NaN u propagates
        // (four chained compares -- FSETP folds the OR -- instead of four subtractions, three max and a compare)
        const bool outside = (u0 < c.lo0) | (u1 < c.lo1) | (u0 > c.hi0) | (u1 > c.hi1);
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
        // running cost and -dH/dx with the (x0,x1), (x3,x4) terms on packed pairs (the pairs of PairMap<Flatquad>)
        using V2 = typename M::V2;
        const V2 x01 = M::mk2(x[0], x[1]), x34 = M::mk2(x[3], x[4]), a34 = M::mk2(ax, ay);
        V2 acc = M::mul2v(M::mul2(c.q_pos, x01), x01);
        acc = M::fma2v(M::mul2(c.q_vel, x34), x34, acc);
        acc = M::fma2v(a34, a34, acc);
        R l = acc.x + acc.y;
        l = M::fma(c.q_ang, M::fma(cm1, cm1, s * s), l);
        l = M::fma(c.q_om * x[5], x[5], l);
        l = M::fma(aw, aw, l);
        vd = -l;
        const V2 l01 = M::mul2(-c.two_qpos, x01);
        const V2 l34 = M::fma2(-c.two_qvel, x34, M::mk2(-lam[0], -lam[1]));
        ld[0] = l01.x;
        ld[1] = l01.y;
        ld[2] = M::fma(Sm, M::fma(lam[3], co, lam[4] * s), -(s * M::fma(c.two_g, Sm, c.two_qang)));
        ld[3] = l34.x;
        ld[4] = l34.y;
        ld[5] = -M::fma(c.two_qom, x[5], lam[2]);
    }

    // ---- value-Hessian branch: algo_params['pontryagin_solver_vxx'], pontryagin_utils.py:460-467 ----
    //   vxx' = gx + glam vxx - vxx fx - vxx flam vxx,  (fx, flam), (gx, glam) = d pmp_rhs / d(x, lambda)
    // differentiated THROUGH u_star_2d as jax does: the active set (which candidate wins) is piecewise
    // constant, the winning candidate is differentiable in z = (Phi, lam3, lam4, lam5):
    //   interior: du = -H_uu^-1 dH_u;  on an edge: only the moving coordinate, d/dz clip(kN H_u[ax] + kQ)
    //   = kN dH_u[ax] strictly inside the edge and 0 in a corner.
    // The four 6x6 Jacobians are sparse and flam has rank <= 2 (flam = f_u du/dlam), so the record below
    // holds every entry that is not a structural 0 / +-1 / cost constant and one vxx' costs ~200 FMAs
    // instead of the 864 of four dense 6x6 products:
    //   fx  : [0][3] = [1][4] = [2][5] = 1, [3..5][2] = f32, f42, f52
    //   gx  : -diag(2q_pos, 2q_pos, ., 2q_vel, 2q_vel, 2q_om), [2][2] = g22
    //   glam: [3][0] = [4][1] = [5][2] = -1, [2][3..5] = l23, l24, l25
    //   flam[a][b] = al[a] sg[b] + be[a] dl[b]  (a, b in 3..5), al = (-s/m, c/m, 0), be = (0, 0, r/I)
    struct JacCoef {
        R f32, f42, f52, g22, l23, l24, l25;
        R al3, al4;  // al[3], al[4]
        R sg[3], dl[3];
    };

    // What the Jacobian coefficients depend on, 5 reals + the active set of u*: the integrator evaluates all
    // base stages first (registers) and keeps one JacPrim per stage for the V_xx stage loop that follows.
    struct JacPrim {
        R s, co, Sm, hxu2, hxx22;  // sin, cos, (u0+u1)/m, H_xu[2][.] = dH_u/dPhi, H_xx[2][2] at fixed u
        int act;                   // 0 interior, 1 / 2 edge with u0 / u1 moving freely, 3 corner
    };

    // the reals of a JacPrim as an array (the integrator keeps one register array per entry across the base stages)
    static constexpr int NPRIM = 5;
    static __device__ __forceinline__ void prim_store(const JacPrim& p, R* a, int& act) {
        a[0] = p.s; a[1] = p.co; a[2] = p.Sm; a[3] = p.hxu2; a[4] = p.hxx22;
        act = p.act;
    }
    static __device__ __forceinline__ void prim_load(JacPrim& p, const R* a, int act) {
        p.s = a[0]; p.co = a[1]; p.Sm = a[2]; p.hxu2 = a[3]; p.hxx22 = a[4];
        p.act = act;
    }

    // rhs<2> (KKT-form u*) that also reports the Jacobian primitives
    static __device__ __forceinline__ void rhs_prim(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld,
                                                    JacPrim& p) {
        using M = Math<R>;
        R s, co;
        M::sincos(x[2], &s, &co);
        const R base = M::fma(M::fma(lam[4], co, -(lam[3] * s)), c.inv_m, -(c.two_g_m * co));
        const R rot = lam[5] * c.r_I;
        const R Hu0 = base - rot, Hu1 = base + rot;
        R u0 = -M::fma(c.hi00, Hu0, c.hi01 * Hu1);
        R u1 = -M::fma(c.hi01, Hu0, c.hi00 * Hu1);
        const bool outside = (u0 < c.lo0) | (u1 < c.lo1) | (u0 > c.hi0) | (u1 > c.hi1);
        // same selection as ustar_sc<2>, additionally tracking which coordinate moves along the winning edge
        const R xb = M::min(M::max(M::fma(c.kN, Hu0, c.kQ[0]), c.lo0), c.hi0);
        const R yl = M::min(M::max(M::fma(c.kN, Hu1, c.kQ[1]), c.lo1), c.hi1);
        const R xt = M::min(M::max(M::fma(c.kN, Hu0, c.kQ[2]), c.lo0), c.hi0);
        const R yr = M::min(M::max(M::fma(c.kN, Hu1, c.kQ[3]), c.lo1), c.hi1);
        const R mb = M::fma(c.h01, xb, Hu1 + c.kM[0]);
        const R ml = M::fma(c.h01, yl, Hu0 + c.kM[1]);
        const R mt = -M::fma(c.h01, xt, Hu1 + c.kM[2]);
        const R mr = -M::fma(c.h01, yr, Hu0 + c.kM[3]);
        R bm = mb, bx = xb, by = c.lo1;
        bool axis1 = false;
        bool gt = ml > bm;
        bm = gt ? ml : bm; bx = gt ? c.lo0 : bx; by = gt ? yl : by; axis1 = gt ? true : axis1;
        gt = mt > bm;
        bm = gt ? mt : bm; bx = gt ? xt : bx; by = gt ? c.hi1 : by; axis1 = gt ? false : axis1;
        gt = mr > bm;
        bx = gt ? c.hi0 : bx; by = gt ? yr : by; axis1 = gt ? true : axis1;
        u0 = outside ? bx : u0;
        u1 = outside ? by : u1;
        const bool free_edge = axis1 ? (by > c.lo1 && by < c.hi1) : (bx > c.lo0 && bx < c.hi0);
        p.act = outside ? (free_edge ? (axis1 ? 2 : 1) : 3) : 0;

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
        const R lcs = M::fma(lam[3], co, lam[4] * s);
        ld[0] = -(c.two_qpos * x[0]);
        ld[1] = -(c.two_qpos * x[1]);
        ld[2] = M::fma(Sm, lcs, -(s * M::fma(c.two_g, Sm, c.two_qang)));
        ld[3] = -M::fma(c.two_qvel, x[3], lam[0]);
        ld[4] = -M::fma(c.two_qvel, x[4], lam[1]);
        ld[5] = -M::fma(c.two_qom, x[5], lam[2]);
        p.s = s; p.co = co; p.Sm = Sm;
        p.hxu2 = M::fma(c.two_g, s, -lcs) * c.inv_m;
        p.hxx22 = M::fma(co, M::fma(c.two_g, Sm, c.two_qang), -(Sm * M::fma(lam[4], co, -(lam[3] * s))));
    }

    static __device__ __forceinline__ void jac_from_prim(const Consts& c, const JacPrim& p, JacCoef& j) {
        using M = Math<R>;
        // d(u0+u1)/dz = MS dbase_z, d(u1-u0)/dz = MD dbase_z for z = Phi, lam3, lam4 (dH_u[0] = dH_u[1] = dbase);
        // z = lam5 (dH_u = -+ r/I): S3, D3
        const bool edge = p.act == 1 || p.act == 2, axis1 = p.act == 2;
        const R ek = edge ? c.kN : R(0), ekr = edge ? c.kN_rI : R(0);
        const R MS = p.act == 0 ? c.jMS : ek, MD = axis1 ? ek : -ek;
        const R S3 = axis1 ? ekr : -ekr, D3 = p.act == 0 ? c.jD3 : ekr;
        const R s = p.s, co = p.co, Sm = p.Sm, hxu2 = p.hxu2;
        // dbase_z: z = Phi equals H_xu[2] (symmetry of the Hessian of H), lam3: -s/m, lam4: c/m
        const R sm = s * c.inv_m, cm = co * c.inv_m;
        const R sum0 = MS * hxu2, dif0 = MD * hxu2;
        j.al3 = -sm; j.al4 = cm;
        j.sg[0] = -(MS * sm); j.sg[1] = MS * cm; j.sg[2] = S3;
        j.dl[0] = -(MD * sm); j.dl[1] = MD * cm; j.dl[2] = D3;
        j.f32 = -M::fma(co, Sm, sm * sum0);
        j.f42 = M::fma(cm, sum0, -(s * Sm));
        j.f52 = c.r_I * dif0;
        j.g22 = -M::fma(hxu2, sum0, p.hxx22);
        j.l23 = M::fma(co, Sm, -(hxu2 * j.sg[0]));
        j.l24 = M::fma(s, Sm, -(hxu2 * j.sg[1]));
        j.l25 = -(hxu2 * j.sg[2]);
    }

    static __device__ __forceinline__ void rhs_jac(const Consts& c, const R* x, const R* lam, R* xd, R& vd, R* ld,
                                                   JacCoef& j) {
        JacPrim p;
        rhs_prim(c, x, lam, xd, vd, ld, p);
        jac_from_prim(c, p, j);
    }

    // vxx' for V stored as VxxMap<6, VXX> says (VXX = 2: upper triangle in, upper triangle out); every entry
    // is handed to put(storage entry, value) as soon as it is complete
    template <int VXX, class Put>
    static __device__ __forceinline__ void vxx_dot(const Consts& c, const JacCoef& j, const R* V, Put&& put) {
        using M = Math<R>;
        using MP = VxxMap<6, VXX>;
        R Z1[6], Z2[6], G2[6];
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const R v3 = V[MP::ent(3, q)], v4 = V[MP::ent(4, q)], v5 = V[MP::ent(5, q)];
            Z1[q] = M::fma(j.sg[0], v3, M::fma(j.sg[1], v4, j.sg[2] * v5));
            Z2[q] = M::fma(j.dl[0], v3, M::fma(j.dl[1], v4, j.dl[2] * v5));
            G2[q] = M::fma(j.l23, v3, M::fma(j.l24, v4, j.l25 * v5));
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const R vi3 = V[MP::ent(i, 3)], vi4 = V[MP::ent(i, 4)], vi5 = V[MP::ent(i, 5)];
            const R W1 = M::fma(j.al3, vi3, j.al4 * vi4), W2 = c.r_I * vi5;
            const R vf2 = M::fma(vi3, j.f32, M::fma(vi4, j.f42, vi5 * j.f52));
#pragma unroll
            for (int q = (VXX == 2 ? i : 0); q < 6; q++) {
                // gx + glam V - V fx
                R a = (q == 2) ? -vf2 : (q > 2 ? -V[MP::ent(i, q - 3)] : R(0));
                if (i == 2) a += G2[q];
                if (i > 2) a -= V[MP::ent(i - 3, q)];
                if (i == q) {
                    const R gd = i == 2 ? j.g22 : (i < 2 ? -c.two_qpos : (i < 5 ? -c.two_qvel : -c.two_qom));
                    a += gd;
                }
                a = M::fma(-W1, Z1[q], a);
                a = M::fma(-W2, Z2[q], a);
                put(MP::ent(i, q), a);
            }
        }
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
