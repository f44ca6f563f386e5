// pmp_ustar2d.cuh -- u_star_2d (pontryagin_utils.py:11-216) for a GENERAL quadratic model of the Hamiltonian,
//     H(u) = H(0) + H_u'u + 1/2 u'H_uu u,   u in [lo, hi]^2,
// with H_u and H_uu supplied per (x, lambda) by generated code (codegen.py: jax.jacobian / jax.hessian of H at u = 0
// become sympy derivatives).  The hand-specialised flatquad functor (pmp_systems.cuh) folds its constant H_uu into
// per-edge constants; this is the same algorithm with H_uu at run time.
#pragma once
#include "pmp_common.cuh"

namespace pmp {

// MODE & 2: KKT-multiplier selection among the four edge minimisers (integrator default);
// otherwise the literal two-pass form: score all candidates with the quadratic model, argmin, re-expand about the
// winner, argmin again (:88-158).  Candidate order = the reference's: unconstrained, then the edges u1 = lo1 (bottom),
// u0 = lo0 (left), u1 = hi1 (top), u0 = hi0 (right) (:78-86).  NaN compares false: a NaN minimiser counts as inside.
// u_star_2d(smooth=True) (pontryagin_utils.py:217-254; a debugging aid of the reference, never inside a solve): instead of the
// argmin, a softmax blend of the five candidates with smoothing length 0.01 -- the unconstrained minimiser carries the penalty
// 1000 softplus(softmax-violation of the box), everything through log-sum-exp.  softmax is invariant to a common shift of the
// scores, so the quadratic model H(u) - H(0) = H_u'u + 1/2 u'H_uu u stands in for the reference's full H(u) (they are equal:
// H is quadratic in u).  cand: unconstrained, bottom, left, top, right -- the reference's order (:78-86).
template <class R>
__device__ __forceinline__ void ustar2d_blend(const R (&cx)[5], const R (&cy)[5], R Hu0, R Hu1, R h00, R h01, R h11, R lo0, R lo1,
                                              R hi0, R hi1, R& u0, R& u1) {
    using M = Math<R>;
    const R ss = R(0.01), inv = R(100.0);
    auto lse2 = [&](R a, R b) {  // log(exp(a) + exp(b)), stable
        const R m = M::max(a, b);
        return m + fx::log(fx::exp(a - m) + fx::exp(b - m));
    };
    // smooth max of the four constraint violations of the unconstrained minimiser, then "clipped" at 0 by another log-sum-exp
    const R v0 = (lo0 - cx[0]) * inv, v1 = (lo1 - cy[0]) * inv, v2 = (cx[0] - hi0) * inv, v3 = (cy[0] - hi1) * inv;
    const R viol = lse2(lse2(v0, v1), lse2(v2, v3)) * ss;
    const R penalty = R(1000) * (lse2(R(0), viol * inv) * ss);
    R sc[5], best = M::inf();
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const R t = M::fma(R(0.5) * h00, cx[k], M::fma(h01, cy[k], Hu0));
        sc[k] = M::fma(cx[k], t, cy[k] * M::fma(R(0.5) * h11, cy[k], Hu1)) + (k == 0 ? penalty : R(0));
        best = M::min(best, sc[k]);
    }
    R wsum = R(0), a0 = R(0), a1 = R(0);
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const R w = fx::exp((best - sc[k]) * inv);  // softmax(-score / ss), shifted by the smallest score
        wsum += w;
        a0 = M::fma(w, cx[k], a0);
        a1 = M::fma(w, cy[k], a1);
    }
    u0 = M::div(a0, wsum);
    u1 = M::div(a1, wsum);
}

// MODE & 4: the smooth blend above instead of a selection (pointwise evaluation only)
template <class R, int MODE>
__device__ __forceinline__ void ustar2d_box(R Hu0, R Hu1, R h00, R h01, R h11, R lo0, R lo1, R hi0, R hi1, R& u0, R& u1) {
    using M = Math<R>;
    // :41 unconstrained minimiser solve(H_uu, -H_u)
    const R idet = M::div(R(1), M::fma(h00, h11, -(h01 * h01)));
    u0 = -(M::fma(h11, Hu0, -(h01 * Hu1)) * idet);
    u1 = -(M::fma(h00, Hu1, -(h01 * Hu0)) * idet);
    const R viol = M::max(M::max(lo0 - u0, lo1 - u1), M::max(u0 - hi0, u1 - hi1));
    const bool outside = viol > R(0);
    // minimiser of H along each edge: root of the tangential derivative, clipped to the edge
    const R n0 = -M::div(R(1), h00), n1 = -M::div(R(1), h11);
    const R xb = M::min(M::max(M::fma(h01, lo1, Hu0) * n0, lo0), hi0);  // bottom: u1 = lo1, u0 moves
    const R yl = M::min(M::max(M::fma(h01, lo0, Hu1) * n1, lo1), hi1);  // left:   u0 = lo0, u1 moves
    const R xt = M::min(M::max(M::fma(h01, hi1, Hu0) * n0, lo0), hi0);  // top
    const R yr = M::min(M::max(M::fma(h01, hi0, Hu1) * n1, lo1), hi1);  // right
    if (MODE & 4) {
        const R bx[5] = {u0, xb, lo0, xt, hi0}, by[5] = {u1, lo1, yl, hi1, yr};
        ustar2d_blend<R>(bx, by, Hu0, Hu1, h00, h01, h11, lo0, lo1, hi0, hi1, u0, u1);
        return;
    }
    if (MODE & 2) {
        // KKT multiplier of the active bound (outward component of -grad H); the box minimiser is the edge
        // minimiser whose multiplier is >= 0: take the largest, first maximum wins
        const R mb = M::fma(h11, lo1, M::fma(h01, xb, Hu1));
        const R ml = M::fma(h00, lo0, M::fma(h01, yl, Hu0));
        const R mt = -M::fma(h11, hi1, M::fma(h01, xt, Hu1));
        const R mr = -M::fma(h00, hi0, M::fma(h01, yr, Hu0));
        R bm = mb, bx = xb, by = lo1;
        bool gt = ml > bm;
        bm = gt ? ml : bm; bx = gt ? lo0 : bx; by = gt ? yl : by;
        gt = mt > bm;
        bm = gt ? mt : bm; bx = gt ? xt : bx; by = gt ? hi1 : by;
        gt = mr > bm;
        bx = gt ? hi0 : bx; by = gt ? yr : by;
        u0 = outside ? bx : u0;
        u1 = outside ? by : u1;
        return;
    }
    const R cx[4] = {xb, lo0, xt, hi0}, cy[4] = {lo1, yl, hi1, yr};
    auto quad = [&](R w0, R w1, R g0, R g1) {  // 1/2 w'H_uu w + g'w
        const R t = M::fma(R(0.5) * h00, w0, M::fma(h01, w1, g0));
        return M::fma(w0, t, w1 * M::fma(R(0.5) * h11, w1, g1));
    };
    R bx = cx[0], by = cy[0], bh = quad(cx[0], cy[0], Hu0, Hu1);
#pragma unroll
    for (int k = 1; k < 4; k++) {
        const R h = quad(cx[k], cy[k], Hu0, Hu1);
        const bool lt = h < bh;
        bh = lt ? h : bh; bx = lt ? cx[k] : bx; by = lt ? cy[k] : by;
    }
    const R g0 = M::fma(h00, bx, M::fma(h01, by, Hu0)), g1 = M::fma(h01, bx, M::fma(h11, by, Hu1));
    R fx = cx[0], fy = cy[0], fh = quad(cx[0] - bx, cy[0] - by, g0, g1);
#pragma unroll
    for (int k = 1; k < 4; k++) {
        const R h = quad(cx[k] - bx, cy[k] - by, g0, g1);
        const bool lt = h < fh;
        fh = lt ? h : fh; fx = lt ? cx[k] : fx; fy = lt ? cy[k] : fy;
    }
    u0 = outside ? fx : u0;
    u1 = outside ? fy : u1;
}

}  // namespace pmp
