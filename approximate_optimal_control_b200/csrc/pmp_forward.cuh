// pmp_forward.cuh -- batched FORWARD closed-loop simulation under a costate feedback law.
//
// The reference simulates the closed loop x' = f(x, u*(x, lambda(x))) forward in time in two places:
//   * levelsets.py:818-836  forward_sim_lqr: lambda(x) = P_lqr x, Tsit5 + PIDController(rtol, atol, dtmin=.05),
//     t: 0 -> 10, SaveAt(steps, t0), max_steps, throw=False  (the NN variants :838-933 differ in lambda(x) only);
//   * eval_utils.py:48-99   closed_loop_eval_general: one extra state records the control cost,
//     z = (x, cost), z' = (f(x,u*), l(x,u*)), Tsit5 with a constant step dt, max_steps = T/dt.
// This kernel covers both, for two costate laws (the `Pol` template parameter):
//   * LinearCostate  lambda(x) = P (x - x_eq), the LQR value function (forward_sim_lqr);
//   * NnCostate      lambda(x) = gradient of the mean of an ensemble of value networks (forward_sim_nn,
//                    forward_sim_nn_until_value levelsets.py:838-933; closed_loop_eval_nn_ensemble eval_utils.py:10-45).
// The state is z = (x, cost); the error norm runs over the x leaf only (the forward_sim_* calls have no cost state,
// and the constant-step evaluation has no error control), so either call site sees its own step sequence.
// Optional terminating event ("inside the value level set"): LinearCostate 1/2 dx'P dx <= level; NnCostate the
// reference's v_mean + 2 v_std <= v_k and v_std <= 0.5 over the ensemble (levelsets.py:897-913).
//
// One thread per trajectory, state and stage derivatives in registers, the same diffrax step-control
// semantics as pmp_solve.cuh.  Forward simulations run by the hundred, not by the million
// (levelsets.py:1201-1215): no work queue, no staging of the dense output.
#pragma once
#include <type_traits>

#include "pmp_solve.cuh"

namespace pmp {

// ---- costate laws lambda(x): eval() also reports an estimate of the value and its spread ------------------
template <class R, int NX> struct LinearCostate {
    R P[NX * NX];  // row-major
    R x_eq[NX];
    // lambda = P (x - x_eq), value 1/2 dx'P dx
    __device__ __forceinline__ void eval(const R* x, R* lam, R& vmean, R& vstd) const {
        using M = Math<R>;
        R dx[NX];
#pragma unroll
        for (int q = 0; q < NX; q++) dx[q] = x[q] - x_eq[q];
        R v2 = R(0);
#pragma unroll
        for (int r = 0; r < NX; r++) {
            R a = R(0);
#pragma unroll
            for (int q = 0; q < NX; q++) a = M::fma(P[r * NX + q], dx[q], a);
            lam[r] = a;
            v2 = M::fma(a, dx[r], v2);
        }
        vmean = R(0.5) * v2;
        vstd = R(0);
    }
    __device__ __forceinline__ bool inside(R vmean, R, R level) const { return vmean <= level; }
    __device__ __forceinline__ int group_lanes() const { return 1; }
    int group_lanes_host() const { return 1; }
};

// 16-byte vector load of four consecutive reals (weights are read at warp-uniform addresses: one broadcast)
__device__ __forceinline__ void load4(const float* p, float* o) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void load4(const double* p, double* o) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// Ensemble of value-function networks (levelsets.py:726-778, nn_utils.py:182-199): each member is the flax MLP
// Dense(nx -> W) softplus, (L-1) x [Dense(W -> W) softplus], Dense(W -> 1), applied to the normalised state,
//     V_e(x) = v_std * MLP_e((x - x_mean) / x_std) + v_mean            (data_normaliser, nn_utils.py:119-152).
// lambda(x) = gradient of the ensemble MEAN (levelsets.py:843-852), by one forward and one backward sweep per
// member; vmean / vstd = mean and (population) standard deviation of V_e(x) over the ensemble (v_meanstd, :789-798).
// Weights: one flat device array, per member K1[nx][W], b1[W], (L-1) x {K[W][W], b[W]}, Kout[W], bout, 3 pad reals
// (flax kernels are [in][out]); W % 4 == 0 and nx*W % 4 == 0 keep every row 16-byte aligned.
// A GROUP of `group` lanes (a power of two <= 16, normally the ensemble size) serves one trajectory: every lane of
// the group carries the same state through the same integrator arithmetic, but evaluates only the members
// e = sub, sub + group, ...; gradients and values are then summed over the group by xor-shuffles (all lanes end up
// with bit-identical sums, so the group never diverges).  Forward simulations come by the hundred or thousand
// (levelsets.py:1201-1215): one thread per trajectory filled a fifth of the machine at one warp per scheduler and
// waited on its own weight loads (21 % issue utilisation); eight lanes per trajectory give 8x the threads and 1/8 of the
// serial work.  Per member: the W accumulators live in registers, the hidden activations and their sigmoids in
// local memory (L1), a weight row is one group-uniform 16-byte load per four FMAs.
constexpr int kNnMaxLayers = 4, kNnMaxNets = 16;
template <class R, int NX, int W> struct NnCostate {
    const R* w;
    int n_nets, n_layers, group;
    R x_mean[NX], inv_x_std[NX], v_mean, v_std, n_sigma, sigma_max;
    __device__ __forceinline__ int group_lanes() const { return group; }
    int group_lanes_host() const { return group; }
    static __host__ __device__ constexpr int net_stride(int L) { return NX * W + W + (L - 1) * (W * W + W) + W + 4; }

    static __device__ __forceinline__ void act(R a, R& sp, R& sg) {  // softplus = logaddexp(a, 0) and its derivative
        const R e = ::exp(-Math<R>::abs(a)), d = Math<R>::div(R(1), R(1) + e);
        sp = Math<R>::max(a, R(0)) + ::log1p(e);
        sg = a >= R(0) ? d : e * d;
    }

    // four consecutive weights: from shared memory (one 16-byte LDS) or from global memory (__ldg)
    template <bool SMEM> static __device__ __forceinline__ void ld4(const R* a, R* o) {
        if (SMEM) {
            typedef typename std::conditional<sizeof(R) == 4, float4, double2>::type V16;
            if (sizeof(R) == 4) {
                const V16 v = *reinterpret_cast<const V16*>(a);
                const R* e = reinterpret_cast<const R*>(&v);
                o[0] = e[0]; o[1] = e[1]; o[2] = e[2]; o[3] = e[3];
            } else {
                const V16 u = *reinterpret_cast<const V16*>(a), v = *(reinterpret_cast<const V16*>(a) + 1);
                const R* e = reinterpret_cast<const R*>(&u);
                const R* f = reinterpret_cast<const R*>(&v);
                o[0] = e[0]; o[1] = e[1]; o[2] = f[0]; o[3] = f[1];
            }
        } else {
            load4(a, o);
        }
    }

    template <bool SMEM> static __device__ __forceinline__ void load_row(const R* a, R* w) {
#pragma unroll
        for (int j = 0; j < W; j += 4) ld4<SMEM>(a + j, w + j);
    }
    static __device__ __forceinline__ R dot_row(const R* w, const R* g) {  // four partial sums
        using M = Math<R>;
        R d0 = R(0), d1 = R(0), d2 = R(0), d3 = R(0);
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            d0 = M::fma(w[j], g[j], d0); d1 = M::fma(w[j + 1], g[j + 1], d1);
            d2 = M::fma(w[j + 2], g[j + 2], d2); d3 = M::fma(w[j + 3], g[j + 3], d3);
        }
        return (d0 + d1) + (d2 + d3);
    }

    // h = softplus(acc), sg = softplus'(acc) for one layer.  The accumulators are registers (compile-time indices), so
    // they are first parked in local memory and the activation runs as a ROLLED loop: unrolled, 64 inlined copies of
    // exp + log1p per layer made a sweep 250-500 KB of code, and the kernels ran at ~20 cycles per instruction out of
    // the instruction cache.
    static __device__ __forceinline__ void activate(const R* acc, R* h, R* sg) {
#pragma unroll
        for (int j = 0; j < W; j++) h[j] = acc[j];
#pragma unroll 1
        for (int j = 0; j < W; j++) {
            R sp, s;
            act(h[j], sp, s);
            h[j] = sp;
            sg[j] = s;
        }
    }

    // One member's forward + backward sweep at the normalised state xn: v = MLP(xn), gx = dMLP/dxn.  p = the member's
    // weights; SMEM: they sit in shared memory (cluster kernel, plain vector loads) instead of global memory (__ldg).
    // NOT inlined: the integrator calls the law once per stage (7 call sites); inlined, every forward kernel carries
    // seven copies of the unrolled sweeps and the library takes a quarter of an hour to compile.
    template <bool SMEM>
    static __device__ __noinline__ void member_sweep(const R* p, const R* xn, int L, R* gx, R& v_out) {
        using M = Math<R>;
        R h[W], sg[kNnMaxLayers][W];  // dynamically indexed: local memory
        R acc[W];
        // ---- forward
#pragma unroll
        for (int j = 0; j < W; j += 4) ld4<SMEM>(p + NX * W + j, acc + j);  // bias
#pragma unroll
        for (int q = 0; q < NX; q++) {
#pragma unroll
            for (int j = 0; j < W; j += 4) {
                R k4[4];
                ld4<SMEM>(p + q * W + j, k4);
#pragma unroll
                for (int c = 0; c < 4; c++) acc[j + c] = M::fma(xn[q], k4[c], acc[j + c]);
            }
        }
        activate(acc, h, sg[0]);
        const R* pl = p + NX * W + W;
        for (int l = 1; l < L; l++, pl += W * W + W) {
#pragma unroll
            for (int j = 0; j < W; j += 4) ld4<SMEM>(pl + W * W + j, acc + j);
            if (sizeof(R) == 4) {
                // software pipeline, two weight rows in flight: the loads of row i+1 are issued before the 64 FMAs of
                // row i (one warp per scheduler cannot hide a 30-cycle shared-memory load behind 4 FMAs)
                R wa[W], wb[W];
                load_row<SMEM>(pl, wa);
                R ha = h[0], hb;
#pragma unroll 1
                for (int i = 0; i < W; i += 2) {
                    load_row<SMEM>(pl + (i + 1) * W, wb);
                    hb = h[i + 1];
#pragma unroll
                    for (int j = 0; j < W; j++) acc[j] = M::fma(ha, wa[j], acc[j]);
                    if (i + 2 < W) {
                        load_row<SMEM>(pl + (i + 2) * W, wa);
                        ha = h[i + 2];
                    }
#pragma unroll
                    for (int j = 0; j < W; j++) acc[j] = M::fma(hb, wb[j], acc[j]);
                }
            } else {
#pragma unroll 2
                for (int i = 0; i < W; i++) {
                    const R hi = h[i];
#pragma unroll
                    for (int j = 0; j < W; j += 4) {
                        R k4[4];
                        ld4<SMEM>(pl + i * W + j, k4);
#pragma unroll
                        for (int c = 0; c < 4; c++) acc[j + c] = M::fma(hi, k4[c], acc[j + c]);
                    }
                }
            }
            activate(acc, h, sg[l]);
        }
        // output layer (pl now points at Kout) and the seed of the backward sweep g = Kout * softplus'(a_L)
        R v = pl[W];
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            R k4[4];
            ld4<SMEM>(pl + j, k4);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                v = M::fma(h[j + c], k4[c], v);
                acc[j + c] = k4[c] * sg[L - 1][j + c];
            }
        }
        // ---- backward through the hidden layers: g_{l-1}[i] = softplus'(a_{l-1})[i] * sum_j K_l[i][j] g_l[j]
        for (int l = L - 1; l >= 1; l--) {
            pl -= W * W + W;
            if (sizeof(R) == 4) {
                R wa[W], wb[W];
                load_row<SMEM>(pl, wa);
#pragma unroll 1
                for (int i = 0; i < W; i += 2) {
                    load_row<SMEM>(pl + (i + 1) * W, wb);
                    const R s0 = sg[l - 1][i], s1 = sg[l - 1][i + 1];
                    const R da = dot_row(wa, acc);
                    if (i + 2 < W) load_row<SMEM>(pl + (i + 2) * W, wa);
                    const R db = dot_row(wb, acc);
                    h[i] = da * s0;
                    h[i + 1] = db * s1;
                }
            } else {
#pragma unroll 2
                for (int i = 0; i < W; i++) {
                    R d0 = R(0), d1 = R(0);
#pragma unroll
                    for (int j = 0; j < W; j += 4) {
                        R k4[4];
                        ld4<SMEM>(pl + i * W + j, k4);
                        d0 = M::fma(k4[0], acc[j], d0); d1 = M::fma(k4[1], acc[j + 1], d1);
                        d0 = M::fma(k4[2], acc[j + 2], d0); d1 = M::fma(k4[3], acc[j + 3], d1);
                    }
                    h[i] = (d0 + d1) * sg[l - 1][i];
                }
            }
#pragma unroll
            for (int j = 0; j < W; j++) acc[j] = h[j];
        }
        // input layer: dMLP/dxn[q] = sum_j K1[q][j] g_0[j]
#pragma unroll
        for (int q = 0; q < NX; q++) {
            R d0 = R(0), d1 = R(0);
#pragma unroll
            for (int j = 0; j < W; j += 4) {
                R k4[4];
                ld4<SMEM>(p + q * W + j, k4);
                d0 = M::fma(k4[0], acc[j], d0); d1 = M::fma(k4[1], acc[j + 1], d1);
                d0 = M::fma(k4[2], acc[j + 2], d0); d1 = M::fma(k4[3], acc[j + 3], d1);
            }
            gx[q] = d0 + d1;
        }
        v_out = v;
    }

    __device__ __forceinline__ void eval(const R* x, R* lam, R& vmean, R& vstd) const {
        using M = Math<R>;
        R xn[NX], gsum[NX], ve[kNnMaxNets];
#pragma unroll
        for (int q = 0; q < NX; q++) { xn[q] = (x[q] - x_mean[q]) * inv_x_std[q]; gsum[q] = R(0); }
        const int L = n_layers;
        const int G = group, lane = threadIdx.x & 31, sub = lane & (G - 1);
        const unsigned gmask = G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
        int mine = 0;
        for (int e = sub; e < n_nets; e += G) {
            R gx[NX], v;
            member_sweep<false>(w + (size_t)e * net_stride(L), xn, L, gx, v);
#pragma unroll
            for (int q = 0; q < NX; q++) gsum[q] = M::fma(gx[q], v_std * inv_x_std[q], gsum[q]);  // chain rule of the normaliser
            ve[mine++] = M::fma(v_std, v, v_mean);
        }
        // sums over the group (xor butterfly: every lane of the group ends with the same bits)
        const R inv_e = M::div(R(1), R(n_nets));
        R m = R(0);
        for (int e = 0; e < mine; e++) m += ve[e];
        for (int off = G >> 1; off > 0; off >>= 1) {
            m += __shfl_xor_sync(gmask, m, off);
#pragma unroll
            for (int q = 0; q < NX; q++) gsum[q] += __shfl_xor_sync(gmask, gsum[q], off);
        }
        m *= inv_e;
        R var = R(0);
        for (int e = 0; e < mine; e++) var = M::fma(ve[e] - m, ve[e] - m, var);
        for (int off = G >> 1; off > 0; off >>= 1) var += __shfl_xor_sync(gmask, var, off);
        vmean = m;
        vstd = M::sqrt(var * inv_e);
#pragma unroll
        for (int q = 0; q < NX; q++) lam[q] = gsum[q] * inv_e;
    }
    // levelsets.py:897-913: very likely inside the value level set AND the ensemble agrees
    __device__ __forceinline__ bool inside(R vmean, R vstd, R level) const {
        return (M_fma(vmean, vstd) <= level) && (vstd <= sigma_max);
    }
    __device__ __forceinline__ R M_fma(R vmean, R vstd) const { return Math<R>::fma(n_sigma, vstd, vmean); }
};

template <class R> struct ForwardIO {
    const R* x0;  // [NX][n]
    R* z_end;     // [NX + 1][n]: x, cost
    R *ts, *xs, *cost;  // dense: [n][S1], [n][S1][NX], [n][S1] (NULL: not saved)
    int32_t *n_accepted, *n_steps, *result;
    long long n;
};

template <class Sys, class R, int METHOD, class Pol>
__global__ void __launch_bounds__(128)
forward_kernel(const typename Sys::Consts sc, const SolveOpts<R> o, const ForwardIO<R> io, const Pol pol,
               const TabParams<R> tp) {
    using M = Math<R>;
    using TB = Tab<METHOD>;
    constexpr int NX = Sys::NX, NZ = NX + 1, S = TB::S;
    // G lanes per trajectory (1 for the linear law); only the first lane of a group writes
    const int G = pol.group_lanes();
    const long long n = io.n, gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, i = gtid / G;
    if (i >= n) return;  // (n * G threads are launched: whole groups leave together)
    const bool writer = (gtid - i * G) == 0;
    const int S1 = o.max_steps + 1;

    // z' = (f(x, u*(x, lambda(x))), l(x, u*)); also the law's value estimate (mean, spread) at the state
    auto field = [&](const R* z, R* k, R& vm, R& vs) {
        R lam[NX], ld[NX], vd;
        pol.eval(z, lam, vm, vs);
        Sys::template rhs<2>(sc, z, lam, k, vd, ld);
        k[NX] = -vd;
    };

    R z[NZ], k[S][NZ];
#pragma unroll
    for (int q = 0; q < NX; q++) z[q] = io.x0[q * n + i];
    z[NX] = R(0);
    R tprev = o.T0, tnext = M::min(o.T0 + o.dt0, o.T1), vlqr = R(0), vsig = R(0);
    int nsteps = 0, nacc = 0, result = 0;
    bool at_dtmin = false;
    const R inf = M::inf();
    auto save = [&](int slot) {
        if (!io.ts || slot >= S1 || !writer) return;
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
    if (TB::FSAL) field(z, k[0], vlqr, vsig);

    while (tprev < o.T1 && result == 0 && nsteps < o.max_steps) {
        const R dt = tnext - tprev, h = dt * o.dir;
        if (!TB::FSAL) field(z, k[0], vlqr, vsig);
        R z1[NZ], v1 = R(0), s1 = R(0);
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
            field(zs, k[s], v1, s1);
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
            if (TB::FSAL) { vlqr = v1; vsig = s1; }
            save(nacc);
        }
        if (!TB::FSAL && o.event != PMP_EVENT_NONE) {  // RK4: the value of the post-step state
            R kk[NZ];
            field(z, kk, vlqr, vsig);
        }
        if (result == 0 && o.event != PMP_EVENT_NONE && pol.inside(vlqr, vsig, o.event_level)) result = PMP_RESULT_EVENT;
        if (result == 0 && nan_state) result = PMP_RESULT_NAN;
    }
    if (result == 0 && tprev < o.T1) result = PMP_RESULT_MAX_STEPS;
    if (!writer) return;
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

// v_meanstds / vx_meanstd (levelsets.py:789-815): the ensemble at n states
template <class R, class Pol, int NX>
__global__ void __launch_bounds__(128) nn_value_kernel(const Pol pol, const R* __restrict__ x_, R* __restrict__ lam_,
                                                       R* __restrict__ vm_, R* __restrict__ vs_, long long n) {
    const int G = pol.group_lanes();
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, i = gtid / G;
    if (i >= n) return;
    R x[NX], lam[NX], vm, vs;
#pragma unroll
    for (int q = 0; q < NX; q++) x[q] = x_[q * n + i];
    pol.eval(x, lam, vm, vs);
    if (gtid - i * G != 0) return;
#pragma unroll
    for (int q = 0; q < NX; q++) lam_[q * n + i] = lam[q];
    vm_[i] = vm;
    vs_[i] = vs;
}

template <class R, int NX, class Fill>
int launch_nn_value_r(const pmp_nn_costate_t* nn, long long n, const void* x, void* lam, void* vm, void* vs, Fill&& fill,
                      cudaStream_t st) {
    int G = 1;
    while (2 * G <= nn->n_nets && 2 * G <= 16) G *= 2;
    const unsigned grid = (unsigned)((n * G + 127) / 128);
    if (nn->width == 64) {
        NnCostate<R, NX, 64> pol; fill(pol);
        nn_value_kernel<R, NnCostate<R, NX, 64>, NX><<<grid, 128, 0, st>>>(pol, (const R*)x, (R*)lam, (R*)vm, (R*)vs, n);
    } else if (nn->width == 32) {
        NnCostate<R, NX, 32> pol; fill(pol);
        nn_value_kernel<R, NnCostate<R, NX, 32>, NX><<<grid, 128, 0, st>>>(pol, (const R*)x, (R*)lam, (R*)vm, (R*)vs, n);
    } else {
        return -1000;
    }
    return (int)cudaGetLastError();
}

template <int NX>
int launch_nn_value(const pmp_nn_costate_t* nn, int dtype, long long n, const void* x, void* lam, void* vm, void* vs,
                    cudaStream_t st) {
    if (n == 0) return 0;
    auto fill = [&](auto& pol) {
        using R = std::remove_cv_t<std::remove_reference_t<decltype(pol.v_mean)>>;
        pol.w = (const R*)nn->weights;
        pol.n_nets = nn->n_nets; pol.n_layers = nn->n_hidden_layers;
        pol.group = 1;
        while (2 * pol.group <= nn->n_nets && 2 * pol.group <= 16) pol.group *= 2;
        for (int q = 0; q < NX; q++) { pol.x_mean[q] = R(nn->x_mean[q]); pol.inv_x_std[q] = R(1.0 / nn->x_std[q]); }
        pol.v_mean = R(nn->v_mean); pol.v_std = R(nn->v_std); pol.n_sigma = R(nn->n_sigma); pol.sigma_max = R(nn->sigma_max);
    };
    if (dtype == PMP_F32) return launch_nn_value_r<float, NX>(nn, n, x, lam, vm, vs, fill, st);
    return launch_nn_value_r<double, NX>(nn, n, x, lam, vm, vs, fill, st);
}

template <class Sys, class R, class Pol>
int launch_forward_pol(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<R>& io, const Pol& pol,
                       cudaStream_t st) {
    const unsigned grid = (unsigned)((n * pol.group_lanes_host() + 127) / 128);
#define PMP_FGO(METHOD)                                                                                          \
    do {                                                                                                          \
        using TB = Tab<METHOD>;                                                                                   \
        forward_kernel<Sys, R, METHOD, Pol><<<grid, 128, 0, st>>>(Sys::make(p), make_solve_opts<R, TB>(o, TB::ORDER, Sys::NX), \
                                                                  io, pol, TabParams<R>::template make<TB>());   \
        return (int)cudaGetLastError();                                                                           \
    } while (0)
    if (o.method == PMP_METHOD_RK4) PMP_FGO(PMP_METHOD_RK4);
    if (o.method == PMP_METHOD_TSIT5) PMP_FGO(PMP_METHOD_TSIT5);
    if (o.method == PMP_METHOD_DOPRI5) PMP_FGO(PMP_METHOD_DOPRI5);
#undef PMP_FGO
    return -1000;
}

template <template <class> class SysT, class R>
int launch_forward_r(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const double* P,
                     const double* x_eq, void* z_end, void* ts, void* xs, void* cost, int32_t* a, int32_t* s, int32_t* r,
                     cudaStream_t st) {
    using Sys = SysT<R>;
    constexpr int NX = Sys::NX;
    ForwardIO<R> io{(const R*)x0, (R*)z_end, (R*)ts, (R*)xs, (R*)cost, a, s, r, n};
    LinearCostate<R, NX> pol;
    for (int q = 0; q < NX * NX; q++) pol.P[q] = R(P[q]);
    for (int q = 0; q < NX; q++) pol.x_eq[q] = R(x_eq ? x_eq[q] : 0.0);
    return launch_forward_pol<Sys, R>(p, o, n, io, pol, st);
}

// defined in a compile unit of its own (pmp_forward_cluster.cu) for every (system, dtype, width)
template <class Sys, class R, int W>
int launch_forward_cluster(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<R>& io,
                           const NnCostate<R, Sys::NX, W>& pol, cudaStream_t st);

// fp32 ensembles of width 64: the layer products on the tensor cores (pmp_forward_tc.cuh; own compile unit pmp_forward_tc.cu)
template <class Sys>
int launch_forward_tc_entry(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<float>& io,
                            const NnCostate<float, Sys::NX, 64>& pol, cudaStream_t st);

// the same on the 5th-generation tensor cores: tcgen05.mma, clusters of one CTA per member (pmp_forward_umma.cuh; pmp_forward_umma.cu)
template <class Sys>
int launch_forward_umma_entry(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<float>& io,
                              const NnCostate<float, Sys::NX, 64>& pol, cudaStream_t st);

template <template <class> class SysT, class R>
int launch_forward_nn_r(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const pmp_nn_costate_t* nn,
                        void* z_end, void* ts, void* xs, void* cost, int32_t* a, int32_t* s, int32_t* r, cudaStream_t st) {
    using Sys = SysT<R>;
    constexpr int NX = Sys::NX;
    ForwardIO<R> io{(const R*)x0, (R*)z_end, (R*)ts, (R*)xs, (R*)cost, a, s, r, n};
    auto fill = [&](auto& pol) {
        pol.w = (const R*)nn->weights;
        pol.n_nets = nn->n_nets; pol.n_layers = nn->n_hidden_layers;
        pol.group = 1;
        while (2 * pol.group <= nn->n_nets && 2 * pol.group <= 16) pol.group *= 2;
        for (int q = 0; q < NX; q++) { pol.x_mean[q] = R(nn->x_mean[q]); pol.inv_x_std[q] = R(1.0 / nn->x_std[q]); }
        pol.v_mean = R(nn->v_mean); pol.v_std = R(nn->v_std); pol.n_sigma = R(nn->n_sigma); pol.sigma_max = R(nn->sigma_max);
    };
    // Two kernels serve an ensemble: the cluster kernel (pmp_forward_cluster.cuh: weights in shared memory, partial results
    // exchanged through distributed shared memory, slots refilled from a per-cluster queue; default) and, with
    // B200PMP_NN_CLUSTER=0, the lane-group kernel above.  2^14 flatquad simulations (8 x 6-64-64-64-1, fp32): 70 ms / 94 ms.
    const char* env = getenv("B200PMP_NN_CLUSTER");
    const bool use_cluster = nn->n_nets >= 2 && !(env && env[0] == '0');
    if (nn->width == 64) {
        NnCostate<R, NX, 64> pol; fill(pol);
        if constexpr (std::is_same<R, float>::value && NX <= 8) {
            // default for fp32: tensor-core kernels (3xTF32 split precision) -- tcgen05 clusters for up to 8 members
            // (B200PMP_NN_UMMA=0 skips them), else warp-level MMAs; B200PMP_NN_TC=0 selects the CUDA-core kernels
            const char* tc = getenv("B200PMP_NN_TC");
            const char* um = getenv("B200PMP_NN_UMMA");
            if (!(tc && tc[0] == '0') && !(um && um[0] == '0') && nn->n_hidden_layers >= 1 && nn->n_hidden_layers <= 3) {
                const int rc = launch_forward_umma_entry<Sys>(p, o, n, io, pol, st);
                if (rc != -1000) return rc;
            }
            if (!(tc && tc[0] == '0') && nn->n_hidden_layers >= 1 && nn->n_hidden_layers <= 3) {
                const int rc = launch_forward_tc_entry<Sys>(p, o, n, io, pol, st);
                if (rc != -1000) return rc;
            }
        }
        return use_cluster ? launch_forward_cluster<Sys, R, 64>(p, o, n, io, pol, st) : launch_forward_pol<Sys, R>(p, o, n, io, pol, st);
    }
    if (nn->width == 32) {
        NnCostate<R, NX, 32> pol; fill(pol);
        return use_cluster ? launch_forward_cluster<Sys, R, 32>(p, o, n, io, pol, st) : launch_forward_pol<Sys, R>(p, o, n, io, pol, st);
    }
    return -1000;
}

template <template <class> class SysT>
int launch_forward(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const double* P,
                   const double* x_eq, void* z_end, void* ts, void* xs, void* cost, int32_t* a, int32_t* s, int32_t* r,
                   cudaStream_t st) {
    if (n == 0) return 0;
    if (o.dtype == PMP_F32) return launch_forward_r<SysT, float>(p, o, n, x0, P, x_eq, z_end, ts, xs, cost, a, s, r, st);
    if (o.dtype == PMP_F64) return launch_forward_r<SysT, double>(p, o, n, x0, P, x_eq, z_end, ts, xs, cost, a, s, r, st);
    return -1000;
}

template <template <class> class SysT>
int launch_forward_nn(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x0, const pmp_nn_costate_t* nn,
                      void* z_end, void* ts, void* xs, void* cost, int32_t* a, int32_t* s, int32_t* r, cudaStream_t st) {
    if (n == 0) return 0;
    if (o.dtype == PMP_F32) return launch_forward_nn_r<SysT, float>(p, o, n, x0, nn, z_end, ts, xs, cost, a, s, r, st);
    if (o.dtype == PMP_F64) return launch_forward_nn_r<SysT, double>(p, o, n, x0, nn, z_end, ts, xs, cost, a, s, r, st);
    return -1000;
}

}  // namespace pmp
