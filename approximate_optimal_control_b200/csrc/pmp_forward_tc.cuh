// pmp_forward_tc.cuh -- forward closed-loop simulation under the value-network-ensemble costate law with the layer
// products on the TENSOR CORES (fp32 ensembles of width 64; levelsets.py:838-933, eval_utils.py:10-45).
//
// The costate law lambda(x) = grad mean_e V_e(x) costs 140 k FMAs per right hand side: per member a forward and a backward
// sweep through a 6-64-64-64-1 softplus MLP.  On CUDA cores (pmp_forward.cuh, pmp_forward_cluster.cuh) that ran at 5-7
// TFLOP/s.  Per layer and member it is a [trajectories x 64] . [64 x 64] product -- a dense contraction once 32
// trajectories are batched -- so here:
//   * a WARP carries 32 trajectories (one per lane: state, stage derivatives and step control stay per thread, as in
//     every other integrator of this library) and evaluates the law for its 32 rows with warp-level MMAs
//     (mma.sync.m16n8k8 TF32: 2 row tiles x 8 column tiles x 8 k-steps per layer product);
//   * SPLIT PRECISION (3xTF32): a = a_hi + a_lo, w = w_hi + w_lo in TF32, a.w ~ a_lo.w_hi + a_hi.w_lo + a_hi.w_hi with
//     fp32 accumulation -- relative error ~2^-21 per product, i.e. fp32 quality: the closed loop is sensitive at the 1e-3
//     level (u* sits on the edges of the input box), plain TF32 (2^-11) is not enough;
//   * activations (softplus, and its derivative for the backward sweep) run on the CUDA cores between the products, on the
//     accumulator fragments; they go back to the next product through a per-warp shared-memory tile (accumulator layout ->
//     A-operand layout), as do the backward sweep's gradients;
//   * the weights of ONE member at a time sit in shared memory, already split into hi / lo and padded (row stride 68:
//     conflict-free B-operand reads for the transposed products of the backward sweep, two-way conflicts for the forward
//     ones); the CTA's four warps load them together between members.
// One CTA = 128 threads = 128 trajectory slots, refilled from the CTA's queue as trajectories end (as in the cluster kernel,
// which remains the fp64 / other-width path).  No cluster, no distributed shared memory.
#pragma once
#include "pmp_forward.cuh"

namespace pmp {

constexpr int kTcW = 64;          // hidden width served by this kernel
constexpr int kTcWS = 68;         // padded row stride of weight matrices and activation tiles (floats)
constexpr int kTcRefillMin = 16;  // idle slots of a CTA that trigger a refill round (each costs one extra law evaluation)

__device__ __forceinline__ uint32_t tf32_of(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_of(x);
    lo = tf32_of(x - __uint_as_float(hi));
}
// d += a . b   (m16n8k8, A row-major 16x8, B "col" 8x8, fp32 accumulators)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// softplus and its derivative from MUFU exp2 / log2 / rcp (absolute error ~1e-7 on values of order one)
__device__ __forceinline__ void act_fast(float a, float& sp, float& sg) {
    const float e = __expf(-fabsf(a));
    const float d = __fdividef(1.0f, 1.0f + e);
    sp = fmaxf(a, 0.0f) + __logf(1.0f + e);
    sg = a >= 0.0f ? d : e * d;
}

// shared-memory image of one member (floats): per hidden layer l >= 1 a [64][68] matrix, hi then lo; layer 0 an [8][68]
// matrix (rows NX..7 zero), hi then lo; then the biases [L][64], w_out [64], b_out
struct TcLayout {
    int L;
    __host__ __device__ int w0_hi() const { return 0; }
    __host__ __device__ int w0_lo() const { return 8 * kTcWS; }
    __host__ __device__ int wl_hi(int l) const { return 16 * kTcWS + (l - 1) * 2 * kTcW * kTcWS; }
    __host__ __device__ int wl_lo(int l) const { return wl_hi(l) + kTcW * kTcWS; }
    __host__ __device__ int bias(int l) const { return 16 * kTcWS + (L - 1) * 2 * kTcW * kTcWS + l * kTcW; }
    __host__ __device__ int wout() const { return bias(L); }
    __host__ __device__ int bout() const { return wout() + kTcW; }
    __host__ __device__ int total() const { return (bout() + 4) & ~3; }
};
// per-warp tiles: H [32][68] (activations / gradients: the A operand of the next product) and SG [L][32][68]
__host__ __device__ inline int tc_warp_floats(int L) { return (1 + L) * 32 * kTcWS; }
template <int NX> inline size_t tc_smem_bytes(int L) {
    TcLayout lay{L};
    return ((size_t)lay.total() + 4 * (size_t)tc_warp_floats(L)) * sizeof(float);
}

// One layer product for the warp's 32 rows: acc[mt][nt][.] += A[32 x K] . B, A from the tile `hA` (row stride 68),
// B = W (TRANS = false: B[k][n] = w[k][n]) or W^T (TRANS = true: B[k][n] = w[n][k]); K = 8 * KSTEPS, N = 8 * NT.
template <int KSTEPS, int NT, bool TRANS>
__device__ __forceinline__ void tc_product(const float* __restrict__ hA, const float* __restrict__ w_hi, const float* __restrict__ w_lo,
                                           int lane, float (&acc)[2][NT][4]) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 2
    for (int ks = 0; ks < KSTEPS; ks++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const float* r0 = hA + (16 * mt + g) * kTcWS + 8 * ks + t;
            split_tf32(r0[0], ah[mt][0], al[mt][0]);
            split_tf32(r0[8 * kTcWS], ah[mt][1], al[mt][1]);
            split_tf32(r0[4], ah[mt][2], al[mt][2]);
            split_tf32(r0[8 * kTcWS + 4], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const int o0 = TRANS ? (8 * nt + g) * kTcWS + 8 * ks + t : (8 * ks + t) * kTcWS + 8 * nt + g;
            const int o1 = TRANS ? o0 + 4 : o0 + 4 * kTcWS;
            const uint32_t bh0 = __float_as_uint(w_hi[o0]), bh1 = __float_as_uint(w_hi[o1]);
            const uint32_t bl0 = __float_as_uint(w_lo[o0]), bl1 = __float_as_uint(w_lo[o1]);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                mma_tf32(acc[mt][nt], al[mt], bh0, bh1);   // small terms first
                mma_tf32(acc[mt][nt], ah[mt], bl0, bl1);
                mma_tf32(acc[mt][nt], ah[mt], bh0, bh1);
            }
        }
    }
}

template <class Sys, int METHOD>
__global__ void __launch_bounds__(128, 1)
forward_tc_kernel(const typename Sys::Consts sc, const SolveOpts<float> o, const ForwardIO<float> io,
                  const NnCostate<float, Sys::NX, kTcW> pol, const long long chunk, const int refill_min) {
    using R = float;
    using M = Math<R>;
    using TB = Tab<METHOD>;
    using Pol = NnCostate<R, Sys::NX, kTcW>;
    constexpr int NX = Sys::NX, NZ = NX + 1, S = TB::S, BLK = 128, W = kTcW, WS = kTcWS;
    static_assert(NX <= 8, "the first layer is one k-step of 8");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long n = io.n;
    const long long cb = (long long)blockIdx.x * chunk, ce = cb + chunk < n ? cb + chunk : n;
    long long next = cb;
    const int S1 = o.max_steps + 1, L = pol.n_layers, E = pol.n_nets;
    const int stride = Pol::net_stride(L);
    const TcLayout lay{L};

    extern __shared__ __align__(16) unsigned char tc_smem[];
    float* sW = reinterpret_cast<float*>(tc_smem);                   // one member's image (TcLayout)
    float* hA = sW + lay.total() + (size_t)warp * tc_warp_floats(L);  // this warp's [32][68] tile
    float* sSG = hA + 32 * WS;                                       // this warp's softplus' tiles, [L][32][68]
    __shared__ int s_cnt[BLK / 32];

    // cooperative load of member e: flax layout in global memory -> split / padded image in shared memory
    auto load_member = [&](int e) {
        const R* src = pol.w + (size_t)e * stride;
        __syncthreads();  // everybody is done with the previous member's image
        for (int q = threadIdx.x; q < 8 * W; q += BLK) {  // layer 0: [NX][W] -> [8][68], rows >= NX zero
            const int r = q / W, c = q - r * W;
            const float v = r < NX ? __ldg(src + r * W + c) : 0.0f;
            uint32_t hi, lo;
            split_tf32(v, hi, lo);
            sW[lay.w0_hi() + r * WS + c] = __uint_as_float(hi);
            sW[lay.w0_lo() + r * WS + c] = __uint_as_float(lo);
        }
        if (threadIdx.x < W) sW[lay.bias(0) + threadIdx.x] = __ldg(src + NX * W + threadIdx.x);
        const R* pl = src + NX * W + W;
        for (int l = 1; l < L; l++, pl += W * W + W) {
            for (int q = threadIdx.x; q < W * W; q += BLK) {
                const int r = q / W, c = q - r * W;
                uint32_t hi, lo;
                split_tf32(__ldg(pl + q), hi, lo);
                sW[lay.wl_hi(l) + r * WS + c] = __uint_as_float(hi);
                sW[lay.wl_lo(l) + r * WS + c] = __uint_as_float(lo);
            }
            if (threadIdx.x < W) sW[lay.bias(l) + threadIdx.x] = __ldg(pl + W * W + threadIdx.x);
        }
        if (threadIdx.x < W) sW[lay.wout() + threadIdx.x] = __ldg(pl + threadIdx.x);
        if (threadIdx.x == 0) sW[lay.bout()] = __ldg(pl + W);
        __syncthreads();
    };

    // accumulator fragments (+ bias) -> softplus -> tile hA (next product's A operand), softplus' -> tile sg
    auto activate_store = [&](float (&acc)[2][8][4], const float* bias, float* sg) {
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const int col = 8 * nt + 2 * t;
                const float b0 = bias[col], b1 = bias[col + 1];
                float2 h0, h1, s0, s1;
                act_fast(acc[mt][nt][0] + b0, h0.x, s0.x);
                act_fast(acc[mt][nt][1] + b1, h0.y, s0.y);
                act_fast(acc[mt][nt][2] + b0, h1.x, s1.x);
                act_fast(acc[mt][nt][3] + b1, h1.y, s1.y);
                const int r0 = (16 * mt + g) * WS + col, r1 = r0 + 8 * WS;
                *reinterpret_cast<float2*>(hA + r0) = h0;
                *reinterpret_cast<float2*>(hA + r1) = h1;
                *reinterpret_cast<float2*>(sg + r0) = s0;
                *reinterpret_cast<float2*>(sg + r1) = s1;
            }
    };

    // lambda(x), v_mean, v_std for this thread's trajectory; collective over the CTA (weight loads) and the warp (MMAs)
    auto law = [&](const R* x, R* lam, R& vmean, R& vstd) {
        R xn[NX], gs[NX], ve[kNnMaxNets], sum = R(0);
#pragma unroll
        for (int q = 0; q < NX; q++) { xn[q] = (x[q] - pol.x_mean[q]) * pol.inv_x_std[q]; gs[q] = R(0); }
        for (int e = 0; e < E; e++) {
            load_member(e);
            // ---- forward sweep
#pragma unroll
            for (int q = 0; q < 8; q++) hA[lane * WS + q] = q < NX ? xn[q] : 0.0f;
            __syncwarp();
            {
                float acc[2][8][4] = {};
                tc_product<1, 8, false>(hA, sW + lay.w0_hi(), sW + lay.w0_lo(), lane, acc);
                __syncwarp();  // every lane has read its A fragments: the tile may be overwritten
                activate_store(acc, sW + lay.bias(0), sSG);
                __syncwarp();
            }
            for (int l = 1; l < L; l++) {
                float acc[2][8][4] = {};
                tc_product<8, 8, false>(hA, sW + lay.wl_hi(l), sW + lay.wl_lo(l), lane, acc);
                __syncwarp();
                activate_store(acc, sW + lay.bias(l), sSG + l * 32 * WS);
                __syncwarp();
            }
            // value: this lane's row of the last activation tile . w_out + b_out (four partial sums)
            {
                const float* hr = hA + lane * WS;
                const float* wo = sW + lay.wout();
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int j = 0; j < W; j += 4) {
                    const float4 hv = *reinterpret_cast<const float4*>(hr + j), wv = *reinterpret_cast<const float4*>(wo + j);
                    d0 = fmaf(hv.x, wv.x, d0); d1 = fmaf(hv.y, wv.y, d1); d2 = fmaf(hv.z, wv.z, d2); d3 = fmaf(hv.w, wv.w, d3);
                }
                const float v = ((d0 + d1) + (d2 + d3)) + sW[lay.bout()];
                const float vv = fmaf(pol.v_std, v, pol.v_mean);
                ve[e] = vv;
                sum += vv;
            }
            __syncwarp();
            // ---- backward sweep: G_L = w_out (.) sg_L;  G_l = (G_{l+1} . W_{l+1}^T) (.) sg_l;  gx = G_1 . W_1^T
            {
                const float* sgL = sSG + (L - 1) * 32 * WS;
                const float* wo = sW + lay.wout();
                for (int q = lane; q < 32 * W; q += 32) {
                    const int r = q / W, c = q - r * W;
                    hA[r * WS + c] = sgL[r * WS + c] * wo[c];
                }
                __syncwarp();
            }
            for (int l = L - 1; l >= 1; l--) {
                float acc[2][8][4] = {};
                tc_product<8, 8, true>(hA, sW + lay.wl_hi(l), sW + lay.wl_lo(l), lane, acc);
                __syncwarp();
                const float* sg = sSG + (l - 1) * 32 * WS;
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        const int r0 = (16 * mt + g) * WS + 8 * nt + 2 * t, r1 = r0 + 8 * WS;
                        const float2 s0 = *reinterpret_cast<const float2*>(sg + r0), s1 = *reinterpret_cast<const float2*>(sg + r1);
                        *reinterpret_cast<float2*>(hA + r0) = make_float2(acc[mt][nt][0] * s0.x, acc[mt][nt][1] * s0.y);
                        *reinterpret_cast<float2*>(hA + r1) = make_float2(acc[mt][nt][2] * s1.x, acc[mt][nt][3] * s1.y);
                    }
                __syncwarp();
            }
            {
                float acc[2][1][4] = {};
                tc_product<8, 1, true>(hA, sW + lay.w0_hi(), sW + lay.w0_lo(), lane, acc);
                __syncwarp();
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    const int r0 = (16 * mt + g) * WS + 2 * t, r1 = r0 + 8 * WS;
                    *reinterpret_cast<float2*>(hA + r0) = make_float2(acc[mt][0][0], acc[mt][0][1]);
                    *reinterpret_cast<float2*>(hA + r1) = make_float2(acc[mt][0][2], acc[mt][0][3]);
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < NX; q++) gs[q] = fmaf(hA[lane * WS + q], pol.v_std * pol.inv_x_std[q], gs[q]);
                __syncwarp();
            }
        }
        const R inv_e = M::div(R(1), R(E));
        vmean = sum * inv_e;
        R var = R(0);
        for (int e = 0; e < E; e++) var = fmaf(ve[e] - vmean, ve[e] - vmean, var);
        vstd = M::sqrt(var * inv_e);
#pragma unroll
        for (int q = 0; q < NX; q++) lam[q] = gs[q] * inv_e;
    };
    auto field = [&](const R* z, R* k, R& vm, R& vs) {
        R lam[NX], ld[NX], vd;
        law(z, lam, vm, vs);
        Sys::template rhs<2>(sc, z, lam, k, vd, ld);
        k[NX] = -vd;
    };

    long long i = -1;  // the slot's trajectory (-1: none)
    R z[NZ], k[S][NZ];
#pragma unroll
    for (int q = 0; q < NZ; q++) z[q] = R(0);
    R tprev = o.T1, tnext = o.T1, vcur = R(0), scur = R(0);
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
    auto is_live = [&]() { return i >= 0 && tprev < o.T1 && result == 0 && nsteps < o.max_steps; };

    while (true) {
        // ---- retire + refill round: when enough slots wait, or nothing is live
        const bool live0 = is_live();
        const bool want = !live0;
        const int n_want = __syncthreads_count(want);
        const bool any_left = next < ce;
        if (n_want == BLK || (n_want >= refill_min && any_left) || (n_want > 0 && !any_left)) {
            const unsigned bal = __ballot_sync(0xffffffffu, want);
            if (lane == 0) s_cnt[warp] = __popc(bal);
            __syncthreads();
            int before = 0;
            for (int w = 0; w < warp; w++) before += s_cnt[w];
            const int my_rank = before + __popc(bal & ((1u << lane) - 1u));
            __syncthreads();
            bool fresh = false;
            if (want) {
                if (i >= 0) {  // retire
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
                i = -1;
                if (next + my_rank < ce) {
                    i = next + my_rank;
                    fresh = true;
#pragma unroll
                    for (int q = 0; q < NX; q++) z[q] = io.x0[q * n + i];
                    z[NX] = R(0);
                    tprev = o.T0; tnext = M::min(o.T0 + o.dt0, o.T1);
                    nsteps = 0; nacc = 0; result = 0; at_dtmin = false;
                    bool bad = false;
#pragma unroll
                    for (int q = 0; q < NX; q++) bad = bad || (z[q] != z[q]);
                    if (bad) result = PMP_RESULT_NAN;  // retires in the next round (it sweeps its NaN state meanwhile)
                    save(0);
                }
            }
            const long long left = ce - next;
            next += n_want < left ? n_want : left;
            if (TB::FSAL && __syncthreads_or(fresh)) {  // k1 of the fresh trajectories: one collective evaluation for everybody
                R kf[NZ], vf, sf;
                field(z, kf, vf, sf);
                if (fresh) {
#pragma unroll
                    for (int q = 0; q < NZ; q++) k[0][q] = kf[q];
                    vcur = vf; scur = sf;
                }
            }
        }
        const bool live = is_live();
        if (!__syncthreads_or(live)) {
            if (next >= ce && __syncthreads_count(i >= 0) == 0) break;  // nothing live, nothing waiting, nothing left
            continue;                                                  // finished slots wait for the retire round above
        }
        const R dt = live ? tnext - tprev : o.dt0, h = dt * o.dir;
        if (!TB::FSAL) field(z, k[0], vcur, scur);
        R z1[NZ], v1 = R(0), s1 = R(0);
#pragma unroll
        for (int s = 1; s < S; s++) {
            R zs[NZ];
#pragma unroll
            for (int q = 0; q < NZ; q++) {
                R acc = z[q];
#pragma unroll
                for (int j = 0; j < s; j++)
                    if (TB::a(s, j) != 0.0) acc = M::fma(h * R(TB::a(s, j)), k[j][q], acc);
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
                    if (TB::b(j) != 0.0) acc = M::fma(h * R(TB::b(j)), k[j][q], acc);
                z1[q] = acc;
            }
        }
        bool keep = true, step_nonfinite = false;
        R next_t0, next_t1;
        bool at_dtmin_new = at_dtmin;
        int result_new = result;
        if (TB::EMBEDDED && o.adaptive) {
            R sumsq = R(0);
#pragma unroll
            for (int q = 0; q < NX; q++) {
                R e = (h * R(TB::berr(0))) * k[0][q];
#pragma unroll
                for (int j = 1; j < S; j++)
                    if (TB::berr(j) != 0.0) e = M::fma(h * R(TB::berr(j)), k[j][q], e);
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
                if (!o.force_dtmin && dtn < o.dtmin) result_new = PMP_RESULT_DT_MIN;
                at_dtmin_new = dtn <= o.dtmin;
                dtn = M::max(dtn, o.dtmin);
            } else {
                at_dtmin_new = false;
            }
            next_t0 = keep ? tnext : tprev;
            next_t1 = next_t0 + dtn;
        } else {
            next_t0 = tnext;
            next_t1 = tnext + dt;
        }
        bool nan_state = false;
        if (live) {
            at_dtmin = at_dtmin_new;
            result = result_new;
            tprev = M::min(next_t0, o.T1);
            if (next_t1 > o.T1 - M::clip_tol()) next_t1 = keep ? o.T1 : M::fma(R(0.5), o.T1 - tprev, tprev);
            tnext = next_t1;
            nsteps += 1;
            if (keep) {
                nacc += 1;
                nan_state = step_nonfinite;
#pragma unroll
                for (int q = 0; q < NZ; q++) {
                    z[q] = z1[q];
                    nan_state = nan_state || (z1[q] != z1[q]);
                    if (TB::FSAL) k[0][q] = k[S - 1][q];
                }
                if (TB::FSAL) { vcur = v1; scur = s1; }
                save(nacc);
            }
        }
        if (!TB::FSAL && o.event != PMP_EVENT_NONE) {  // RK4: the law's value at the post-step state (collective)
            R kk[NZ];
            field(z, kk, vcur, scur);
        }
        if (live && result == 0 && o.event != PMP_EVENT_NONE && pol.inside(vcur, scur, o.event_level)) result = PMP_RESULT_EVENT;
        if (live && result == 0 && nan_state) result = PMP_RESULT_NAN;
    }
}

template <class Sys>
int launch_forward_tc(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<float>& io,
                      const NnCostate<float, Sys::NX, kTcW>& pol, cudaStream_t st) {
    const size_t smem = tc_smem_bytes<Sys::NX>(pol.n_layers);
    if (smem > 227 * 1024 || pol.n_layers < 1 || pol.n_layers > 3 || pol.n_nets > kNnMaxNets) return -1000;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // one CTA per SM (shared memory); each owns a queue of `chunk` trajectories for its 128 slots
    long long chunk = ((n + sms - 1) / sms + 127) / 128 * 128;
    if (chunk < 128) chunk = 128;
    if (const char* e = getenv("B200PMP_NN_CHUNK")) chunk = (atoll(e) + 127) / 128 * 128 > 0 ? (atoll(e) + 127) / 128 * 128 : chunk;  // tuning
    int refill_min = kTcRefillMin;
    if (const char* e = getenv("B200PMP_NN_REFILL")) refill_min = atoi(e) > 0 ? atoi(e) : refill_min;  // tuning
    const unsigned grid = (unsigned)((n + chunk - 1) / chunk);
#define PMP_TCGO(METHOD)                                                                                            \
    do {                                                                                                            \
        using TB = Tab<METHOD>;                                                                                     \
        auto kern = forward_tc_kernel<Sys, METHOD>;                                                                 \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return (int)e;                                                                        \
        kern<<<grid, 128, smem, st>>>(Sys::make(p), make_solve_opts<float, TB>(o, TB::ORDER, Sys::NX), io, pol, chunk, refill_min); \
        return (int)cudaGetLastError();                                                                             \
    } while (0)
    if (o.method == PMP_METHOD_RK4) PMP_TCGO(PMP_METHOD_RK4);
    if (o.method == PMP_METHOD_TSIT5) PMP_TCGO(PMP_METHOD_TSIT5);
    if (o.method == PMP_METHOD_DOPRI5) PMP_TCGO(PMP_METHOD_DOPRI5);
#undef PMP_TCGO
    return -1000;
}

}  // namespace pmp
