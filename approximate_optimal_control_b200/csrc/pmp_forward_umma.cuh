// pmp_forward_umma.cuh -- forward closed-loop simulation under the value-network-ensemble costate law with the layer products
// on the 5th-generation TENSOR CORES (tcgen05.mma, accumulators and A operands in tensor memory): fp32 ensembles of width 64,
// up to 8 members (levelsets.py:838-933, eval_utils.py:10-45).
//
// Structure (the cluster protocol of pmp_forward_cluster.cuh, the sweep replaced):
//   * a THREAD-BLOCK CLUSTER of C = members CTAs carries 128 trajectory slots; CTA r keeps member r's weights in shared memory
//     for the whole kernel, as K-major UMMA images split into TF32 hi / lo parts -- every 64 x 64 matrix twice (W for the
//     forward product, W^T for the backward one: 4 x 16 KB per layer), plus the first layer (K = 8) and its transpose (N = 16);
//   * the 128 slots are the 128 LANES of tensor memory: thread r owns row r of every operand.  Per layer the threads read
//     their accumulator row (tcgen05.ld), apply bias + softplus on the CUDA cores, split the activations into hi / lo and
//     write them back as the next product's A operand (tcgen05.st); softplus' is parked in tensor memory for the backward
//     sweep (192 + 64 l columns).  One elected thread issues the 3 x K/8 tcgen05.mma of a product (a_lo.w_hi + a_hi.w_lo +
//     a_hi.w_hi, fp32 accumulation: fp32 quality, scripts/tc_probe.cu: 2.4e-7 relative) and commits them to an mbarrier the
//     CTA waits on.  With 256 threads the two halves of the CTA share a row: columns 0-31 / 32-63;
//   * per right hand side: 1 + (L-1) forward products, L-1 backward products, the input gradient (N = 16), then the members'
//     (gradient, value) meet through DISTRIBUTED SHARED MEMORY: every CTA PUSHES its row results into a slot of every peer's
//     exchange buffer with st.async (remote stores that count bytes on the DESTINATION's mbarrier: no cluster barrier, no
//     fence; remote loads cost ~1 us each and were 45 % of the first version), waits for its own barrier to have received
//     the C x 3.5 KB of the evaluation, then adds up its local copy in rank order -- every CTA holds bit-identical costates
//     and takes identical steps.  The buffer is double-buffered by evaluation parity: nobody can be two evaluations ahead
//     of a peer, because evaluation n+1's data of a peer only leaves after that peer has consumed evaluation n;
//   * slots are refilled from the cluster's queue as trajectories end (same rounds as the cluster kernel).
// Tensor-memory map (columns, 128 lanes x 32 bit each): D 0-63 (a_hi.w_hi + a_lo.w_hi) | D2 64-127 (a_hi.w_lo) | A_hi 128-191 |
// A_lo 192-255 | softplus' of layer l at 256 + 64 l.
#pragma once
#include <cooperative_groups.h>

#include "pmp_forward.cuh"
#include "pmp_umma.cuh"

namespace pmp {
namespace cg = cooperative_groups;

constexpr int kUmW = 64;          // hidden width served by this kernel
constexpr int kUmRefillMin = 16;  // idle slots that trigger a refill round (each costs one extra law evaluation)
constexpr uint32_t kUmColD = 0, kUmColD2 = 64, kUmColAh = 128, kUmColAl = 192, kUmColSg = 256, kUmCols = 512;

// shared-memory image of one member (float offsets; every matrix image is 16-byte aligned)
struct UmLayout {
    int L;  // hidden layers, 1..3
    __host__ __device__ int w0() const { return 0; }                                                          // first layer, fp32 [8][64] (rows >= NX zero)
    __host__ __device__ int wf(int l, int part) const { return 512 + (l - 1) * 16384 + part * 4096; }         // B(n = out, k = in)
    __host__ __device__ int wb(int l, int part) const { return 512 + (l - 1) * 16384 + 8192 + part * 4096; }  // B(n = in, k = out)
    __host__ __device__ int bias(int l) const { return 512 + (L - 1) * 16384 + l * kUmW; }
    __host__ __device__ int wout() const { return bias(L); }
    __host__ __device__ int bout() const { return wout() + kUmW; }
    __host__ __device__ int part() const { return bout() + 4; }           // [2 parities][NX + 1][2][128] partial gradients / values of the two column halves
    __host__ __device__ int xch(int NX) const { return part() + 2 * (NX + 1) * 256; }  // [2 parities][8 ranks][NX + 1][128] exchange buffer
    __host__ __device__ int total(int NX) const { return xch(NX) + 2 * 8 * (NX + 1) * 128; }
};

// what the law needs besides the state: lives in registers / constant bank of the caller
template <int NX> struct UmCtx {
    float* sm;                  // the member image (UmLayout)
    unsigned long long* bar;    // completion barrier of the tensor core
    uint32_t tbase;             // tensor memory base
    uint32_t phase;             // parity of the next completion
    uint32_t xcnt;              // exchanges so far (buffer = count & 1, barrier phase = (count >> 1) & 1)
    unsigned long long* xbar;   // [2] arrival barriers of the two exchange buffers
    int rank;                   // this CTA's rank = its member
    int L, C, E;
    float x_mean[NX], x_scale[NX], g_scale[NX], v_mean, v_std, inv_e;
};

// softplus and its derivative from three MUFU operations (ex2, rcp, lg2 in their .ftz.approx forms: the arguments are
// e = exp(-|a|) in (0, 1] -- a denormal e may flush to zero, softplus then returns max(a, 0) exactly as it should to fp32
// accuracy -- and 1 + e in [1, 2]); absolute error ~1e-7 on values of order one
__device__ __forceinline__ void um_act(float a, float& sp, float& sg) {
    float e, d, lg;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * fabsf(a)));
    const float y = 1.0f + e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(y));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(y));
    sp = fmaf(0.6931471805599453f, lg, fmaxf(a, 0.0f));
    sg = a >= 0.0f ? d : 1.0f - d;  // (1 - d = e d)
}

// a = hi + lo with hi = a rounded to TF32's 11 significant bits (Veltkamp splitting: three FMA-pipe operations, where
// cvt.rna.tf32 is emulated with five integer ones) and lo = a - hi exactly; the tensor core reads lo's leading 11 bits:
// |a - hi - lo_read| <= 2^-21 |a|
__device__ __forceinline__ void um_split(float a, uint32_t& hi, uint32_t& lo) {
    const float g = __fmul_rn(a, 8193.0f);       // 2^13 + 1 (no contraction into the sums below)
    const float h = __fadd_rn(g, __fsub_rn(a, g));
    hi = __float_as_uint(h);
    lo = __float_as_uint(__fsub_rn(a, h));
}

// everybody has staged its operands -> one thread issues the product -> everybody waits for the tensor core
template <int N, int K>
__device__ __forceinline__ void um_product(uint32_t tbase, unsigned long long* bar, uint32_t& phase, const float* b_hi, const float* b_lo) {
    using namespace umma;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        // (b_lo == b_hi + N K floats: the two images form one [B_hi | B_lo] operand of 2N rows)
        issue_3xtf32_wide<N, K>(tbase + kUmColD, tbase + kUmColAh, tbase + kUmColAl, smem_u32(b_hi), make_idesc_tf32(128, 2 * N),
                                make_idesc_tf32(128, N));
        umma_commit(bar);
    }
    if (!mbar_wait(bar, phase)) __trap();  // a lost completion must not hang the GPU
    phase ^= 1u;
    tc_fence_after();
}

// lambda(x), v_mean, v_std of the ensemble for the calling thread's row; collective over the cluster.  NWG = warpgroups per CTA.
// Not inlined: the integrator calls it once per stage.
template <int NX, int NWG>
__device__ __noinline__ void um_law(UmCtx<NX>& c, const float* x, float* lam, float& vmean, float& vstd) {
    using namespace umma;
    constexpr int CH = kUmW / NWG;  // columns per thread
    cg::cluster_group cluster = cg::this_cluster();
    const int r = threadIdx.x & 127, wg = threadIdx.x >> 7, c0 = wg * CH;
    const uint32_t lane_base = c.tbase + ((uint32_t)(threadIdx.x & 96) << 16);
    const UmLayout lay{c.L};
    extern __shared__ __align__(128) unsigned char um_smem[];  // (named here so that the loads below are LDS, not generic)
    float* const smw = reinterpret_cast<float*>(um_smem);
    const float* sm = smw;
    const int L = c.L;

    const int C = c.C;
    // ---- the first layer and the input gradient run on the CUDA cores (K = NX: 6 FMAs per unit from broadcast LDS.128 of the
    // fp32 weights; as tensor-core products each cost a full stage -> issue -> wait round trip of ~2000 cycles)
    float xn[NX], gxp[NX];
#pragma unroll
    for (int q = 0; q < NX; q++) { xn[q] = (x[q] - c.x_mean[q]) * c.x_scale[q]; gxp[q] = 0.0f; }
    const float* w0 = sm + lay.w0();
    // G = dV/d(pre-activation of layer 0) for 16 of this thread's columns: its share of the input gradient
    auto input_grad = [&](int col, const float (&g)[16]) {
#pragma unroll
        for (int q = 0; q < NX; q++) {
            float wr[16];
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) *reinterpret_cast<float4*>(wr + jj) = *reinterpret_cast<const float4*>(w0 + q * kUmW + col + jj);
#pragma unroll
            for (int jj = 0; jj < 16; jj++) gxp[q] = fmaf(g[jj], wr[jj], gxp[q]);
        }
    };
    // ---- forward sweep: D = pre-activations of layer l >= 1 (without bias)
    float vpart = 0.0f;
    for (int l = 0; l < L; l++) {
        const bool last = l == L - 1;
        const float* bias = sm + lay.bias(l);
        const float* wo = sm + lay.wout();
#pragma unroll
        for (int cc = 0; cc < CH; cc += 16) {
            const int col = c0 + cc;
            uint32_t d[16], hi[16], lo[16], sg[16];
            float bv[16], wv[16], g[16];
            uint32_t d2[16];
            if (l > 0) {
                tmem_ld16(lane_base + kUmColD + col, d);
                tmem_ld16(lane_base + kUmColD2 + col, d2);
            }
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4) {
                *reinterpret_cast<float4*>(bv + jj) = *reinterpret_cast<const float4*>(bias + col + jj);
                if (last) *reinterpret_cast<float4*>(wv + jj) = *reinterpret_cast<const float4*>(wo + col + jj);
            }
            if (l == 0) {
#pragma unroll
                for (int q = 0; q < NX; q++) {
                    float wr[16];
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4) *reinterpret_cast<float4*>(wr + jj) = *reinterpret_cast<const float4*>(w0 + q * kUmW + col + jj);
#pragma unroll
                    for (int jj = 0; jj < 16; jj++) bv[jj] = fmaf(xn[q], wr[jj], bv[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < 16; jj++) d[jj] = 0u;  // (+0.0f)
            } else {
                tmem_wait_ld();
#pragma unroll
                for (int jj = 0; jj < 16; jj++) d[jj] = __float_as_uint(__uint_as_float(d[jj]) + __uint_as_float(d2[jj]));
            }
#pragma unroll
            for (int jj = 0; jj < 16; jj++) {
                float h, sd;
                um_act(__uint_as_float(d[jj]) + bv[jj], h, sd);
                if (last) {  // value, and the backward sweep's first operand G = w_out (.) softplus'
                    vpart = fmaf(h, wv[jj], vpart);
                    h = wv[jj] * sd;
                }
                g[jj] = h;
                um_split(h, hi[jj], lo[jj]);
                sg[jj] = __float_as_uint(sd);
            }
            if (last && l == 0) {  // one hidden layer: G is already the gradient at the first layer
                input_grad(col, g);
            } else {
                tmem_st16(lane_base + kUmColAh + col, hi);
                tmem_st16(lane_base + kUmColAl + col, lo);
                if (!last) tmem_st16(lane_base + kUmColSg + 64 * l + col, sg);
            }
        }
        if (!(last && l == 0)) tmem_wait_st();
        if (!last) um_product<64, 64>(c.tbase, c.bar, c.phase, sm + lay.wf(l + 1, 0), sm + lay.wf(l + 1, 1));
    }
    // ---- backward sweep: G_{l-1} = (G_l . W_l^T) (.) softplus'_{l-1}; G_0 goes straight into the input gradient
    for (int l = L - 1; l >= 1; l--) {
        um_product<64, 64>(c.tbase, c.bar, c.phase, sm + lay.wb(l, 0), sm + lay.wb(l, 1));
#pragma unroll
        for (int cc = 0; cc < CH; cc += 16) {
            const int col = c0 + cc;
            uint32_t d[16], sg[16], hi[16], lo[16];
            float g[16];
            uint32_t d2[16];
            tmem_ld16(lane_base + kUmColD + col, d);
            tmem_ld16(lane_base + kUmColD2 + col, d2);
            tmem_ld16(lane_base + kUmColSg + 64 * (l - 1) + col, sg);
            tmem_wait_ld();
#pragma unroll
            for (int jj = 0; jj < 16; jj++) g[jj] = (__uint_as_float(d[jj]) + __uint_as_float(d2[jj])) * __uint_as_float(sg[jj]);
            if (l == 1) {
                input_grad(col, g);
            } else {
#pragma unroll
                for (int jj = 0; jj < 16; jj++) um_split(g[jj], hi[jj], lo[jj]);
                tmem_st16(lane_base + kUmColAh + col, hi);
                tmem_st16(lane_base + kUmColAl + col, lo);
            }
        }
        if (l > 1) tmem_wait_st();
    }
    // the two column halves of a row meet in shared memory (fixed order: both warpgroups hold identical results)
    float v = vpart;
    float gx[NX];
#pragma unroll
    for (int q = 0; q < NX; q++) gx[q] = gxp[q];
    if (NWG > 1) {
        float* pt = smw + lay.part() + (c.xcnt & 1u) * (NX + 1) * 256;  // (by evaluation parity: with one hidden layer this is the only barrier)
#pragma unroll
        for (int q = 0; q < NX; q++) pt[q * 256 + wg * 128 + r] = gxp[q];
        pt[NX * 256 + wg * 128 + r] = vpart;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NX; q++) gx[q] = pt[q * 256 + r] + pt[q * 256 + 128 + r];
        v = pt[NX * 256 + r] + pt[NX * 256 + 128 + r];
    }
    v += sm[lay.bout()];
    // ---- the members meet: push (scaled gradient, value) of this row into slot `rank` of every CTA's buffer, one barrier,
    // then add up the local copy in rank order
    constexpr int XS = (NX + 1) * 128;
    const uint32_t xb = c.xcnt & 1u, xph = (c.xcnt >> 1) & 1u;
    c.xcnt++;
    float* xbuf = smw + lay.xch(NX) + xb * 8 * XS;
    unsigned long long* xbar = c.xbar + xb;
    if (threadIdx.x == 0) mbar_arrive_expect_tx(xbar, (uint32_t)(C * XS * 4));
    {
        uint32_t val[NX + 1];
#pragma unroll
        for (int q = 0; q < NX; q++) val[q] = __float_as_uint(gx[q] * c.g_scale[q]);
        val[NX] = __float_as_uint(fmaf(c.v_std, v, c.v_mean));
        constexpr int PER = 8 / NWG;  // destination ranks served by one warpgroup
        const uint32_t slot = smem_u32(xbuf + c.rank * XS + r), lbar = smem_u32(xbar);
#pragma unroll
        for (int dd = 0; dd < PER; dd++) {
            const int d = wg * PER + dd;
            if (d < C) {
                const uint32_t dst = mapa_u32(slot, (uint32_t)d), dbar = mapa_u32(lbar, (uint32_t)d);
#pragma unroll
                for (int q = 0; q <= NX; q++) st_async_b32(dst + q * 512, val[q], dbar);
            }
        }
    }
    if (!mbar_wait(xbar, xph)) __trap();
    float ve[8], sum = 0.0f;
#pragma unroll
    for (int q = 0; q < NX; q++) lam[q] = 0.0f;
    if (C == 8) {  // the usual ensemble: straight sums
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
#pragma unroll
            for (int q = 0; q < NX; q++) lam[q] += xbuf[rr * XS + q * 128 + r];
            ve[rr] = xbuf[rr * XS + NX * 128 + r];
            sum += ve[rr];
        }
    } else {
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
            ve[rr] = 0.0f;
            if (rr < C) {
#pragma unroll
                for (int q = 0; q < NX; q++) lam[q] += xbuf[rr * XS + q * 128 + r];
                ve[rr] = xbuf[rr * XS + NX * 128 + r];
                sum += ve[rr];
            }
        }
    }
    const float inv_e = c.inv_e;
    vmean = sum * inv_e;
    float var = 0.0f;
#pragma unroll
    for (int rr = 0; rr < 8; rr++)
        if (rr < C) var = fmaf(ve[rr] - vmean, ve[rr] - vmean, var);
    vstd = sqrtf(var * inv_e);
#pragma unroll
    for (int q = 0; q < NX; q++) lam[q] *= inv_e;
}

template <class Sys, int METHOD, int NWG>
__global__ void __launch_bounds__(128 * NWG, 1)
forward_umma_kernel(const typename Sys::Consts sc, const SolveOpts<float> o, const ForwardIO<float> io,
                    const NnCostate<float, Sys::NX, kUmW> pol, const TabParams<float> tp, const long long chunk, const int refill_min) {
    using namespace umma;
    using R = float;
    using M = Math<R>;
    using TB = Tab<METHOD>;
    using Pol = NnCostate<R, Sys::NX, kUmW>;
    constexpr int NX = Sys::NX, NZ = NX + 1, S = TB::S, BLK = 128, W = kUmW;
    static_assert(NX <= 8, "the first layer is one k-step of 8");
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31, wq = (threadIdx.x >> 5) & 3, wg = threadIdx.x >> 7;
    const long long n = io.n;
    const long long cb = (long long)(blockIdx.x / C) * chunk, ce = cb + chunk < n ? cb + chunk : n;
    long long next = cb;
    const bool is_writer = rank == 0 && wg == 0;
    const int S1 = o.max_steps + 1, L = pol.n_layers;
    const UmLayout lay{L};

    extern __shared__ __align__(128) unsigned char um_smem[];
    float* sm = reinterpret_cast<float*>(um_smem);
    __shared__ __align__(8) unsigned long long s_bar, s_xbar[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_cnt[4];

    // ---- one-time setup: tensor memory, barrier, this CTA's member as UMMA images
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, kUmCols);
    if (threadIdx.x == 32) { mbar_init(&s_bar, 1); mbar_init(&s_xbar[0], 1); mbar_init(&s_xbar[1], 1); }
    {
        const R* src = pol.w + (size_t)rank * Pol::net_stride(L);
        constexpr int NT = 128 * NWG;
        for (int q = threadIdx.x; q < 8 * W; q += NT)  // first layer: plain fp32 rows (used by the CUDA cores), rows >= NX zero
            sm[lay.w0() + q] = q < NX * W ? __ldg(src + q) : 0.0f;
        if (threadIdx.x < W) sm[lay.bias(0) + threadIdx.x] = __ldg(src + NX * W + threadIdx.x);
        const R* pl = src + NX * W + W;
        for (int l = 1; l < L; l++, pl += W * W + W) {
            for (int q = threadIdx.x; q < W * W; q += NT) {  // Wl[in][out]
                const int in = q / W, out = q - in * W;
                uint32_t hi, lo;
                split_tf32(__ldg(pl + q), hi, lo);
                const int of = b_off_k<64, 64>(out, in), ob = b_off_k<64, 64>(in, out);
                sm[lay.wf(l, 0) + of] = __uint_as_float(hi);
                sm[lay.wf(l, 1) + of] = __uint_as_float(lo);
                sm[lay.wb(l, 0) + ob] = __uint_as_float(hi);
                sm[lay.wb(l, 1) + ob] = __uint_as_float(lo);
            }
            if (threadIdx.x < W) sm[lay.bias(l) + threadIdx.x] = __ldg(pl + W * W + threadIdx.x);
        }
        if (threadIdx.x < W) sm[lay.wout() + threadIdx.x] = __ldg(pl + threadIdx.x);
        if (threadIdx.x == 0) sm[lay.bout()] = __ldg(pl + W);
    }
    fence_proxy_async();  // the images are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    UmCtx<NX> ctx;
    ctx.sm = sm; ctx.bar = &s_bar; ctx.tbase = s_tmem; ctx.phase = 0u; ctx.xcnt = 0u; ctx.xbar = s_xbar; ctx.rank = rank; ctx.L = L; ctx.C = C; ctx.E = pol.n_nets;
#pragma unroll
    for (int q = 0; q < NX; q++) {
        ctx.x_mean[q] = pol.x_mean[q]; ctx.x_scale[q] = pol.inv_x_std[q]; ctx.g_scale[q] = pol.v_std * pol.inv_x_std[q];
    }
    ctx.v_mean = pol.v_mean; ctx.v_std = pol.v_std; ctx.inv_e = __fdiv_rn(1.0f, (float)pol.n_nets);
    cluster.sync();  // every CTA of the cluster is resident before anybody reads remote shared memory

    auto field = [&](const R* z, R* k, R& vm, R& vs) {
        R lam[NX], ld[NX], vd;
        um_law<NX, NWG>(ctx, z, lam, vm, vs);
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
        if (!io.ts || slot >= S1 || !is_writer) return;
        io.ts[i * S1 + slot] = tprev * o.dir;
#pragma unroll
        for (int q = 0; q < NX; q++) io.xs[(i * S1 + slot) * NX + q] = z[q];
        io.cost[i * S1 + slot] = z[NX];
    };
    auto is_live = [&]() { return i >= 0 && tprev < o.T1 && result == 0 && nsteps < o.max_steps; };

    while (true) {
        // ---- retire + refill round (identical decisions in every warpgroup and CTA of the cluster: they hold identical states)
        const bool live0 = is_live();
        const bool want = !live0;
        const int n_want = __syncthreads_count(want) / NWG;
        const bool any_left = next < ce;
        if (n_want == BLK || (n_want >= refill_min && any_left) || (n_want > 0 && !any_left)) {
            const unsigned bal = __ballot_sync(0xffffffffu, want);
            if (lane == 0 && wg == 0) s_cnt[wq] = __popc(bal);
            __syncthreads();
            int before = 0;
            for (int w = 0; w < wq; w++) before += s_cnt[w];
            const int my_rank = before + __popc(bal & ((1u << lane) - 1u));
            __syncthreads();
            bool fresh = false;
            if (want) {
                if (i >= 0 && is_writer) {  // retire
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
        bool at_dtmin_new = at_dtmin;
        int result_new = result;
        if (TB::EMBEDDED && o.adaptive) {
            R sumsq = R(0);
#pragma unroll
            for (int q = 0; q < NX; q++) {
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
    // ---- teardown: nobody may still read this CTA's shared memory, and the tensor memory goes back
    tc_fence_before();
    cluster.sync();
    if (threadIdx.x < 32) tmem_dealloc(ctx.tbase, kUmCols);
}

template <int NX> inline size_t um_smem_bytes(int L) {
    UmLayout lay{L};
    return (size_t)lay.total(NX) * sizeof(float);
}

template <class Sys, int METHOD, int NWG>
int launch_forward_umma_m(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<float>& io,
                          const NnCostate<float, Sys::NX, kUmW>& pol, cudaStream_t st) {
    using TB = Tab<METHOD>;
    const int C = pol.n_nets;
    const size_t smem = um_smem_bytes<Sys::NX>(pol.n_layers);
    auto kern = forward_umma_kernel<Sys, METHOD, NWG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)C, 1, 1);
    cfg.blockDim = dim3(128 * NWG, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // clusters that are resident at once (a cluster lives inside one GPC: fewer than SMs / C in general); every cluster gets a
    // queue of equal length for its 128 slots, so that there is ONE wave
    int resident = 0;
    if (cudaOccupancyMaxActiveClusters(&resident, kern, &cfg) != cudaSuccess || resident < 1) {
        cudaGetLastError();
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident = sms / C > 1 ? sms / C : 1;
    }
    long long chunk = ((n + resident - 1) / resident + 127) / 128 * 128;
    if (chunk < 128) chunk = 128;
    if (const char* ev = getenv("B200PMP_NN_CHUNK")) chunk = (atoll(ev) + 127) / 128 * 128 > 0 ? (atoll(ev) + 127) / 128 * 128 : chunk;  // tuning
    int refill_min = kUmRefillMin;
    if (const char* ev = getenv("B200PMP_NN_REFILL")) refill_min = atoi(ev) > 0 ? atoi(ev) : refill_min;  // tuning
    cfg.gridDim = dim3((unsigned)(((n + chunk - 1) / chunk) * C), 1, 1);
    e = cudaLaunchKernelEx(&cfg, kern, Sys::make(p), make_solve_opts<float, TB>(o, TB::ORDER, Sys::NX), io, pol,
                           TabParams<float>::template make<TB>(), chunk, refill_min);
    return (int)e;
}

// -1000: configuration not served by this kernel (the caller falls back to the warp-MMA / CUDA-core kernels)
template <class Sys>
int launch_forward_umma(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<float>& io,
                        const NnCostate<float, Sys::NX, kUmW>& pol, cudaStream_t st) {
    const size_t smem = um_smem_bytes<Sys::NX>(pol.n_layers);
    if (pol.n_nets < 1 || pol.n_nets > 8 || pol.n_layers < 1 || pol.n_layers > 3 || smem > 227 * 1024) return -1000;
    int nwg = 2;
    if (const char* ev = getenv("B200PMP_NN_WG")) nwg = atoi(ev) == 1 ? 1 : 2;  // tuning: warpgroups per CTA
#define PMP_UGO(METHOD) \
    return nwg == 1 ? launch_forward_umma_m<Sys, METHOD, 1>(p, o, n, io, pol, st) : launch_forward_umma_m<Sys, METHOD, 2>(p, o, n, io, pol, st)
    if (o.method == PMP_METHOD_RK4) PMP_UGO(PMP_METHOD_RK4);
    if (o.method == PMP_METHOD_TSIT5) PMP_UGO(PMP_METHOD_TSIT5);
    if (o.method == PMP_METHOD_DOPRI5) PMP_UGO(PMP_METHOD_DOPRI5);
#undef PMP_UGO
    return -1000;
}

}  // namespace pmp
