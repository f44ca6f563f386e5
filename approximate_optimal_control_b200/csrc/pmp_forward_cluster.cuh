// pmp_forward_cluster.cuh -- forward closed-loop simulation under the value-network-ensemble costate law, one
// THREAD-BLOCK CLUSTER per 128 trajectories (levelsets.py:838-933, eval_utils.py:10-45).
//
// The ensemble's weights (8 x 35 KB in fp32) do not fit an SM's L1, and in the one-lane-per-member kernel
// (pmp_forward.cuh) every weight row is an L2 round trip.  Here CTA r of a cluster of C = min(members, 8) CTAs keeps
// the weights of ITS member(s) (r, r + C, ...) in shared memory for the whole kernel and all C CTAs carry the SAME 128
// trajectories through the same integrator arithmetic:
//   * per right hand side every thread sweeps its CTA's member(s) for its trajectory from shared memory (broadcast
//     16-byte loads, ~4 FMAs per load, nothing leaves the SM), publishes (sum of gradients, values) in its CTA's
//     exchange buffer, the cluster synchronises, and every thread adds up the C partial results through DISTRIBUTED
//     SHARED MEMORY in rank order -- so all CTAs hold bit-identical costates and take identical steps;
//   * the 128 slots of a CTA advance in lockstep (one step attempt per iteration for everybody, `live` gates every
//     update), because the cluster barrier needs all threads; a slot whose trajectory has ended is REFILLED from the
//     cluster's queue of trajectories -- in rounds (at least kClusterRefillMin waiting slots, each round costs one extra
//     sweep for the fresh trajectories' first stage), with ranks from a block-wide scan, so that every CTA of the
//     cluster takes the same decisions;
//   * CTA 0 of the cluster writes the outputs.
#pragma once
#include <cooperative_groups.h>

#include "pmp_forward.cuh"

namespace pmp {
namespace cg = cooperative_groups;

constexpr int kClusterMaxMembers = 2;  // members per CTA: up to 16 members on clusters of 8

constexpr int kClusterRefillMin = 16;  // idle slots of a CTA that trigger a refill round (each costs one extra sweep)

template <class Sys, class R, int METHOD, int W>
__global__ void __launch_bounds__(128, 1)
forward_cluster_kernel(const typename Sys::Consts sc, const SolveOpts<R> o, const ForwardIO<R> io,
                       const NnCostate<R, Sys::NX, W> pol, const TabParams<R> tp, const long long chunk, const int refill_min) {
    using M = Math<R>;
    using TB = Tab<METHOD>;
    using Pol = NnCostate<R, Sys::NX, W>;
    constexpr int NX = Sys::NX, NZ = NX + 1, S = TB::S, BLK = 128;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n = io.n;
    // this cluster's trajectories [cb, ce): 128 slots, refilled from `next` as trajectories end
    const long long cb = (long long)(blockIdx.x / C) * chunk, ce = cb + chunk < n ? cb + chunk : n;
    long long next = cb;
    const bool is_writer_cta = rank == 0;
    const int S1 = o.max_steps + 1, L = pol.n_layers, E = pol.n_nets;
    const int stride = Pol::net_stride(L);

    extern __shared__ __align__(16) unsigned char fc_smem[];
    R* sW = reinterpret_cast<R*>(fc_smem);                       // this CTA's member weights
    int my = 0;
    for (int e = rank; e < E; e += C) my++;
    R* xch = sW + (size_t)kClusterMaxMembers * stride;            // [NX + kClusterMaxMembers][BLK]
    __shared__ int s_cnt[BLK / 32];
    for (int m = 0; m < my; m++) {
        const R* src = pol.w + (size_t)(rank + m * C) * stride;
        for (int q = threadIdx.x; q < stride; q += BLK) sW[(size_t)m * stride + q] = __ldg(src + q);
    }
    __syncthreads();

    // lambda(x), v_mean, v_std: collective over the cluster -- every thread of every CTA must call it
    auto law = [&](const R* x, R* lam, R& vmean, R& vstd) {
        R xn[NX], gs[NX];
#pragma unroll
        for (int q = 0; q < NX; q++) { xn[q] = (x[q] - pol.x_mean[q]) * pol.inv_x_std[q]; gs[q] = R(0); }
        for (int m = 0; m < my; m++) {
            R gx[NX], v;
            Pol::template member_sweep<true>(sW + (size_t)m * stride, xn, L, gx, v);
#pragma unroll
            for (int q = 0; q < NX; q++) gs[q] = M::fma(gx[q], pol.v_std * pol.inv_x_std[q], gs[q]);
            xch[(NX + m) * BLK + threadIdx.x] = M::fma(pol.v_std, v, pol.v_mean);
        }
#pragma unroll
        for (int q = 0; q < NX; q++) xch[q * BLK + threadIdx.x] = gs[q];
        cluster.sync();
        R ve[kNnMaxNets], sum = R(0);
#pragma unroll
        for (int q = 0; q < NX; q++) lam[q] = R(0);
        int cnt = 0;
        for (int r = 0; r < C; r++) {  // rank order: the same additions in the same order in every CTA
            const R* rx = cluster.map_shared_rank(xch, r);
#pragma unroll
            for (int q = 0; q < NX; q++) lam[q] += rx[q * BLK + threadIdx.x];
            for (int e = r, m = 0; e < E; e += C, m++) {
                const R v = rx[(NX + m) * BLK + threadIdx.x];
                ve[cnt++] = v;
                sum += v;
            }
        }
        cluster.sync();  // every CTA has read every exchange buffer: they may be overwritten
        const R inv_e = M::div(R(1), R(E));
        vmean = sum * inv_e;
        R var = R(0);
        for (int e = 0; e < cnt; e++) var = M::fma(ve[e] - vmean, ve[e] - vmean, var);
        vstd = M::sqrt(var * inv_e);
#pragma unroll
        for (int q = 0; q < NX; q++) lam[q] *= inv_e;
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
        if (!io.ts || slot >= S1 || !is_writer_cta) return;
        io.ts[i * S1 + slot] = tprev * o.dir;
#pragma unroll
        for (int q = 0; q < NX; q++) io.xs[(i * S1 + slot) * NX + q] = z[q];
        io.cost[i * S1 + slot] = z[NX];
    };
    auto is_live = [&]() { return i >= 0 && tprev < o.T1 && result == 0 && nsteps < o.max_steps; };

    while (true) {
        // ---- retire + refill round: when enough slots wait, or nothing is live (identical decisions in every CTA of
        // the cluster: they hold identical states)
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
                if (i >= 0 && is_writer_cta) {  // retire
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
            if (TB::FSAL && __syncthreads_or(fresh)) {  // k1 of the fresh trajectories: one collective sweep for everybody
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
}

// smem bytes of one CTA
template <class R, int NX, int W> inline size_t cluster_smem_bytes(int L) {
    return ((size_t)kClusterMaxMembers * NnCostate<R, NX, W>::net_stride(L) + (size_t)(NX + kClusterMaxMembers) * 128) * sizeof(R);
}

template <class Sys, class R, int W>
int launch_forward_cluster_w(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const ForwardIO<R>& io,
                             const NnCostate<R, Sys::NX, W>& pol, cudaStream_t st) {
    const int C = pol.n_nets < 8 ? pol.n_nets : 8;
    const size_t smem = cluster_smem_bytes<R, Sys::NX, W>(pol.n_layers);
    // trajectories per cluster: 128 slots, refilled; enough clusters to fill the machine twice over, then longer queues
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long clusters_target = (long long)(2 * sms / C > 1 ? 2 * sms / C : 1);
    long long chunk = ((n + clusters_target - 1) / clusters_target + 127) / 128 * 128;
    if (chunk < 128) chunk = 128;
    if (const char* e = getenv("B200PMP_NN_CHUNK")) chunk = (atoll(e) + 127) / 128 * 128 > 0 ? (atoll(e) + 127) / 128 * 128 : chunk;  // tuning
    int refill_min = kClusterRefillMin;
    if (const char* e = getenv("B200PMP_NN_REFILL")) refill_min = atoi(e) > 0 ? atoi(e) : refill_min;  // tuning
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((n + chunk - 1) / chunk) * C), 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define PMP_CGO(METHOD)                                                                                             \
    do {                                                                                                            \
        using TB = Tab<METHOD>;                                                                                     \
        auto kern = forward_cluster_kernel<Sys, R, METHOD, W>;                                                      \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return (int)e;                                                                        \
        e = cudaLaunchKernelEx(&cfg, kern, Sys::make(p), make_solve_opts<R, TB>(o, TB::ORDER, Sys::NX), io, pol,     \
                               TabParams<R>::template make<TB>(), chunk, refill_min);                               \
        return (int)e;                                                                                              \
    } while (0)
    if (o.method == PMP_METHOD_RK4) PMP_CGO(PMP_METHOD_RK4);
    if (o.method == PMP_METHOD_TSIT5) PMP_CGO(PMP_METHOD_TSIT5);
    if (o.method == PMP_METHOD_DOPRI5) PMP_CGO(PMP_METHOD_DOPRI5);
#undef PMP_CGO
    return -1000;
}

}  // namespace pmp
