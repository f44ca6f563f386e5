// pmp_solve.cuh -- the batched PMP backward-shooting integrator (sm_100a).
//
// Replaces jax.vmap(solve_backward) (pontryagin_utils.py:488-527 as called at levelsets.py:601,1254):
// diffrax.diffeqsolve(ODETerm(f_extended), Tsit5(), t0=0, t1=-T, dt0=-0.1, y0, PIDController(rtol,
// atol, dtmin, dtmax), SaveAt(steps=True, t0=True), max_steps, throw=False), one solve per
// trajectory, with the diffrax 0.4.1 step-control semantics restated in DESIGN.md.
//
// Design (one thread per trajectory, everything in registers):
//   * y (2nx+1 dynamic scalars), the stage derivatives k1..kS and the stage state live in registers;
//     tableau coefficients are compile-time immediates (pmp_tableau.cuh), system constants sit in
//     the kernel parameter bank.  No shared memory, no local memory for the fp32 flatquad kernel
//     (the dense write-out and the value-Hessian branch add shared-memory staging, see below).
//   * Adaptive step control is per lane: every lane of a warp executes one step ATTEMPT per loop
//     iteration (accept/reject are selects, not branches), so the only divergence is trajectories
//     finishing at different attempt counts.
//   * Lane refill: a persistent grid pulls trajectory indices from a global counter in warp-sized
//     chunks.  After every attempt the warp ballots the lanes whose trajectory finished (reached
//     t1, event, max_steps, NaN); once refill_min of them wait (parked lanes run dummy attempts
//     meanwhile) it retires them (endpoint + stats written) and hands each a new trajectory, so
//     the lanes stay busy until the queue is empty.  The first-stage derivative
//     k1 = F(y_f) of a fresh trajectory (FSAL) is NOT computed in the divergent refill branch: a
//     coalesced prep kernel evaluates it for the whole batch up front (1 RHS per trajectory,
//     ~0.7 % of the work) and the refill just loads it.
//   * Time reversal as in diffrax: integrate tau = dir*t from T0 to T1 with positive steps; k holds
//     the forward-time field and the signed step h = dir*dt multiplies it.
//   * The reference's 't' leaf (dt/dt = 1) is not integrated numerically: t = t_f + dir*(tau - T0).
//     It still counts as one leaf of the RMS error norm (its own error is ~1e-9 of the tolerance).
#pragma once
#include <cstdlib>
#include <type_traits>

#include "pmp_common.cuh"
#include "pmp_hostq.h"
#include "pmp_systems.cuh"
#include "pmp_tableau.cuh"

namespace pmp {

#ifndef PMP_KBLOCK
#define PMP_KBLOCK 128
#endif
constexpr int kBlock = PMP_KBLOCK;  // threads per CTA
#ifndef PMP_KCHUNK
#define PMP_KCHUNK 32
#endif
constexpr int kChunk = PMP_KCHUNK;   // trajectory indices claimed per atomicAdd (one per lane: with 64 a warp at the queue tail,
                                     // or at the frontier of the host queue, sat on a second round of work while others idled:
                                     // 2^20 rollouts 1.190 -> 1.166 ms, 2^17 0.244 -> 0.229 ms)
constexpr long long kWatchdogCycles = 4000000000LL;  // HOSTIO: ~2 s without new input => give up (never hang the GPU)
// lanes that must wait before a warp takes the retire + refill detour (measured on 2^20 flatquad trajectories:
// fp32 1.41 ms at 1, 1.33 at 4, 1.38 at 8; fp64 4.09 / 4.03 at 3 / 4.27 at 8; vxx 5.49 / 4.90 at 6 / 5.11 at 12)
constexpr int kRefillMin32 = 4, kRefillMin64 = 3;
#ifdef PMP_EXP_NOPAD  /* experiment builds (scripts/exp_dense_prefill.py): the dense kernel writes real nodes only */
constexpr bool kExpNoPad = true;
#else
constexpr bool kExpNoPad = false;
#endif
constexpr int kRefillMinVxx = 6;  // with the vxx leaf the detour is heavier: 72 more loads, 36 more stores

// ---- value-Hessian branch (PMP_FLAG_VXX, pontryagin_utils.py:460-467) ---------------------------------
// The state gains NV = nx*nx scalars (50 in all for flatquad): with seven stage derivatives that is 288
// more values per trajectory, which no register file holds.  The base state (x, lambda, v and its k's)
// stays in registers exactly as above; V_xx and its stage derivatives live in SHARED memory, laid out
// [entry][thread] so that a warp's access to one entry is one conflict-free 128-byte row.  Per stage the
// thread forms the stage matrix in 36 registers, evaluates vxx' (sparse Jacobians, ~200 FMAs, see
// Flatquad::vxx_dot) and streams the result back.  (1 + stages) * NV reals per thread bound the CTA size:
// 64 threads (fp32) / 32 threads (fp64) = 72 KB per CTA, three CTAs per SM (six / three warps).  Shared-memory
// BANDWIDTH is what bounds this kernel (43 word accesses per entry and attempt against ~90 flops), which is why
// the symmetric storage (VXX = 2, VxxMap: 21 instead of 36 entries, 42 KB per CTA, five CTAs per SM) exists.
template <class R, int VXX> struct BlockOf { static constexpr int value = VXX ? (sizeof(R) == 4 ? 64 : 32) : kBlock; };

// Runtime-indexable Butcher tableau for the V_xx stage loop (kernel parameter bank; empty without VXX)
template <class R, int VXX> struct VTab {
    R a[7][7], b[7], e[7];
    template <class TB> static VTab make() {
        VTab t;
        for (int i = 0; i < 7; i++) {
            t.b[i] = i < TB::S ? R(TB::b(i)) : R(0);
            t.e[i] = i < TB::S ? R(TB::berr(i)) : R(0);
            for (int j = 0; j < 7; j++) t.a[i][j] = (i < TB::S && j < TB::S) ? R(TB::a(i, j)) : R(0);
        }
        return t;
    }
};
template <class R> struct VTab<R, 0> {
    template <class TB> static VTab make() { return VTab(); }
};

// a[i] for a runtime i with `a` in registers (select chain)
template <class R, int S> __device__ __forceinline__ R pick(const R (&a)[S], int i) {
    R r = a[0];
#pragma unroll
    for (int m = 1; m < S; m++) r = (i == m) ? a[m] : r;
    return r;
}

// One V_xx stage on this thread's shared-memory column (entry e of [V | k1 .. kS] at sV[e * BLK]):
//     Vs = V + sum_{j<i} (h a_ij) k_j,   k_{i+1} = vxx'(jc, Vs).
// The stage index i is a RUNTIME value and the sum over j a runtime loop (coefficients from the parameter
// bank): the integrator runs ONE copy of this code for all stages.  Fully unrolled, the seven-stage body was
// 6000 instructions (96 KB) -- three times the instruction cache -- and ran 8x slower than this.
// last (FSAL methods, i = S-1; warp-uniform): Vs is the candidate V_xx of the attempt; it is parked in the
// column of k2 and the error partial sum_{j<i} (h e_j) k_j in the column of k3 (k2..k6 are dead once the last
// stage matrix exists), so nothing of it has to stay in registers across the error estimate.
template <class Sys, class R, int BLK, int VXX>
__device__ __forceinline__ void vxx_stage(const typename Sys::Consts& sc, const typename Sys::JacCoef& jc, R* sV, int i,
                                          bool last, R h, const VTab<R, VXX>& vt) {
    using M = Math<R>;
    constexpr int NV = VxxMap<Sys::NX, VXX>::NV;
    // compiler fences: without them nvcc forwards values stored to the columns to their later loads, i.e. tries
    // to keep all (1 + stages) * NV entries in registers and spills them to local memory
    asm volatile("" ::: "memory");
    R Vs[NV], ep[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) {
        Vs[q] = sV[q * BLK];
        ep[q] = R(0);
    }
    const R* col = sV + NV * BLK;
#pragma unroll 1
    for (int j = 0; j < i; j++, col += NV * BLK) {
        const R ha = h * vt.a[i][j], hej = h * vt.e[j];
        R kj[NV];
#pragma unroll
        for (int q = 0; q < NV; q++) {
            kj[q] = col[q * BLK];
            Vs[q] = M::fma(ha, kj[q], Vs[q]);
        }
        if (last) {
#pragma unroll
            for (int q = 0; q < NV; q++) ep[q] = M::fma(hej, kj[q], ep[q]);
        }
    }
    if (last) {
#pragma unroll
        for (int q = 0; q < NV; q++) {
            sV[(NV * 2 + q) * BLK] = Vs[q];
            sV[(NV * 3 + q) * BLK] = ep[q];
        }
    }
    R* kv = sV + NV * (1 + i) * BLK;
    Sys::template vxx_dot<VXX>(sc, jc, Vs, [&](int q, R val) { kv[q * BLK] = val; });
    asm volatile("" ::: "memory");
}

// ---- forward-time field on the integrator's vector layout y = (x[NX], lam[NX], aux) -------------
// time mode : aux = v,  k = (f, -dH/dx, -l)
// value mode: aux = t,  k = (f, -dH/dx, 1) / (dv/dt)      (experiment_orbits.py:109-111)
// UMODE: u* selection mode handed to Sys::rhs (bit 0 warp vote, bit 1 KKT form)
template <class Sys, class R, bool VALUE, int UMODE>
__device__ __forceinline__ void field(const typename Sys::Consts& sc, const R* y, R* k) {
    constexpr int NX = Sys::NX;
    R vd;
    Sys::template rhs<UMODE>(sc, y, y + NX, k, vd, k + NX);
    if (VALUE) {
        const R iv = Math<R>::div(R(1), vd);
#pragma unroll
        for (int q = 0; q < 2 * NX; q++) k[q] *= iv;
        k[2 * NX] = iv;
    } else {
        k[2 * NX] = vd;
    }
}

// rows of the public [NS][n] layout: x -> 0.., t -> NX, v -> NX+1, vx -> NX+2..
// rd(row) reads one row of the trajectory's terminal state (from global memory, or from its prefetch slot)
template <class R, int NX, bool VALUE, class Rd>
__device__ __forceinline__ void load_state_rd(Rd&& rd, R* y, R& clock0, R& aux_fixed) {
#pragma unroll
    for (int q = 0; q < NX; q++) y[q] = rd(q);
#pragma unroll
    for (int q = 0; q < NX; q++) y[NX + q] = rd(NX + 2 + q);
    const R t = rd(NX), v = rd(NX + 1);
    if (VALUE) { y[2 * NX] = t; clock0 = v; aux_fixed = R(0); }
    else { y[2 * NX] = v; clock0 = R(0); aux_fixed = t; }
}
template <class R, int NX, bool VALUE>
__device__ __forceinline__ void load_state(const R* __restrict__ y_f, long long n, long long i, R* y, R& clock0,
                                           R& aux_fixed) {
    load_state_rd<R, NX, VALUE>([&](int row) { return __ldg(y_f + row * n + i); }, y, clock0, aux_fixed);
}

// ---- prep kernel: k1 = field(y_f) for every trajectory (coalesced), and reset of the work counter
template <class Sys, class R, bool VALUE, bool ULIT, int VXX>
__global__ void __launch_bounds__(kBlock) prep_kernel(const typename Sys::Consts sc, const R* __restrict__ y_f,
                                                      R* __restrict__ k1, long long n,
                                                      unsigned long long* counter) {
    constexpr int NX = Sys::NX, ND = 2 * NX + 1, NS = 2 * NX + 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *counter = 0ull;
    if (i >= n) return;
    R y[ND], k[ND], c0, af;
    load_state<R, NX, VALUE>(y_f, n, i, y, c0, af);
    if constexpr (VXX != 0) {
        using MP = VxxMap<NX, VXX>;
        constexpr int NV = MP::NV;
        R V[NV];
#pragma unroll
        for (int q = 0; q < NV; q++) V[q] = __ldg(y_f + (NS + MP::pub(q)) * n + i);
        typename Sys::JacCoef jc;
        R vd;
        Sys::rhs_jac(sc, y, y + NX, k, vd, k + NX, jc);
        k[2 * NX] = vd;
        Sys::template vxx_dot<VXX>(sc, jc, V, [&](int q, R val) { k1[(ND + q) * n + i] = val; });  // storage order
    } else {
        field<Sys, R, VALUE, ULIT ? 0 : 2>(sc, y, k);
    }
#pragma unroll
    for (int q = 0; q < ND; q++) k1[q * n + i] = k[q];
}

static __global__ void reset_counter_kernel(unsigned long long* counter) { *counter = 0ull; }


// ---- dense ("SaveAt(steps=True)") write-out through shared memory -----------------------------------
// One thread owns one trajectory, so its accepted steps land 24 B (x), 24 B (vx) and 3 x 4 B apart
// from every other lane's.  Written directly, each warp store instruction touches 32 different
// sectors with 4 bytes each -- and, worse, every such store is a PARTIAL-sector write to a sector
// that is not in L2: with ECC the L2 must first fetch the sector from DRAM, the store occupies a
// miss slot for a DRAM round trip, the L1/LSU backs up behind it and even the refill loads of the
// other warps stall for microseconds (measured: 2.9 us per refill, 17-28 M sector fills per launch).
// So nodes are written in SECTOR-ALIGNED GROUPS instead.  Slots are numbered globally
// (g = trajectory * S1 + slot); 4 consecutive slots of a vector leaf (x, vx) and 8 of a scalar leaf
// (ts, t, v) start on a 32-byte boundary whenever g is a multiple of 4 / 8.  Each lane keeps the
// nodes of its current group in a shared-memory ring; when a lane completes a group the WARP writes
// it with one store instruction per leaf (whole sectors, no fill).  When a trajectory retires, its
// last, partial group is completed with the +inf padding that follows it anyway, and the rest of
// the tail leaves as bulk copies executed by the TMA unit (bulk_fill below).  Only the first and last sector of a
// trajectory's row can be partial (rows of S1 slots are not sector aligned).
constexpr int kGrpV = 4;  // slots per vector-leaf group
constexpr int kGrpS = 8;  // slots per scalar-leaf group

template <class R, int NX> struct StageLayout {
    static_assert((NX * sizeof(R) * kGrpV) % 32 == 0, "vector groups must be whole sectors");
    static_assert((sizeof(R) * kGrpS) % 32 == 0, "scalar groups must be whole sectors");
    static_assert(kGrpV * NX <= 32 && 3 * kGrpS <= 32, "one warp store per group");
    static constexpr int XS = (kGrpV * NX) | 1;  // per-owner stride, odd => conflict-free owner writes
    static constexpr int SS = (kGrpS * 3) | 1;
    static constexpr int ORD = 2 * 32 * sizeof(int) / sizeof(R);  // two rank -> owner-lane tables (see flush_groups)
    static constexpr int WORDS_PER_WARP = 32 * (2 * XS + SS) + ORD;
    static constexpr size_t bytes_per_warp() { return (size_t)WORDS_PER_WARP * sizeof(R); }
};

// store one node (global slot g) of this lane's trajectory in its rings
template <class R, int NX>
__device__ __forceinline__ void stage_put(R* sX, R* sVX, R* sSC, int lane, unsigned g, const R* x, const R* vx, R ts, R t, R v) {
    using SL = StageLayout<R, NX>;
    R* px = sX + lane * SL::XS + (g & (kGrpV - 1)) * NX;
    R* pv = sVX + lane * SL::XS + (g & (kGrpV - 1)) * NX;
#pragma unroll
    for (int q = 0; q < NX; q++) { px[q] = x[q]; pv[q] = vx[q]; }
    R* ps = sSC + lane * SL::SS + (g & (kGrpS - 1)) * 3;
    ps[0] = ts; ps[1] = t; ps[2] = v;
}

// the warp writes vector-leaf group G (slots 4G..4G+3) of owner lane o: one store per leaf.
// Slots outside [row_lo, row_hi) belong to a neighbouring trajectory and are skipped; slots >= valid_hi
// are past the last node and get +inf (this is how a partial last group is closed).
template <class R, int NX>
__device__ __forceinline__ void write_group_v(const DensePtrs<R>& d, const R* sX, const R* sVX, int lane, int o, unsigned G,
                                              unsigned row_lo, unsigned row_hi, unsigned valid_hi, R inf) {
    using SL = StageLayout<R, NX>;
    if (lane < kGrpV * NX) {
        const unsigned slot = G * kGrpV + lane / NX;
        if (slot >= row_lo && slot < row_hi) {
            const size_t off = (size_t)G * (kGrpV * NX) + lane;
            const bool have = slot < valid_hi;
            d.p[kDenseX][off] = have ? sX[o * SL::XS + lane] : inf;
            d.p[kDenseVx][off] = have ? sVX[o * SL::XS + lane] : inf;
        }
    }
}
// scalar-leaf group G (slots 8G..8G+7) of owner o: lanes 0-7 -> ts, 8-15 -> t, 16-23 -> v
template <class R, int NX>
__device__ __forceinline__ void write_group_s(const DensePtrs<R>& d, const R* sSC, int lane, int o, unsigned G,
                                              unsigned row_lo, unsigned row_hi, unsigned valid_hi, R inf, R inf_ts) {
    using SL = StageLayout<R, NX>;
    if (lane < 3 * kGrpS) {
        const int f = lane >> 3, j = lane & (kGrpS - 1);
        const unsigned slot = G * kGrpS + j;
        if (slot >= row_lo && slot < row_hi) {
            const R pad = f == 0 ? inf_ts : inf;
            // (select, not d.p[f]: a dynamically indexed kernel parameter is copied to LOCAL memory, and those loads
            //  were 8 % of the dense kernel's stall samples)
            R* const leaf = f == 0 ? d.p[0] : (f == 1 ? d.p[1] : d.p[2]);  // leaves 0,1,2 = ts,t,v
            leaf[slot] = slot < valid_hi ? sSC[o * SL::SS + j * 3 + f] : pad;
        }
    }
}

// ---- +inf padding by the TMA unit ------------------------------------------------------------------------
// 85 % of the dense output is padding (mean 26 of 129 slots hold nodes).  Written by the warp with 16-byte stores it
// cost ~370 instructions per step attempt (five fills per retiring trajectory, each with its own address and
// remainder logic) and every one of those stores went through the SM's load/store path in front of the refill's
// loads.  Instead every CTA keeps two 4 KB shared-memory blocks of +inf (and dir*inf for `ts`) and a retiring
// trajectory's tails leave as BULK ASYNC COPIES shared -> global (cp.async.bulk, SASS UBLKCP): one instruction per
// leaf, issued by one lane each, executed by the TMA unit.  Tails start on 16-byte boundaries (they begin where
// the last sector-aligned group ends); the <= 3 reals that do not fill a 16-byte unit at the end of a row are stored
// directly.
constexpr int kInfBytes = 4096;
// (L2 evict-first: the padding is never read back by this kernel and must not push the inputs of the refill out of L2)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(src_smem);
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(s),
                 "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// dst[0, len) = val for one thread: 16-byte units by bulk copies from the shared-memory block `blk` (kInfBytes of
// val), the remainder directly.  dst must be 16-byte aligned.
template <class R> __device__ __forceinline__ void bulk_fill(R* dst, int len, const R* blk, R val) {
    constexpr int VEC = 16 / sizeof(R), BLKW = kInfBytes / sizeof(R);
    const int body = len & ~(VEC - 1);
    for (int e = 0; e < body; e += BLKW) {
        const int m = body - e < BLKW ? body - e : BLKW;
        bulk_s2g(dst + e, blk, (unsigned)(m * sizeof(R)));
    }
    for (int e = body; e < len; e++) __stcs(dst + e, val);
}

// In-loop flush of the groups completed by the last attempt.  About 7 of a warp's 32 lanes complete a vector group
// and 3.4 a scalar group per attempt; written one owner at a time (write_group_*: 24 lanes, 4-byte stores) those
// loops were 40 % of the dense kernel's stall samples and a third of its instructions.  Here a group leaves as
// 16-byte chunks: CH chunks per owner and leaf, 32 / CH owners per store instruction (fp32 flatquad: a 96-byte vector
// group = 6 chunks, 5 owners at once).  Owners publish their lane in a rank table; lane (k, c) then serves chunk c of
// the k-th ready owner.  Only groups that lie wholly inside their row may come here.
template <class R, int NX>
__device__ __forceinline__ void flush_groups(const DensePtrs<R>& d, const R* sX, const R* sVX, const R* sSC, int* sOrd,
                                             int lane, bool rdy_v, bool rdy_s, unsigned g_cur) {
    using SL = StageLayout<R, NX>;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int VEC = 16 / sizeof(R);                       // reals per chunk
    constexpr int CHV = kGrpV * NX / VEC, OWV = 32 / CHV;     // vector leaves: chunks per owner, owners per instruction
    constexpr int CHL = kGrpS / VEC, CHS = 3 * CHL, OWS = 32 / CHS;  // scalar leaves: ts, t, v together
    static_assert((kGrpV * NX) % VEC == 0 && kGrpS % VEC == 0 && CHV <= 32 && CHS <= 32, "16-byte chunks");
    const unsigned mv = __ballot_sync(FULL, rdy_v), ms = __ballot_sync(FULL, rdy_s);
    if (!(mv | ms)) return;
    const unsigned lt = (1u << lane) - 1u;
    if (rdy_v) sOrd[__popc(mv & lt)] = lane;
    if (rdy_s) sOrd[32 + __popc(ms & lt)] = lane;
    __syncwarp();  // rank tables and the owners' ring stores are visible to the serving lanes
    typedef typename std::conditional<sizeof(R) == 4, float4, double2>::type V16;
    {
        const int nv = __popc(mv), k = lane / CHV, c = lane - k * CHV;
        for (int base = 0; base < nv; base += OWV) {
            const bool act = k < OWV && base + k < nv;
            const int o = act ? sOrd[base + k] : lane;
            const unsigned G = __shfl_sync(FULL, g_cur, o) / kGrpV;
            if (act) {
                R a[VEC], b[VEC];
#pragma unroll
                for (int e = 0; e < VEC; e++) { a[e] = sX[o * SL::XS + VEC * c + e]; b[e] = sVX[o * SL::XS + VEC * c + e]; }
                const size_t off = (size_t)G * (kGrpV * NX) + VEC * c;
                *reinterpret_cast<V16*>(d.p[kDenseX] + off) = *reinterpret_cast<const V16*>(a);
                *reinterpret_cast<V16*>(d.p[kDenseVx] + off) = *reinterpret_cast<const V16*>(b);
            }
        }
    }
    {
        const int ns = __popc(ms), k = lane / CHS, i = lane - k * CHS, f = i / CHL, c = i - f * CHL;
        R* const leaf = f == 0 ? d.p[0] : (f == 1 ? d.p[1] : d.p[2]);  // ts, t, v
        for (int base = 0; base < ns; base += OWS) {
            const bool act = k < OWS && base + k < ns;
            const int o = act ? sOrd[32 + base + k] : lane;
            const unsigned G = __shfl_sync(FULL, g_cur, o) / kGrpS;
            if (act) {
                R a[VEC];
#pragma unroll
                for (int e = 0; e < VEC; e++) a[e] = sSC[o * SL::SS + (VEC * c + e) * 3 + f];
                *reinterpret_cast<V16*>(leaf + (size_t)G * kGrpS + VEC * c) = *reinterpret_cast<const V16*>(a);
            }
        }
    }
    __syncwarp();  // the rings may be overwritten by the next attempt's nodes only after every chunk has been read
}

// Which two components of the integrator's vector y = (x[NX], lam[NX], aux) share a packed pair: position p of the
// pair sequence holds component PairMap<Sys>::idx(p), pairs are positions (0,1), (2,3), ...  Any pairing gives the same
// numbers; a good one makes the copies inside the field (x' = velocity components of x) land on whole pairs, so that they
// stay register aliases instead of becoming MOVs.  Flatquad: x' = (x3, x4, x5, ..) => pairs (0,1) (3,4) (2,5), same for lambda.
template <class Sys> struct PairMap {
    __host__ __device__ static constexpr int idx(int p) { return p; }
};
template <class R> struct PairMap<Flatquad<R>> {
    __host__ __device__ static constexpr int idx(int p) {
        constexpr int m[12] = {0, 1, 3, 4, 2, 5, 6, 7, 9, 10, 8, 11};
        return m[p];
    }
};

// Stage state ys = y + sum_{j<i} (h a_ij) k_j of stage i (ha[j] = h a_ij).  The 2nx dynamic components of x and lambda
// go through packed pairs (Math<R>::fma2: one FFMA2 per pair and coefficient in fp32, half the issue slots of the scalar
// form, identical rounding); the aux scalar (v, or t in value mode) does not enter the field, so only the last stage of
// an FSAL method (whose stage state is the candidate y1) forms it.
template <class Sys, class R, class TB, int ND>
__device__ __forceinline__ void stage_state(int i, bool with_aux, const R* ha, const R* y, const R (*k)[ND], R* ys) {
    using M = Math<R>;
    using PM = PairMap<Sys>;
    constexpr int AUX = ND - 1;
    static_assert(AUX % 2 == 0, "x and lambda make an even number of components");
#pragma unroll
    for (int p = 0; p < AUX; p += 2) {
        const int a = PM::idx(p), b = PM::idx(p + 1);
        typename M::V2 acc = M::mk2(y[a], y[b]);
#pragma unroll
        for (int j = 0; j < TB::S; j++)
            if (j < i && TB::a(i, j) != 0.0) acc = M::fma2(ha[j], M::mk2(k[j][a], k[j][b]), acc);
        ys[a] = acc.x; ys[b] = acc.y;
    }
    R acc = y[AUX];
    if (with_aux) {
#pragma unroll
        for (int j = 0; j < TB::S; j++)
            if (j < i && TB::a(i, j) != 0.0) acc = M::fma(ha[j], k[j][AUX], acc);
    }
    ys[AUX] = acc;
}

// Squared scaled error of the embedded pair, summed over the components of y (diffrax rms_norm numerator):
//   e_q = sum_j (h (b_j - bhat_j)) k_j[q],  scale_q = atol + rtol max(|y_q|, |y1_q|),  sum_q (e_q / scale_q)^2
// x and lambda in packed pairs (two partial sums, even and odd members), the aux scalar on its own.
template <class Sys, class R, class TB, int ND>
__device__ __forceinline__ R error_sumsq(const R* he, const R* y, const R* y1, const R (*k)[ND], R rtol, R atol) {
    using M = Math<R>;
    using PM = PairMap<Sys>;
    constexpr int AUX = ND - 1, S = TB::S;
    typename M::V2 part = M::mk2(R(0), R(0));
#pragma unroll
    for (int p = 0; p < AUX; p += 2) {
        const int a = PM::idx(p), b = PM::idx(p + 1);
        typename M::V2 e = M::mul2(he[0], M::mk2(k[0][a], k[0][b]));
#pragma unroll
        for (int j = 1; j < S; j++)
            if (TB::berr(j) != 0.0) e = M::fma2(he[j], M::mk2(k[j][a], k[j][b]), e);
        // scale = atol + rtol*max(|y0|,|y1|); fmax drops a NaN y1 (diffrax substitutes y0)
        const typename M::V2 mx = M::mk2(M::max(M::abs(y[a]), M::abs(y1[a])), M::max(M::abs(y[b]), M::abs(y1[b])));
        const typename M::V2 scale = M::fma2(rtol, mx, M::mk2(atol, atol));
        const typename M::V2 ratio = M::mk2(M::ctl_div(e.x, scale.x), M::ctl_div(e.y, scale.y));
        part = M::fma2v(ratio, ratio, part);
    }
    R e = he[0] * k[0][AUX];
#pragma unroll
    for (int j = 1; j < S; j++)
        if (TB::berr(j) != 0.0) e = M::fma(he[j], k[j][AUX], e);
    const R scale = M::fma(rtol, M::max(M::abs(y[AUX]), M::abs(y1[AUX])), atol);
    const R ratio = M::ctl_div(e, scale);
    return M::fma(ratio, ratio, part.x + part.y);
}

// ---- the integrator -----------------------------------------------------------------------------
// HOSTIO: the persistent launch behind pmp_solve_host (HostQueue, pmp_common.cuh): inputs and endpoints in the state
// dict's own layout, work released piece by piece by the host->device stream, k1 evaluated in the refill.
template <class Sys, class R, int METHOD, bool DENSE, bool VALUE, bool ULIT, int MINB, int VXX = 0, bool HOSTIO = false, bool PUSH = false>
#ifdef PMP_MAXNREG  /* tuning builds: an explicit register cap instead of the residency hint */
__global__ void __maxnreg__(PMP_MAXNREG)
#else
__global__ void __launch_bounds__((BlockOf<R, VXX>::value), MINB)
#endif
solve_kernel(const typename Sys::Consts sc, const SolveOpts<R> o, const SolveIO<R> io, const TabParams<R> tp,
             const VTab<R, VXX> vt, const HostQueue<R> hq) {
    static_assert(!HOSTIO || (!DENSE && !VXX && !VALUE), "the host queue serves time-parameterised endpoint solves");
    static_assert(!PUSH || (!DENSE && !VXX && !HOSTIO), "the fused gather pushes endpoint records");
    static_assert(!PUSH || kChunk == 32, "a completed work-queue chunk is pushed as one 32-lane row segment");
    using M = Math<R>;
    using TB = Tab<METHOD>;
    constexpr int NX = Sys::NX, ND = 2 * NX + 1, NS = 2 * NX + 2, S = TB::S;
    constexpr int AUX = 2 * NX;
    constexpr int BLK = BlockOf<R, VXX>::value;
    using MP = VxxMap<NX, VXX>;
    constexpr int NV = MP::NV;    // stored entries of V_xx
    constexpr int NVP = VXX ? NX * NX : 0;  // entries of the public layout (rows NS.., dense leaf)
    static_assert(!(VXX && VALUE), "the vxx branch exists for time-parameterised solves only");
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long n = io.n;
    const int S1 = o.max_steps + 1;
    // leaves of the RMS norm: x, t, v, vx (, vxx) in time mode (NS + NV); x, lambda, t in value mode (NS-1)
    const R inv_nleaf = VALUE ? R(1.0 / (NS - 1)) : R(1.0 / (NS + NVP));

    // per-lane trajectory state
    R y[ND], k[S][ND];
    R tprev = R(0), tnext = R(0), clock0 = R(0), aux_fixed = R(0);
    long long idx = -1;
    int nsteps = 0, nacc = 0, result = 0;
    bool at_dtmin = false, has_work = false, finished = false;
#pragma unroll
    for (int q = 0; q < ND; q++) {
        y[q] = R(0);
#pragma unroll
        for (int j = 0; j < S; j++) k[j][q] = R(0);
    }
    // warp-local slice of the work queue (uniform across the warp)
    long long w_next = 0, w_end = 0;
    bool exhausted = false;
    bool starved = false;       // HOSTIO: the next chunk of the queue has not been delivered yet
    bool chunk_ok = !HOSTIO;    // HOSTIO: the inputs of the chunk [w_next, w_end) this warp owns have landed
    long long starve_t0 = 0;    // HOSTIO: start of the current wait (watchdog)

    // dense write-out staging (shared memory, one private region per warp)
    extern __shared__ __align__(16) unsigned char pmp_smem[];
    R* sX = nullptr; R* sVX = nullptr; R* sSC = nullptr;
    int* sOrd = nullptr;
    unsigned g_cur = 0, row0 = 0;           // global slot of the newest node; first slot of the row
    bool ready_v = false, ready_s = false;  // the newest node completed a vector / scalar group
    const R* sInf = nullptr;  // DENSE: kInfBytes of +inf, then kInfBytes of dir*inf (the padding of ts)
    constexpr size_t kInfWords = DENSE ? 2 * kInfBytes / sizeof(R) : 0;
    if (DENSE) {
        R* blk = reinterpret_cast<R*>(pmp_smem);
        for (int e = threadIdx.x; e < kInfBytes / (int)sizeof(R); e += BLK) {
            blk[e] = M::inf();
            blk[kInfBytes / sizeof(R) + e] = M::inf() * o.dir;
        }
        fence_async_smem();  // generic-proxy writes -> visible to the async proxy (TMA) that reads the blocks
        __syncthreads();
        sInf = blk;
    }
    if (DENSE) {
        using SL = StageLayout<R, NX>;
        R* wbase = reinterpret_cast<R*>(pmp_smem) + kInfWords + (size_t)(threadIdx.x >> 5) * SL::WORDS_PER_WARP;
        sX = wbase; sVX = wbase + 32 * SL::XS; sSC = wbase + 64 * SL::XS;
        sOrd = reinterpret_cast<int*>(wbase + 32 * (2 * SL::XS + SL::SS));
    }
    // VXX: this thread's column of [V | k1 .. kS] (entry e at sV[e * BLK]); the newest V is a pending dense node
    R* sV = nullptr;
    R vnorm2 = R(0);
    bool new_node = false;
    if (VXX) {
        const size_t off = DENSE ? (size_t)(BLK / 32) * StageLayout<R, NX>::WORDS_PER_WARP : 0;
        sV = reinterpret_cast<R*>(pmp_smem) + kInfWords + off + threadIdx.x;
    }
    // the warp writes the pending V_xx nodes of the lanes in `mask` (slot g_cur of each): 36 contiguous reals
    auto write_vxx_nodes = [&](unsigned mask) {
        __syncwarp();
        const R* col0 = sV - lane;  // column of lane 0 of this warp
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const unsigned g = __shfl_sync(FULL, g_cur, src);
            R* dst = io.dense.p[kDenseVxx] + (size_t)g * NVP;
            for (int q = lane; q < NVP; q += 32) dst[q] = col0[MP::ent(q / NX, q % NX) * BLK + src];
        }
        __syncwarp();
    };

    while (true) {
        // =============================== retire + refill =========================================
        // Retiring and refilling is a divergent detour of a few hundred instructions for the whole warp, and with
        // ~25 attempts per trajectory some lane of a warp finishes in almost every iteration.  So finished lanes
        // are parked (they run the next attempts as dummies, `live` is false) until at least o.refill_min lanes
        // wait, or nothing else is left to do: fewer, fuller detours for a few percent of idle lanes.
        const bool fin = has_work && finished;
        unsigned want = __ballot_sync(FULL, fin || !has_work);
        const bool any_live = __any_sync(FULL, has_work && !finished);
        if (want && (__popc(want) >= o.refill_min || !any_live || exhausted)) {
            if (fin) {
                // endpoint = last accepted state (== ys[num_accepted_steps], levelsets.py:1215)
                const R clock = tprev * o.dir;  // t (time mode) or v (value mode)
                if constexpr (HOSTIO) {
                    // state-dict layout: a lane's x / vx row is NX contiguous reals
                    if (hq.out_x) {
#pragma unroll
                        for (int q = 0; q < NX; q++) hq.out_x[idx * NX + q] = y[q];
                    }
                    if (hq.out_vx) {
#pragma unroll
                        for (int q = 0; q < NX; q++) hq.out_vx[idx * NX + q] = y[NX + q];
                    }
                    if (hq.out_t) hq.out_t[idx] = aux_fixed + (clock - clock0);
                    if (hq.out_v) hq.out_v[idx] = y[AUX];
                    if (hq.out_pk) hq.out_pk[idx] = (uint32_t)nacc | ((uint32_t)nsteps << 14) | ((uint32_t)result << 28);
                    if (hq.out_na) hq.out_na[idx] = nacc;
                    if (hq.out_ns) hq.out_ns[idx] = nsteps;
                    if (hq.out_res) hq.out_res[idx] = result;
                    __threadfence();  // the record is visible device-wide before the block counter says so
                    int piece = 0;  // the piece whose range holds idx (bounds ascending, <= 32 pieces)
                    for (int c = 1; c < hq.n_pieces; c++) piece += idx >= hq.bounds[c] ? 1 : 0;
                    atomicAdd(hq.done + piece, 1u);
                }
                R* ye = io.y_end;
                if constexpr (!HOSTIO) {
#pragma unroll
                for (int q = 0; q < NX; q++) ye[q * n + idx] = y[q];
#pragma unroll
                for (int q = 0; q < NX; q++) ye[(NX + 2 + q) * n + idx] = y[NX + q];
                if (VALUE) { ye[NX * n + idx] = y[AUX]; ye[(NX + 1) * n + idx] = clock; }
                else { ye[NX * n + idx] = aux_fixed + (clock - clock0); ye[(NX + 1) * n + idx] = y[AUX]; }
                if (VXX) {
#pragma unroll
                    for (int q = 0; q < NVP; q++) ye[(NS + q) * n + idx] = sV[MP::ent(q / NX, q % NX) * BLK];
                }
                if (io.n_accepted) io.n_accepted[idx] = nacc;
                if (io.n_steps) io.n_steps[idx] = nsteps;
                if (io.result) io.result[idx] = result;
                }
                has_work = false;
                ready_v = false; ready_s = false;  // DENSE: the retirement code writes the last group
            }
            if constexpr (PUSH) {
                // ---- the gather, fused: a work-queue chunk (32 consecutive trajectory indices) is integrated by the lanes of
                // ONE warp; the lane that retires its last trajectory makes the warp copy the chunk's endpoint rows -- 128
                // contiguous bytes per row (fp32) -- from this rank's slot into the same slot of every peer's gather buffer with
                // plain stores to peer memory (NVLink).  The records travel while the rest of the shard is still being
                // integrated: when the kernel ends, the gather has ended, and no copy engine is involved.
                __syncwarp();  // the endpoint stores of the lanes that retired just now are ordered before the loads below
                // Which chunk is complete needs no counter: a chunk is done when the warp has handed out all of its indices
                // and no lane still carries one of its trajectories (parked lanes count as carrying).
                const long long cb = idx & ~(long long)(kChunk - 1);  // (meaningful on retiring lanes only)
                unsigned fm = __ballot_sync(FULL, fin);               // the lanes that retired just now (has_work is clear)
                while (fm) {
                    const int src = __ffs(fm) - 1;
                    const long long b = __shfl_sync(FULL, cb, src);
                    fm &= ~__ballot_sync(FULL, fin && cb == b);       // this chunk once, whatever the number of its retiring lanes
                    const bool carried = __any_sync(FULL, has_work && (idx & ~(long long)(kChunk - 1)) == b);
                    const bool unclaimed = w_next < w_end && (w_next & ~(long long)(kChunk - 1)) == b;
                    if (carried || unclaimed) continue;
                    const long long j = b + lane;
                    if (j < n) {
                        const R* ye = io.y_end;
                        R rec[NS];
#pragma unroll
                        for (int q = 0; q < NS; q++) rec[q] = __ldcg(ye + q * n + j);  // (L2: written by other lanes of this warp)
                        if (io.mc) {  // one store per row: the switch replicates it into every GPU's buffer (this one included)
#pragma unroll
                            for (int q = 0; q < NS; q++) mc_store(io.mc + q * n + j, rec[q]);
                        } else
#pragma unroll
                        for (int pr = 0; pr < kMaxPeers; pr++) {  // (unrolled: a run-time index would move the pointers to local memory)
                            if (pr < io.n_peers) {
                                R* dst = io.peer[pr];
#pragma unroll
                                for (int q = 0; q < NS; q++) dst[q * n + j] = rec[q];
                            }
                        }
                    }
                }
            }
            // DENSE: the +inf tails of the retiring trajectories are written after the refill below, so
            // that the refill's loads are in flight while the padding stores are issued
            const unsigned pad_mask = DENSE ? __ballot_sync(FULL, fin) : 0u;
            const unsigned pad_g = g_cur, pad_row0 = row0;  // newest node and row start of the retiring lane
            if (DENSE && VXX) {  // the last V_xx node of a retiring lane must leave its column before the refill
                write_vxx_nodes(__ballot_sync(FULL, fin && new_node));
                if (fin) new_node = false;
            }
            if (DENSE) {
                // The last group of a retiring trajectory (complete or not) is written by the retirement
                // code below: it must leave the rings before the refill reuses them.
                const R inf = M::inf();
                __syncwarp();
                unsigned rm = pad_mask;
                while (rm) {
                    const int src = __ffs(rm) - 1;
                    rm &= rm - 1;
                    const unsigned gl = __shfl_sync(FULL, pad_g, src), r0 = __shfl_sync(FULL, pad_row0, src);
                    write_group_v<R, NX>(io.dense, sX, sVX, lane, src, gl / kGrpV, r0, r0 + S1, gl + 1, inf);
                    write_group_s<R, NX>(io.dense, sSC, lane, src, gl / kGrpS, r0, r0 + S1, gl + 1, inf, inf * o.dir);
                }
                __syncwarp();
            }
            // start trajectory i on this lane
            auto start_traj = [&](long long i) {
                idx = i;
                if constexpr (HOSTIO) {
                    // L2 loads (ld.global.cg): a neighbour's earlier read may have pulled this 128-byte line into L1 when
                    // the copy engine had not yet delivered this trajectory's part of it
#pragma unroll
                    for (int q = 0; q < NX; q++) y[q] = __ldcg(hq.in_x + idx * NX + q);
                    if (hq.term) {  // LQR terminal set: lambda = P (x - x_eq), v = 1/2 (x - x_eq)' lambda
                        R dx[NX], v2 = R(0);
#pragma unroll
                        for (int q = 0; q < NX; q++) dx[q] = y[q] - __ldg(hq.term + NX * NX + q);
#pragma unroll
                        for (int q = 0; q < NX; q++) {
                            R acc = R(0);
#pragma unroll
                            for (int j = 0; j < NX; j++) acc = M::fma(__ldg(hq.term + q * NX + j), dx[j], acc);
                            y[NX + q] = acc;
                            v2 = M::fma(dx[q], acc, v2);
                        }
                        y[AUX] = R(0.5) * v2;
                    } else {
#pragma unroll
                        for (int q = 0; q < NX; q++) y[NX + q] = __ldcg(hq.in_vx + idx * NX + q);
                        y[AUX] = __ldcg(hq.in_v + idx);
                    }
                    clock0 = R(0);
                    aux_fixed = hq.in_t ? __ldcg(hq.in_t + idx) : hq.t_fill;
                    if (TB::FSAL) field<Sys, R, VALUE, ULIT ? 0 : 2>(sc, y, k[0]);  // what the prep kernel does for device batches
                } else {
                load_state<R, NX, VALUE>(io.y_f, n, idx, y, clock0, aux_fixed);
                if (TB::FSAL) {
#pragma unroll
                    for (int q = 0; q < ND; q++) k[0][q] = __ldg(io.k1 + q * n + idx);
                }
                }
                // diffrax: t0, t1, dt0 are multiplied by the direction; tnext = min(t0+dt0, t1)
                const R T0 = VALUE ? clock0 * o.dir : o.T0;
                clock0 = T0 * o.dir;  // keep the un-normalised start for the t output
                tprev = T0;
                tnext = M::min(T0 + o.dt0, o.T1);
                nsteps = 0; nacc = 0; result = 0; at_dtmin = false;
                has_work = true;
                bool bad = false;
#pragma unroll
                for (int q = 0; q < ND; q++) bad = bad || (y[q] != y[q]);
                if (VXX) {
                    R n2 = R(0);
#pragma unroll
                    for (int q = 0; q < NV; q++) {
                        const R v = __ldg(io.y_f + (NS + MP::pub(q)) * n + idx);
                        sV[q * BLK] = v;
                        n2 = M::fma(MP::offdiag(q) ? v + v : v, v, n2);
                        if (TB::FSAL) sV[(NV + q) * BLK] = __ldg(io.k1 + (ND + q) * n + idx);
                    }
                    bad = bad || (n2 != n2);
                    vnorm2 = n2;
                    new_node = DENSE;
                }
                // NaN terminal state (levelsets.py:1233): retire at once with result 4
                finished = bad || !(tprev < o.T1);
                if (bad) result = PMP_RESULT_NAN;
                if (DENSE) {  // slot 0 = the terminal state (SaveAt(t0=True))
                    row0 = (unsigned)(idx * S1);
                    g_cur = row0;
                    const R c0 = tprev * o.dir;
                    stage_put<R, NX>(sX, sVX, sSC, lane, g_cur, y, y + NX, c0, VALUE ? y[AUX] : aux_fixed, VALUE ? c0 : y[AUX]);
                    ready_v = (g_cur & (kGrpV - 1)) == kGrpV - 1;
                    ready_s = (g_cur & (kGrpS - 1)) == kGrpS - 1;
                }
            };
            // hand out new trajectories to the lanes in `want`, in rank order
            while (want && !exhausted) {
                if (w_next >= w_end) {
                    unsigned long long base = 0;
                    if constexpr (HOSTIO) {
                        // claim a chunk only if its inputs have landed; otherwise go on integrating (or, with nothing to
                        // integrate, sleep below) and ask again after the next attempt
                        unsigned long long rdy = 0, cnt = 0;
                        if (lane == 0) {
                            rdy = *(volatile const unsigned long long*)hq.ready;
                            cnt = *(volatile unsigned long long*)io.counter;
                        }
                        rdy = __shfl_sync(FULL, rdy, 0);
                        cnt = __shfl_sync(FULL, cnt, 0);
                        const unsigned long long need = cnt + kChunk < (unsigned long long)n ? cnt + kChunk : (unsigned long long)n;
                        if (cnt < (unsigned long long)n && rdy < need) { starved = true; break; }
                    }
                    if (lane == 0) base = atomicAdd(io.counter, (unsigned long long)kChunk);
                    base = __shfl_sync(FULL, base, 0);
                    w_next = (long long)base;
                    w_end = w_next + kChunk < n ? w_next + kChunk : n;
                    if (w_next >= n) { exhausted = true; break; }
                    chunk_ok = !HOSTIO;
                    if (DENSE) {
                        // the chunk's terminal states and k1 rows are needed over the next ~50 attempts: pull them
                        // into L2 now (the refill's loads otherwise miss to DRAM behind the padding's write traffic)
                        constexpr int PER = kChunk * sizeof(R) / 128;  // 128-byte lines per row of the chunk
                        for (int q = lane; q < (NS + ND) * PER; q += 32) {
                            const int row = q / PER, part = q - row * PER;
                            const R* a = row < NS ? io.y_f + (size_t)row * n : io.k1 + (size_t)(row - NS) * n;
                            if (w_next + part * (long long)(128 / sizeof(R)) < n)
                                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a + w_next + part * (128 / sizeof(R))));
                        }
                    }
                }
                if constexpr (HOSTIO) {
                    // Several warps can pass the frontier check at once and one of them then owns a chunk that has not
                    // landed yet.  It keeps the claim but starts nothing from it until it has (it must NOT spin here: its
                    // live lanes may belong to the piece the host is waiting for before it sends the next one -- pageable
                    // output buffers make the host's device->host copies blocking).
                    if (!chunk_ok) {
                        unsigned long long rdy = 0;
                        if (lane == 0) rdy = *(volatile const unsigned long long*)hq.ready;
                        rdy = __shfl_sync(FULL, rdy, 0);
                        if (rdy < (unsigned long long)w_end) { starved = true; break; }
                        chunk_ok = true;
                    }
                }
                const int avail = (int)(w_end - w_next);
                const bool mine = (want >> lane) & 1u;
                const int rank = __popc(want & ((1u << lane) - 1u));
                const bool take = mine && rank < avail;
                if (take) start_traj(w_next + rank);
                const unsigned took = __ballot_sync(FULL, take);
                want &= ~took;
                w_next += __popc(took);
            }
            if (DENSE && !kExpNoPad) {
                // +inf tails of the retired trajectories (ts: dir*inf), from the end of their last group to the end of
                // the row: lane f issues the bulk copy of leaf f (bulk_fill); every slot of the output is written once.
                __syncwarp();
                unsigned rm = pad_mask;
                const R inf = M::inf();
                constexpr int NLEAF = VXX ? 6 : 5;
                while (rm) {
                    const int src = __ffs(rm) - 1;
                    rm &= rm - 1;
                    const unsigned gl = __shfl_sync(FULL, pad_g, src), row_hi = __shfl_sync(FULL, pad_row0, src) + S1;
                    const unsigned pv = min((gl / kGrpV + 1) * kGrpV, row_hi), ps = min((gl / kGrpS + 1) * kGrpS, row_hi);
                    if (lane < NLEAF) {
                        const bool sc = lane < 3, vx = lane == 5;  // leaves 0..2 = ts, t, v; 3, 4 = x, vx; 5 = vxx
                        const size_t start = sc ? (size_t)ps : (vx ? (size_t)(gl + 1) * NVP : (size_t)pv * NX);
                        const int len = sc ? (int)(row_hi - ps) : (vx ? (int)(row_hi - gl - 1) * NVP : (int)(row_hi - pv) * NX);
                        R* const leaf = lane == 0 ? io.dense.p[0] : lane == 1 ? io.dense.p[1] : lane == 2 ? io.dense.p[2]
                                      : lane == 3 ? io.dense.p[3] : lane == 4 ? io.dense.p[4] : io.dense.p[5];
                        bulk_fill<R>(leaf + start, len, lane == 0 ? sInf + kInfBytes / sizeof(R) : sInf,
                                     lane == 0 ? inf * o.dir : inf);
                    }
                }
                if (pad_mask) bulk_commit();
            }
        }
        if (!__any_sync(FULL, has_work)) {
            if (!HOSTIO || !starved) break;
            // nothing to integrate and the next chunk has not landed yet: wait for the host->device stream
            __nanosleep(500);
            if (starve_t0 == 0) starve_t0 = clock64();
            if (clock64() - starve_t0 > kWatchdogCycles || *(volatile unsigned int*)hq.abort) {
                // give up: flag the call as failed and release the device->host stream's waits on the piece counters
                *hq.abort = 1u;
                if (lane == 0)
                    for (int c = 0; c < hq.n_pieces; c++) atomicExch(hq.done + c, 0x7fffffffu);
                return;
            }
            starved = false;
            continue;
        }
        starved = false;
        starve_t0 = 0;
        const bool live = has_work && !finished;
        if (DENSE) {
            // write the groups completed by the previous attempt or by the refill above.  This is the
            // only in-loop call site (code size: the loop body must stay inside the I-cache).
            const R inf = M::inf();
            // Groups that lie wholly inside their trajectory's row (all but the first of a row) leave with 16-byte
            // stores, several owners per instruction (flush_groups); the others take the per-owner path below.
            const bool edge_v = ready_v && (g_cur / kGrpV) * kGrpV < row0, edge_s = ready_s && (g_cur / kGrpS) * kGrpS < row0;
            flush_groups<R, NX>(io.dense, sX, sVX, sSC, sOrd, lane, ready_v && !edge_v, ready_s && !edge_s, g_cur);
            unsigned mv = __ballot_sync(FULL, edge_v), ms = __ballot_sync(FULL, edge_s);
            if (mv | ms) __syncwarp();  // the owners' ring stores must be visible to the lanes that write the group
            while (mv) {
                const int src = __ffs(mv) - 1;
                mv &= mv - 1;
                write_group_v<R, NX>(io.dense, sX, sVX, lane, src, __shfl_sync(FULL, g_cur, src) / kGrpV,
                                     __shfl_sync(FULL, row0, src), 0xffffffffu, 0xffffffffu, inf);
            }
            while (ms) {
                const int src = __ffs(ms) - 1;
                ms &= ms - 1;
                write_group_s<R, NX>(io.dense, sSC, lane, src, __shfl_sync(FULL, g_cur, src) / kGrpS,
                                     __shfl_sync(FULL, row0, src), 0xffffffffu, 0xffffffffu, inf, inf * o.dir);
            }
        }
        ready_v = false; ready_s = false;
        if (DENSE && VXX) {  // V_xx nodes accepted by the previous attempt and the slot-0 nodes of the refill
            write_vxx_nodes(__ballot_sync(FULL, new_node));
            new_node = false;
        }

        // =============================== one step attempt ========================================
        // parked / idle lanes run the attempt as dummies: give them an ordinary step length (their own interval
        // is empty, and a zero step sends the fp64 pow() and divisions of the controller down their slow paths)
        const R dt = live ? tnext - tprev : o.dt0;
        const R h = dt * o.dir;
        R y1[ND];
        R V1[(VXX && !TB::FSAL) ? NV : 1];  // VXX, non-FSAL methods (RK4): the candidate V_xx (FSAL: see vxx_stage)
        R sumsq_base = R(0);                // VXX: base leaves' share of the squared error norm
        if constexpr (VXX != 0) {
            // ---- phase A: every base stage, in registers, exactly as without the vxx leaf; one JacPrim per stage
            constexpr int NP = Sys::NPRIM;  // reals of a JacPrim (flatquad: 5; pmp_vxx_dual.cuh: the stage state, 2 nx)
            R pr[NP][S];
            int acts = 0;
            auto base_stage = [&](int i, const R* yy) {
                typename Sys::JacPrim p;
                R vd, pa[NP];
                int act;
                Sys::rhs_prim(sc, yy, yy + NX, k[i], vd, k[i] + NX, p);
                k[i][AUX] = vd;
                Sys::prim_store(p, pa, act);
#pragma unroll
                for (int m = 0; m < NP; m++) pr[m][i] = pa[m];
                acts |= act << (2 * i);
            };
            if (!TB::FSAL) base_stage(0, y);
#pragma unroll
            for (int i = 1; i < S; i++) {
                R ys[ND], ha[S];
                const bool last_fsal = TB::FSAL && i == S - 1;
#pragma unroll
                for (int j = 0; j < i; j++) ha[j] = h * tp.template a<TB>(i, j);
                stage_state<Sys, R, TB, ND>(i, last_fsal, ha, y, k, ys);
                base_stage(i, ys);
                if (last_fsal) {
#pragma unroll
                    for (int q = 0; q < ND; q++) y1[q] = ys[q];
                }
            }
            if (!TB::FSAL) {
#pragma unroll
                for (int q = 0; q < ND; q++) {
                    R acc = y[q];
#pragma unroll
                    for (int j = 0; j < S; j++)
                        if (TB::b(j) != 0.0) acc = M::fma(h * tp.template b<TB>(j), k[j][q], acc);
                    y1[q] = acc;
                }
            }
            if (TB::EMBEDDED) {  // the base leaves' error terms now, so that k2..k6 are dead during phase B
                R he[S];
#pragma unroll
                for (int j = 0; j < S; j++) he[j] = h * tp.template e<TB>(j);
                sumsq_base = error_sumsq<Sys, R, TB, ND>(he, y, y1, k, o.rtol, o.atol);
            }
            // ---- phase B: the V_xx stages, a RUNTIME loop over one copy of the stage code (see vxx_stage)
            auto coef_of = [&](int i, typename Sys::JacCoef& jc) {
                typename Sys::JacPrim p;
                R pa[NP];
#pragma unroll
                for (int m = 0; m < NP; m++) pa[m] = pick<R, S>(pr[m], i);
                Sys::prim_load(p, pa, (acts >> (2 * i)) & 3);
                Sys::jac_from_prim(sc, p, jc);
            };
#pragma unroll 1
            for (int i = (TB::FSAL ? 1 : 0); i < S; i++) {
                typename Sys::JacCoef jc;
                coef_of(i, jc);
                vxx_stage<Sys, R, BLK, VXX>(sc, jc, sV, i, TB::FSAL && i == S - 1, h, vt);
            }
            if (!TB::FSAL) {
#pragma unroll
                for (int q = 0; q < NV; q++) V1[q] = sV[q * BLK];
                const R* col = sV + NV * BLK;
#pragma unroll 1
                for (int j = 0; j < S; j++, col += NV * BLK) {
                    const R hb = h * vt.b[j];
#pragma unroll
                    for (int q = 0; q < NV; q++) V1[q] = M::fma(hb, col[q * BLK], V1[q]);
                }
            }
        } else {
        if (!TB::FSAL) field<Sys, R, VALUE, ULIT ? 1 : 2>(sc, y, k[0]);
#pragma unroll
        for (int i = 1; i < S; i++) {
            R ys[ND];
            const bool last_fsal = TB::FSAL && i == S - 1;
            // ys = y + sum_j (h a_ij) k_j : i scalar products, then i FMAs per component
            R ha[S];
#pragma unroll
            for (int j = 0; j < i; j++) ha[j] = h * tp.template a<TB>(i, j);
            stage_state<Sys, R, TB, ND>(i, last_fsal, ha, y, k, ys);
            field<Sys, R, VALUE, ULIT ? 1 : 2>(sc, ys, k[i]);
            if (last_fsal) {
#pragma unroll
                for (int q = 0; q < ND; q++) y1[q] = ys[q];
            }
        }
        if (!TB::FSAL) {
            R hb[S];
#pragma unroll
            for (int j = 0; j < S; j++) hb[j] = h * tp.template b<TB>(j);
#pragma unroll
            for (int q = 0; q < ND; q++) {
                R acc = y[q];
#pragma unroll
                for (int j = 0; j < S; j++)
                    if (TB::b(j) != 0.0) acc = M::fma(hb[j], k[j][q], acc);
                y1[q] = acc;
            }
        }
        }

        // ---- error estimate + I-controller (diffrax PIDController, pcoeff=dcoeff=0) --------------
        bool keep = true;
        // an ACCEPTED step (force-accept at dtmin) whose scaled error is not finite means the state
        // has left the representable range / turned NaN: the trajectory retires with result 4
        bool step_nonfinite = false;
        R next_t0, next_t1;
        if (TB::EMBEDDED && o.adaptive) {
            R sumsq = sumsq_base;
            R he[S];
#pragma unroll
            for (int j = 0; j < S; j++) he[j] = h * tp.template e<TB>(j);
            if constexpr (VXX == 0) sumsq = error_sumsq<Sys, R, TB, ND>(he, y, y1, k, o.rtol, o.atol);
            if constexpr (VXX) {  // fifth leaf of the norm (embedded methods are FSAL: see vxx_stage)
                static_assert(!VXX || !TB::EMBEDDED || (TB::FSAL && S >= 4), "column reuse needs an FSAL method");
                R part[4] = {R(0), R(0), R(0), R(0)};  // four partial sums: a single chain is latency bound
#pragma unroll
                for (int q = 0; q < NV; q++) {
                    const R e = M::fma(he[S - 1], sV[(NV * S + q) * BLK], sV[(NV * 3 + q) * BLK]);
                    const R scale = M::fma(o.rtol, M::max(M::abs(sV[q * BLK]), M::abs(sV[(NV * 2 + q) * BLK])), o.atol);
                    const R ratio = M::ctl_div(e, scale);
                    part[q & 3] = M::fma(MP::offdiag(q) ? ratio + ratio : ratio, ratio, part[q & 3]);
                }
                sumsq += (part[0] + part[1]) + (part[2] + part[3]);
            }
            // integrate.py maps a NaN error estimate to +inf (=> reject, shrink by factormin); a NaN
            // in any component makes the whole sum NaN, so one test on the sum is equivalent
            sumsq = (sumsq != sumsq) ? M::inf() : sumsq;
            const R msq = sumsq * inv_nleaf;  // scaled_error^2 (rms_norm over all leaves)
            step_nonfinite = !(msq < M::inf());
            keep = (msq < R(1)) || at_dtmin;
            // factor = clip(safety * scaled_error^(-1/order), keep ? 1 : factormin, factormax)
            R factor = o.safety * M::pow_neg_inv(msq, R(0.5) * o.inv_order);
            factor = M::min(M::max(factor, keep ? R(1) : o.factormin), o.factormax);
            R dtn = dt * factor;
            if (o.use_dtmax) dtn = M::min(dtn, o.dtmax);
            bool dt_min_hit = false;
            if (o.use_dtmin) {
                dt_min_hit = dtn < o.dtmin;
                at_dtmin = dtn <= o.dtmin;
                dtn = M::max(dtn, o.dtmin);
            } else {
                at_dtmin = false;
            }
            if (live && !o.force_dtmin && dt_min_hit) result = PMP_RESULT_DT_MIN;
            next_t0 = keep ? tnext : tprev;
            next_t1 = next_t0 + dtn;
        } else {  // diffrax ConstantStepSize: next interval as long as the last one
            next_t0 = tnext;
            next_t1 = tnext + dt;
        }
        if (live) {
            // integrate.py: tprev = min(tprev, t1); tnext = _clip_to_end(tprev, tnext, t1, keep)
            tprev = M::min(next_t0, o.T1);
            if (next_t1 > o.T1 - M::clip_tol()) next_t1 = keep ? o.T1 : M::fma(R(0.5), o.T1 - tprev, tprev);
            tnext = next_t1;
            nsteps += 1;
            bool nan_state = false;
            if (keep) {
                nacc += 1;
                bool bad = step_nonfinite;
#pragma unroll
                for (int q = 0; q < ND; q++) {
                    y[q] = M::copy(y1[q], o.one);
                    if (!(TB::EMBEDDED)) bad = bad || (y1[q] != y1[q]);
                    if (TB::FSAL) k[0][q] = M::copy(k[S - 1][q], o.one);
                }
                if constexpr (VXX) {
                    R np[4] = {R(0), R(0), R(0), R(0)};
#pragma unroll
                    for (int q = 0; q < NV; q++) {
                        const R v1 = TB::FSAL ? sV[(NV * 2 + q) * BLK] : V1[q];
                        sV[q * BLK] = v1;
                        np[q & 3] = M::fma(MP::offdiag(q) ? v1 + v1 : v1, v1, np[q & 3]);
                        if (TB::FSAL) sV[(NV + q) * BLK] = sV[(NV * S + q) * BLK];
                    }
                    const R n2 = (np[0] + np[1]) + (np[2] + np[3]);
                    vnorm2 = n2;
                    if (!(TB::EMBEDDED)) bad = bad || (n2 != n2);
                    new_node = DENSE && nacc < S1;
                }
                if (DENSE && nacc < S1) {
                    const R clock = tprev * o.dir;
                    g_cur = row0 + (unsigned)nacc;
                    stage_put<R, NX>(sX, sVX, sSC, lane, g_cur, y, y + NX, clock,
                                     VALUE ? y[AUX] : aux_fixed + (clock - clock0), VALUE ? clock : y[AUX]);
                    ready_v = (g_cur & (kGrpV - 1)) == kGrpV - 1;
                    ready_s = (g_cur & (kGrpS - 1)) == kGrpS - 1;
                }
                nan_state = bad;
            }
            // DiscreteTerminatingEvent, evaluated on the post-step state after every attempt
            if (!VALUE && result == 0 && o.event == PMP_EVENT_VALUE_GE && y[AUX] >= o.event_level)
                result = PMP_RESULT_EVENT;
            // np.linalg.norm(state.y['vxx']) > vxx_max_norm  (pontryagin_utils.py:508-512)
            if (VXX && result == 0 && o.event == PMP_EVENT_VXX_NORM && M::sqrt(vnorm2) > o.event_level)
                result = PMP_RESULT_EVENT;
            if (result == 0 && nan_state) result = PMP_RESULT_NAN;  // diffrax would idle to max_steps
            const bool more = tprev < o.T1;
            if (result == 0 && more && nsteps >= o.max_steps) result = PMP_RESULT_MAX_STEPS;
            finished = !more || result != 0;
        }
    }
    if (DENSE) bulk_wait_all();  // the CTA's +inf blocks must outlive every bulk copy that reads them
}

// ---- host-side launcher for one system ----------------------------------------------------------
template <class R, class TB> inline SolveOpts<R> make_solve_opts(const pmp_opts_t& o, int order, int nx) {
    SolveOpts<R> s;
    const double dir = (o.indep == PMP_INDEP_VALUE) ? 1.0 : ((o.t0 < o.t1) ? 1.0 : -1.0);
    s.dir = R(dir);
    s.T0 = R(o.t0 * dir);
    s.T1 = R(o.t1 * dir);
    s.dt0 = R(o.dt0 * dir);
    s.rtol = R(o.rtol); s.atol = R(o.atol);
    s.dtmin = R(o.dtmin); s.dtmax = R(o.dtmax);
    s.use_dtmin = o.dtmin > 0; s.use_dtmax = o.dtmax > 0;
    s.safety = R(o.safety); s.factormin = R(o.factormin); s.factormax = R(o.factormax);
    s.inv_order = R(1.0 / order);
    s.event_level = R(o.event_level);
    s.one = R(1);
    s.max_steps = o.max_steps; s.adaptive = o.adaptive; s.force_dtmin = o.force_dtmin; s.event = o.event;
    s.refill_min = (o.flags & PMP_FLAG_VXX) ? kRefillMinVxx : (sizeof(R) == 4 ? kRefillMin32 : kRefillMin64);
    if (nx <= 2) s.refill_min = 1;  // two-state systems: the detour is as cheap as an attempt (orbits: 0.33 ms at 1, 0.36 at 3)
    if (const char* e = getenv("B200PMP_REFILL_MIN")) s.refill_min = atoi(e);  // tuning knob, 1 = retire at once
    if (s.refill_min < 1) s.refill_min = 1;
    if (s.refill_min > 32) s.refill_min = 32;
    return s;
}

struct LaunchPlan {
    int grid_solve, grid_prep, sm_count, blocks_per_sm;
};

// workspace layout: [counter: 256 B][k1: ND*n reals]
inline size_t workspace_bytes_for(int nd, long long n, size_t real_size) {
    return 256 + (size_t)nd * (size_t)n * real_size;
}

// the fused gather of pmp_solve_batch_push_cuda: where this rank's records go on the peers
struct PushArgs {
    void* peer[kMaxPeers];
    int n_peers;
    void* mc;  // multicast address of this rank's block, or NULL
};

template <class Sys, class R, int METHOD, bool DENSE, bool VALUE, bool ULIT, int MINB, int VXX = 0, bool PUSH = false>
int launch_one(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
               const pmp_dense_t* dense, int32_t* n_acc, int32_t* n_steps, int32_t* result,
               void* workspace, cudaStream_t st, const PushArgs* push = nullptr) {
    using TB = Tab<METHOD>;
    constexpr int ND = 2 * Sys::NX + 1, BLK = BlockOf<R, VXX>::value, NV = VxxMap<Sys::NX, VXX>::NV;
    auto kern = solve_kernel<Sys, R, METHOD, DENSE, VALUE, ULIT, MINB, VXX, false, PUSH>;
    int dev = 0, sms = 0, occ = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (DENSE ? 2 * kInfBytes + (size_t)(BLK / 32) * StageLayout<R, Sys::NX>::bytes_per_warp() : 0) +
                        (size_t)BLK * NV * (1 + TB::S) * sizeof(R);
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    if (VXX) {  // shared memory bounds the residency of the vxx kernels: take the largest carve-out
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLK, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) occ = 1;
    const long long need_prep = (n + kBlock - 1) / kBlock;
    long long need = (n + BLK - 1) / BLK;
    long long cap = (long long)sms * occ;
    const int grid = (int)(need < cap ? need : cap);

    const typename Sys::Consts sc = Sys::make(p);
    SolveIO<R> io;
    io.y_f = (const R*)y_f;
    io.counter = (unsigned long long*)workspace;
    io.k1 = (const R*)((char*)workspace + 256);
    io.y_end = (R*)y_end;
    io.n_accepted = n_acc; io.n_steps = n_steps; io.result = result;
    io.n = n;
    io.n_peers = 0;
    io.mc = nullptr;
    for (int q = 0; q < kMaxPeers; q++) io.peer[q] = nullptr;
    if (PUSH) {
        if (!push || push->n_peers < 0 || push->n_peers > kMaxPeers) return (int)cudaErrorInvalidValue;
        io.n_peers = push->n_peers;
        for (int q = 0; q < push->n_peers; q++) io.peer[q] = (R*)push->peer[q];
        io.mc = (R*)push->mc;
    }
    io.dense = DensePtrs<R>{{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
    if (DENSE && dense)
        io.dense = DensePtrs<R>{{(R*)dense->ts, (R*)dense->t, (R*)dense->v, (R*)dense->x, (R*)dense->vx,
                                 VXX ? (R*)dense->vxx : nullptr}};
    if (TB::FSAL) {
        prep_kernel<Sys, R, VALUE, ULIT, VXX><<<(unsigned)need_prep, kBlock, 0, st>>>(
            sc, (const R*)y_f, (R*)((char*)workspace + 256), n, io.counter);
    } else {
        reset_counter_kernel<<<1, 1, 0, st>>>(io.counter);
    }
    const SolveOpts<R> so = make_solve_opts<R, TB>(o, TB::ORDER, Sys::NX);
    kern<<<grid, BLK, smem, st>>>(sc, so, io, TabParams<R>::template make<TB>(), VTab<R, VXX>::template make<TB>(), HostQueue<R>());
    return (int)cudaGetLastError();
}

// ---- the persistent launch of pmp_solve_host (HOSTIO) ----------------------------------------------------------------
template <class Sys, class R, int METHOD, int MINB>
int launch_hostq_one(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st) {
    using TB = Tab<METHOD>;
    auto kern = solve_kernel<Sys, R, METHOD, false, false, false, MINB, 0, true>;
    int dev = 0, sms = 0, occ = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, 0);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) occ = 1;
    const long long need = (n + kBlock - 1) / kBlock, cap = (long long)sms * occ;
    const int grid = (int)(need < cap ? need : cap);
    SolveIO<R> io = {};
    io.counter = q.counter;
    io.n = n;
    HostQueue<R> hq;
    hq.in_x = (const R*)q.in_x; hq.in_vx = (const R*)q.in_vx; hq.in_t = (const R*)q.in_t; hq.in_v = (const R*)q.in_v;
    hq.term = (const R*)q.term;
    hq.out_x = (R*)q.out_x; hq.out_vx = (R*)q.out_vx; hq.out_t = (R*)q.out_t; hq.out_v = (R*)q.out_v;
    hq.out_pk = q.out_pk; hq.out_na = q.out_na; hq.out_ns = q.out_ns; hq.out_res = q.out_res;
    hq.ready = q.ready; hq.done = q.done; hq.abort = q.abort;
    hq.n_pieces = q.n_pieces;
    for (int c = 0; c <= q.n_pieces; c++) hq.bounds[c] = q.bounds[c];
    hq.t_fill = R(o.t0);
    const SolveOpts<R> so = make_solve_opts<R, TB>(o, TB::ORDER, Sys::NX);
    kern<<<grid, kBlock, 0, st>>>(Sys::make(p), so, io, TabParams<R>::template make<TB>(), VTab<R, 0>::template make<TB>(), hq);
    return (int)cudaGetLastError();
}

// -1000: this (method, dtype) has no host-queue instantiation (the caller falls back to per-piece launches)
template <template <class> class SysT, int MINB32, int MINB64>
int launch_hostq(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const HostQueueRaw& q, cudaStream_t st) {
#define PMP_HQ(R, MINB)                                                                                                      \
    do {                                                                                                                     \
        if (o.method == PMP_METHOD_RK4) return launch_hostq_one<SysT<R>, R, PMP_METHOD_RK4, MINB>(p, o, n, q, st);           \
        if (o.method == PMP_METHOD_TSIT5) return launch_hostq_one<SysT<R>, R, PMP_METHOD_TSIT5, MINB>(p, o, n, q, st);       \
        if (o.method == PMP_METHOD_DOPRI5) return launch_hostq_one<SysT<R>, R, PMP_METHOD_DOPRI5, MINB>(p, o, n, q, st);     \
    } while (0)
    if (o.dtype == PMP_F32) PMP_HQ(float, MINB32);
    if (o.dtype == PMP_F64) PMP_HQ(double, MINB64);
#undef PMP_HQ
    return -1000;
}

// The integrator with the gather fused in (time-parameterised endpoint solves, KKT-form u*): pmp_solve_batch_push_cuda
template <template <class> class SysT, int MINB32, int MINB64>
int launch_solve_push(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end, int32_t* a, int32_t* s,
                      int32_t* r, void* ws, const PushArgs& push, cudaStream_t st) {
#define PMP_PUSH(R, MINB)                                                                                                                        \
    do {                                                                                                                                         \
        if (o.method == PMP_METHOD_RK4) return launch_one<SysT<R>, R, PMP_METHOD_RK4, false, false, false, MINB, 0, true>(p, o, n, y_f, y_end, nullptr, a, s, r, ws, st, &push);       \
        if (o.method == PMP_METHOD_TSIT5) return launch_one<SysT<R>, R, PMP_METHOD_TSIT5, false, false, false, MINB, 0, true>(p, o, n, y_f, y_end, nullptr, a, s, r, ws, st, &push);   \
        if (o.method == PMP_METHOD_DOPRI5) return launch_one<SysT<R>, R, PMP_METHOD_DOPRI5, false, false, false, MINB, 0, true>(p, o, n, y_f, y_end, nullptr, a, s, r, ws, st, &push); \
    } while (0)
    if (o.dtype == PMP_F32) PMP_PUSH(float, MINB32);
    if (o.dtype == PMP_F64) PMP_PUSH(double, MINB64);
#undef PMP_PUSH
    return -1000;
}

// PMP_FLAG_VXX instantiations (time-parameterised, KKT-form u*): a compile unit of their own
// METHODS: bit PMP_METHOD_x set = instantiate that method
// SYM_MINB32: CTAs per SM the symmetric fp32 instantiation is compiled for (a register cap: 5 suits the hand-written flatquad
// Jacobians; the forward-mode adaptor of pmp_vxx_dual.cuh needs the registers more than the residency)
template <template <class> class SysT, int METHODS = 7, int SYM_MINB32 = 5>
int launch_solve_vxx(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
                     const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    const bool dn = o.save_dense != 0, sym = (o.flags & PMP_FLAG_VXX_SYMMETRIC) != 0;
#define PMP_GOV(R, METHOD)                                                                                          \
    do {                                                                                                            \
        /* symmetric fp32: 42 KB per CTA lets five CTAs share an SM if they stay within 204 registers */            \
        if (dn && sym) return launch_one<SysT<R>, R, METHOD, true, false, false, (sizeof(R) == 4 ? SYM_MINB32 : 1), 2>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);  \
        if (dn) return launch_one<SysT<R>, R, METHOD, true, false, false, 1, 1>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);         \
        if (sym) return launch_one<SysT<R>, R, METHOD, false, false, false, (sizeof(R) == 4 ? SYM_MINB32 : 1), 2>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);       \
        return launch_one<SysT<R>, R, METHOD, false, false, false, 1, 1>(p, o, n, y_f, y_end, dense, a, s, r, ws, st);                \
    } while (0)
#define PMP_BY_METHODV(R)                                              \
    do {                                                               \
        if constexpr ((METHODS >> PMP_METHOD_RK4) & 1) if (o.method == PMP_METHOD_RK4) PMP_GOV(R, PMP_METHOD_RK4);          \
        if constexpr ((METHODS >> PMP_METHOD_TSIT5) & 1) if (o.method == PMP_METHOD_TSIT5) PMP_GOV(R, PMP_METHOD_TSIT5);    \
        if constexpr ((METHODS >> PMP_METHOD_DOPRI5) & 1) if (o.method == PMP_METHOD_DOPRI5) PMP_GOV(R, PMP_METHOD_DOPRI5); \
    } while (0)
    if (o.dtype == PMP_F32) PMP_BY_METHODV(float);
    if (o.dtype == PMP_F64) PMP_BY_METHODV(double);
#undef PMP_BY_METHODV
#undef PMP_GOV
    return -1000;
}

template <template <class> class SysT, int MINB32, int MINB64>
int launch_solve(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* y_f, void* y_end,
                 const pmp_dense_t* dense, int32_t* a, int32_t* s, int32_t* r, void* ws, cudaStream_t st) {
    const bool dn = o.save_dense != 0, val = o.indep == PMP_INDEP_VALUE;
    // the literal two-pass u* only exists for nu = 2 systems
    const bool ulit = (o.flags & PMP_FLAG_USTAR_LITERAL) != 0 && SysT<float>::NU == 2;
#define PMP_GO(R, METHOD, DENSE, VALUE, MINB)                                                                      \
    do {                                                                                                           \
        if (SysT<float>::NU == 2 && ulit)                                                                          \
            return launch_one<SysT<R>, R, METHOD, DENSE, VALUE, (SysT<float>::NU == 2), MINB>(p, o, n, y_f, y_end, dense, a, s, r, ws, st); \
        return launch_one<SysT<R>, R, METHOD, DENSE, VALUE, false, MINB>(p, o, n, y_f, y_end, dense, a, s, r, ws, st); \
    } while (0)
#define PMP_BY_FLAGS(R, METHOD, MINB)                           \
    do {                                                        \
        if (dn && val) PMP_GO(R, METHOD, true, true, MINB);     \
        if (dn && !val) PMP_GO(R, METHOD, true, false, MINB);   \
        if (!dn && val) PMP_GO(R, METHOD, false, true, MINB);   \
        PMP_GO(R, METHOD, false, false, MINB);                  \
    } while (0)
#define PMP_BY_METHOD(R, MINB)                                                   \
    do {                                                                         \
        if (o.method == PMP_METHOD_RK4) PMP_BY_FLAGS(R, PMP_METHOD_RK4, MINB);   \
        if (o.method == PMP_METHOD_TSIT5) PMP_BY_FLAGS(R, PMP_METHOD_TSIT5, MINB); \
        if (o.method == PMP_METHOD_DOPRI5) PMP_BY_FLAGS(R, PMP_METHOD_DOPRI5, MINB); \
    } while (0)
    if (o.dtype == PMP_F32) PMP_BY_METHOD(float, MINB32);
    if (o.dtype == PMP_F64) PMP_BY_METHOD(double, MINB64);
#undef PMP_BY_METHOD
#undef PMP_BY_FLAGS
#undef PMP_GO
    return -1000;
}

}  // namespace pmp
