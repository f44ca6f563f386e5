// pmp_host.cu -- host in -> host out pipeline behind pmp_solve_host (include/b200pmp.h).
//
// The reference keeps every terminal state and every result in HOST memory (numpy / CPU jax arrays handed to
// jax.vmap(solve_backward), levelsets.py:595-601), in the array-of-structs layout of its state dict ('x' [n][nx], ...).
// A binding that starts there pays PCIe for every trajectory: 2 (2nx+2) reals + 3 ints, 132 B for flatquad fp32, i.e.
// ~2.4 ms per 2^20 trajectories one way after the other against 1.3 ms of integration.  This file cuts the batch into
// pieces and runs, per piece, host->device copies (stream s_in), layout change + integration + layout change
// (one compute stream per staging set) and device->host copies (stream s_out); both copy engines and the SMs are busy
// at the same time.  The issue cost of a piece is ~15 driver calls (the same pipeline written with torch calls costs
// 0.2 ms of CPU per piece).  Measured (B200PMP_HOST_TRACE=1, 2^20 flatquad fp32 rollouts, PCIe 5 x16): the
// device->host stream is the bottleneck -- 80 MB at 44 GB/s while the 59 MB of inputs share the link -- and it can
// only start once the first piece has been integrated, hence the small first piece of host_plan().
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>
#include <type_traits>

#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>

#include "../../include/b200pmp.h"
#include "pmp_hostq.h"

namespace pmp {
namespace {

constexpr int kSets = 3;     // staging sets in flight
constexpr int kMaxDev = 16;

struct HostPipe {
    bool ready = false;
    cudaStream_t s_in = nullptr, s_out = nullptr, s_comp[kSets] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[kSets], ev_pack[kSets], ev_comp[kSets], ev_out[kSets];
    char* buf[kSets] = {nullptr, nullptr, nullptr};
    size_t cap = 0;  // bytes per set
    uint32_t* h_packed = nullptr;  // pinned host staging of the packed counters, [n]
    size_t h_cap = 0;
    char* small = nullptr;  // whole-batch device arrays of the per-trajectory scalars: t_f, v_f, t_end, v_end, packed counters
    size_t small_cap = 0;   // bytes per array
    // persistent launch (solve_host_persistent): whole-batch leaf arrays, control block, pinned frontier values
    char* big = nullptr;
    size_t big_cap = 0;
    char* ctrl = nullptr;                       // HostCtrl in device memory
    unsigned long long* h_frontier = nullptr;   // pinned: bounds[c+1] of every piece (source of the 8-byte `ready` copies)
    cudaEvent_t ev_ctrl = nullptr;
    char *term = nullptr, *h_term = nullptr;    // LQR terminal set: P, x_eq (device / pinned staging)
    std::vector<cudaEvent_t> ev_piece;  // device->host copies of piece c have landed
};
HostPipe g_pipe[kMaxDev];
std::mutex g_mu[kMaxDev];  // one per device: calls on different GPUs of one process run concurrently

inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

// state-dict leaves (array of structs) -> extended state rows (struct of arrays), and back
template <typename R>
__global__ void pack_kernel(int nx, long long m, const R* __restrict__ x, const R* __restrict__ vx, const R* __restrict__ t,
                            const R* __restrict__ v, R t_fill, R* __restrict__ y, const R* __restrict__ term) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (term) {  // LQR terminal set (see HostQueue::term): only x crossed the link
        R v2 = R(0);
        for (int q = 0; q < nx; q++) y[q * m + i] = x[i * nx + q];
        for (int q = 0; q < nx; q++) {
            R acc = R(0);
            for (int j = 0; j < nx; j++) acc = fma(term[q * nx + j], x[i * nx + j] - term[nx * nx + j], acc);
            y[(nx + 2 + q) * m + i] = acc;
            v2 = fma(x[i * nx + q] - term[nx * nx + q], acc, v2);
        }
        y[nx * m + i] = t ? t[i] : t_fill;
        y[(nx + 1) * m + i] = R(0.5) * v2;
        return;
    }
    if ((nx & 1) == 0) {  // rows of an even number of reals: two per load (a warp covers whole 128-byte lines in nx/2 loads)
        typedef typename std::conditional<sizeof(R) == 4, float2, double2>::type V2;
        const V2* x2 = reinterpret_cast<const V2*>(x + i * nx);
        const V2* v2 = reinterpret_cast<const V2*>(vx + i * nx);
        for (int q = 0; q < nx; q += 2) {
            const V2 a = x2[q >> 1], b = v2[q >> 1];
            y[q * m + i] = a.x; y[(q + 1) * m + i] = a.y;
            y[(nx + 2 + q) * m + i] = b.x; y[(nx + 3 + q) * m + i] = b.y;
        }
    } else {
        for (int q = 0; q < nx; q++) {
            y[q * m + i] = x[i * nx + q];
            y[(nx + 2 + q) * m + i] = vx[i * nx + q];
        }
    }
    y[nx * m + i] = t ? t[i] : t_fill;
    y[(nx + 1) * m + i] = v[i];
}

// also folds the three per-trajectory counters into one word (accepted: bits 0-13, attempts: 14-27, result: 28-31;
// max_steps < 2^14) so that they cross the link as 4 bytes instead of 12; unpack_counters() undoes it on the host
constexpr int kPackBits = 14;
template <typename R>
__global__ void unpack_kernel(int nx, long long m, const R* __restrict__ y, R* __restrict__ x, R* __restrict__ vx,
                              const int32_t* __restrict__ na, const int32_t* __restrict__ ns, const int32_t* __restrict__ res,
                              uint32_t* __restrict__ packed, R* __restrict__ t_out, R* __restrict__ v_out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (t_out) t_out[i] = y[nx * m + i];
    if (v_out) v_out[i] = y[(nx + 1) * m + i];
    if (x) {
        if ((nx & 1) == 0) {
            typedef typename std::conditional<sizeof(R) == 4, float2, double2>::type V2;
            V2* x2 = reinterpret_cast<V2*>(x + i * nx);
            V2* v2 = reinterpret_cast<V2*>(vx + i * nx);
            for (int q = 0; q < nx; q += 2) {
                V2 a, b;
                a.x = y[q * m + i]; a.y = y[(q + 1) * m + i];
                b.x = y[(nx + 2 + q) * m + i]; b.y = y[(nx + 3 + q) * m + i];
                x2[q >> 1] = a; v2[q >> 1] = b;
            }
        } else {
            for (int q = 0; q < nx; q++) {
                x[i * nx + q] = y[q * m + i];
                vx[i * nx + q] = y[(nx + 2 + q) * m + i];
            }
        }
    }
    if (packed) packed[i] = (uint32_t)na[i] | ((uint32_t)ns[i] << kPackBits) | ((uint32_t)res[i] << (2 * kPackBits));
}

// (packed may BE na: the packed words of a piece can land in the n_accepted leaf itself, see pair_tvk below)
void unpack_counters(const uint32_t* packed, long long m, int32_t* na, int32_t* __restrict__ ns, int32_t* __restrict__ res) {
    constexpr uint32_t mask = (1u << kPackBits) - 1u;
    if (na && ns && res) {  // one pass: a load and three stores per trajectory
        for (long long i = 0; i < m; i++) {
            const uint32_t w = packed[i];
            na[i] = (int32_t)(w & mask);
            ns[i] = (int32_t)((w >> kPackBits) & mask);
            res[i] = (int32_t)(w >> (2 * kPackBits));
        }
        return;
    }
    if (na) for (long long i = 0; i < m; i++) na[i] = (int32_t)(packed[i] & mask);
    if (ns) for (long long i = 0; i < m; i++) ns[i] = (int32_t)((packed[i] >> kPackBits) & mask);
    if (res) for (long long i = 0; i < m; i++) res[i] = (int32_t)(packed[i] >> (2 * kPackBits));
}

#define PMP_TRY(call)                               \
    do {                                            \
        cudaError_t e_ = (call);                    \
        if (e_ != cudaSuccess) return (int)e_;      \
    } while (0)

int ensure_packed(HostPipe& P, size_t n, int pieces) {
    if (n * 4 > P.h_cap) {
        if (P.h_packed) PMP_TRY(cudaFreeHost(P.h_packed));
        P.h_packed = nullptr; P.h_cap = 0;
        PMP_TRY(cudaHostAlloc((void**)&P.h_packed, n * 4, cudaHostAllocDefault));
        P.h_cap = n * 4;
    }
    while ((int)P.ev_piece.size() < pieces) {
        cudaEvent_t e;
        PMP_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        P.ev_piece.push_back(e);
    }
    return 0;
}

int ensure_small(HostPipe& P, size_t bytes_per_array) {
    if (bytes_per_array > P.small_cap) {
        if (P.small) PMP_TRY(cudaFree(P.small));
        P.small = nullptr; P.small_cap = 0;
        PMP_TRY(cudaMalloc((void**)&P.small, 5 * up256(bytes_per_array)));
        P.small_cap = up256(bytes_per_array);
    }
    return 0;
}

int ensure(HostPipe& P, size_t need) {
    if (!P.ready) {
        PMP_TRY(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
        PMP_TRY(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
        for (auto& s : P.s_comp) PMP_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        for (int k = 0; k < kSets; k++) {
            PMP_TRY(cudaEventCreateWithFlags(&P.ev_in[k], cudaEventDisableTiming));
            PMP_TRY(cudaEventCreateWithFlags(&P.ev_pack[k], cudaEventDisableTiming));
            PMP_TRY(cudaEventCreateWithFlags(&P.ev_comp[k], cudaEventDisableTiming));
            PMP_TRY(cudaEventCreateWithFlags(&P.ev_out[k], cudaEventDisableTiming));
        }
        P.ready = true;
    }
    if (need > P.cap) {  // every earlier call has drained its streams before returning
        for (auto& b : P.buf) {
            if (b) PMP_TRY(cudaFree(b));
            b = nullptr;
        }
        P.cap = 0;
        for (auto& b : P.buf) PMP_TRY(cudaMalloc((void**)&b, need));
        P.cap = need;
    }
    return 0;
}

// ---- the persistent launch -------------------------------------------------------------------------------------
// Per-piece launches made the INTEGRATION the slowest stage of the pipeline: a piece of 2^17 trajectories is 2.3
// trajectories per resident lane, every launch ends in a tail in which most lanes idle, and the pieces' kernels took
// 0.17-0.30 ms each instead of their 0.147 ms share of the whole-batch launch (profiles/r1d_host_pipeline_trace.txt).
// Here ONE launch integrates the whole batch while the copy engines are still delivering it:
//   s_in    per piece: the leaves into whole-batch device arrays, then an 8-byte copy of the new frontier into ctrl->ready;
//   s_comp  the integrator (HOSTIO instantiation of solve_kernel): starts trajectories below *ready only, reads the leaves
//           in the state dict's layout, evaluates k1 itself, writes endpoints in the state dict's layout, counts retired
//           trajectories per piece in ctrl->done[];
//   s_out   per piece: cuStreamWaitValue32(done[c] >= size of piece c), then the copies back.
// No other kernel runs meanwhile and nothing the integrator waits for needs an SM (copy engines only), so it cannot
// deadlock; a watchdog in the kernel gives up after ~2 s without new input (ctrl->abort) rather than hang the GPU.
struct HostCtrl {
    unsigned long long counter, ready;
    unsigned int abort, pad;
    unsigned int done[kHostMaxPieces];
};

typedef int (*WaitValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
WaitValue32Fn wait_value32() {  // cuStreamWaitValue32 through the runtime's driver entry point lookup (no libcuda link)
    static WaitValue32Fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return (WaitValue32Fn)f;
    }();
    return fn;
}

// returns -1000 when this call cannot take the persistent path (the caller then uses per-piece launches)
int solve_host_persistent(HostPipe& P, const pmp_problem_t& p, const pmp_opts_t& o, long long n, const std::vector<long long>& bounds,
                          const void* x_f, const void* t_f, const void* v_f, const void* vx_f, void* x_end, void* t_end,
                          void* v_end, void* vx_end, int32_t* n_accepted, int32_t* n_steps, int32_t* result, bool pack, bool trace,
                          const void* d_term) {
    const int n_pieces = (int)bounds.size() - 1;
    const auto h0 = std::chrono::steady_clock::now();  // host clock of the trace
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count(); };
    double h_launched = 0, h_enqueued = 0, h_synced = 0;
    WaitValue32Fn wait32 = wait_value32();
    if (!wait32 || n_pieces > kHostMaxPieces || o.indep != PMP_INDEP_TIME || o.flags != 0 || o.save_dense) return -1000;
    const int nx = p.nx;
    const size_t es = o.dtype == PMP_F32 ? 4 : 8;
    // whole-batch arrays: x, vx, t, v in; x, vx, t, v out; packed word or three counters
    const size_t sv = up256((size_t)n * nx * es), ss = up256((size_t)n * es), si = up256((size_t)n * 4);
    const size_t total = 4 * sv + 4 * ss + 3 * si;
    if (total > P.big_cap) {
        if (P.big) PMP_TRY(cudaFree(P.big));
        P.big = nullptr; P.big_cap = 0;
        PMP_TRY(cudaMalloc((void**)&P.big, total));
        P.big_cap = total;
    }
    if (!P.ctrl) {
        PMP_TRY(cudaMalloc((void**)&P.ctrl, sizeof(HostCtrl)));
        PMP_TRY(cudaHostAlloc((void**)&P.h_frontier, sizeof(unsigned long long) * (kHostMaxPieces + 1), cudaHostAllocDefault));
        PMP_TRY(cudaEventCreateWithFlags(&P.ev_ctrl, cudaEventDisableTiming));
    }
    char* q = P.big;
    char *d_ix = q, *d_ivx = q + sv, *d_ox = q + 2 * sv, *d_ovx = q + 3 * sv;
    q += 4 * sv;
    char *d_it = q, *d_iv = q + ss, *d_ot = q + 2 * ss, *d_ov = q + 3 * ss;
    q += 4 * ss;
    uint32_t* d_pk = (uint32_t*)q;
    int32_t *d_na = (int32_t*)q, *d_ns = (int32_t*)(q + si), *d_res = (int32_t*)(q + 2 * si);
    HostCtrl* ctrl = (HostCtrl*)P.ctrl;

    HostQueueRaw hq = {};
    hq.in_x = d_ix; hq.in_vx = d_ivx; hq.in_t = t_f ? d_it : nullptr; hq.in_v = d_iv;
    hq.term = d_term;
    hq.out_x = x_end ? d_ox : nullptr; hq.out_vx = vx_end ? d_ovx : nullptr;
    hq.out_t = t_end ? d_ot : nullptr; hq.out_v = v_end ? d_ov : nullptr;
    hq.out_pk = pack ? d_pk : nullptr;
    hq.out_na = (!pack && n_accepted) ? d_na : nullptr;
    hq.out_ns = (!pack && n_steps) ? d_ns : nullptr;
    hq.out_res = (!pack && result) ? d_res : nullptr;
    hq.counter = &ctrl->counter; hq.ready = &ctrl->ready; hq.done = ctrl->done; hq.abort = &ctrl->abort;
    hq.n_pieces = n_pieces;
    for (int c = 0; c <= n_pieces; c++) hq.bounds[c] = bounds[c];
    for (int c = 0; c < n_pieces; c++) P.h_frontier[c] = (unsigned long long)bounds[c + 1];

    cudaStream_t sc = P.s_comp[0];
    std::vector<cudaEvent_t> tev;
    auto stamp = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tev.push_back(e);
    };
    auto drain = [&](int rc) {
        // unblock a kernel that may be waiting for input that will never arrive, then wait for everything
        const unsigned int one = 1u;
        cudaMemcpyAsync(&ctrl->abort, &one, 4, cudaMemcpyHostToDevice, P.s_in);
        cudaStreamSynchronize(P.s_in);
        cudaStreamSynchronize(sc);
        cudaStreamSynchronize(P.s_out);
        for (auto e : tev) cudaEventDestroy(e);
        cudaGetLastError();
        return rc;
    };
#undef PMP_TRY
#define PMP_TRY(call)                                  \
    do {                                               \
        cudaError_t e_ = (call);                       \
        if (e_ != cudaSuccess) return drain((int)e_);  \
    } while (0)
    stamp(P.s_in);
    PMP_TRY(cudaMemsetAsync(ctrl, 0, sizeof(HostCtrl), sc));
    PMP_TRY(cudaEventRecord(P.ev_ctrl, sc));
    {
        const int rc = solve_hostq_dispatch(p, o, n, hq, sc);
        if (rc == -1000) {
            cudaStreamSynchronize(sc);
            return -1000;
        }
        if (rc) return drain(rc);
    }
    stamp(sc);  // (recorded when the kernel has finished)
    h_launched = host_ms();
    const char *hx = (const char*)x_f, *hvx = (const char*)vx_f, *ht = (const char*)t_f, *hv = (const char*)v_f;
    PMP_TRY(cudaStreamWaitEvent(P.s_out, P.ev_ctrl, 0));
    int mode2d = 2;  // 0: one copy per leaf, 1: x+vx and t+v as two-row copies, 2: the packed counters as a third row (fp32)
    if (const char* e = getenv("B200PMP_HOST_2D")) mode2d = atoi(e);
    const bool use2d = mode2d != 0;
    const long long kMaxPitch = 0x7fffffffLL;  // cudaDeviceProp::memPitch
    const long long pitch_xvx = (x_end && vx_end) ? (long long)((char*)vx_end - (char*)x_end) : 0;
    const long long pitch_tv = (t_end && v_end) ? (long long)((char*)v_end - (char*)t_end) : 0;
    // (a pitched copy must stay inside ONE allocation on either side: leaves of separate allocations that merely lie close to
    // each other are refused by the runtime at enqueue time -- the first refusal switches the pairing off for the call)
    bool pair_xvx = use2d && pitch_xvx >= (long long)((size_t)n * nx * es) && pitch_xvx <= kMaxPitch && (long long)sv <= kMaxPitch;
    bool pair_tv = use2d && pitch_tv >= (long long)((size_t)n * es) && pitch_tv <= kMaxPitch;
    // fp32: the packed counters are a third row of the same pitched copy when the n_accepted leaf follows v at the same
    // distance; they land in that leaf and are unfolded in place
    bool pair_tvk = mode2d >= 2 && pair_tv && pack && es == 4 && n_accepted && n_steps && result &&
                          (long long)((char*)n_accepted - (char*)v_end) == pitch_tv && (size_t)((char*)d_pk - d_ov) == ss;
    const uint32_t* h_pk = pair_tvk ? (const uint32_t*)n_accepted : P.h_packed;
    auto refused = [](cudaError_t e) {
        if (e != cudaErrorInvalidValue && e != cudaErrorInvalidPitchValue) return false;
        cudaGetLastError();
        return true;
    };
    long long group_min = 1;  // every piece its own group
    if (const char* e = getenv("B200PMP_HOST_GROUP")) group_min = 1LL << atoi(e);  // tuning knob
    int g0 = 0;  // first piece of the open group
    std::vector<int> group_end;
    for (int c = 0; c < n_pieces; c++) {
        const long long a = bounds[c], m = bounds[c + 1] - a;
        // host -> device, then the new frontier
        PMP_TRY(cudaMemcpyAsync(d_ix + a * nx * es, hx + a * nx * es, m * nx * es, cudaMemcpyHostToDevice, P.s_in));
        if (!d_term) PMP_TRY(cudaMemcpyAsync(d_ivx + a * nx * es, hvx + a * nx * es, m * nx * es, cudaMemcpyHostToDevice, P.s_in));
        if (ht) PMP_TRY(cudaMemcpyAsync(d_it + a * es, ht + a * es, m * es, cudaMemcpyHostToDevice, P.s_in));
        if (!d_term) PMP_TRY(cudaMemcpyAsync(d_iv + a * es, hv + a * es, m * es, cudaMemcpyHostToDevice, P.s_in));
        if (c == 0) PMP_TRY(cudaStreamWaitEvent(P.s_in, P.ev_ctrl, 0));  // the control block has been zeroed
        PMP_TRY(cudaMemcpyAsync(&ctrl->ready, &P.h_frontier[c], 8, cudaMemcpyHostToDevice, P.s_in));
        stamp(P.s_in);
        // device -> host once every trajectory of the piece has retired
        if (int rc = wait32(P.s_out, (unsigned long long)(uintptr_t)&ctrl->done[c], (unsigned int)m, 0u /* CU_STREAM_WAIT_VALUE_GEQ */))
            return drain(-2000 - rc);
        stamp(P.s_out);
        // Two leaves that sit at a constant distance from each other on both sides leave as ONE two-row pitched copy (the
        // device arrays are `sv` / `ss` apart by construction; the Python wrapper carves its pinned outputs out of one block):
        // a piece costs three enqueued copies instead of five, and the copy engine switches direction less often while the
        // input pieces are still arriving.  B200PMP_HOST_2D=0 restores one copy per leaf (A/B knob).
        if (pair_xvx) {
            const cudaError_t e2 = cudaMemcpy2DAsync((char*)x_end + a * nx * es, (size_t)pitch_xvx, d_ox + a * nx * es, sv, m * nx * es, 2, cudaMemcpyDeviceToHost, P.s_out);
            if (refused(e2)) pair_xvx = false;
            else PMP_TRY(e2);
        }
        if (!pair_xvx) {
            if (x_end) PMP_TRY(cudaMemcpyAsync((char*)x_end + a * nx * es, d_ox + a * nx * es, m * nx * es, cudaMemcpyDeviceToHost, P.s_out));
            if (vx_end) PMP_TRY(cudaMemcpyAsync((char*)vx_end + a * nx * es, d_ovx + a * nx * es, m * nx * es, cudaMemcpyDeviceToHost, P.s_out));
        }
        // The per-trajectory scalars (t, v, counters: 12 of the 60 bytes) can leave once per GROUP of pieces (B200PMP_HOST_GROUP:
        // a 0.5 MB copy takes 13 us against 55 us for 3 MB).  Measured, 2^20 rollouts: 2.03 ms per piece (default), 2.12 at
        // 2^17, 2.22 at 2^18, 2.40 at 2^19 -- what the copies gain, the calling thread loses unfolding the counters of a
        // large group at the very end.
        const bool close = c == n_pieces - 1 || c == n_pieces - 2 || bounds[c + 1] - bounds[g0] >= group_min;
        if (close) {
            const long long ga = bounds[g0], gm = bounds[c + 1] - ga;
            if (pair_tv) {
                const cudaError_t e2 = cudaMemcpy2DAsync((char*)t_end + ga * es, (size_t)pitch_tv, d_ot + ga * es, ss, gm * es, pair_tvk ? 3 : 2, cudaMemcpyDeviceToHost, P.s_out);
                if (refused(e2)) {  // (can only happen at the first group: the leaves' relation to each other is the same for all)
                    pair_tv = pair_tvk = false;
                    h_pk = P.h_packed;
                } else {
                    PMP_TRY(e2);
                }
            }
            if (!pair_tv) {
                if (t_end) PMP_TRY(cudaMemcpyAsync((char*)t_end + ga * es, d_ot + ga * es, gm * es, cudaMemcpyDeviceToHost, P.s_out));
                if (v_end) PMP_TRY(cudaMemcpyAsync((char*)v_end + ga * es, d_ov + ga * es, gm * es, cudaMemcpyDeviceToHost, P.s_out));
            }
            if (pair_tvk) {
            } else if (pack) {
                PMP_TRY(cudaMemcpyAsync(P.h_packed + ga, d_pk + ga, gm * 4, cudaMemcpyDeviceToHost, P.s_out));
            } else {
                if (n_accepted) PMP_TRY(cudaMemcpyAsync(n_accepted + ga, d_na + ga, gm * 4, cudaMemcpyDeviceToHost, P.s_out));
                if (n_steps) PMP_TRY(cudaMemcpyAsync(n_steps + ga, d_ns + ga, gm * 4, cudaMemcpyDeviceToHost, P.s_out));
                if (result) PMP_TRY(cudaMemcpyAsync(result + ga, d_res + ga, gm * 4, cudaMemcpyDeviceToHost, P.s_out));
            }
            PMP_TRY(cudaEventRecord(P.ev_piece[c], P.s_out));
            group_end.push_back(c);
            g0 = c + 1;
        }
        stamp(P.s_out);
    }
    // the watchdog flag travels behind the last piece into the spare slot of the pinned frontier array (the kernel sets it
    // before it releases the piece counters): no blocking copy on the calling thread at the end.  NOT enqueued behind the
    // kernel on its own stream: a device->host copy that sits in the engine's queue until the kernel ends holds up the
    // pieces' copies enqueued after it.
    unsigned int* h_abort = (unsigned int*)&P.h_frontier[kHostMaxPieces];
    *h_abort = 0u;
    PMP_TRY(cudaMemcpyAsync(h_abort, &ctrl->abort, 4, cudaMemcpyDeviceToHost, P.s_out));
    h_enqueued = host_ms();
    {
        int first = 0;
        for (int c : group_end) {
            PMP_TRY(cudaEventSynchronize(P.ev_piece[c]));
            if (pack) {
                const long long a = bounds[first], m = bounds[c + 1] - a;
                unpack_counters(h_pk + a, m, n_accepted ? n_accepted + a : nullptr, n_steps ? n_steps + a : nullptr,
                                result ? result + a : nullptr);
            }
            first = c + 1;
        }
    }
    PMP_TRY(cudaStreamSynchronize(P.s_out));
    PMP_TRY(cudaStreamSynchronize(sc));
    PMP_TRY(cudaStreamSynchronize(P.s_in));
    const unsigned int aborted = *h_abort;
    h_synced = host_ms();
    if (trace) {
        fprintf(stderr, "[pmp_solve_host] calling thread: kernel launched at %.3f ms, every copy enqueued at %.3f, all results on the host at %.3f\n",
                h_launched, h_enqueued, h_synced);
        float t;
        cudaEventElapsedTime(&t, tev[0], tev[1]);
        fprintf(stderr, "[pmp_solve_host] persistent launch: integrator finished at %.3f ms\n", t);
        for (int c = 0; c < n_pieces; c++) {
            float u[3];
            for (int k = 0; k < 3; k++) cudaEventElapsedTime(&u[k], tev[0], tev[2 + 3 * c + k]);
            fprintf(stderr, "[pmp_solve_host] piece %d: h2d done %.3f | piece complete, d2h %.3f -> %.3f ms\n", c, u[0], u[1], u[2]);
        }
    }
    for (auto e : tev) cudaEventDestroy(e);
    return aborted ? (int)cudaErrorLaunchTimeout : 0;
}
#undef PMP_TRY
#define PMP_TRY(call)                               \
    do {                                            \
        cudaError_t e_ = (call);                    \
        if (e_ != cudaSuccess) return (int)e_;      \
    } while (0)

}  // namespace

// Piece boundaries for a batch of n (sizes are multiples of 128 trajectories).  n_pieces > 0: that many equal pieces.
// n_pieces == 0: the device->host stream is the bottleneck of the pipeline; it idles until the first piece has been
// copied in and integrated, and the call ends with the last piece's copy-out and counter unfolding.  So the plan starts
// and ends with small pieces (a = n/32, at least 2^15): a, 2a, [pieces of <= n/8], 2a, a.
int host_plan(long long n, int n_pieces, std::vector<long long>* bounds, long long* cs_max) {
    auto up128 = [](long long v) { return std::max<long long>(128, (v + 127) & ~127LL); };
    std::vector<long long> sizes;
    const long long a = std::max<long long>(1LL << 15, up128(n / 32));
    if (const char* e = (n_pieces == 0 ? getenv("B200PMP_HOST_PLAN") : nullptr)) {
        // tuning knob: comma-separated log2 sizes, e.g. "15,16,17,18,18,18,15"; the remainder (if any) is appended
        long long done = 0;
        for (const char* q = e; *q && done < n;) {
            const long long m = std::min(n - done, up128(1LL << atoi(q)));
            sizes.push_back(m);
            done += m;
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        if (done < n) sizes.push_back(n - done);
    } else if (n_pieces > 0 || n < 8 * a) {
        if (n_pieces <= 0) n_pieces = (int)std::max<long long>(1, n >> 16);  // 2^16 .. 2^18: pieces of 2^16
        const long long cs = up128((n + n_pieces - 1) / n_pieces);
        for (long long done = 0; done < n; done += cs) sizes.push_back(std::min(cs, n - done));
    } else {
        const long long cap = std::max<long long>(1LL << 16, up128(n / 8)), mid = n - 6 * a;
        const long long k = (mid + cap - 1) / cap, cs = up128((mid + k - 1) / k);
        sizes = {a, 2 * a};
        for (long long done = 0; done < mid; done += cs) sizes.push_back(std::min(cs, mid - done));
        sizes.push_back(2 * a);
        sizes.push_back(a);
    }
    std::vector<long long> b{0};
    long long biggest = 0;
    for (long long m : sizes) {
        b.push_back(b.back() + m);
        biggest = std::max(biggest, m);
    }
    if (cs_max) *cs_max = up128(biggest);
    const int count = (int)sizes.size();
    if (bounds) *bounds = std::move(b);
    return count;
}

// 0 = ok, > 0 = cudaError_t, < 0 = PMP_ERR_* already recorded by pmp_solve_batch_cuda
int solve_host(const pmp_problem_t& p, const pmp_opts_t& o, long long n, const void* x_f, const void* t_f, const void* v_f,
               const void* vx_f, void* x_end, void* t_end, void* v_end, void* vx_end, int32_t* n_accepted, int32_t* n_steps,
               int32_t* result, int n_pieces, const double* term_P, const double* term_xeq) {
    int dev = 0;
    PMP_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return (int)cudaErrorInvalidDevice;
    const int nx = p.nx, NS = 2 * nx + 2;
    const size_t es = o.dtype == PMP_F32 ? 4 : 8;
    long long cs = 0;  // capacity of a staging set, in trajectories
    std::vector<long long> bounds;
    n_pieces = host_plan(n, n_pieces, &bounds, &cs);

    const int64_t ws_bytes = pmp_solve_workspace_bytes(&p, &o, cs);
    if (ws_bytes < 0) return (int)ws_bytes;
    // one staging set: vector leaves in (x, vx), y, y_end, vector leaves out (x, vx), three int arrays, solver workspace
    const size_t o_in = 0, o_y = o_in + up256(NS * cs * es), o_ye = o_y + up256(NS * cs * es), o_ox = o_ye + up256(NS * cs * es),
                 o_int = o_ox + up256(2 * nx * cs * es), o_ws = o_int + up256(4 * cs * 4), total = o_ws + up256((size_t)ws_bytes);

    std::lock_guard<std::mutex> lock(g_mu[dev]);
    HostPipe& P = g_pipe[dev];
    if (int e = ensure(P, total)) return e;
    // counters travel packed when at least two of them are wanted and they fit the fields
    const bool pack = ((n_accepted != nullptr) + (n_steps != nullptr) + (result != nullptr)) >= 2 && o.max_steps < (1 << kPackBits);
    if (pack)
        if (int e = ensure_packed(P, (size_t)n, n_pieces)) return e;
    // The per-trajectory scalars (t_f, v_f in; t_end, v_end, packed counters out) are 4 of every 56 / 60 bytes, but as
    // per-piece copies they were 5 of the 9 copies of a piece, each half a megabyte: with both directions busy the link
    // moved 43 GB/s in 512 KiB copies against 78 GB/s in 3 MiB copies (profiles/r2_pcie_probe.jsonl).  They live in
    // whole-batch device arrays instead and cross the link once per GROUP of consecutive pieces (>= 2^18 trajectories,
    // i.e. >= 1 MiB per copy); the first and the last piece are groups of their own so that neither the start of the
    // integration nor the end of the call waits for a long copy.
    if (int e = ensure_small(P, (size_t)n * es)) return e;
    // LQR terminal set (pmp_solve_host_lqr): P and x_eq in the kernel's precision, in a small device array
    const void* d_term = nullptr;
    if (term_P) {
        if (!P.term) {
            PMP_TRY(cudaMalloc((void**)&P.term, 8 * (16 * 16 + 16)));
            PMP_TRY(cudaHostAlloc((void**)&P.h_term, 8 * (16 * 16 + 16), cudaHostAllocDefault));
        }
        const int nx2 = p.nx * p.nx;
        for (int q = 0; q < nx2 + p.nx; q++) {
            const double val = q < nx2 ? term_P[q] : (term_xeq ? term_xeq[q - nx2] : 0.0);
            if (o.dtype == PMP_F32) ((float*)P.h_term)[q] = (float)val;
            else ((double*)P.h_term)[q] = val;
        }
        PMP_TRY(cudaMemcpyAsync(P.term, P.h_term, (size_t)(nx2 + p.nx) * es, cudaMemcpyHostToDevice, P.s_in));
        PMP_TRY(cudaStreamSynchronize(P.s_in));  // (a few hundred bytes; every stream below may read it)
        d_term = P.term;
    }
    char* const s_tin = P.small, *const s_vin = s_tin + P.small_cap, *const s_tout = s_vin + P.small_cap,
               *const s_vout = s_tout + P.small_cap, *const s_pk = s_vout + P.small_cap;
    std::vector<int> g_first(n_pieces, 0), g_last(n_pieces, 0);  // first / last piece of the group a piece belongs to
    long long group_min = 1LL << 17;  // (measured: 2.38 ms at 2^17, 2.41 per piece, 2.48 at 2^18 for 2^20 rollouts: noise level)
    if (const char* e = getenv("B200PMP_HOST_GROUP")) group_min = 1LL << atoi(e);  // tuning knob (0: every piece its own group)
    {
        int c = 0;
        while (c < n_pieces) {
            int e = c;
            if (c > 0 && c < n_pieces - 1)
                while (e + 1 < n_pieces - 1 && bounds[e + 1] - bounds[c] < group_min) e++;
            for (int q = c; q <= e; q++) { g_first[q] = c; g_last[q] = e; }
            c = e + 1;
        }
    }

    // B200PMP_HOST_TRACE=1: timestamps of every piece's stages on stderr (diagnostics; adds timed events)
    const bool trace_p = getenv("B200PMP_HOST_TRACE") != nullptr;
    if (n >= (1LL << 16) && !getenv("B200PMP_HOST_PER_PIECE")) {  // one persistent launch over the whole batch
        if (!pack)
            if (int e = ensure_packed(P, 4, n_pieces)) return e;  // (the per-piece events)
        const int rc = solve_host_persistent(P, p, o, n, bounds, x_f, t_f, v_f, vx_f, x_end, t_end, v_end, vx_end, n_accepted, n_steps,
                                             result, pack, trace_p, d_term);
        if (rc != -1000) return rc;
    }
    const bool trace = getenv("B200PMP_HOST_TRACE") != nullptr;  // (one lookup per call: ~0.1 us against a call of milliseconds)
    std::vector<cudaEvent_t> tev;
    auto stamp = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tev.push_back(e);
    };
    stamp(P.s_in);
    const char *hx = (const char*)x_f, *hvx = (const char*)vx_f, *ht = (const char*)t_f, *hv = (const char*)v_f;
    // Any failure after the first enqueue must not return while copies of earlier pieces are still in flight: the caller
    // frees or reuses its host buffers as soon as the call returns.  `drain` waits for all three kinds of streams (errors
    // ignored: the first error is what the caller sees) and drops the trace events.
    auto drain = [&](int rc) {
        cudaStreamSynchronize(P.s_in);
        for (auto& s : P.s_comp) cudaStreamSynchronize(s);
        cudaStreamSynchronize(P.s_out);
        for (auto e : tev) cudaEventDestroy(e);
        tev.clear();
        cudaGetLastError();
        return rc;
    };
#undef PMP_TRY
#define PMP_TRY(call)                                  \
    do {                                               \
        cudaError_t e_ = (call);                       \
        if (e_ != cudaSuccess) return drain((int)e_);  \
    } while (0)
    for (int c = 0; c < n_pieces; c++) {
        const long long a = bounds[c], m = bounds[c + 1] - a;
        const int k = c % kSets;
        cudaStream_t sc = P.s_comp[k];
        char* B = P.buf[k];
        char *d_x = B + o_in, *d_vx = d_x + cs * nx * es;
        char *d_y = B + o_y, *d_ye = B + o_ye, *d_ox = B + o_ox, *d_ovx = d_ox + cs * nx * es;
        int32_t *d_na = (int32_t*)(B + o_int), *d_ns = d_na + cs, *d_res = d_ns + cs;
        const long long ga = bounds[g_first[c]], gm = bounds[g_last[c] + 1] - ga;  // the group's range of trajectories

        // host -> device; the leaves-in area of this set is free once piece c - kSets has been packed
        if (c >= kSets) PMP_TRY(cudaStreamWaitEvent(P.s_in, P.ev_pack[k], 0));
        if (g_first[c] == c) {  // scalars of the whole group ahead of its first piece
            if (ht) PMP_TRY(cudaMemcpyAsync(s_tin + ga * es, ht + ga * es, gm * es, cudaMemcpyHostToDevice, P.s_in));
            if (!d_term) PMP_TRY(cudaMemcpyAsync(s_vin + ga * es, hv + ga * es, gm * es, cudaMemcpyHostToDevice, P.s_in));
        }
        PMP_TRY(cudaMemcpyAsync(d_x, hx + a * nx * es, m * nx * es, cudaMemcpyHostToDevice, P.s_in));
        if (!d_term) PMP_TRY(cudaMemcpyAsync(d_vx, hvx + a * nx * es, m * nx * es, cudaMemcpyHostToDevice, P.s_in));
        PMP_TRY(cudaEventRecord(P.ev_in[k], P.s_in));
        stamp(P.s_in);

        // layout change, integration, layout change; y / y_end / out areas are free once piece c - kSets has left
        PMP_TRY(cudaStreamWaitEvent(sc, P.ev_in[k], 0));
        if (c >= kSets) PMP_TRY(cudaStreamWaitEvent(sc, P.ev_out[k], 0));
        const unsigned blocks = (unsigned)((m + 255) / 256);
        stamp(sc);
        if (o.dtype == PMP_F32)
            pack_kernel<float><<<blocks, 256, 0, sc>>>(nx, m, (const float*)d_x, (const float*)d_vx,
                                                        ht ? (const float*)(s_tin + a * es) : nullptr, (const float*)(s_vin + a * es),
                                                        (float)o.t0, (float*)d_y, (const float*)d_term);
        else
            pack_kernel<double><<<blocks, 256, 0, sc>>>(nx, m, (const double*)d_x, (const double*)d_vx,
                                                         ht ? (const double*)(s_tin + a * es) : nullptr,
                                                         (const double*)(s_vin + a * es), o.t0, (double*)d_y, (const double*)d_term);
        PMP_TRY(cudaGetLastError());
        PMP_TRY(cudaEventRecord(P.ev_pack[k], sc));
        if (int rc = pmp_solve_batch_cuda(&p, &o, m, d_y, d_ye, nullptr, d_na, d_ns, d_res, B + o_ws, ws_bytes, sc)) return drain(rc);
        {
            const bool xs = x_end || vx_end;
            uint32_t* pk = pack ? (uint32_t*)s_pk + a : nullptr;
            if (o.dtype == PMP_F32)
                unpack_kernel<float><<<blocks, 256, 0, sc>>>(nx, m, (const float*)d_ye, xs ? (float*)d_ox : nullptr, (float*)d_ovx,
                                                              d_na, d_ns, d_res, pk, t_end ? (float*)(s_tout + a * es) : nullptr,
                                                              v_end ? (float*)(s_vout + a * es) : nullptr);
            else
                unpack_kernel<double><<<blocks, 256, 0, sc>>>(nx, m, (const double*)d_ye, xs ? (double*)d_ox : nullptr,
                                                               (double*)d_ovx, d_na, d_ns, d_res, pk,
                                                               t_end ? (double*)(s_tout + a * es) : nullptr,
                                                               v_end ? (double*)(s_vout + a * es) : nullptr);
            PMP_TRY(cudaGetLastError());
        }
        PMP_TRY(cudaEventRecord(P.ev_comp[k], sc));
        stamp(sc);

        // device -> host
        PMP_TRY(cudaStreamWaitEvent(P.s_out, P.ev_comp[k], 0));
        stamp(P.s_out);
        if (x_end) PMP_TRY(cudaMemcpyAsync((char*)x_end + a * nx * es, d_ox, m * nx * es, cudaMemcpyDeviceToHost, P.s_out));
        if (vx_end) PMP_TRY(cudaMemcpyAsync((char*)vx_end + a * nx * es, d_ovx, m * nx * es, cudaMemcpyDeviceToHost, P.s_out));
        if (!pack) {  // counters that do not fit the packed word: three plain copies per piece
            if (n_accepted) PMP_TRY(cudaMemcpyAsync(n_accepted + a, d_na, m * 4, cudaMemcpyDeviceToHost, P.s_out));
            if (n_steps) PMP_TRY(cudaMemcpyAsync(n_steps + a, d_ns, m * 4, cudaMemcpyDeviceToHost, P.s_out));
            if (result) PMP_TRY(cudaMemcpyAsync(result + a, d_res, m * 4, cudaMemcpyDeviceToHost, P.s_out));
        }
        if (g_last[c] == c) {  // every piece of the group has been integrated (s_out waited for them in order)
            if (t_end) PMP_TRY(cudaMemcpyAsync((char*)t_end + ga * es, s_tout + ga * es, gm * es, cudaMemcpyDeviceToHost, P.s_out));
            if (v_end) PMP_TRY(cudaMemcpyAsync((char*)v_end + ga * es, s_vout + ga * es, gm * es, cudaMemcpyDeviceToHost, P.s_out));
            if (pack) {
                PMP_TRY(cudaMemcpyAsync(P.h_packed + ga, (uint32_t*)s_pk + ga, gm * 4, cudaMemcpyDeviceToHost, P.s_out));
                PMP_TRY(cudaEventRecord(P.ev_piece[c], P.s_out));
            }
        }
        PMP_TRY(cudaEventRecord(P.ev_out[k], P.s_out));
        stamp(P.s_out);
    }
    // the waiting host thread unfolds the counters of each piece as soon as its copies have landed
    if (pack)
        for (int c = 0; c < n_pieces; c++) {
            if (g_last[c] != c) continue;  // one event per group, recorded behind its last piece
            PMP_TRY(cudaEventSynchronize(P.ev_piece[c]));
            const long long a = bounds[g_first[c]], m = bounds[c + 1] - a;
            unpack_counters(P.h_packed + a, m, n_accepted ? n_accepted + a : nullptr, n_steps ? n_steps + a : nullptr,
                            result ? result + a : nullptr);
        }
    // every stream's work precedes the last record on s_out, except the tail of s_in/s_comp when all outputs are NULL
    PMP_TRY(cudaStreamSynchronize(P.s_out));
    for (auto& s : P.s_comp) PMP_TRY(cudaStreamSynchronize(s));
    if (trace) {  // per piece: h2d end | compute start, end | d2h start, end   (ms since the first copy was enqueued)
        for (int c = 0; c < n_pieces; c++) {
            float t[5];
            for (int q = 0; q < 5; q++) cudaEventElapsedTime(&t[q], tev[0], tev[1 + 5 * c + q]);
            fprintf(stderr, "[pmp_solve_host] piece %d: h2d done %.3f | compute %.3f -> %.3f | d2h %.3f -> %.3f ms\n", c, t[0], t[1],
                    t[2], t[3], t[4]);
        }
        for (auto e : tev) cudaEventDestroy(e);
    }
    return 0;
}
#undef PMP_TRY
#define PMP_TRY(call)                               \
    do {                                            \
        cudaError_t e_ = (call);                    \
        if (e_ != cudaSuccess) return (int)e_;      \
    } while (0)

// Diagnostics (scripts/pcie_probe.py): a plain 16-byte-per-thread copy kernel between any two unified-address pointers --
// with a pinned host pointer on one side it measures what the SMs themselves can move over the host link ("zero copy"),
// next to the copy engines.
__global__ void probe_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long long n16) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}
// blocks < 0: FILL with the 16-byte pattern at src (streaming stores, -blocks CTAs): scripts/exp_dense_prefill.py
__global__ void __launch_bounds__(256) probe_fill_kernel(uint4* __restrict__ dst, const uint4* __restrict__ pat, long long n16) {
    const uint4 v = *pat;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        __stcs(dst + i, v);
}
int host_probe_copy(void* dst, const void* src, long long bytes, int blocks, cudaStream_t st) {
    if (blocks < 0) probe_fill_kernel<<<-blocks, 256, 0, st>>>((uint4*)dst, (const uint4*)src, bytes / 16);
    else probe_copy_kernel<<<blocks, 256, 0, st>>>((uint4*)dst, (const uint4*)src, bytes / 16);
    return (int)cudaGetLastError();
}

int host_release() {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int d = 0; d < kMaxDev; d++) {
        std::lock_guard<std::mutex> lock(g_mu[d]);
        HostPipe& P = g_pipe[d];
        if (!P.ready) continue;
        PMP_TRY(cudaSetDevice(d));
        for (auto& b : P.buf) { if (b) cudaFree(b); b = nullptr; }
        P.cap = 0;
        if (P.h_packed) cudaFreeHost(P.h_packed);
        P.h_packed = nullptr; P.h_cap = 0;
        if (P.small) cudaFree(P.small);
        P.small = nullptr; P.small_cap = 0;
        if (P.big) cudaFree(P.big);
        P.big = nullptr; P.big_cap = 0;
        if (P.ctrl) { cudaFree(P.ctrl); cudaFreeHost(P.h_frontier); cudaEventDestroy(P.ev_ctrl); }
        P.ctrl = nullptr; P.h_frontier = nullptr; P.ev_ctrl = nullptr;
        if (P.term) { cudaFree(P.term); cudaFreeHost(P.h_term); }
        P.term = nullptr; P.h_term = nullptr;
        for (auto e : P.ev_piece) cudaEventDestroy(e);
        P.ev_piece.clear();
        cudaStreamDestroy(P.s_in); cudaStreamDestroy(P.s_out);
        for (auto& s : P.s_comp) cudaStreamDestroy(s);
        for (int k = 0; k < kSets; k++) {
            cudaEventDestroy(P.ev_in[k]); cudaEventDestroy(P.ev_pack[k]); cudaEventDestroy(P.ev_comp[k]); cudaEventDestroy(P.ev_out[k]);
        }
        P.ready = false;
    }
    PMP_TRY(cudaSetDevice(cur));
    return 0;
}

}  // namespace pmp
