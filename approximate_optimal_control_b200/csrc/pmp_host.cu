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
#include <vector>

#include "../../include/b200pmp.h"

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
    std::vector<cudaEvent_t> ev_piece;  // device->host copies of piece c have landed
};
HostPipe g_pipe[kMaxDev];
std::mutex g_mu[kMaxDev];  // one per device: calls on different GPUs of one process run concurrently

inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

// state-dict leaves (array of structs) -> extended state rows (struct of arrays), and back
template <typename R>
__global__ void pack_kernel(int nx, long long m, const R* __restrict__ x, const R* __restrict__ vx, const R* __restrict__ t,
                            const R* __restrict__ v, R t_fill, R* __restrict__ y) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
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
                              uint32_t* __restrict__ packed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
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

void unpack_counters(const uint32_t* __restrict__ packed, long long m, int32_t* __restrict__ na, int32_t* __restrict__ ns,
                     int32_t* __restrict__ res) {
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

}  // namespace

// Piece boundaries for a batch of n (sizes are multiples of 128 trajectories).  n_pieces > 0: that many equal pieces.
// n_pieces == 0: the device->host stream is the bottleneck of the pipeline; it idles until the first piece has been
// copied in and integrated, and the call ends with the last piece's copy-out and counter unfolding.  So the plan starts
// and ends with small pieces (a = n/32, at least 2^15): a, 2a, [pieces of <= n/8], 2a, a.
int host_plan(long long n, int n_pieces, std::vector<long long>* bounds, long long* cs_max) {
    auto up128 = [](long long v) { return std::max<long long>(128, (v + 127) & ~127LL); };
    std::vector<long long> sizes;
    const long long a = std::max<long long>(1LL << 15, up128(n / 32));
    if (n_pieces > 0 || n < 8 * a) {
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
               int32_t* result, int n_pieces) {
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
    // one staging set: leaves in, y, y_end, leaves out (x, vx), three int arrays + their packed form, solver workspace
    const size_t o_in = 0, o_y = o_in + up256(NS * cs * es), o_ye = o_y + up256(NS * cs * es), o_ox = o_ye + up256(NS * cs * es),
                 o_int = o_ox + up256(2 * nx * cs * es), o_ws = o_int + up256(4 * cs * 4), total = o_ws + up256((size_t)ws_bytes);

    std::lock_guard<std::mutex> lock(g_mu[dev]);
    HostPipe& P = g_pipe[dev];
    if (int e = ensure(P, total)) return e;
    // counters travel packed when at least two of them are wanted and they fit the fields
    const bool pack = ((n_accepted != nullptr) + (n_steps != nullptr) + (result != nullptr)) >= 2 && o.max_steps < (1 << kPackBits);
    if (pack)
        if (int e = ensure_packed(P, (size_t)n, n_pieces)) return e;

    // B200PMP_HOST_TRACE=1: timestamps of every piece's stages on stderr (diagnostics; adds timed events)
    static const bool trace = getenv("B200PMP_HOST_TRACE") != nullptr;  // read once per process
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
        char *d_x = B + o_in, *d_vx = d_x + cs * nx * es, *d_t = d_vx + cs * nx * es, *d_v = d_t + cs * es;
        char *d_y = B + o_y, *d_ye = B + o_ye, *d_ox = B + o_ox, *d_ovx = d_ox + cs * nx * es;
        int32_t *d_na = (int32_t*)(B + o_int), *d_ns = d_na + cs, *d_res = d_ns + cs;
        uint32_t* d_pk = (uint32_t*)(d_res + cs);

        // host -> device; the leaves-in area of this set is free once piece c - kSets has been packed
        if (c >= kSets) PMP_TRY(cudaStreamWaitEvent(P.s_in, P.ev_pack[k], 0));
        PMP_TRY(cudaMemcpyAsync(d_x, hx + a * nx * es, m * nx * es, cudaMemcpyHostToDevice, P.s_in));
        PMP_TRY(cudaMemcpyAsync(d_vx, hvx + a * nx * es, m * nx * es, cudaMemcpyHostToDevice, P.s_in));
        if (ht) PMP_TRY(cudaMemcpyAsync(d_t, ht + a * es, m * es, cudaMemcpyHostToDevice, P.s_in));
        PMP_TRY(cudaMemcpyAsync(d_v, hv + a * es, m * es, cudaMemcpyHostToDevice, P.s_in));
        PMP_TRY(cudaEventRecord(P.ev_in[k], P.s_in));
        stamp(P.s_in);

        // layout change, integration, layout change; y / y_end / out areas are free once piece c - kSets has left
        PMP_TRY(cudaStreamWaitEvent(sc, P.ev_in[k], 0));
        if (c >= kSets) PMP_TRY(cudaStreamWaitEvent(sc, P.ev_out[k], 0));
        const unsigned blocks = (unsigned)((m + 255) / 256);
        stamp(sc);
        if (o.dtype == PMP_F32)
            pack_kernel<float><<<blocks, 256, 0, sc>>>(nx, m, (const float*)d_x, (const float*)d_vx, ht ? (const float*)d_t : nullptr,
                                                        (const float*)d_v, (float)o.t0, (float*)d_y);
        else
            pack_kernel<double><<<blocks, 256, 0, sc>>>(nx, m, (const double*)d_x, (const double*)d_vx,
                                                         ht ? (const double*)d_t : nullptr, (const double*)d_v, o.t0, (double*)d_y);
        PMP_TRY(cudaGetLastError());
        PMP_TRY(cudaEventRecord(P.ev_pack[k], sc));
        if (int rc = pmp_solve_batch_cuda(&p, &o, m, d_y, d_ye, nullptr, d_na, d_ns, d_res, B + o_ws, ws_bytes, sc)) return drain(rc);
        if (x_end || vx_end || pack) {
            const bool xs = x_end || vx_end;
            if (o.dtype == PMP_F32)
                unpack_kernel<float><<<blocks, 256, 0, sc>>>(nx, m, (const float*)d_ye, xs ? (float*)d_ox : nullptr, (float*)d_ovx,
                                                              d_na, d_ns, d_res, pack ? d_pk : nullptr);
            else
                unpack_kernel<double><<<blocks, 256, 0, sc>>>(nx, m, (const double*)d_ye, xs ? (double*)d_ox : nullptr,
                                                               (double*)d_ovx, d_na, d_ns, d_res, pack ? d_pk : nullptr);
            PMP_TRY(cudaGetLastError());
        }
        PMP_TRY(cudaEventRecord(P.ev_comp[k], sc));
        stamp(sc);

        // device -> host
        PMP_TRY(cudaStreamWaitEvent(P.s_out, P.ev_comp[k], 0));
        stamp(P.s_out);
        if (x_end) PMP_TRY(cudaMemcpyAsync((char*)x_end + a * nx * es, d_ox, m * nx * es, cudaMemcpyDeviceToHost, P.s_out));
        if (vx_end) PMP_TRY(cudaMemcpyAsync((char*)vx_end + a * nx * es, d_ovx, m * nx * es, cudaMemcpyDeviceToHost, P.s_out));
        if (t_end) PMP_TRY(cudaMemcpyAsync((char*)t_end + a * es, d_ye + nx * m * es, m * es, cudaMemcpyDeviceToHost, P.s_out));
        if (v_end) PMP_TRY(cudaMemcpyAsync((char*)v_end + a * es, d_ye + (nx + 1) * m * es, m * es, cudaMemcpyDeviceToHost, P.s_out));
        if (pack) {
            PMP_TRY(cudaMemcpyAsync(P.h_packed + a, d_pk, m * 4, cudaMemcpyDeviceToHost, P.s_out));
            PMP_TRY(cudaEventRecord(P.ev_piece[c], P.s_out));
        } else {
            if (n_accepted) PMP_TRY(cudaMemcpyAsync(n_accepted + a, d_na, m * 4, cudaMemcpyDeviceToHost, P.s_out));
            if (n_steps) PMP_TRY(cudaMemcpyAsync(n_steps + a, d_ns, m * 4, cudaMemcpyDeviceToHost, P.s_out));
            if (result) PMP_TRY(cudaMemcpyAsync(result + a, d_res, m * 4, cudaMemcpyDeviceToHost, P.s_out));
        }
        PMP_TRY(cudaEventRecord(P.ev_out[k], P.s_out));
        stamp(P.s_out);
    }
    // the waiting host thread unfolds the counters of each piece as soon as its copies have landed
    if (pack)
        for (int c = 0; c < n_pieces; c++) {
            PMP_TRY(cudaEventSynchronize(P.ev_piece[c]));
            const long long a = bounds[c], m = bounds[c + 1] - a;
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
