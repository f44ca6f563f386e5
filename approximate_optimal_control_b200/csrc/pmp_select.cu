// pmp_select.cu -- the data-selection step right after the backward solve, on the device-resident dense Solution.
//
//   select_train_pts(value_interval, sols)            levelsets.py:667-709
//       bool_train_idx = (v >= v_lower) & (v <= v_upper) [& ||vxx||_F < vxx_max_norm];  all_ys = node[bool_train_idx]
//   data_normaliser(train_ode_states)                 nn_utils.py:112-134
//       x_means, x_stds (per component), v_mean, v_std over the selected nodes (population std, numpy's default)
//
// The reference builds an (N, S+1) boolean mask over the inf-padded SaveAt(steps) buffers and gathers every leaf with it
// (row-major order: trajectory by trajectory, slot by slot), then makes separate passes for the statistics.  Here a WARP
// owns a trajectory and reads only its n_accepted + 1 saved nodes (15 % of the 8 GB buffer at the headline size):
//   pass 1  select_count_kernel   per-trajectory count of selected nodes (one coalesced read of v, optionally vxx);
//           the caller turns the counts into offsets (exclusive prefix sum);
//   pass 2  select_write_kernel   stable compaction of every leaf into flat arrays (same order as node[bool_train_idx])
//           fused with the normaliser's sums (x, x^2 per component, v, v^2; fp64 accumulators, one set of atomics per
//           CTA) and the NaN check of levelsets.py:700-706.
// HBM-bound: bytes = valid nodes x (1 + 2 nx + 3) reals read + selected nodes written.
#include "pmp_common.cuh"

namespace pmp {

constexpr int kSelWarps = 8;  // warps (= trajectories in flight) per CTA

template <class R>
__device__ __forceinline__ bool node_selected(const R* __restrict__ v, const R* __restrict__ vxx, long long node, int nvxx, R lo, R hi,
                                              R vxx_max) {
    const R val = v[node];
    bool sel = val >= lo && val <= hi;  // NaN / inf padding compare false
    if (sel && vxx) {
        R n2 = R(0);
        for (int q = 0; q < nvxx; q++) {
            const R e = vxx[node * nvxx + q];
            n2 = Math<R>::fma(e, e, n2);
        }
        sel = Math<R>::sqrt(n2) < vxx_max;
    }
    return sel;
}

template <class R>
__global__ void __launch_bounds__(kSelWarps * 32) select_count_kernel(long long n, int s1, const R* __restrict__ v,
                                                                      const R* __restrict__ vxx, int nvxx,
                                                                      const int32_t* __restrict__ n_accepted, R lo, R hi, R vxx_max,
                                                                      int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    for (long long i = blockIdx.x * (long long)kSelWarps + (threadIdx.x >> 5); i < n; i += (long long)gridDim.x * kSelWarps) {
        const int valid = n_accepted ? min(n_accepted[i] + 1, s1) : s1;
        int c = 0;
        for (int s0 = 0; s0 < valid; s0 += 32) {
            const int s = s0 + lane;
            const bool sel = s < valid && node_selected<R>(v, vxx, i * s1 + s, nvxx, lo, hi, vxx_max);
            c += __popc(__ballot_sync(0xffffffffu, sel));
        }
        if (lane == 0) counts[i] = c;
    }
}

// stats: [2 nx + 2] doubles = sum x[nx], sum x^2 [nx], sum v, sum v^2;  flags[0] |= 1 if a selected node holds a NaN
template <class R>
__global__ void __launch_bounds__(kSelWarps * 32) select_write_kernel(long long n, int s1, int nx, const R* __restrict__ x,
                                                                      const R* __restrict__ t, const R* __restrict__ v,
                                                                      const R* __restrict__ vx, const R* __restrict__ vxx, int nvxx,
                                                                      const int32_t* __restrict__ n_accepted,
                                                                      const long long* __restrict__ offsets, R lo, R hi, R vxx_max,
                                                                      R* __restrict__ ox, R* __restrict__ ot, R* __restrict__ ov,
                                                                      R* __restrict__ ovx, R* __restrict__ ovxx,
                                                                      double* __restrict__ stats, int32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int MAXNX = 16;
    double sx[MAXNX], sxx[MAXNX], sv = 0.0, svv = 0.0;
    for (int q = 0; q < MAXNX; q++) { sx[q] = 0.0; sxx[q] = 0.0; }
    bool nan_seen = false;
    for (long long i = blockIdx.x * (long long)kSelWarps + warp; i < n; i += (long long)gridDim.x * kSelWarps) {
        const int valid = n_accepted ? min(n_accepted[i] + 1, s1) : s1;
        long long base = offsets[i];
        for (int s0 = 0; s0 < valid; s0 += 32) {
            const int s = s0 + lane;
            const long long node = i * s1 + s;
            const bool sel = s < valid && node_selected<R>(v, vxx, node, nvxx, lo, hi, vxx_max);
            const unsigned m = __ballot_sync(0xffffffffu, sel);
            if (sel) {
                const long long o = base + __popc(m & ((1u << lane) - 1u));
                const R val = v[node], tt = t[node];
                ov[o] = val;
                ot[o] = tt;
                nan_seen |= tt != tt;
                sv += (double)val;
                svv += (double)val * (double)val;
                for (int q = 0; q < nx; q++) {
                    const R a = x[node * nx + q], b = vx[node * nx + q];
                    ox[o * nx + q] = a;
                    ovx[o * nx + q] = b;
                    nan_seen |= (a != a) | (b != b);
                    if (q < MAXNX) { sx[q] += (double)a; sxx[q] += (double)a * (double)a; }
                }
                if (ovxx)
                    for (int q = 0; q < nvxx; q++) ovxx[o * nvxx + q] = vxx[node * nvxx + q];
            }
            base += __popc(m);
        }
    }
    // CTA reduction of the statistics: warp shuffles, then one lane per warp into shared memory, then one set of atomics
    __shared__ double red[kSelWarps][2 * MAXNX + 2];
    auto wsum = [](double a) {
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        return a;
    };
    for (int q = 0; q < nx && q < MAXNX; q++) {
        const double a = wsum(sx[q]), b = wsum(sxx[q]);
        if (lane == 0) { red[warp][q] = a; red[warp][MAXNX + q] = b; }
    }
    {
        const double a = wsum(sv), b = wsum(svv);
        if (lane == 0) { red[warp][2 * MAXNX] = a; red[warp][2 * MAXNX + 1] = b; }
    }
    if (__any_sync(0xffffffffu, nan_seen) && lane == 0) atomicOr(flags, 1);
    __syncthreads();
    if (threadIdx.x < 2 * nx + 2) {
        const int q = threadIdx.x;
        const int col = q < nx ? q : (q < 2 * nx ? MAXNX + (q - nx) : 2 * MAXNX + (q - 2 * nx));
        double a = 0.0;
        for (int w = 0; w < kSelWarps; w++) a += red[w][col];
        if (a != 0.0) atomicAdd(stats + q, a);
    }
}

int select_grid(long long n) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long need = (n + kSelWarps - 1) / kSelWarps, cap = (long long)sms * 8;  // 8 CTAs of 256 threads per SM
    return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

int select_count(int dtype, long long n, int s1, const void* v, const void* vxx, int nvxx, const int32_t* n_accepted, double lo,
                 double hi, double vxx_max, int32_t* counts, cudaStream_t st) {
    if (n == 0) return 0;
    const int grid = select_grid(n);
    if (dtype == PMP_F32)
        select_count_kernel<float><<<grid, kSelWarps * 32, 0, st>>>(n, s1, (const float*)v, (const float*)vxx, nvxx, n_accepted, (float)lo,
                                                                    (float)hi, (float)vxx_max, counts);
    else
        select_count_kernel<double><<<grid, kSelWarps * 32, 0, st>>>(n, s1, (const double*)v, (const double*)vxx, nvxx, n_accepted, lo, hi,
                                                                     vxx_max, counts);
    return (int)cudaGetLastError();
}

int select_write(int dtype, long long n, int s1, int nx, const void* x, const void* t, const void* v, const void* vx, const void* vxx,
                 int nvxx, const int32_t* n_accepted, const long long* offsets, double lo, double hi, double vxx_max, void* ox, void* ot,
                 void* ov, void* ovx, void* ovxx, double* stats, int32_t* flags, cudaStream_t st) {
    if (n == 0) return 0;
    const int grid = select_grid(n);
    if (dtype == PMP_F32)
        select_write_kernel<float><<<grid, kSelWarps * 32, 0, st>>>(n, s1, nx, (const float*)x, (const float*)t, (const float*)v,
                                                                    (const float*)vx, (const float*)vxx, nvxx, n_accepted, offsets, (float)lo,
                                                                    (float)hi, (float)vxx_max, (float*)ox, (float*)ot, (float*)ov,
                                                                    (float*)ovx, (float*)ovxx, stats, flags);
    else
        select_write_kernel<double><<<grid, kSelWarps * 32, 0, st>>>(n, s1, nx, (const double*)x, (const double*)t, (const double*)v,
                                                                     (const double*)vx, (const double*)vxx, nvxx, n_accepted, offsets, lo, hi,
                                                                     vxx_max, (double*)ox, (double*)ot, (double*)ov, (double*)ovx,
                                                                     (double*)ovxx, stats, flags);
    return (int)cudaGetLastError();
}

}  // namespace pmp
