// tc_probe.cu -- stand-alone probe of the tcgen05 building block used by pmp_forward_umma.cuh: one CTA of 128 threads,
// D[128 x N] (TMEM, fp32) = A[128 x K] (TMEM, written by the row-owning threads with tcgen05.st) . B (shared memory, UMMA
// descriptor, no swizzle), kind::tf32 in 3xTF32 split precision.  Checks K-major and MN-major B images, N = 64 / 16, K = 64 / 8
// against a double-precision product and times the st -> mma -> ld round trip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tc_probe scripts/tc_probe.cu && gpurun_out/tc_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../approximate_optimal_control_b200/csrc/pmp_umma.cuh"

using namespace pmp::umma;

// B images in shared memory (floats): [hi | lo], each N*K floats
template <int N, int K, bool MN_MAJOR, bool SWAP, bool SPIN>
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ Wt /* [N][K] */,
                                                       float* __restrict__ D, long long* cycles, int* err, long long* marks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sB = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) mbar_init(&bar, 1);
    for (int q = tid; q < N * K; q += 128) {
        const int n = q / K, k = q - n * K;
        uint32_t hi, lo;
        split_tf32(Wt[q], hi, lo);
        const int off = b_off_k<N, K>(n, k);
        sB[off] = __uint_as_float(hi);
        sB[N * K + off] = __uint_as_float(lo);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = tbase + ((uint32_t)(32 * warp) << 16);
    const uint32_t tD = 0, tAh = 256, tAl = 320;
    uint32_t phase = 0;
    long long t_acc = 0;
    float arow[K];
#pragma unroll
    for (int c = 0; c < K; c++) arow[c] = A[tid * K + c];
    float keep = 0.f;
    for (int rep = 0; rep < 5; rep++) {
        __syncthreads();
        const long long t0 = clock64();
        // stage A: this thread's row
#pragma unroll
        for (int c = 0; c < K; c += 8) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; j++) split_tf32(arow[c + j] + keep * 0.f, hi[j], lo[j]);
            tmem_st8(lane_base + tAh + c, hi);
            tmem_st8(lane_base + tAl + c, lo);
        }
        tmem_wait_st();
        const long long m1 = clock64();
        tc_fence_before();
        __syncthreads();
        const long long m2 = clock64();
        long long m3 = m2;
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = make_idesc_tf32(128, N);
            const uint32_t sb = smem_u32(sB);
            issue_3xtf32<N, K>(tbase + tD, tbase + tAh, tbase + tAl, sb, sb + N * K * 4, idesc);
            umma_commit(&bar);
            m3 = clock64();
        }
        if (!(SPIN ? mbar_spin(&bar, phase) : mbar_wait(&bar, phase))) { if (tid == 0) *err = 1; break; }
        const long long m4 = clock64();
        phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < N; c += 8) {
            uint32_t v[8];
            tmem_ld8(lane_base + tD + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (rep == 4) D[tid * N + c + j] = __uint_as_float(v[j]);
                keep += __uint_as_float(v[j]);
            }
        }
        const long long m5 = clock64();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const long long t1 = clock64();
        if (tid == 0 && rep == 4) { marks[0] = m1 - t0; marks[1] = m2 - m1; marks[2] = m3 - m2; marks[3] = m4 - m3; marks[4] = m5 - m4; marks[5] = t1 - m5; }
        if (rep > 0) t_acc += t1 - t0;
    }
    if (tid == 0) *cycles = t_acc / 4;
    if (keep == 12345.678f) D[0] = keep;
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int N, int K, bool MN, bool SWAP = false, bool SPIN = false>
static int run(const char* name) {
    std::vector<float> A(128 * K), Wt(N * K), D(128 * N);
    srand(1234 + N + K);
    for (auto& a : A) a = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& w : Wt) w = (float)rand() / RAND_MAX - 0.5f;
    float *dA, *dW, *dD;
    long long* dc;
    int* de;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, Wt.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMalloc(&dc, 8); cudaMalloc(&de, 4);
    long long* dm; cudaMalloc(&dm, 64); cudaMemset(dm, 0, 64);
    cudaMemset(de, 0, 4); cudaMemset(dD, 0, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, Wt.data(), Wt.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 2 * N * K * 4;
    cudaFuncSetAttribute(probe_kernel<N, K, MN, SWAP, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<N, K, MN, SWAP, SPIN><<<1, 128, smem>>>(dA, dW, dD, dc, de, dm);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    int err = 0;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&err, de, 4, cudaMemcpyDeviceToHost);
    long long mk[8]; cudaMemcpy(mk, dm, 64, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            double r = 0;
            for (int k = 0; k < K; k++) r += (double)A[m * K + k] * (double)Wt[n * K + k];
            maxerr = fmax(maxerr, fabs(r - (double)D[m * N + n]));
            maxref = fmax(maxref, fabs(r));
        }
    printf("{\"probe\": \"%s\", \"N\": %d, \"K\": %d, \"b_major\": \"%s\", \"cuda\": \"%s\", \"timeout\": %d, \"max_abs_err\": %.3e, \"max_ref\": %.3f, "
           "\"round_trip_cycles\": %lld, \"wait\": \"%s\", \"stage_st\": %lld, \"sync\": %lld, \"issue_commit\": %lld, \"mma_wait\": %lld, \"ld\": %lld, \"sync2\": %lld}\n", name, N, K, MN ? "MN" : "K", cudaGetErrorString(e), err, maxerr, maxref, cyc, SPIN ? "test_wait" : "try_wait", mk[0], mk[1], mk[2], mk[3], mk[4], mk[5]);
    fflush(stdout);
    return e != cudaSuccess || err || !(maxerr < 1e-4);
}

int main() {
    int bad = 0;
    bad += run<64, 64, false>("layer product, B K-major");
    bad += run<16, 64, false>("input gradient (N = 16), B K-major");
    bad += run<64, 8, false>("first layer (K = 8), B K-major");
    bad += run<64, 64, false, false, true>("layer product, B K-major");
    bad += run<16, 64, false, false, true>("input gradient (N = 16), B K-major");
    bad += run<64, 8, false, false, true>("first layer (K = 8), B K-major");
    bad += run<128, 64, false, false, true>("N = 128 (cost of a wider instruction)");
    bad += run<256, 64, false, false, true>("N = 256 (cost of a wider instruction)");
    printf("%s\n", bad ? "PROBE FAILED" : "PROBE OK");
    return bad;
}
