// pmp_umma.cuh -- the tcgen05 (5th-generation tensor core) building blocks of the value-network costate law
// (pmp_forward_umma.cuh): tensor-memory allocation, tcgen05.st / tcgen05.ld of a thread's own row, shared-memory matrix
// descriptors without swizzle, the kind::tf32 instruction descriptor, single-thread MMA issue and the mbarrier it commits to.
//
// Conventions (one CTA = 128 threads = the 128 lanes of tensor memory; thread r owns row r of every operand/accumulator):
//   * A operand [128 x K] in TENSOR MEMORY: lane = row, one 32-bit column per element (K-major), written by the owning
//     threads with tcgen05.st.32x32b (warp w can only touch lanes 32w .. 32w+31);
//   * B operand in SHARED MEMORY as 8 x 16-byte core matrices (128 contiguous bytes each), no swizzle, K-major:
//       B(n, k) at byte offset  (n/8) SBO + (k/4) LBO + (n%8) 16 + (k%4) 4,  LBO = 128, SBO = (K/4) 128
//     (a product with W^T therefore needs its own image of W^T; an MN-major image of W -- same bytes as row-major W cut into
//     8 x 4 blocks -- was tried in scripts/tc_probe.cu and did not give the product, so every square matrix is kept twice);
//   * D accumulator [128 x N] fp32 in tensor memory: lane = row, column = n;
//   * split precision: a = a_hi + a_lo, b = b_hi + b_lo (TF32 each), D = a_lo.b_hi + a_hi.b_lo + a_hi.b_hi accumulated in
//     fp32 by the tensor core (relative error ~2^-21 per product).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pmp {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}

// ---- tensor memory ------------------------------------------------------------------------------------------------------
// executed by ONE whole warp; the base address (lane 0, first column) lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// the calling thread's row (lane), 8 / 16 consecutive columns starting at taddr = base + (32 warp << 16) + column
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// ---- mbarrier (the tensor core's completion signal) -----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// spins until the phase with the given parity has completed; false after ~2^26 polls (a lost commit must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 26); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
// the same with the non-suspending test_wait (lowest wake-up latency; the waiting warps have nothing else to do)
__device__ __forceinline__ bool mbar_spin(unsigned long long* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 26); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
// all tcgen05.mma issued so far by this thread arrive on the barrier when they have completed
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- distributed shared memory: remote stores that complete a transaction count on the destination CTA's mbarrier ------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta_rank) {  // shared::cta address -> the same offset in a peer
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(v), "r"(remote_bar)
                 : "memory");
}
// one arrival + `bytes` expected on the local barrier (the phase completes when the bytes have landed)
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------------------------------
// float index of B(n, k) inside an image
template <int N, int K> __host__ __device__ constexpr int b_off_k(int n, int k) { return (n / 8) * (K / 4) * 32 + (k / 4) * 32 + (n % 8) * 4 + (k % 4); }

// shared-memory matrix descriptor (no swizzle): start address, leading-dimension and stride byte offsets in 16-byte units,
// descriptor version 1 (bit 46)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulator: c_format F32 (bits 4-5 = 1), a/b format TF32 (2 at bits 7-9 / 10-12), both operands K-major
// (bits 15 / 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem descriptor]: one instruction = 128 x N x 8
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one 3xTF32 product D[128 x N] = A[128 x K] . B, A hi / lo at tensor-memory columns a_hi / a_lo, B hi / lo images at shared
// addresses b_hi / b_lo; issued by ONE thread
template <int N, int K>
__device__ __forceinline__ void issue_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc) {
    constexpr uint32_t LBO = 128, SBO = (K / 4) * 128;
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < 3; term++) {  // small terms first
        const uint32_t a = term == 0 ? a_lo : a_hi, b = term == 1 ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < K / 8; ks++) {  // a k-step of 8 = two core matrices along K
            umma_tf32_ts(d_tmem, a + 8 * ks, make_sdesc(b + ks * 2 * LBO, LBO, SBO), idesc, acc);
            acc = 1;
        }
    }
}

// The same product with the two terms that share A_hi issued as ONE instruction of twice the width: the hi and lo images of B
// are contiguous in shared memory (the lo image starts where row group N/8 of the hi image would), so a descriptor of 2N rows
// at b_hi reads [B_hi | B_lo]:   D[:, 0:N) = A_hi.B_hi + A_lo.B_hi,   D[:, N:2N) = A_hi.B_lo   (the caller adds the two blocks).
// 2 K/8 instead of 3 K/8 instructions; an instruction costs ~48 cycles at N = 64 and ~70 at N = 128 (scripts/tc_probe.cu).
template <int N, int K>
__device__ __forceinline__ void issue_3xtf32_wide(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t idesc_2n, uint32_t idesc_n) {
    constexpr uint32_t LBO = 128, SBO = (K / 4) * 128;
#pragma unroll
    for (int ks = 0; ks < K / 8; ks++)
        umma_tf32_ts(d_tmem, a_hi + 8 * ks, make_sdesc(b_hi + ks * 2 * LBO, LBO, SBO), idesc_2n, ks > 0 ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < K / 8; ks++)
        umma_tf32_ts(d_tmem, a_lo + 8 * ks, make_sdesc(b_hi + ks * 2 * LBO, LBO, SBO), idesc_n, 1u);
}

}  // namespace umma
}  // namespace pmp
