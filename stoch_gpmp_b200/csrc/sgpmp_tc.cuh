// tcgen05 (5th-generation tensor core) helpers shared by the two GEMM-shaped kernels: the weighted covariance (sgpmp_cov.cu)
// and the dense-L sampling variant (sgpmp_sample_dense.cu).  Hand-built descriptors, TF32 operands from shared memory in the
// canonical no-swizzle K-major core-matrix layout, fp32 accumulator in tensor memory (TMEM).
#pragma once
#include "sgpmp_common.cuh"

namespace sgpmp {
namespace tc {

constexpr int TILE = 128;         // UMMA M = UMMA N = 128 (one accumulator tile = 128 TMEM lanes x 128 columns)
constexpr int KC = 32;            // K elements staged per chunk (4 UMMA K steps of 8 tf32)
// A core matrix is 8 rows x 16 bytes, contiguous (128 B).  K-adjacent core matrices are placed 144 B apart (16 B of padding):
// the 8 lanes that stage one row write 16-byte pieces 144 B apart = 8 different bank groups (128 B apart they would all
// hit the same four banks, an 8-way conflict on every STS.128).
constexpr uint32_t LBO = 144;                      // bytes between core matrices adjacent in K
// 8-row groups follow each other 8 LBO + 16 B apart: the extra 16 B keep the 4-byte transposing stores of the dense-L sampler
// (consecutive lanes = rows 4 apart, i.e. every second lane in the next group) off a common bank.
constexpr uint32_t SBO = (KC / 4) * LBO + 16;      // bytes between 8-row groups
constexpr int OP_BYTES = (TILE / 8) * SBO;         // one 128 x 32 operand buffer (18.25 KiB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, K block kb = k / 4) inside an operand buffer
__device__ __forceinline__ uint32_t op_offset(int r, int kb) { return (uint32_t)(r >> 3) * SBO + (uint32_t)(r & 7) * 16 + (uint32_t)kb * LBO; }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, no swizzle, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // start address, bits [0,14)
    d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;             // leading-dimension byte offset, bits [16,30)
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;             // stride-dimension byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version 1
    return d;                                               // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC_TF32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}\n" ::"r"(mbar), "r"(parity) : "memory");
}

// TMEM: 128 columns x 128 lanes of fp32 (one accumulator tile); called by one whole warp
__device__ __forceinline__ void tmem_alloc128(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free128(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(128) : "memory");
}

// D[tmem] (+)= A[smem] B[smem]^T, 128 x 128 x 8, TF32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(IDESC_TF32_128x128), "r"(accumulate) : "memory");
}
// 3xTF32 over one staged chunk (KC = 32): (hi, hi) + (hi, lo) + (lo, hi) per K step of 8; first = this is the first chunk
__device__ __forceinline__ void mma_chunk_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, bool first) {
#pragma unroll
    for (int kk = 0; kk < KC / 8; ++kk) {
        const uint32_t adv = (uint32_t)kk * 2 * LBO;              // 8 tf32 = 2 core matrices along K
        const uint64_t dah = smem_desc(a_hi + adv), dal = smem_desc(a_lo + adv);
        const uint64_t dbh = smem_desc(b_hi + adv), dbl = smem_desc(b_lo + adv);
        mma_tf32(tmem_d, dah, dbh, (first && kk == 0) ? 0u : 1u);
        mma_tf32(tmem_d, dah, dbl, 1u);
        mma_tf32(tmem_d, dal, dbh, 1u);
    }
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive accumulator columns of this thread's TMEM lane (warp w may read lanes [32 (w % 4), 32 (w % 4) + 32))
__device__ __forceinline__ void tmem_load32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace tc
}  // namespace sgpmp
