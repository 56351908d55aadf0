// Weighted sample covariance of every particle,  C_p = sum_s w_s (x_s - mu)(x_s - mu)^T  in R^{M x M},  on the 5th-generation
// tensor cores (tcgen05.mma, accumulator in TMEM).
//
// NO REFERENCE COUNTERPART: the reference never updates the covariance (Sigma^-1 is fixed, planner.py:226); the north-star
// lists "the covariance outer-product sum on tensor cores" next to the weighted-mean update, and SURVEY §8 a-13 asks for it
// as an optional diagnostic (parity unpinned; checked against numpy / the fp64 CUDA-core kernel below).  It is the only
// GEMM-shaped operation of the whole path:  C = Y Y^T  with  Y[i][s] = sqrt(w_s) (x[i][s] - mu[i]),  and the S-minor sample
// layout [M][S] is exactly the K-major operand layout the tensor core wants for both A and B.
//
// fp32 kernel (sm_100a):
//   * one CTA (128 threads) per 128 x 128 tile of C and particle; K = S is consumed in chunks of 32 samples;
//   * per chunk every thread prepares ONE row of the A block and one of the B block: subtract the mean, scale by sqrt(w_s),
//     split into a TF32 head and a TF32 tail (3xTF32: hi*hi + hi*lo + lo*hi keeps ~2^-21 relative accuracy, plain TF32 would
//     give 2^-11) and stores them in the canonical no-swizzle K-major core-matrix layout (8 rows x 16 bytes per core matrix);
//   * fence.proxy.async + __syncthreads, then ONE thread issues 4 (k steps of 8) x 3 tcgen05.mma.kind::tf32 with shared-memory
//     descriptors for A and B, accumulating into 128 lanes x 128 columns of TMEM, and commits them to an mbarrier;
//   * everybody waits on the mbarrier (the shared-memory chunk may then be overwritten);
//   * epilogue: each warp reads its 32 TMEM lanes with tcgen05.ld.32x32b.x32 and writes rows of C.
// The operation is tiny (2 M^2 S = 0.8 GFLOP per Panda particle), so the kernel favours a simple, serial load -> MMA
// chunk loop over a TMA/mbarrier pipeline.
// fp64: CUDA-core kernel, one thread per entry (fp64 planners are parity configurations).
#include "sgpmp_common.cuh"

namespace sgpmp {

constexpr int COV_TILE = 128;     // rows of the A block = rows of the B block = UMMA M = UMMA N
constexpr int COV_KC = 32;        // samples per chunk (4 UMMA K steps of 8 tf32)
constexpr uint32_t COV_LBO = 128;                    // bytes between core matrices adjacent in K
constexpr uint32_t COV_SBO = (COV_KC / 4) * 128;     // bytes between 8-row groups: all K core matrices of a group are contiguous
constexpr int COV_OP_BYTES = COV_TILE * COV_KC * 4;  // one operand buffer (16 KiB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, no swizzle, version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // start address, bits [0,14)
    d |= (uint64_t)((COV_LBO >> 4) & 0x3FFF) << 16;         // leading-dimension byte offset, bits [16,30)
    d |= (uint64_t)((COV_SBO >> 4) & 0x3FFF) << 32;         // stride-dimension byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version 1
    return d;                                               // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t COV_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COV_TILE >> 3) << 17) | ((uint32_t)(COV_TILE >> 4) << 24);

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
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

__global__ void __launch_bounds__(128, 1)
weighted_cov_tc_kernel(int M, int S, int tiles, const float* __restrict__ samples, const float* __restrict__ means,
                       const float* __restrict__ weights, float* __restrict__ cov) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* a_hi = smem_raw;
    unsigned char* a_lo = a_hi + COV_OP_BYTES;
    unsigned char* b_hi = a_lo + COV_OP_BYTES;
    unsigned char* b_lo = b_hi + COV_OP_BYTES;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int bp = blockIdx.y, ti = blockIdx.x / tiles, tj = blockIdx.x - ti * tiles;
    const int i0 = ti * COV_TILE, j0 = tj * COV_TILE;
    const float* xs = samples + (size_t)bp * M * S;
    const float* mu = means + (size_t)bp * M;
    const float* w = weights + (size_t)bp * S;
    float* C = cov + (size_t)bp * M * M;

    if (warp == 0) {      // TMEM: 128 columns x 128 lanes of fp32 for the accumulator tile
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = tmem_base_s;

    // this thread's operand rows: row tid of the A block (state component i0 + tid) and of the B block (j0 + tid)
    const int ia = i0 + tid, jb = j0 + tid;
    const float mua = ia < M ? mu[ia] : 0.f, mub = jb < M ? mu[jb] : 0.f;
    const uint32_t row_off = (uint32_t)(tid >> 3) * COV_SBO + (uint32_t)(tid & 7) * 16;   // core-matrix row of this thread
    const bool same = (ti == tj);
    uint32_t phase = 0;
    const int n_chunks = (S + COV_KC - 1) / COV_KC;
    for (int c = 0; c < n_chunks; ++c) {
        const int s0 = c * COV_KC;
        // ---- stage Y = sqrt(w) (x - mu) for this chunk, split into TF32 head and tail ---------------------------------
#pragma unroll
        for (int kb = 0; kb < COV_KC / 4; ++kb) {         // one 16-byte core-matrix row (4 samples) at a time
            float ah[4], al[4], bh[4], bl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int s = s0 + 4 * kb + q;
                const float sw = s < S ? sqrtf(w[s]) : 0.f;
                const float ya = (ia < M && s < S) ? sw * (xs[(size_t)ia * S + s] - mua) : 0.f;
                ah[q] = to_tf32(ya);
                al[q] = to_tf32(ya - ah[q]);
                if (!same) {
                    const float yb = (jb < M && s < S) ? sw * (xs[(size_t)jb * S + s] - mub) : 0.f;
                    bh[q] = to_tf32(yb);
                    bl[q] = to_tf32(yb - bh[q]);
                }
            }
            const uint32_t off = row_off + (uint32_t)kb * COV_LBO;
            *reinterpret_cast<float4*>(a_hi + off) = make_float4(ah[0], ah[1], ah[2], ah[3]);
            *reinterpret_cast<float4*>(a_lo + off) = make_float4(al[0], al[1], al[2], al[3]);
            if (!same) {
                *reinterpret_cast<float4*>(b_hi + off) = make_float4(bh[0], bh[1], bh[2], bh[3]);
                *reinterpret_cast<float4*>(b_lo + off) = make_float4(bl[0], bl[1], bl[2], bl[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        // ---- one thread issues the MMAs of this chunk and commits them to the mbarrier -------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo);
            const uint32_t sb_hi = same ? sa_hi : smem_u32(b_hi), sb_lo = same ? sa_lo : smem_u32(b_lo);
#pragma unroll
            for (int kk = 0; kk < COV_KC / 8; ++kk) {
                const uint32_t adv = (uint32_t)kk * 2 * COV_LBO;          // 8 tf32 = 2 core matrices along K
                const uint64_t dah = umma_smem_desc(sa_hi + adv), dal = umma_smem_desc(sa_lo + adv);
                const uint64_t dbh = umma_smem_desc(sb_hi + adv), dbl = umma_smem_desc(sb_lo + adv);
                const uint32_t acc0 = (c > 0 || kk > 0) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                             ::"r"(tmem_acc), "l"(dah), "l"(dbh), "r"(COV_IDESC), "r"(acc0) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                             ::"r"(tmem_acc), "l"(dah), "l"(dbl), "r"(COV_IDESC), "r"(1u) : "memory");
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                             ::"r"(tmem_acc), "l"(dal), "l"(dbh), "r"(COV_IDESC), "r"(1u) : "memory");
            }
            // arrives on the mbarrier when every MMA issued so far has completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        mbar_wait(smem_u32(&mbar), phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- epilogue: TMEM -> registers -> C.  Warp w owns TMEM lanes [32 w, 32 w + 32) = rows i0 + 32 w + lane -----------
    const int row = i0 + tid;
#pragma unroll 1
    for (int cb = 0; cb < COV_TILE / 32; ++cb) {
        uint32_t r[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32);
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
        if (row < M) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int col = j0 + cb * 32 + q;
                if (col < M) C[(size_t)row * M + col] = __uint_as_float(r[q]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(128) : "memory");
}

// fp64 (and cross-check) kernel on the CUDA cores: one thread per entry of C
template <typename real>
__global__ void weighted_cov_simple_kernel(int M, int S, const real* __restrict__ samples, const real* __restrict__ means,
                                           const real* __restrict__ weights, real* __restrict__ cov) {
    const int bp = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= M || j >= M) return;
    const real* xi = samples + ((size_t)bp * M + i) * S;
    const real* xj = samples + ((size_t)bp * M + j) * S;
    const real mi = means[(size_t)bp * M + i], mj = means[(size_t)bp * M + j];
    const real* w = weights + (size_t)bp * S;
    double acc = 0.0;
    for (int s = 0; s < S; ++s) acc += (double)w[s] * (double)(xi[s] - mi) * (double)(xj[s] - mj);
    cov[((size_t)bp * M + i) * M + j] = (real)acc;
}

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int sgpmp_weighted_cov(const sgpmp_shape_t* shape, const void* samples, const void* means, const void* weights,
                                  void* cov, int32_t use_tensor_cores, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_weighted_cov: invalid shape");
    SGPMP_REQUIRE(samples && means && weights && cov, "sgpmp_weighted_cov: null pointer");
    const int BP = shape->B * shape->G * shape->K, M = shape->T * 2 * shape->n_dof, S = shape->S;
    SGPMP_REQUIRE(BP <= 65535, "sgpmp_weighted_cov: at most 65535 particles per call (got %d)", BP);
    cudaStream_t st = (cudaStream_t)stream;
    if (shape->dtype == SGPMP_F32 && use_tensor_cores) {
        const int tiles = (M + COV_TILE - 1) / COV_TILE;
        const size_t smem = 4 * (size_t)COV_OP_BYTES + 1024;
        cudaFuncSetAttribute(weighted_cov_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        weighted_cov_tc_kernel<<<dim3((unsigned)(tiles * tiles), (unsigned)BP), 128, smem, st>>>(
            M, S, tiles, (const float*)samples, (const float*)means, (const float*)weights, (float*)cov);
        SGPMP_CHECK_LAUNCH("sgpmp_weighted_cov");
        return SGPMP_OK;
    }
    dim3 block(32, 8), grid((unsigned)((M + 31) / 32), (unsigned)((M + 7) / 8), (unsigned)BP);
    if (shape->dtype == SGPMP_F32)
        weighted_cov_simple_kernel<float><<<grid, block, 0, st>>>(M, S, (const float*)samples, (const float*)means, (const float*)weights, (float*)cov);
    else
        weighted_cov_simple_kernel<double><<<grid, block, 0, st>>>(M, S, (const double*)samples, (const double*)means, (const double*)weights, (double*)cov);
    SGPMP_CHECK_LAUNCH("sgpmp_weighted_cov");
    return SGPMP_OK;
}
