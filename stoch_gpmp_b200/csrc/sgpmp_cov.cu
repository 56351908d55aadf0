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
//   * one CTA (256 threads) per 128 x 128 tile of the upper triangle of C and particle; K = S is consumed in chunks of 32 samples;
//   * per chunk the 256 threads stage the A and B blocks with coalesced 128-byte row segments (8 lanes per row, one 16-byte
//     core-matrix row per lane): subtract the mean, scale by sqrt(w_s), split into a TF32 head and a TF32 tail (3xTF32: hi*hi + hi*lo + lo*hi keeps ~2^-21 relative accuracy, plain TF32 would
//     give 2^-11) and stores them in the canonical no-swizzle K-major core-matrix layout (8 rows x 16 bytes per core matrix);
//   * fence.proxy.async + __syncthreads, then ONE thread issues 4 (k steps of 8) x 3 tcgen05.mma.kind::tf32 with shared-memory
//     descriptors for A and B, accumulating into 128 lanes x 128 columns of TMEM, and commits them to an mbarrier;
//   * two shared-memory stages: chunk c+1 is staged while the MMAs of chunk c run; a stage is reused after waiting on its mbarrier;
//   * epilogue: each warp reads its 32 TMEM lanes (two 32-column blocks per warp) with tcgen05.ld.32x32b.x32 and writes rows of C.
// The operation is tiny (2 M^2 S = 0.8 GFLOP per Panda particle) and its operands need arithmetic before they reach the tensor
// core (mean subtraction, weighting, TF32 split), so the staging is done by the CTA's threads rather than by TMA.
// fp64: CUDA-core kernel, one thread per entry (fp64 planners are parity configurations).
#include "sgpmp_tc.cuh"

namespace sgpmp {

using tc::smem_u32;
constexpr int COV_TILE = tc::TILE;     // rows of the A block = rows of the B block = UMMA M = UMMA N
constexpr int COV_KC = tc::KC;         // samples per chunk (4 UMMA K steps of 8 tf32)
constexpr uint32_t COV_LBO = tc::LBO, COV_SBO = tc::SBO;
constexpr int COV_OP_BYTES = tc::OP_BYTES;
constexpr int COV_THREADS = 256;       // 8 warps stage; one thread issues the MMAs; all 8 warps drain TMEM
constexpr int COV_STAGES = 2;          // shared-memory stages: chunk c+1 is staged while the MMAs of chunk c run

__global__ void __launch_bounds__(COV_THREADS, 1)
weighted_cov_tc_kernel(int M, int S, int tiles, const float* __restrict__ samples, const float* __restrict__ means,
                       const float* __restrict__ weights, float* __restrict__ cov) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar[COV_STAGES];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    // C is symmetric: only the tiles with tj >= ti are computed (tiles (tiles + 1) / 2 CTAs per particle); off-diagonal tiles
    // are stored twice (as computed and transposed)
    const int bp = blockIdx.y;
    int ti = 0, tj = (int)blockIdx.x;
    while (tj >= tiles - ti) { tj -= tiles - ti; ++ti; }
    tj += ti;
    const int i0 = ti * COV_TILE, j0 = tj * COV_TILE;
    const float* xs = samples + (size_t)bp * M * S;
    const float* mu = means + (size_t)bp * M;
    const float* w = weights + (size_t)bp * S;
    float* C = cov + (size_t)bp * M * M;

    if (warp == 0) tc::tmem_alloc128(&tmem_base_s);      // TMEM: 128 columns x 128 lanes of fp32 for the accumulator tile
    if (tid == 0) {
        for (int k = 0; k < COV_STAGES; ++k) tc::mbar_init(&mbar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_acc = tmem_base_s;

    // Staging map: thread = (row group rg = tid / 8, K block kb = tid % 8).  In pass it the thread handles row
    // RPP it + rg of the block (RPP = COV_THREADS / 8 rows per pass) and its samples [s0 + 4 kb, s0 + 4 kb + 4): the 8 lanes of a row read one contiguous 128-byte line
    // (coalesced, every line touched exactly once), and what a thread holds is exactly one 16-byte core-matrix row.
    const int rg = tid >> 3, kb = tid & 7;
    const bool same = (ti == tj);
    const bool vec = (S & 3) == 0;
    const int n_chunks = (S + COV_KC - 1) / COV_KC;
    for (int c = 0; c < n_chunks; ++c) {
        // stage buffers of this chunk; they are free once the MMAs of chunk c - COV_STAGES have completed (k-th arrival on the
        // stage's mbarrier completes its phase k)
        const int stg = c % COV_STAGES;
        unsigned char* a_hi = smem_raw + (size_t)stg * 4 * COV_OP_BYTES;
        unsigned char* a_lo = a_hi + COV_OP_BYTES;
        unsigned char* b_hi = a_lo + COV_OP_BYTES;
        unsigned char* b_lo = b_hi + COV_OP_BYTES;
        const int s0 = c * COV_KC + 4 * kb;
        float sw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) sw[q] = (s0 + q < S) ? sqrtf(w[s0 + q]) : 0.f;
        // ---- stage Y = sqrt(w) (x - mu) for this chunk, split into TF32 head and tail ---------------------------------
        // all 16 row segments of this thread are requested before any is consumed: one HBM/L2 round trip per chunk
        constexpr int RPP = COV_THREADS / 8, NIT = COV_TILE / RPP;
        float4 v[2][NIT];
        float m[2][NIT];
#pragma unroll
        for (int op = 0; op < 2; ++op) {
            if (op == 1 && same) break;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int row = (op ? j0 : i0) + RPP * it + rg;
                v[op][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                m[op][it] = 0.f;
                if (row < M) {
                    m[op][it] = mu[row];
                    const float* src = xs + (size_t)row * S + s0;
                    if (vec) {
                        if (s0 < S) v[op][it] = *reinterpret_cast<const float4*>(src);
                    } else {
                        if (s0 + 0 < S) v[op][it].x = src[0];
                        if (s0 + 1 < S) v[op][it].y = src[1];
                        if (s0 + 2 < S) v[op][it].z = src[2];
                        if (s0 + 3 < S) v[op][it].w = src[3];
                    }
                }
            }
        }
        if (c >= COV_STAGES) {      // (the global loads above are already in flight)
            tc::mbar_wait(smem_u32(&mbar[stg]), (uint32_t)(((c / COV_STAGES) - 1) & 1));
            tc::fence_after_sync();
        }
#pragma unroll
        for (int op = 0; op < 2; ++op) {
            if (op == 1 && same) break;
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int r = RPP * it + rg;
                const uint32_t off = tc::op_offset(r, kb);
                const float x4[4] = {v[op][it].x, v[op][it].y, v[op][it].z, v[op][it].w};
                float hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float y = sw[q] * (x4[q] - m[op][it]);       // rows >= M hold x = m = 0, samples >= S have sw = 0
                    hi[q] = tc::to_tf32(y);
                    lo[q] = tc::to_tf32(y - hi[q]);
                }
                *reinterpret_cast<float4*>((op ? b_hi : a_hi) + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>((op ? b_lo : a_lo) + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        tc::fence_async_smem();                  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        // ---- one thread issues the MMAs of this chunk and commits them to the stage's mbarrier -----------------------
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo);
            tc::mma_chunk_3xtf32(tmem_acc, sa_hi, sa_lo, same ? sa_hi : smem_u32(b_hi), same ? sa_lo : smem_u32(b_lo), c == 0);
            tc::mma_commit(&mbar[stg]);
        }
    }
    // the last commit covers every MMA issued before it
    tc::mbar_wait(smem_u32(&mbar[(n_chunks - 1) % COV_STAGES]), (uint32_t)(((n_chunks - 1) / COV_STAGES) & 1));
    tc::fence_after_sync();
    // ---- epilogue: TMEM -> registers -> C.  A warp may read the TMEM lanes [32 (w % 4), 32 (w % 4) + 32) = rows i0 + 32 (w % 4) +
    // lane; warps 0-3 take the column blocks 0 and 1, warps 4-7 the blocks 2 and 3 ---------------------------------------------
    const int lane_grp = warp & 3;
    const int row = i0 + lane_grp * 32 + (tid & 31);
    constexpr int CB_PER_WARP = (COV_TILE / 32) / (COV_THREADS / 128);
#pragma unroll 1
    for (int cb = (warp >> 2) * CB_PER_WARP; cb < (warp >> 2) * CB_PER_WARP + CB_PER_WARP; ++cb) {
        uint32_t r[32];
        tc::tmem_load32(tmem_acc + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(cb * 32), r);
        if (row < M) {
            const int c0 = j0 + cb * 32;
            if ((M & 3) == 0 && c0 + 32 <= M) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<float4*>(C + (size_t)row * M + c0 + q) =
                        make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    if (c0 + q < M) C[(size_t)row * M + c0 + q] = __uint_as_float(r[q]);
            }
            if (!same) {            // transposed copy: consecutive lanes write consecutive addresses
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    if (c0 + q < M) C[(size_t)(c0 + q) * M + row] = __uint_as_float(r[q]);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free128(tmem_acc);
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
        const size_t smem = (size_t)COV_STAGES * 4 * COV_OP_BYTES + 1024;
        cudaFuncSetAttribute(weighted_cov_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        weighted_cov_tc_kernel<<<dim3((unsigned)(tiles * (tiles + 1) / 2), (unsigned)BP), COV_THREADS, smem, st>>>(
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
