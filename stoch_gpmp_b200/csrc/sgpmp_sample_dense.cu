// K2, dense-L variant on the tensor cores: x = mu + L eps as a GEMM per DoF,
//   X_i[(t,a)][s] = sum_k L1[(t,a)][k] E_i[k][s],   L1 = the per-DoF 2T x 2T scale_tril, k = (t', a'), i = DoF, s = sample.
//
// This is what the reference computes (MultivariateNormal.rsample: loc + L @ eps, torch multivariate_normal.py:251-254, with the
// dense M x M L; because P decouples per DoF, SURVEY §8a-2, L is block-diagonal over the DoFs after a permutation and the dense
// contraction is n independent 2T x 2T ones).  The north-star names it as the tensor-core variant of the sampler and SURVEY §8(d)
// as the second K2 variant.  It is NOT the product path: it issues 2 (2T)^2 = 32,768 flop per (sample, DoF) (x3 for the 3xTF32
// split that holds 1e-5) against 16 T = 1,024 for the banded recurrence of sgpmp_sample.cu, and bench_kernels.py shows the
// banded kernel ahead on the same inputs; it exists to put a measured number behind that design decision.
//
// One CTA (256 threads) per (particle, DoF, 128-sample tile) and per 128-row tile of L1; K is consumed in chunks of 32:
//   A chunk = L1[m0 .. m0+127][k0 .. k0+31]   row-major = K-major, staged with coalesced float4 loads
//   B chunk = E_i[k0 .. k0+31][s0 .. s0+127]  S-minor in global memory = "MN-major"; transposed into the K-major core-matrix
//             layout while staging (coalesced float4 loads along s, four 4-byte shared stores each)
// both split into TF32 head and tail (tc::mma_chunk_3xtf32), accumulator in TMEM, epilogue adds the mean and writes 128
// contiguous samples per row.  L1 is lower triangular: K chunks beyond the row tile are skipped.
#include "sgpmp_tc.cuh"

namespace sgpmp {

constexpr int SD_THREADS = 256;
constexpr int SD_STAGES = 2;

__global__ void __launch_bounds__(SD_THREADS, 1)
sample_dense_tc_kernel(int T, int n, int S, const float* __restrict__ L1, const float* __restrict__ means,
                       const float* __restrict__ eps, float* __restrict__ samples) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar[SD_STAGES];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int d = 2 * n, K2 = 2 * T;                       // per-DoF system size
    const int s_tiles = (S + TILE - 1) / TILE;
    const int s0 = (int)(blockIdx.x % s_tiles) * TILE, i = (int)(blockIdx.x / s_tiles);     // sample tile, DoF
    const int m0 = (int)blockIdx.y * TILE;                 // row tile of L1
    const int bp = blockIdx.z;
    const float* E = eps + (size_t)bp * T * d * S;
    float* X = samples + (size_t)bp * T * d * S;
    const float* mu = means + (size_t)bp * T * d;
    // per-DoF index r = 2 t + a  ->  row (t, a n + i) of the [T, d, S] arrays
    auto grow = [&](int r) { return (size_t)((r >> 1) * d + (r & 1) * n + i) * S; };

    if (warp == 0) tmem_alloc128(&tmem_base_s);
    if (tid == 0) {
        for (int k = 0; k < SD_STAGES; ++k) mbar_init(&mbar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_acc = tmem_base_s;

    const int rg = tid >> 3, kb = tid & 7;                 // A staging: (row group, K block of 4)
    const int kq = tid >> 5, sq = tid & 31;                // B staging: (k row within a pass, group of 4 samples)
    const int k_end = min(K2, m0 + TILE);                  // lower triangular: columns beyond the row tile are zero
    const int n_chunks = (k_end + KC - 1) / KC;
    const bool vec = (S & 3) == 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int stg = c % SD_STAGES, k0 = c * KC;
        unsigned char* a_hi = smem_raw + (size_t)stg * 4 * OP_BYTES;
        unsigned char* a_lo = a_hi + OP_BYTES;
        unsigned char* b_hi = a_lo + OP_BYTES;
        unsigned char* b_lo = b_hi + OP_BYTES;
        // ---- global loads of this chunk (all in flight before the stage is waited for) --------------------------------
        constexpr int RPP = SD_THREADS / 8, NIT = TILE / RPP;      // A: 32 rows per pass, 4 passes
        float4 va[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int row = m0 + RPP * it + rg, k = k0 + 4 * kb;
            va[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < K2 && k < K2) va[it] = *reinterpret_cast<const float4*>(L1 + (size_t)row * K2 + k);   // K2 is even; k % 4 == 0
        }
        constexpr int KPP = SD_THREADS / 32, NKT = KC / KPP;       // B: 8 k rows per pass, 4 passes; 32 lanes x 4 samples = 128
        float4 vb[NKT];
#pragma unroll
        for (int it = 0; it < NKT; ++it) {
            const int k = k0 + KPP * it + kq, s = s0 + 4 * sq;
            vb[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K2) {
                const float* src = E + grow(k) + s;
                if (vec) {
                    if (s < S) vb[it] = *reinterpret_cast<const float4*>(src);
                } else {
                    if (s + 0 < S) vb[it].x = src[0];
                    if (s + 1 < S) vb[it].y = src[1];
                    if (s + 2 < S) vb[it].z = src[2];
                    if (s + 3 < S) vb[it].w = src[3];
                }
            }
        }
        if (c >= SD_STAGES) {
            mbar_wait(smem_u32(&mbar[stg]), (uint32_t)(((c / SD_STAGES) - 1) & 1));
            fence_after_sync();
        }
        // ---- A: one 16-byte core-matrix row per (row, K block) -----------------------------------------------------------
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const float x4[4] = {va[it].x, va[it].y, va[it].z, va[it].w};
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { hi[q] = to_tf32(x4[q]); lo[q] = to_tf32(x4[q] - hi[q]); }
            const uint32_t off = op_offset(RPP * it + rg, kb);
            *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        // ---- B: transpose while staging: B row = sample, K = k; this thread holds 4 samples of one k ---------------------
#pragma unroll
        for (int it = 0; it < NKT; ++it) {
            const int kl = KPP * it + kq;                            // k within the chunk
            const float x4[4] = {vb[it].x, vb[it].y, vb[it].z, vb[it].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float h = to_tf32(x4[q]), l = to_tf32(x4[q] - h);
                const uint32_t off = op_offset(4 * sq + q, kl >> 2) + (uint32_t)(kl & 3) * 4;
                *reinterpret_cast<float*>(b_hi + off) = h;
                *reinterpret_cast<float*>(b_lo + off) = l;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            mma_chunk_3xtf32(tmem_acc, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), c == 0);
            mma_commit(&mbar[stg]);
        }
    }
    mbar_wait(smem_u32(&mbar[(n_chunks - 1) % SD_STAGES]), (uint32_t)(((n_chunks - 1) / SD_STAGES) & 1));
    fence_after_sync();
    // ---- epilogue: lane = row (t, a) of this DoF, columns = samples; x = mu + L eps -------------------------------------------
    const int lane_grp = warp & 3;
    const int r = m0 + lane_grp * 32 + (tid & 31);
    constexpr int CB_PER_WARP = (TILE / 32) / (SD_THREADS / 128);
#pragma unroll 1
    for (int cb = (warp >> 2) * CB_PER_WARP; cb < (warp >> 2) * CB_PER_WARP + CB_PER_WARP; ++cb) {
        uint32_t acc[32];
        tmem_load32(tmem_acc + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(cb * 32), acc);
        if (r < K2) {
            const float m = mu[(r >> 1) * d + (r & 1) * n + i];
            float* dst = X + grow(r) + s0 + cb * 32;
            if (vec && s0 + cb * 32 + 32 <= S) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<float4*>(dst + q) = make_float4(m + __uint_as_float(acc[q]), m + __uint_as_float(acc[q + 1]),
                                                                      m + __uint_as_float(acc[q + 2]), m + __uint_as_float(acc[q + 3]));
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    if (s0 + cb * 32 + q < S) dst[q] = m + __uint_as_float(acc[q]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free128(tmem_acc);
}

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int sgpmp_sample_dense_tc(const sgpmp_shape_t* shape, const void* L1, const void* means, const void* eps_in,
                                     void* samples, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_sample_dense_tc: invalid shape");
    SGPMP_REQUIRE(L1 && means && eps_in && samples, "sgpmp_sample_dense_tc: null pointer (the dense variant takes injected eps)");
    if (shape->dtype != SGPMP_F32) { set_error("sgpmp_sample_dense_tc: fp32 only (TF32 tensor cores)"); return SGPMP_ERR_UNSUPPORTED; }
    const int BP = shape->B * shape->G * shape->K, T = shape->T, n = shape->n_dof, S = shape->S;
    SGPMP_REQUIRE(BP <= 65535, "sgpmp_sample_dense_tc: at most 65535 particles per call (got %d)", BP);
    SGPMP_REQUIRE((2 * T) % 4 == 0, "sgpmp_sample_dense_tc: traj_len must be even");
    const int s_tiles = (S + tc::TILE - 1) / tc::TILE, m_tiles = (2 * T + tc::TILE - 1) / tc::TILE;
    SGPMP_REQUIRE(m_tiles <= 65535, "sgpmp_sample_dense_tc: traj_len too large");
    const size_t smem = (size_t)SD_STAGES * 4 * tc::OP_BYTES + 1024;
    cudaFuncSetAttribute(sample_dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sample_dense_tc_kernel<<<dim3((unsigned)(s_tiles * n), (unsigned)m_tiles, (unsigned)BP), SD_THREADS, smem, (cudaStream_t)stream>>>(
        T, n, S, (const float*)L1, (const float*)means, (const float*)eps_in, (float*)samples);
    SGPMP_CHECK_LAUNCH("sgpmp_sample_dense_tc");
    return SGPMP_OK;
}
