// K4 — per-particle softmax weights + weighted-mean update, and the split-particle statistics.
//
// Reference being replaced: StochGPMP._update_distribution (planner.py:263-275)
//     w = softmax(-costs/tau, dim=1);  grad = sum_s w_s (x_s - mu);  mu += step_size * grad
// (the set_mean() that follows, planner.py:273 -> mp_priors_multi.py:120-123, rebuilds the whole
// MultivariateNormal — Cholesky + PD validation of NP dense MxM copies — although the precision never
// changes; nothing of that remains here).
//
// Mapping: one CTA per (problem, particle).  Weights live in shared memory; each warp owns rows (t, j)
// of the S-minor sample block and reduces over s with coalesced loads + shuffles.  HBM-bound: M*S reals
// read per particle.
#include "sgpmp_common.cuh"

namespace sgpmp {

// block-wide max / sum via warp shuffles + one shared-memory hop (red: >= 32 reals)
template <typename real>
__device__ __forceinline__ real block_max(real v, real* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    real r = red[0];
    for (int k = 1; k < nw; ++k) r = sg_max(r, red[k]);
    return r;
}
template <typename real>
__device__ __forceinline__ real block_sum(real v, real* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    real r = 0;
    for (int k = 0; k < nw; ++k) r += red[k];
    return r;
}

// e_s = exp(-c_s/tau - m) into wsm[S]; returns (m, Z).  normalise=true divides by Z.
template <typename real>
__device__ __forceinline__ void block_softmax(const real* __restrict__ c, int S, real tau, real* wsm, real* red,
                                              bool normalise, real* m_out, real* Z_out) {
    real m = -INFINITY;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const real z = -c[s] / tau;
        wsm[s] = z;
        m = sg_max(m, z);
    }
    m = block_max(m, red);
    real Z = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const real e = sg_exp(wsm[s] - m);
        wsm[s] = e;
        Z += e;
    }
    Z = block_sum(Z, red);
    if (normalise)
        for (int s = threadIdx.x; s < S; s += blockDim.x) wsm[s] = wsm[s] / Z;
    __syncthreads();
    *m_out = m;
    *Z_out = Z;
}

template <typename real>
__global__ void __launch_bounds__(256)
update_kernel(int S, int M, real tau, real step, const real* __restrict__ costs, const real* __restrict__ samples,
              real* __restrict__ means, real* __restrict__ grad, real* __restrict__ weights, real* __restrict__ means_pre) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* wsm = reinterpret_cast<real*>(smem_raw);   // [S]
    real* red = wsm + S;                             // [32]
    const size_t bp = blockIdx.x;
    real m, Z;
    pdl_launch_dependents();
    pdl_wait();                  // everything below reads what the cost kernel wrote
    block_softmax<real>(costs + bp * S, S, tau, wsm, red, true, &m, &Z);
    if (weights && blockIdx.y == 0)
        for (int s = threadIdx.x; s < S; s += blockDim.x) weights[bp * S + s] = wsm[s];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const real* xs = samples + bp * (size_t)M * S;
    // gridDim.y > 1 (few particles): the state rows of a particle are divided over gridDim.y CTAs, each of which recomputes
    // the (cheap) softmax — with one planning problem a CTA per particle would leave all but NP SMs idle
    const int rows_per = ((M + (int)gridDim.y - 1) / (int)gridDim.y + 3) & ~3;
    const int row_lo = (int)blockIdx.y * rows_per, row_hi = min(M, row_lo + rows_per);
    // four rows per warp pass: four independent load streams in flight (the kernel is HBM-latency bound otherwise)
    for (int r0 = row_lo + warp * 4; r0 < row_hi; r0 += nw * 4) {
        real acc[4] = {0, 0, 0, 0}, mu[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) mu[q] = (r0 + q < row_hi) ? means[bp * M + r0 + q] : (real)0;
        // unrolled by four: 16 independent loads in flight per lane instead of 4 (the loop is latency-bound otherwise)
#pragma unroll 4
        for (int s = lane; s < S; s += 32) {
            const real w = wsm[s];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (r0 + q < row_hi) acc[q] += w * (xs[(size_t)(r0 + q) * S + s] - mu[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const real a = warp_sum(acc[q]);
            if (lane == 0 && r0 + q < row_hi) {
                if (grad) grad[bp * M + r0 + q] = a;
                if (means_pre) means_pre[bp * M + r0 + q] = mu[q];        // the means this iteration sampled from (planner.py:252-253)
                means[bp * M + r0 + q] = mu[q] + step * a;
            }
        }
    }
}

// split-particle mode: stats[bp] = (m, Z, A[M]),  A = sum_s exp(-c_s/tau - m) eps_s
template <typename real>
__global__ void __launch_bounds__(256)
local_stats_kernel(int S, int M, real tau, const real* __restrict__ costs, const real* __restrict__ eps,
                   real* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* wsm = reinterpret_cast<real*>(smem_raw);
    real* red = wsm + S;
    const size_t bp = blockIdx.x;
    real m, Z;
    block_softmax<real>(costs + bp * S, S, tau, wsm, red, false, &m, &Z);
    real* out = stats + bp * (size_t)(M + 2);
    if (threadIdx.x == 0) { out[0] = m; out[1] = Z; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const real* es = eps + bp * (size_t)M * S;
    for (int r = warp; r < M; r += nw) {
        real acc = 0;
        for (int s = lane; s < S; s += 32) acc += wsm[s] * es[(size_t)r * S + s];
        acc = warp_sum(acc);
        if (lane == 0) out[2 + r] = acc;
    }
}

// mu += step * L (A / Z).  One CTA per particle, three phases: (1) all threads merge and normalise the M entries of A — coalesced
// over the entry index — and stage the particle's mean and the (G, H) tables in shared memory; (2) one thread per DoF runs the banded
// recurrence over t out of shared memory; (3) all threads write grad and the updated mean back, coalesced.
// n_ranks > 1: `stats` holds the gathered blocks [n_ranks][n_particles][M + 2] of the split-particle exchange; they are merged
// here by log-sum-exp in a fixed rank order (m = max_r m_r, Z = sum_r Z_r e^(m_r - m), A = sum_r A_r e^(m_r - m)) — every rank
// computes the identical update, no host arithmetic in between.
// (Round 2's first form gave a (particle, DoF) to ONE thread that walked T steps with 2 R strided loads and a read-modify-write
// of the mean in every step: 64 problems, R = 8: 176 us per launch — more than half of a split-mode iteration; same expressions,
// same results here.)
template <typename real>
__global__ void __launch_bounds__(256)
apply_stats_kernel(int n_particles, int T, int n, const double* __restrict__ tab, real step,
                   const real* __restrict__ stats, int n_ranks, real* __restrict__ means, real* __restrict__ grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = 2 * n, M = T * d;
    real* a = reinterpret_cast<real*>(smem_raw);     // [M] merged A / Z, then grad
    real* mu = a + M;                                // [M]
    real* gh = mu + M;                               // [T][7]
    real* scale = gh + (size_t)T * 7;                // [n_ranks] e^(m_r - m)
    __shared__ real s_invZ;
    const int bp = blockIdx.x;
    const size_t rank_stride = (size_t)n_particles * (M + 2);
    const real* st = stats + (size_t)bp * (M + 2);
    if (threadIdx.x == 0) {
        real m = st[0];
        for (int r = 1; r < n_ranks; ++r) m = sg_max(m, st[r * rank_stride]);
        real Z = 0;
        for (int r = 0; r < n_ranks; ++r) {
            const real sc = sg_exp(st[r * rank_stride] - m);
            scale[r] = sc;
            Z += st[r * rank_stride + 1] * sc;
        }
        s_invZ = (real)1 / Z;
    }
    for (int k = threadIdx.x; k < T * 7; k += blockDim.x) gh[k] = (real)tab[(size_t)(k / 7) * SGPMP_TABLE_STRIDE + (k % 7)];
    for (int k = threadIdx.x; k < M; k += blockDim.x) mu[k] = means[(size_t)bp * M + k];
    __syncthreads();
    const real invZ = s_invZ;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        real e = 0;
        if (n_ranks == 1) {
            e = st[2 + k];
        } else {
            for (int q = 0; q < n_ranks; ++q) e += st[q * rank_stride + 2 + k] * scale[q];
        }
        a[k] = e * invZ;
    }
    __syncthreads();
    if (threadIdx.x < n) {
        const int i = threadIdx.x;
        real yp = 0, yv = 0;
        for (int t = 0; t < T; ++t) {
            const real* r = gh + t * 7;          // g11 g21 g22 h11 h12 h21 h22
            const real ep = a[t * d + i], ev = a[t * d + n + i];
            const real np_ = r[0] * ep - (r[3] * yp + r[4] * yv);
            const real nv_ = r[1] * ep + r[2] * ev - (r[5] * yp + r[6] * yv);
            yp = np_; yv = nv_;
            a[t * d + i] = yp;
            a[t * d + n + i] = yv;
            mu[t * d + i] += step * yp;
            mu[t * d + n + i] += step * yv;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        if (grad) grad[(size_t)bp * M + k] = a[k];
        means[(size_t)bp * M + k] = mu[k];
    }
}

template <typename real>
static int launch_update(const sgpmp_shape_t& sh, double tau, double step, const void* costs, const void* samples,
                         void* means, void* grad, void* weights, cudaStream_t st, int row_chunks = 1, void* means_pre = nullptr,
                         bool pdl = false) {
    const int NP = sh.G * sh.K, M = sh.T * 2 * sh.n_dof;
    const size_t smem = ((size_t)sh.S + 32) * sizeof(real);
    if (smem > SGPMP_SMEM_OPTIN) cudaFuncSetAttribute(update_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_kernel(update_kernel<real>, dim3((unsigned)(sh.B * NP), (unsigned)row_chunks), dim3(256), smem, st, pdl, sh.S, M, (real)tau,
                  (real)step, (const real*)costs, (const real*)samples, (real*)means, (real*)grad, (real*)weights, (real*)means_pre);
    SGPMP_CHECK_LAUNCH("sgpmp_update");
    return SGPMP_OK;
}

template <typename real>
static int launch_local_stats(const sgpmp_shape_t& sh, double tau, const void* costs, const void* eps, void* stats,
                              cudaStream_t st) {
    const int NP = sh.G * sh.K, M = sh.T * 2 * sh.n_dof;
    const size_t smem = ((size_t)sh.S + 32) * sizeof(real);
    if (smem > SGPMP_SMEM_OPTIN) cudaFuncSetAttribute(local_stats_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    local_stats_kernel<real><<<(unsigned)(sh.B * NP), 256, smem, st>>>(sh.S, M, (real)tau, (const real*)costs,
                                                                      (const real*)eps, (real*)stats);
    SGPMP_CHECK_LAUNCH("sgpmp_local_stats");
    return SGPMP_OK;
}

template <typename real>
static int launch_apply_stats(const sgpmp_shape_t& sh, const double* tables, double step, const void* stats, void* means,
                              void* grad, cudaStream_t st, int n_ranks = 1) {
    const int n_particles = sh.B * sh.G * sh.K;
    const size_t M = (size_t)sh.T * 2 * sh.n_dof;
    const size_t smem = (2 * M + (size_t)sh.T * 7 + (size_t)n_ranks) * sizeof(real);
    if (smem > SGPMP_SMEM_OPTIN) {
        if (smem > 227 * 1024) { set_error("sgpmp_apply_stats: T=%d too large for shared memory", sh.T); return SGPMP_ERR_UNSUPPORTED; }
        cudaFuncSetAttribute(apply_stats_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    apply_stats_kernel<real><<<(unsigned)n_particles, 256, smem, st>>>(n_particles, sh.T, sh.n_dof, tables, (real)step,
                                                                      (const real*)stats, n_ranks, (real*)means, (real*)grad);
    SGPMP_CHECK_LAUNCH("sgpmp_apply_stats");
    return SGPMP_OK;
}

int update_launch(const sgpmp_shape_t& sh, double tau, double step, const void* costs, const void* samples, void* means,
                  void* grad, void* weights, int row_chunks, cudaStream_t st, void* means_pre, bool pdl) {
    if (sh.dtype == SGPMP_F32) return launch_update<float>(sh, tau, step, costs, samples, means, grad, weights, st, row_chunks, means_pre, pdl);
    return launch_update<double>(sh, tau, step, costs, samples, means, grad, weights, st, row_chunks, means_pre, pdl);
}

int merge_apply_stats_launch(const sgpmp_shape_t& sh, const double* tables, double step, const void* stats_all, int n_ranks,
                             void* means, void* grad, cudaStream_t st) {
    if (sh.dtype == SGPMP_F32) return launch_apply_stats<float>(sh, tables, step, stats_all, means, grad, st, n_ranks);
    return launch_apply_stats<double>(sh, tables, step, stats_all, means, grad, st, n_ranks);
}

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int sgpmp_merge_apply_stats(const sgpmp_shape_t* shape, const double* tables, double step_size, const void* stats_all,
                                       int32_t n_ranks, void* means, void* grad, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_merge_apply_stats: invalid shape");
    SGPMP_REQUIRE(tables && stats_all && means && n_ranks >= 1, "sgpmp_merge_apply_stats: invalid argument");
    return merge_apply_stats_launch(*shape, tables, step_size, stats_all, n_ranks, means, grad, (cudaStream_t)stream);
}

extern "C" int sgpmp_update(const sgpmp_shape_t* shape, double temperature, double step_size, const void* costs,
                            const void* samples, void* means, void* grad, void* weights, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_update: invalid shape");
    SGPMP_REQUIRE(costs && samples && means, "sgpmp_update: null pointer");
    SGPMP_REQUIRE(temperature > 0, "sgpmp_update: temperature must be > 0");
    if (shape->dtype == SGPMP_F32)
        return launch_update<float>(*shape, temperature, step_size, costs, samples, means, grad, weights, (cudaStream_t)stream);
    return launch_update<double>(*shape, temperature, step_size, costs, samples, means, grad, weights, (cudaStream_t)stream);
}

extern "C" int sgpmp_local_stats(const sgpmp_shape_t* shape, double temperature, const void* costs, const void* eps,
                                 void* stats, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_local_stats: invalid shape");
    SGPMP_REQUIRE(costs && eps && stats, "sgpmp_local_stats: null pointer");
    SGPMP_REQUIRE(temperature > 0, "sgpmp_local_stats: temperature must be > 0");
    if (shape->dtype == SGPMP_F32)
        return launch_local_stats<float>(*shape, temperature, costs, eps, stats, (cudaStream_t)stream);
    return launch_local_stats<double>(*shape, temperature, costs, eps, stats, (cudaStream_t)stream);
}

extern "C" int sgpmp_apply_stats(const sgpmp_shape_t* shape, const double* tables, double step_size, const void* stats,
                                 void* means, void* grad, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_apply_stats: invalid shape");
    SGPMP_REQUIRE(tables && stats && means, "sgpmp_apply_stats: null pointer");
    if (shape->dtype == SGPMP_F32)
        return launch_apply_stats<float>(*shape, tables, step_size, stats, means, grad, (cudaStream_t)stream);
    return launch_apply_stats<double>(*shape, tables, step_size, stats, means, grad, (cudaStream_t)stream);
}
