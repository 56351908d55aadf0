// K2 — trajectory sampling x = mu + L eps as a banded recurrence.
//
// Reference being replaced: MultiMPPrior.sample (mp_priors_multi.py:204-207) ->
// MultivariateNormal.rsample (multivariate_normal.py:251-254): eps = normal(S,NP,M);
// loc + L @ eps with a dense M x M L per particle (2 NP S M^2 flop, 84-91 % of the reference's
// iteration).  L^-1 is block-bidiagonal, so  y_t = G_t eps_t - H_t y_{t-1}  (16 flop per state).
//
// Mapping: one thread per (sample, DoF pair) — the unit of the RNG stream (sgpmp_rng.cuh); the S-minor layout
// [B,NP,T,d,S] makes every load/store of a warp one contiguous 128-byte (fp32) line.  The G/H tables (7 reals per
// step) sit in shared memory.
#include "sgpmp_common.cuh"
#include "sgpmp_rng.cuh"

namespace sgpmp {

template <typename real>
__global__ void __launch_bounds__(256)
sample_kernel(int NP, int S, int T, int n, int64_t particle_gid0, uint32_t sample_gid0, const double* __restrict__ tab,
              const real* __restrict__ means, const real* __restrict__ eps_in, RngKey key,
              real* __restrict__ samples, real* __restrict__ eps_out, int mu_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* gh = reinterpret_cast<real*>(smem_raw);   // [T][7]
    const int bp = blockIdx.x;                       // flat (problem, particle)
    const int d = 2 * n;
    // the particle's mean is staged in shared memory when the launcher reserved room for it (mu_in_smem): read from global
    // memory it puts one L2 round trip per time step on the in-order path of a thread that walks T steps (one problem, B = 1:
    // 29 us -> 9 us for the whole kernel)
    real* mu_s = gh + (size_t)T * 7;
    for (int k = threadIdx.x; k < T * 7; k += blockDim.x)
        gh[k] = (real)tab[(size_t)(k / 7) * SGPMP_TABLE_STRIDE + (k % 7)];
    if (mu_in_smem)
        for (int k = threadIdx.x; k < T * d; k += blockDim.x) mu_s[k] = means[(size_t)bp * T * d + k];
    __syncthreads();

    const int n_pairs = (n + 1) >> 1;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= S * n_pairs) return;
    const int k = idx / S, s = idx - k * S;
    const int i0 = 2 * k;
    const bool full = (i0 + 1 < n);
    const size_t base = (size_t)bp * T * d * S;      // samples / eps
    const real* mu = mu_in_smem ? mu_s : means + (size_t)bp * T * d;
    const uint32_t pgid = (uint32_t)(particle_gid0 + bp);

    real yp[2] = {0, 0}, yv[2] = {0, 0};
    for (int t = 0; t < T; ++t) {
        real ep[2] = {0, 0}, ev[2] = {0, 0};
        if (eps_in) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (h == 0 || full) {
                    ep[h] = eps_in[base + ((size_t)t * d + i0 + h) * S + s];
                    ev[h] = eps_in[base + ((size_t)t * d + n + i0 + h) * S + s];
                }
        } else {
            normal_pair<real>(key, (uint32_t)t, (uint32_t)k, full, sample_gid0 + (uint32_t)s, pgid, ep[0], ep[1], ev[0], ev[1]);
        }
        const real* r = gh + t * 7;
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (h == 0 || full) {
                const real np_ = r[0] * ep[h] - (r[3] * yp[h] + r[4] * yv[h]);
                const real nv_ = r[1] * ep[h] + r[2] * ev[h] - (r[5] * yp[h] + r[6] * yv[h]);
                yp[h] = np_; yv[h] = nv_;
                const size_t op = base + ((size_t)t * d + i0 + h) * S + s;
                const size_t ov = base + ((size_t)t * d + n + i0 + h) * S + s;
                samples[op] = mu[t * d + i0 + h] + np_;
                samples[ov] = mu[t * d + n + i0 + h] + nv_;
                if (eps_out) { eps_out[op] = ep[h]; eps_out[ov] = ev[h]; }
            }
    }
}

// Few-samples variant (one planning problem; C5: long horizons): one CTA per (particle, DoF pair, block of SB samples), time in
// CHUNKS of TC steps.  Per chunk the normals of all TC steps are drawn first, by all 256 threads in parallel over (time step,
// sample) — the counter-based stream does not care who draws — into shared memory; only then does one thread per (sample, DoF of
// the pair) walk the (cheap) recurrence over the chunk, carrying (y_pos, y_vel) to the next chunk in registers.  In sample_kernel a
// thread draws AND walks, i.e. T dependent Philox / Box-Muller chains back to back: with 2,048 samples that is a few lone warps
// (20 us at B = 1, T = 64; 1.37 ms at T = 1024 in fp64).  Same stream, same recurrence expressions: bit-identical output.
// SB (samples per CTA) is chosen by the launcher so that the grid covers the SMs (64 ... 8), TC so that the chunk fits in
// shared memory with several CTAs per SM (one CTA's walk then runs under another's draw).
template <typename real, int SB>
__global__ void __launch_bounds__(256)
sample_tiled_kernel(int NP, int S, int T, int TC, int n, int64_t particle_gid0, uint32_t sample_gid0, const double* __restrict__ tab,
                    const real* __restrict__ means, RngKey key, real* __restrict__ samples) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* gh = reinterpret_cast<real*>(smem_raw);       // [TC][7]
    real* mu_k = gh + (size_t)TC * 7;                   // [TC][4]  (pos 2k, pos 2k+1, vel 2k, vel 2k+1) mean of this pair
    real* eps = mu_k + (size_t)TC * 4;                  // [TC][4][SB]  same component order
    const int bp = blockIdx.x, k = blockIdx.y, s0 = blockIdx.z * SB;
    const int d = 2 * n, i0 = 2 * k;
    const bool full = (i0 + 1 < n);
    const uint32_t pgid = (uint32_t)(particle_gid0 + bp);
    const int h = threadIdx.x / SB, sl = threadIdx.x - h * SB, s = s0 + sl;
    const bool walker = !(h >= 2 || s >= S || (h == 1 && !full));
    const size_t base = (size_t)bp * T * d * S;
    real yp = 0, yv = 0;
    pdl_launch_dependents();
    for (int t0 = 0; t0 < T; t0 += TC) {
        const int tc = min(TC, T - t0);
        if (t0) __syncthreads();                         // the walkers are done with the previous chunk
        for (int q = threadIdx.x; q < tc * 7; q += blockDim.x)
            gh[q] = (real)tab[(size_t)(t0 + q / 7) * SGPMP_TABLE_STRIDE + (q % 7)];
        for (int item = threadIdx.x; item < tc * SB; item += blockDim.x) {
            const int tl = item / SB, isl = item - tl * SB, is = s0 + isl;
            if (is < S) {
                real p0, p1, v0, v1;
                normal_pair<real>(key, (uint32_t)(t0 + tl), (uint32_t)k, full, sample_gid0 + (uint32_t)is, pgid, p0, p1, v0, v1);
                real* e = eps + (size_t)tl * 4 * SB + isl;
                e[0] = p0; e[SB] = p1; e[2 * SB] = v0; e[3 * SB] = v1;
            }
        }
        // PDL: the tables and the draw above depend on no other kernel; the means (written by the previous iteration's update)
        // and the sample buffer (read by it) do — the first chunk's draw runs under the tail of the predecessor
        if (t0 == 0) pdl_wait();
        for (int q = threadIdx.x; q < tc * 4; q += blockDim.x) {
            const int t = t0 + (q >> 2), c = q & 3, hh = c & 1, a = c >> 1;
            mu_k[q] = (hh == 0 || full) ? means[(size_t)bp * T * d + (size_t)t * d + a * n + i0 + hh] : (real)0;
        }
        __syncthreads();
        if (walker) {
            for (int tl = 0; tl < tc; ++tl) {
                const real* r = gh + tl * 7;
                const real e0 = eps[((size_t)tl * 4 + h) * SB + sl], e1 = eps[((size_t)tl * 4 + 2 + h) * SB + sl];
                const real np_ = r[0] * e0 - (r[3] * yp + r[4] * yv);
                const real nv_ = r[1] * e0 + r[2] * e1 - (r[5] * yp + r[6] * yv);
                yp = np_; yv = nv_;
                const int t = t0 + tl;
                samples[base + ((size_t)t * d + i0 + h) * S + s] = mu_k[4 * tl + h] + np_;
                samples[base + ((size_t)t * d + n + i0 + h) * S + s] = mu_k[4 * tl + 2 + h] + nv_;
            }
        }
    }
}

// In-kernel-RNG variant for the instantiated DoF counts: one thread per SAMPLE with all DoFs in registers (the
// mapping of the fused kernel's pass 1).  The generic kernel above spends 44 k instructions per trajectory sample
// (per-thread index arithmetic for one DoF pair); this one ~23 k, which moves the kernel from 35 % to >60 % of the HBM
// write roofline.  Row (t, j) of the S-minor output is written by consecutive threads: fully coalesced.
template <typename real, int N>
__global__ void __launch_bounds__(128)
sample_rng_kernel(int S, int T, int64_t particle_gid0, uint32_t sample_gid0, const double* __restrict__ tab,
                  const real* __restrict__ means, RngKey key, real* __restrict__ samples, real* __restrict__ eps_out) {
    constexpr int d = 2 * N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* gh = reinterpret_cast<real*>(smem_raw);   // [T][8]
    real* mu = gh + (size_t)T * 8;                  // [T][d]
    const int bp = blockIdx.x;
    for (int k = threadIdx.x; k < T * 8; k += blockDim.x) gh[k] = (k & 7) < 7 ? (real)tab[(size_t)(k >> 3) * SGPMP_TABLE_STRIDE + (k & 7)] : (real)0;
    for (int k = threadIdx.x; k < T * d; k += blockDim.x) mu[k] = means[(size_t)bp * T * d + k];
    __syncthreads();
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const uint32_t pgid = (uint32_t)(particle_gid0 + bp);
    real* out = samples + (size_t)bp * T * d * S + s;
    real* eo = eps_out ? eps_out + (size_t)bp * T * d * S + s : nullptr;
    real yp[N], yv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { yp[i] = 0; yv[i] = 0; }
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        real e[d + 2];
#pragma unroll
        for (int k = 0; k < (N + 1) / 2; ++k) {
            real p1, v1;
            normal_pair<real>(key, (uint32_t)t, (uint32_t)k, 2 * k + 1 < N, sample_gid0 + (uint32_t)s, pgid, e[2 * k], p1, e[N + 2 * k], v1);
            if (2 * k + 1 < N) { e[2 * k + 1] = p1; e[N + 2 * k + 1] = v1; }
        }
        const real* r = gh + t * 8;
        const real* m = mu + t * d;
        real* row = out + (size_t)t * d * S;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const real np_ = r[0] * e[i] - (r[3] * yp[i] + r[4] * yv[i]);
            const real nv_ = r[1] * e[i] + r[2] * e[N + i] - (r[5] * yp[i] + r[6] * yv[i]);
            yp[i] = np_; yv[i] = nv_;
            row[(size_t)i * S] = m[i] + np_;
            row[(size_t)(N + i) * S] = m[N + i] + nv_;
        }
        if (eo) {
            real* er = eo + (size_t)t * d * S;
#pragma unroll
            for (int j = 0; j < d; ++j) er[(size_t)j * S] = e[j];
        }
    }
}

template <typename real, int N>
static int launch_sample_rng(const sgpmp_shape_t& sh, const double* tables, const void* means, uint64_t seed, uint32_t draw,
                             void* samples, void* eps_out, cudaStream_t st) {
    const int NP = sh.G * sh.K, bs = 128;
    dim3 grid((unsigned)(sh.B * NP), (unsigned)((sh.S + bs - 1) / bs));
    const size_t smem = (size_t)sh.T * (8 + 2 * N) * sizeof(real);
    if (smem > SGPMP_SMEM_OPTIN) {
        if (smem > 227 * 1024) return SGPMP_ERR_UNSUPPORTED;      // caller falls back to the generic kernel
        cudaFuncSetAttribute(sample_rng_kernel<real, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    RngKey key = make_rng_key(seed, draw);
    sample_rng_kernel<real, N><<<grid, bs, smem, st>>>(sh.S, sh.T, sh.problem_gid0 * NP, (uint32_t)sh.sample_gid0, tables,
                                                       (const real*)means, key, (real*)samples, (real*)eps_out);
    SGPMP_CHECK_LAUNCH("sgpmp_sample");
    return SGPMP_OK;
}

template <typename real>
static int launch_sample(const sgpmp_shape_t& sh, const double* tables, const void* means, const void* eps_in,
                         uint64_t seed, uint32_t draw, void* samples, void* eps_out, cudaStream_t st, bool few = false,
                         bool pdl = false) {
    const long n_samples_total = (long)sh.B * sh.G * sh.K * sh.S;
    if (!eps_in && !eps_out && (few || n_samples_total < 148L * 512)) {
        const int n_pairs = (sh.n_dof + 1) / 2;
        const int NPf = sh.G * sh.K;
        const long per_block = (long)sh.B * NPf * n_pairs;
        int SB = 64;                                     // the largest tile that still gives every SM a CTA
        while (SB > 8 && per_block * ((sh.S + SB - 1) / SB) < 148) SB >>= 1;
        const size_t per_step = (size_t)(11 + 4 * SB) * sizeof(real);
        // the whole horizon as one chunk while that leaves two CTAs per SM, else chunks of <= 40 KB (five CTAs per SM)
        int TC = sh.T;
        if ((size_t)sh.T * per_step > 96 * 1024) TC = (int)((40 * 1024) / per_step);
        const size_t smem = (size_t)TC * per_step;
        if (TC >= 8 && n_pairs <= 65535 && (sh.S + SB - 1) / SB <= 65535) {
            RngKey keyf = make_rng_key(seed, draw);
            const dim3 grid((unsigned)(sh.B * NPf), (unsigned)n_pairs, (unsigned)((sh.S + SB - 1) / SB));
#define SGPMP_TILED(SBV)                                                                                                              \
            do {                                                                                                                      \
                if (smem > SGPMP_SMEM_OPTIN) cudaFuncSetAttribute(sample_tiled_kernel<real, SBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                launch_kernel(sample_tiled_kernel<real, SBV>, grid, dim3(256), smem, st, pdl, NPf, sh.S, sh.T, TC, sh.n_dof,          \
                              (int64_t)sh.problem_gid0 * NPf, (uint32_t)sh.sample_gid0, tables, (const real*)means, keyf, (real*)samples); \
            } while (0)
            if (SB == 64) SGPMP_TILED(64); else if (SB == 32) SGPMP_TILED(32); else if (SB == 16) SGPMP_TILED(16); else SGPMP_TILED(8);
#undef SGPMP_TILED
            SGPMP_CHECK_LAUNCH("sgpmp_sample(tiled)");
            return SGPMP_OK;
        }
    }
    // In-kernel RNG: thread-per-sample kernel for the instantiated DoF counts — when there are enough samples to fill the
    // machine with one thread each.  Few samples with long horizons (C5: one problem, T up to 1024) are latency-bound on the
    // sequential recurrence, so they take the thread-per-(sample, DoF) kernel below: n times the threads, the same stream
    // (T = 1024, n = 7, fp64, 2,048 samples: 2.69 ms -> see profiles/r1/c5_sweep.jsonl).
    if (!eps_in && n_samples_total >= 148L * 512) {
        int rc = SGPMP_ERR_UNSUPPORTED;
        switch (sh.n_dof) {
#define SGPMP_DOF_CASE(N) case N: rc = launch_sample_rng<real, N>(sh, tables, means, seed, draw, samples, eps_out, st); break;
#include "sgpmp_dof_list.inc"
#undef SGPMP_DOF_CASE
            default: break;
        }
        if (rc != SGPMP_ERR_UNSUPPORTED) return rc;
    }
    const int NP = sh.G * sh.K;
    const int bs = 256;
    dim3 grid((unsigned)(sh.B * NP), (unsigned)((sh.S * ((sh.n_dof + 1) / 2) + bs - 1) / bs));
    size_t smem = (size_t)sh.T * 7 * sizeof(real);
    const size_t mu_bytes = (size_t)sh.T * 2 * sh.n_dof * sizeof(real);
    const int mu_in_smem = (smem + mu_bytes <= 40 * 1024) ? 1 : 0;
    if (mu_in_smem) smem += mu_bytes;
    if (smem > SGPMP_SMEM_OPTIN) {
        if (smem > 227 * 1024) { set_error("sgpmp_sample: T=%d too large for the shared-memory tables", sh.T); return SGPMP_ERR_UNSUPPORTED; }
        cudaFuncSetAttribute(sample_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    RngKey key = make_rng_key(seed, draw);
    sample_kernel<real><<<grid, bs, smem, st>>>(NP, sh.S, sh.T, sh.n_dof, sh.problem_gid0 * NP, (uint32_t)sh.sample_gid0, tables,
                                                (const real*)means, (const real*)eps_in, key, (real*)samples,
                                                (real*)eps_out, mu_in_smem);
    SGPMP_CHECK_LAUNCH("sgpmp_sample");
    return SGPMP_OK;
}

int sample_launch(const sgpmp_shape_t& sh, const double* tables, const void* means, const void* eps_in, uint64_t seed,
                  uint32_t draw, void* samples, cudaStream_t st, bool pdl) {
    // called by the low-latency iteration only: few samples, so the tiled kernel (parallel draw, then the recurrence)
    if (sh.dtype == SGPMP_F32) return launch_sample<float>(sh, tables, means, eps_in, seed, draw, samples, nullptr, st, true, pdl);
    return launch_sample<double>(sh, tables, means, eps_in, seed, draw, samples, nullptr, st, true, pdl);
}

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int sgpmp_sample(const sgpmp_shape_t* shape, const double* tables, const void* means,
                            const void* eps_in, uint64_t seed, uint32_t draw, void* samples, void* eps_out,
                            void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_sample: invalid shape");
    SGPMP_REQUIRE(tables && means && samples, "sgpmp_sample: null pointer");
    SGPMP_REQUIRE((int64_t)shape->B * shape->G * shape->K <= 0x7fffffffLL, "sgpmp_sample: too many particles for one launch");
    if (shape->dtype == SGPMP_F32)
        return launch_sample<float>(*shape, tables, means, eps_in, seed, draw, samples, eps_out, (cudaStream_t)stream);
    return launch_sample<double>(*shape, tables, means, eps_in, seed, draw, samples, eps_out, (cudaStream_t)stream);
}
