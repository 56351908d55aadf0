// Fused StochGPMP optimisation loop: sample -> cost -> softmax -> weighted update, n_iters times, in
// ONE launch.  Replaces the Python loop of StochGPMP.optimize (planner.py:277-317) and everything it
// calls per iteration (sample_and_eval :239-261, _get_costs :229-237, _update_distribution :263-275).
//
// Design (B200-first, not a translation):
//   * one CTA per (problem, particle); a particle's S samples never leave the SM.  Every thread owns one
//     trajectory sample at a time and marches over t with the whole state in registers: Philox normals
//     -> banded recurrence y_t = G_t eps_t - H_t y_{t-1} -> x_t = mu_t + y_t -> cost factors/FK/obstacles.
//     No sample is written to HBM unless the caller asks for the last iteration's samples
//     (get_recent_samples, planner.py:330-337, only ever observes those).
//   * softmax over the particle's S costs in shared memory.
//   * the update uses grad = L (sum_s w_s eps_s)  (x_s - mu = L eps_s): the weighted eps-sum is
//     re-generated from the counter-based RNG with the transposed mapping thread <-> (time pair, DoF),
//     looping over s and skipping samples whose weight is exactly zero (a CTA-uniform branch: in the
//     reference's shipped configurations the softmax is one-hot, so this pass is almost free), then one
//     banded recurrence per DoF applies L.  mu and b = Sigma^-1 mu stay in shared memory across
//     iterations; particles are independent, so no grid-wide synchronisation exists anywhere.
//   * FP32-pipe/MUFU bound; HBM traffic is O(NP*M) per iteration instead of 3*M*S*NP.
#include <stdlib.h>

#include "sgpmp_common.cuh"
#include "sgpmp_cost.cuh"
#include "sgpmp_rng.cuh"

#ifndef SGPMP_MINB128
#define SGPMP_MINB128 5
#endif
#ifndef SGPMP_MINB256
#define SGPMP_MINB256 3
#endif

namespace sgpmp {

template <typename real>
struct IterArgs {
    int G, K, S, T, n_iters;
    uint32_t particle_gid0, sample_gid0;
    real step;
    RngKey key;            // key.draw = draw index of iteration 0
    const double* tab;
    const real* eps_in;    // [n_iters][B*NP][T][d][S] or null
    real* means;           // [B*NP][T][d] in/out
    real* means_pre;       // optional
    real* samples;         // optional, last iteration
    real* costs;           // optional, last iteration
    real* weights;         // optional, last iteration
    real* grad;            // optional, last iteration
};

template <typename real, int N, int BS, int CHAIN>
__global__ void __launch_bounds__(BS, (sizeof(real) == 4 ? (BS == 128 ? SGPMP_MINB128 : SGPMP_MINB256) : 1))
iterate_kernel(const __grid_constant__ CostParams<real> P, const __grid_constant__ IterArgs<real> A) {
    constexpr int d = 2 * N;
    constexpr int DP = (d + 3) & ~3;      // padded row length of mu / b: rows stay 16-byte aligned for LDS.128
    const int T = A.T, S = A.S, G = A.G, K = A.K;
    const int TP = (T + 1) >> 1;
    const int M = T * d;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* sph = reinterpret_cast<real*>(smem_raw);            // [MAX_SPHERES][8] + coll_const (16-byte aligned)
    real* tabGH = sph + SPH_SMEM;                             // [T][8]  (g11 g21 g22 h11 | h12 h21 h22 -)
    real* mu = tabGH + (size_t)T * 8;                         // [T][DP]
    real* bvec = mu + (size_t)T * DP;                         // [T][DP]
    double* tabDO = reinterpret_cast<double*>(bvec + (size_t)T * DP);   // [T][7]
    real* acc = reinterpret_cast<real*>(tabDO + (size_t)T * 7);         // [T][d]  sum_s w_s eps_s, then grad
    real* wsm = acc + M;                                      // [S]
    real* part = wsm + S;                                     // [4*BS] partial sums of pass 2
    real* red = part + 4 * BS;                                // [32]
    real* start = red + 32;                                   // [d]
    real* goal = start + d;                                   // [d]

    const int NP = G * K;
    const int bp = blockIdx.x, b = bp / NP, p = bp - b * NP;
    const int tid = threadIdx.x;
    const uint32_t pgid = A.particle_gid0 + (uint32_t)bp;

    for (int k = tid; k < T * 8; k += BS) {
        const int t = k >> 3, c = k & 7;
        const double* row = A.tab + (size_t)t * SGPMP_TABLE_STRIDE;
        tabGH[k] = c < 7 ? (real)row[c] : (real)0;
        if (c < 7) tabDO[t * 7 + c] = row[SGPMP_TAB_D11 + c];
    }
    for (int k = tid; k < T * DP; k += BS) {
        const int t = k / DP, j = k - t * DP;
        mu[k] = j < d ? A.means[(size_t)bp * M + t * d + j] : (real)0;
        bvec[k] = 0;
    }
    stage_cta_constants<real, N, CHAIN>(P, b, p / K, G, start, goal, sph);
    CostSmem<real> sm;
    sm.start = start; sm.goal = goal; sm.bvec = bvec; sm.sph = sph;
    sm.coll_const = sph[SPH_STRIDE * SGPMP_MAX_SPHERES];
    sm.self_const = sph[SPH_STRIDE * SGPMP_MAX_SPHERES + 1];
    sm.map = P.has_map ? P.occ_map + (size_t)(P.map_of_problem ? P.map_of_problem[b] : 0) * P.map_h * P.map_w : nullptr;
    __syncthreads();

    for (int it = 0; it < A.n_iters; ++it) {
        const bool last = (it == A.n_iters - 1);
        RngKey key = A.key;
        key.draw += (uint32_t)it;
        const real* eps = A.eps_in ? A.eps_in + ((size_t)it * gridDim.x + bp) * (size_t)M * S : nullptr;

        // ---- b = Sigma^-1 mu (fp64 accumulate) ---------------------------------------------------------
        for (int k = tid; k < T * N; k += BS) {
            const int t = k / N, i = k - t * N;
            precision_times_row<real, DP>(tabDO, mu, T, N, t, i, &bvec[t * DP + i], &bvec[t * DP + N + i]);
        }
        if (last && A.means_pre)
            for (int k = tid; k < M; k += BS) A.means_pre[(size_t)bp * M + k] = mu[(k / d) * DP + (k % d)];
        __syncthreads();

        // ---- pass 1: sample + cost, one thread per trajectory sample -----------------------------------
        const bool emit = last && A.samples != nullptr;
        for (int s = tid; s < S; s += BS) {
            TrajCost<real, N, CHAIN> tc;
            tc.begin();
            real yp[N], yv[N];
#pragma unroll
            for (int i = 0; i < N; ++i) { yp[i] = 0; yv[i] = 0; }
            // One Philox call per DoF yields the normals of TWO time steps; the step body is kept as a single
            // (not 2x unrolled) copy so that the hot loop stays inside the instruction cache.
            real en[d];
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                real e[d];
                if (eps) {
#pragma unroll
                    for (int j = 0; j < d; ++j) e[j] = eps[((size_t)t * d + j) * S + s];
                } else if ((t & 1) == 0) {
#pragma unroll
                    for (int i = 0; i < N; ++i) normal4<real>(key, t >> 1, i, A.sample_gid0 + (uint32_t)s, pgid, e[i], e[N + i], en[i], en[N + i]);
                } else {
#pragma unroll
                    for (int j = 0; j < d; ++j) e[j] = en[j];
                }
                real r[8], m[DP];
                load4(tabGH + t * 8, r[0], r[1], r[2], r[3]);
                load4(tabGH + t * 8 + 4, r[4], r[5], r[6], r[7]);
#pragma unroll
                for (int k = 0; k < DP / 4; ++k) load4(mu + t * DP + 4 * k, m[4 * k], m[4 * k + 1], m[4 * k + 2], m[4 * k + 3]);
                real x[d];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const real np_ = r[0] * e[i] - (r[3] * yp[i] + r[4] * yv[i]);
                    const real nv_ = r[1] * e[i] + r[2] * e[N + i] - (r[5] * yp[i] + r[6] * yv[i]);
                    yp[i] = np_; yv[i] = nv_;
                    x[i] = m[i] + np_;
                    x[N + i] = m[N + i] + nv_;
                }
                tc.step(P, sm, t, T, x, bvec + t * DP);
                if (emit) {
#pragma unroll
                    for (int j = 0; j < d; ++j) A.samples[((size_t)bp * M + (size_t)t * d + j) * S + s] = x[j];
                }
            }
            tc.finish(P, sm, T);
            const real c = tc.total();
            wsm[s] = c;
            if (last && A.costs) A.costs[(size_t)bp * S + s] = c;
        }
        __syncthreads();

        // ---- softmax over the S samples of this particle ------------------------------------------------
        {
            real m = -INFINITY;
            for (int s = tid; s < S; s += BS) {
                const real z = -wsm[s] / P.temperature;
                wsm[s] = z;
                m = sg_max(m, z);
            }
            m = warp_max(m);
            if ((tid & 31) == 0) red[tid >> 5] = m;
            __syncthreads();
            m = red[0];
            for (int k = 1; k < BS / 32; ++k) m = sg_max(m, red[k]);
            __syncthreads();
            real Z = 0;
            for (int s = tid; s < S; s += BS) {
                const real ex = sg_exp(wsm[s] - m);
                wsm[s] = ex;
                Z += ex;
            }
            Z = warp_sum(Z);
            if ((tid & 31) == 0) red[tid >> 5] = Z;
            __syncthreads();
            Z = 0;
            for (int k = 0; k < BS / 32; ++k) Z += red[k];
            for (int s = tid; s < S; s += BS) {
                const real w = wsm[s] / Z;
                wsm[s] = w;
                if (last && A.weights) A.weights[(size_t)bp * S + s] = w;
            }
            __syncthreads();
        }

        // ---- pass 2: acc = sum_s w_s eps_s --------------------------------------------------------------
        if (eps) {
            const int warp = tid >> 5, lane = tid & 31;
            for (int r = warp; r < M; r += BS / 32) {
                real a = 0;
                for (int s = lane; s < S; s += 32) a += wsm[s] * eps[(size_t)r * S + s];
                a = warp_sum(a);
                if (lane == 0) acc[r] = a;
            }
        } else {
            const int n_items = TP * N;
            const int nsplit = (n_items < BS) ? (BS / n_items) : 1;
            for (int base = 0; base < n_items; base += BS) {
                const int item = base + (tid % (nsplit > 1 ? n_items : BS));
                const int split = (nsplit > 1) ? tid / n_items : 0;
                const bool active = item < n_items && split < nsplit;
                real a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                const int tp = item / N, i = item - tp * N;
                if (active) {
                    for (int s = split; s < S; s += nsplit) {
                        const real w = wsm[s];
                        if (w == (real)0) continue;
                        real p0, v0, p1, v1;
                        normal4<real>(key, tp, i, A.sample_gid0 + (uint32_t)s, pgid, p0, v0, p1, v1);
                        a0 += w * p0; a1 += w * v0; a2 += w * p1; a3 += w * v1;
                    }
                }
                if (nsplit > 1) {
                    part[4 * tid + 0] = a0; part[4 * tid + 1] = a1; part[4 * tid + 2] = a2; part[4 * tid + 3] = a3;
                    __syncthreads();
                    if (active && split == 0) {
                        for (int q = 1; q < nsplit; ++q) {
                            const real* pp = part + 4 * (q * n_items + item);
                            a0 += pp[0]; a1 += pp[1]; a2 += pp[2]; a3 += pp[3];
                        }
                    }
                }
                if (active && split == 0) {
                    const int t0 = 2 * tp;
                    acc[t0 * d + i] = a0;
                    acc[t0 * d + N + i] = a1;
                    if (t0 + 1 < T) {
                        acc[(t0 + 1) * d + i] = a2;
                        acc[(t0 + 1) * d + N + i] = a3;
                    }
                }
            }
        }
        __syncthreads();

        // ---- grad = L acc (banded recurrence, one thread per DoF); mu += step * grad --------------------
        if (tid < N) {
            const int i = tid;
            real gp_ = 0, gv_ = 0;
            for (int t = 0; t < T; ++t) {
                const real* r = tabGH + t * 8;
                const real ep = acc[t * d + i], ev = acc[t * d + N + i];
                const real np_ = r[0] * ep - (r[3] * gp_ + r[4] * gv_);
                const real nv_ = r[1] * ep + r[2] * ev - (r[5] * gp_ + r[6] * gv_);
                gp_ = np_; gv_ = nv_;
                acc[t * d + i] = gp_;
                acc[t * d + N + i] = gv_;
                mu[t * DP + i] += A.step * gp_;
                mu[t * DP + N + i] += A.step * gv_;
            }
        }
        __syncthreads();
        if (last && A.grad)
            for (int k = tid; k < M; k += BS) A.grad[(size_t)bp * M + k] = acc[k];
    }
    for (int k = tid; k < M; k += BS) A.means[(size_t)bp * M + k] = mu[(k / d) * DP + (k % d)];
}

template <typename real, int N, int BS, int CHAIN>
static int launch_iterate_nb(const sgpmp_shape_t& sh, const CostParams<real>& P, const IterArgs<real>& A, cudaStream_t st) {
    const int d = 2 * N, M = sh.T * d, DP = (d + 3) & ~3;
    const size_t smem = (size_t)sh.T * 7 * sizeof(double) +
                        ((size_t)sh.T * (8 + 2 * DP) + (size_t)M + sh.S + 4 * BS + 32 + 2 * d + SPH_SMEM) * sizeof(real);
    if (smem > 227 * 1024) {
        set_error("sgpmp_iterate: T=%d, S=%d need %zu bytes of shared memory (> 227 KiB)", sh.T, sh.S, smem);
        return SGPMP_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(iterate_kernel<real, N, BS, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    iterate_kernel<real, N, BS, CHAIN><<<(unsigned)(sh.B * sh.G * sh.K), BS, smem, st>>>(P, A);
    SGPMP_CHECK_LAUNCH("sgpmp_iterate");
    return SGPMP_OK;
}

template <typename real, int N, int CHAIN>
static int launch_iterate_n(const sgpmp_shape_t& sh, const CostParams<real>& P, const IterArgs<real>& A, cudaStream_t st) {
    static const char* force_bs = getenv("SGPMP_ITERATE_BS");   // tuning aid
    if (sh.S > 128 && !(force_bs && atoi(force_bs) == 128)) return launch_iterate_nb<real, N, 256, CHAIN>(sh, P, A, st);
    return launch_iterate_nb<real, N, 128, CHAIN>(sh, P, A, st);
}

template <typename real>
static int launch_iterate(const sgpmp_shape_t& sh, const sgpmp_cost_desc_t& desc, const double* tables, double step,
                          int n_iters, const void* eps_in, uint64_t seed, uint32_t draw0, void* means, void* means_pre,
                          void* samples, void* costs, void* weights, void* grad, cudaStream_t st) {
    CostParams<real> P;
    int rc = lower_cost_desc<real>(sh, desc, P);
    if (rc != SGPMP_OK) return rc;
    if (!(desc.temperature > 0)) { set_error("sgpmp_iterate: temperature must be > 0"); return SGPMP_ERR_INVALID_ARG; }
    IterArgs<real> A;
    A.G = sh.G; A.K = sh.K; A.S = sh.S; A.T = sh.T; A.n_iters = n_iters;
    A.particle_gid0 = (uint32_t)(sh.problem_gid0 * sh.G * sh.K);
    A.sample_gid0 = (uint32_t)sh.sample_gid0;
    A.step = (real)step;
    A.key = RngKey{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), draw0};
    A.tab = tables;
    A.eps_in = (const real*)eps_in;
    A.means = (real*)means; A.means_pre = (real*)means_pre; A.samples = (real*)samples;
    A.costs = (real*)costs; A.weights = (real*)weights; A.grad = (real*)grad;
    if constexpr (sizeof(real) == 4) {
        if ((P.has_spheres || P.has_self) && chain_is_panda_structure(desc, sh.n_dof))
            return P.has_self ? launch_iterate_n<real, 7, 2>(sh, P, A, st) : launch_iterate_n<real, 7, 1>(sh, P, A, st);
    }
    switch (sh.n_dof) {
#define SGPMP_DOF_CASE(N) case N: return launch_iterate_n<real, N, 0>(sh, P, A, st);
#include "sgpmp_dof_list.inc"
#undef SGPMP_DOF_CASE
        default:
            set_error("sgpmp_iterate: n_dof=%d is not instantiated (see sgpmp_dof_list.inc)", sh.n_dof);
            return SGPMP_ERR_UNSUPPORTED;
    }
}

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int sgpmp_iterate(const sgpmp_shape_t* shape, const sgpmp_cost_desc_t* desc, const double* tables,
                             double step_size, int32_t n_iters, const void* eps_in, uint64_t seed, uint32_t draw0,
                             void* means, void* means_pre, void* samples, void* costs, void* weights, void* grad,
                             void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_iterate: invalid shape");
    SGPMP_REQUIRE(desc && tables && means, "sgpmp_iterate: null pointer");
    SGPMP_REQUIRE(n_iters >= 1, "sgpmp_iterate: n_iters must be >= 1");
    if (shape->dtype == SGPMP_F32)
        return launch_iterate<float>(*shape, *desc, tables, step_size, n_iters, eps_in, seed, draw0, means, means_pre,
                                     samples, costs, weights, grad, (cudaStream_t)stream);
    return launch_iterate<double>(*shape, *desc, tables, step_size, n_iters, eps_in, seed, draw0, means, means_pre,
                                  samples, costs, weights, grad, (cudaStream_t)stream);
}
