// Shared declarations of the fused-iteration kernels (sgpmp_iterate.cu, sgpmp_iterate_split.cu).
#pragma once
#include "sgpmp_common.cuh"
#include "sgpmp_rng.cuh"

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
    real* stats_out;       // split-particle mode: [B*NP][M+2] = (m, Z, A) of this launch's samples; the update is skipped
};

// Record-layout fused iteration (fp32): sgpmp_iterate_split.cu.
// chain = 1 (Panda structure, sphere field only) or 2 (+ self-collision field): role-split, state warps + link warps;
// chain = 0: state warps only (no link fields: planar occupancy map or no obstacle cost), n_dof 2, 3, 4, 6.
// Returns SGPMP_ERR_UNSUPPORTED when the shape does not fit (the caller then takes the single-role kernel).
int launch_iterate_split(const sgpmp_shape_t& sh, const CostParams<float>& P, const IterArgs<float>& A, int chain, cudaStream_t st);

}  // namespace sgpmp
