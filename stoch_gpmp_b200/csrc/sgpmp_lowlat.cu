// Low-latency form of the optimisation loop for FEW planning problems (the reference's own use: one problem per planner).
//
// The fused kernel (sgpmp_iterate.cu) gives every trajectory sample to one thread that walks its T time steps: with 4096 problems
// that is 8 million threads, with ONE problem it is 2,048 — a handful of lone warps, each executing ~52 k dependent
// instructions per iteration (89 us per iteration on a B200, of which the SMs are ~98 % idle).  Here an iteration is three short
// launches whose parallelism does not come from the number of problems:
//   K2  sgpmp_sample        thread per (sample, DoF): the same Philox stream, samples to a workspace that stays in L2
//   K3' cost_st_kernel      thread per (sample, time slice): 16 warps share a trajectory's time steps (sgpmp_cost.cu)
//   K4  update_kernel       state rows of a particle divided over CTAs (sgpmp_update.cu)
// all n_iters iterations are enqueued by ONE call from C (a Python-side loop would pay ~15 us of ctypes per launch).
#include <stdlib.h>

#include "sgpmp_common.cuh"

using namespace sgpmp;

extern "C" int sgpmp_iterate_lowlat(const sgpmp_shape_t* shape, const sgpmp_cost_desc_t* desc, const double* tables,
                                    double step_size, int32_t n_iters, const void* eps_in, uint64_t seed, uint32_t draw0,
                                    void* means, void* means_pre, void* samples_ws, void* costs, void* weights, void* grad,
                                    void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_iterate_lowlat: invalid shape");
    SGPMP_REQUIRE(desc && tables && means && samples_ws && costs, "sgpmp_iterate_lowlat: null pointer (samples_ws and costs are required)");
    SGPMP_REQUIRE(n_iters >= 1, "sgpmp_iterate_lowlat: n_iters must be >= 1");
    SGPMP_REQUIRE(desc->temperature > 0, "sgpmp_iterate_lowlat: temperature must be > 0");
    cudaStream_t st = (cudaStream_t)stream;
    const sgpmp_shape_t& sh = *shape;
    const size_t w = sh.dtype == SGPMP_F32 ? 4 : 8;
    const size_t BP = (size_t)sh.B * sh.G * sh.K, M = (size_t)sh.T * 2 * sh.n_dof;
    // state rows per update CTA: enough CTAs to cover the machine, at least 32 rows each
    int row_chunks = (int)((2 * 148 + BP - 1) / BP);
    row_chunks = row_chunks < 1 ? 1 : row_chunks;
    if ((size_t)row_chunks > (M + 31) / 32) row_chunks = (int)((M + 31) / 32);
    // Programmatic dependent launch along the chain K2 -> K3' -> K4 -> K2 ...: each kernel is allowed to start while its predecessor
    // drains, does what depends on no other kernel (table staging, cost constants, K2's whole Philox / Box-Muller draw) and only then
    // waits for the predecessor's completion (pdl_wait, sgpmp_common.cuh).  The FIRST kernel of a call is launched normally: what
    // precedes it in the stream is the caller's (e.g. a torch kernel that wrote the obstacle spheres).  $SGPMP_PDL=0 disables it.
    static const char* pdl_env = getenv("SGPMP_PDL");
    const bool pdl = !(pdl_env && pdl_env[0] == '0');
    for (int it = 0; it < n_iters; ++it) {
        const bool last = (it == n_iters - 1);
        const void* eps = eps_in ? (const char*)eps_in + (size_t)it * BP * M * sh.S * w : nullptr;
        int rc = sample_launch(sh, tables, means, eps, seed, draw0 + (uint32_t)it, samples_ws, st, pdl && it > 0);
        if (rc != SGPMP_OK) return rc;
        rc = cost_st_launch(sh, *desc, tables, samples_ws, means, costs, st, pdl);
        if (rc != SGPMP_OK) return rc;
        rc = update_launch(sh, desc->temperature, step_size, costs, samples_ws, means, last ? grad : nullptr, last ? weights : nullptr,
                           row_chunks, st, last ? means_pre : nullptr, pdl);      // the update kernel also writes the pre-update means
        if (rc != SGPMP_OK) return rc;
    }
    return SGPMP_OK;
}
