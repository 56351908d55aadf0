// Shared declarations of the sm_100a kernels behind include/stoch_gpmp_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stoch_gpmp_b200.h"

namespace sgpmp {

// ---- status / error text (thread-local, returned by sgpmp_last_error) -------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SGPMP_REQUIRE(cond, ...)                     \
    do {                                             \
        if (!(cond)) {                               \
            ::sgpmp::set_error(__VA_ARGS__);         \
            return SGPMP_ERR_INVALID_ARG;            \
        }                                            \
    } while (0)

// Dynamic shared memory above which a launcher opts in with cudaFuncAttributeMaxDynamicSharedMemorySize.  The 48 KB default
// limit counts a kernel's STATIC shared memory too, so the threshold leaves room for it (a launch with 47.9 KB dynamic + 256 B
// static fails with "invalid argument" otherwise); opting in when it was not strictly needed costs nothing.
#define SGPMP_SMEM_OPTIN (40 * 1024)

#define SGPMP_CHECK_LAUNCH(name)                                                          \
    do {                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                              \
        if (e_ != cudaSuccess) {                                                          \
            ::sgpmp::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));    \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)         \
                       ? SGPMP_ERR_NO_DEVICE                                              \
                       : SGPMP_ERR_CUDA;                                                  \
        }                                                                                 \
        ::sgpmp::count_launch();                                                          \
    } while (0)

// ---- device-side cost parameters (passed by value as a kernel argument: constant bank, uniform) -----
template <typename real>
struct CostParams {
    real dt;
    real inv_sig_start2;   // 1/sigma_start^2
    real q11, q12x2, q22;  // GP weights: q11 ep^2 + q12x2 ep ev + q22 ev^2
    real inv_sig_goal2;    // 0 if absent
    real temperature;
    const real* start;     // [B,d]
    const real* goals;     // [B,G,d] or null
    // map
    const real* occ_map;
    const uint8_t* occ_map_u8;   // optional byte copy of occ_map (gathered instead of it when non-null)
    const int32_t* map_of_problem;
    int32_t map_h, map_w, origin_xi, origin_yi, n_maps;
    real map_inv_cell, map_origin_x, map_origin_y, map_w_coll;  // w_coll = 1/sigma_coll^2
    // spheres
    const real* spheres;
    int32_t n_spheres, spheres_per_problem;
    int32_t sphere_mode;   // SGPMP_FIELD_*
    real sphere_w_coll;
    // FK chain
    int32_t n_frames, include_base;
    real R[SGPMP_MAX_FRAMES][9];
    real p[SGPMP_MAX_FRAMES][3];
    int32_t joint[SGPMP_MAX_FRAMES];
    int32_t has_goal, has_map, has_spheres, has_self;
    real self_k, self_w_coll;   // self_k = -0.5/margin^2 (* log2 e in fp32)
    // link interpolation (generic chain path only)
    int32_t sphere_interp_n, sphere_interp_lo, sphere_interp_hi;
    int32_t self_interp_n, self_interp_lo, self_interp_hi;
    real sphere_alpha[SGPMP_MAX_INTERP], self_alpha[SGPMP_MAX_INTERP];
    // EE SE(3) goal
    int32_t has_ee, ee_square;
    real ee_R[9], ee_p[3], ee_w_pos, ee_w_rot, ee_w;   // ee_w = 1/sigma_goal^2
};

template <typename real>
int lower_cost_desc(const sgpmp_shape_t& sh, const sgpmp_cost_desc_t& d, CostParams<real>& out);

// Programmatic dependent launch (PDL) for the chain of short kernels of the low-latency iteration (sgpmp_lowlat.cu): a kernel
// launched with the attribute may START while its predecessor in the stream is still running; everything it does before
// pdl_wait() must therefore touch only data that no kernel of the chain writes (prior tables, cost constants, its own shared
// memory, the counter-based RNG), and every kernel of the chain executes pdl_wait() — which returns when the predecessor grid has
// completed and its writes are visible — before its first dependent access, so completion is transitive along the chain.
// Both instructions are no-ops in a kernel that was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// cross-file launchers used by the low-latency iteration (sgpmp_lowlat.cu)
int sample_launch(const sgpmp_shape_t& sh, const double* tables, const void* means, const void* eps_in, uint64_t seed,
                  uint32_t draw, void* samples, cudaStream_t st, bool pdl = false);
int cost_st_launch(const sgpmp_shape_t& sh, const sgpmp_cost_desc_t& desc, const double* tables, const void* samples,
                   const void* means, void* costs, cudaStream_t st, bool pdl = false);
int update_launch(const sgpmp_shape_t& sh, double tau, double step, const void* costs, const void* samples, void* means,
                  void* grad, void* weights, int row_chunks, cudaStream_t st, void* means_pre = nullptr, bool pdl = false);

int iterate_stats_launch(const sgpmp_shape_t& sh, const sgpmp_cost_desc_t& desc, const double* tables, uint64_t seed,
                         uint32_t draw, const void* means, void* costs, void* stats, cudaStream_t st);
int merge_apply_stats_launch(const sgpmp_shape_t& sh, const double* tables, double step, const void* stats_all, int n_ranks,
                             void* means, void* grad, cudaStream_t st);

inline bool shape_ok(const sgpmp_shape_t* s) {
    return s && s->B > 0 && s->G > 0 && s->K > 0 && s->S > 0 && s->T >= 2 && s->n_dof > 0 && s->n_dof <= 255 &&
           (s->dtype == SGPMP_F32 || s->dtype == SGPMP_F64);
}

// precision-matched math wrappers ---------------------------------------------------------------------
__device__ __forceinline__ float sg_exp(float x) { return expf(x); }
__device__ __forceinline__ double sg_exp(double x) { return exp(x); }
__device__ __forceinline__ float sg_floor(float x) { return floorf(x); }
__device__ __forceinline__ double sg_floor(double x) { return floor(x); }
__device__ __forceinline__ float sg_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double sg_max(double a, double b) { return fmax(a, b); }
// multiply then add with two roundings (no FMA contraction): the reference's map index arithmetic
__device__ __forceinline__ float sg_mul_add_2r(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
__device__ __forceinline__ double sg_mul_add_2r(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }

template <typename real>
__device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename real>
__device__ __forceinline__ real warp_max(real v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = sg_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace sgpmp
