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
    const int32_t* map_of_problem;
    int32_t map_h, map_w, origin_xi, origin_yi, n_maps;
    real map_inv_cell, map_origin_x, map_origin_y, map_w_coll;  // w_coll = 1/sigma_coll^2
    // spheres
    const real* spheres;
    int32_t n_spheres, spheres_per_problem;
    real sphere_w_coll;
    // FK chain
    int32_t n_frames, include_base;
    real R[SGPMP_MAX_FRAMES][9];
    real p[SGPMP_MAX_FRAMES][3];
    int32_t joint[SGPMP_MAX_FRAMES];
    int32_t has_goal, has_map, has_spheres, has_self;
    real self_k, self_w_coll;   // self_k = -0.5/margin^2 (* log2 e in fp32)
};

template <typename real>
int lower_cost_desc(const sgpmp_shape_t& sh, const sgpmp_cost_desc_t& d, CostParams<real>& out);

inline bool shape_ok(const sgpmp_shape_t* s) {
    return s && s->B > 0 && s->G > 0 && s->K > 0 && s->S > 0 && s->T >= 2 && s->n_dof > 0 && s->n_dof <= 255 &&
           (s->dtype == SGPMP_F32 || s->dtype == SGPMP_F64);
}

// precision-matched math wrappers ---------------------------------------------------------------------
__device__ __forceinline__ float sg_exp(float x) { return expf(x); }
__device__ __forceinline__ double sg_exp(double x) { return exp(x); }
// fp32 sincos for joint angles: branch-free, FMA pipe only (no MUFU, no slow path).  2-term Cody-Waite
// reduction by pi/2 (exact enough for |x| < 1e3 rad) + the classic cephes single-precision minimax
// polynomials on [-pi/4, pi/4]; max abs error 7e-8 on |x| <= 12 (emulated against fp64; test_fk_known_answers_cuda checks FK to 2e-6).
// libdevice sincosf costs ~28 issue slots incl. a divergent slow-path guard; this is ~19.
__device__ __forceinline__ void sg_sincos(float x, float* sp, float* cp) {
    const float t = fmaf(x, 0.636619772367581f, 12582912.0f);   // 1.5 * 2^23: rint(x * 2/pi) lands in the low mantissa bits
    const int q = __float_as_int(t);
    const float j = t - 12582912.0f;
    float r = fmaf(j, -1.5707963705062866f, x);        // pi/2 = C1 + C2 (+ 1.7e-15)
    r = fmaf(j, 4.371138828673793e-08f, r);
    const float z = r * r;
    const float s = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, r, r);
    const float c = fmaf(fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z, -0.5f), z, 1.0f);
    const float ss = (q & 1) ? c : s;
    const float cc = (q & 1) ? s : c;
    *sp = __int_as_float(__float_as_int(ss) ^ ((q << 30) & 0x80000000));
    *cp = __int_as_float(__float_as_int(cc) ^ (((q + 1) << 30) & 0x80000000));
}
__device__ __forceinline__ void sg_sincos(double x, double* s, double* c) { sincos(x, s, c); }
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
