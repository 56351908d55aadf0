// Per-trajectory-sample cost accumulation shared by the standalone cost kernel (K3) and the fused
// iteration kernel.  One thread owns one trajectory sample and is fed its states x_t in time order.
//
// Reference being replaced (all under /root/reference/stoch_gpmp):
//   CostComposite.eval      costs/cost_functions.py:47-58  (FK hook, sum of children in list order)
//   CostGP.eval             costs/cost_functions.py:128-146 + GPFactor.get_error costs/factors/gp_factor.py:54-58
//   CostGoalPrior.eval      costs/cost_functions.py:376-388 (goal of particle p is p // K)
//   CostCollision.eval      costs/cost_functions.py:247-261 + FieldFactor costs/factors/field_factor.py:18-32
//                           (steps 1..T-1 only)
//   ObstacleMap.get_collisions   envs/obst_map.py:164-182
//   LinkDistanceField 'rbf'      costs/fields.py:63-79
//   FK hook                      costs/cost_functions.py:51-52 (torch_robotics chain, restated from the URDF)
//   IS term                      planner.py:233-236   tau * x^T Sigma^-1 mu  ==  tau * sum_t x_t . b_t, b = P mu
#pragma once
#include <math.h>

#include "sgpmp_common.cuh"

namespace sgpmp {

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp(k * d2) for the sphere RBF.  fp32: k is pre-multiplied by log2(e) and evaluated with ex2.approx
// (2 ulp; the RBF terms are in (0,1]); fp64: libdevice exp.
__device__ __forceinline__ float rbf_exp(float k_log2e, float d2) { return fast_exp2(k_log2e * d2); }
__device__ __forceinline__ double rbf_exp(double k, double d2) { return exp(k * d2); }

// Per-CTA constants staged in shared memory by the caller.
template <typename real>
struct CostSmem {
    const real* start;   // [d]
    const real* goal;    // [d] goal of this particle (or null)
    const real* bvec;    // [T][d]   b = Sigma^-1 mu of this particle
    const real* sph;     // [O][8]   (cx, cy, cz, k, ax, ay, az, b): k = -0.5/r^2 (* log2 e in fp32),
                         //          a = -2 k c, b = k |c|^2  so that  k |p - c|^2 = k |p|^2 + a.p + b
    const real* map;     // occupancy map of this problem (global memory)
    real coll_const;     // RBF sum of the link frames whose position does not depend on q (structured chains)
    real self_const;     // q-independent part of the self-collision sum (structured chains)
};

constexpr int SPH_STRIDE = 8;

// 4 consecutive reals from a 16-byte (fp32) / 32-byte (fp64) aligned shared-memory address in one LDS
__device__ __forceinline__ void load4(const float* p, float& a, float& b, float& c, float& d) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void load4(const double* p, double& a, double& b, double& c, double& d) {
    const double2 v = *reinterpret_cast<const double2*>(p), w = *reinterpret_cast<const double2*>(p + 2);
    a = v.x; b = v.y; c = w.x; d = w.y;
}

// Stage one sphere (cx, cy, cz, r) into the shared table row.
template <typename real>
__device__ __forceinline__ void stage_sphere(const real* s4, real* row) {
    const double cx = s4[0], cy = s4[1], cz = s4[2], r = s4[3];
    const double k = (sizeof(real) == 4 ? -0.5 * 1.4426950408889634 : -0.5) / (r * r);
    row[0] = (real)cx; row[1] = (real)cy; row[2] = (real)cz; row[3] = (real)k;
    row[4] = (real)(-2.0 * k * cx); row[5] = (real)(-2.0 * k * cy); row[6] = (real)(-2.0 * k * cz);
    row[7] = (real)(k * (cx * cx + cy * cy + cz * cz));
}

// Forward kinematics of a serial arm (frames 0..N-1: fixed transform then revolute-z joint f; frames
// N..n_frames-1 fixed), calling visit(x, y, z) for every link-frame origin in chain order
// (base first if include_base).  Chain constants come from the kernel-parameter constant bank.
template <typename real, int N, typename F>
__device__ __forceinline__ void fk_visit_links(const CostParams<real>& P, const real* q, F&& visit) {
    real R0 = 1, R1 = 0, R2 = 0, R3 = 0, R4 = 1, R5 = 0, R6 = 0, R7 = 0, R8 = 1;
    real px = 0, py = 0, pz = 0;
    if (P.include_base) visit(px, py, pz);
#pragma unroll
    for (int f = 0; f < N; ++f) {
        const real* F_ = P.R[f];
        const real* tf = P.p[f];
        px += R0 * tf[0] + R1 * tf[1] + R2 * tf[2];
        py += R3 * tf[0] + R4 * tf[1] + R5 * tf[2];
        pz += R6 * tf[0] + R7 * tf[1] + R8 * tf[2];
        // R <- R * F
        const real a0 = R0 * F_[0] + R1 * F_[3] + R2 * F_[6], a1 = R0 * F_[1] + R1 * F_[4] + R2 * F_[7], a2 = R0 * F_[2] + R1 * F_[5] + R2 * F_[8];
        const real a3 = R3 * F_[0] + R4 * F_[3] + R5 * F_[6], a4 = R3 * F_[1] + R4 * F_[4] + R5 * F_[7], a5 = R3 * F_[2] + R4 * F_[5] + R5 * F_[8];
        const real a6 = R6 * F_[0] + R7 * F_[3] + R8 * F_[6], a7 = R6 * F_[1] + R7 * F_[4] + R8 * F_[7], a8 = R6 * F_[2] + R7 * F_[5] + R8 * F_[8];
        // R <- R * Rz(q_f): col0' = c col0 + s col1, col1' = -s col0 + c col1
        real s, c;
        sg_sincos(q[f], &s, &c);
        R0 = c * a0 + s * a1; R1 = c * a1 - s * a0; R2 = a2;
        R3 = c * a3 + s * a4; R4 = c * a4 - s * a3; R5 = a5;
        R6 = c * a6 + s * a7; R7 = c * a7 - s * a6; R8 = a8;
        visit(px, py, pz);
    }
    // fixed tail frames (only positions are needed downstream)
    for (int f = N; f < P.n_frames; ++f) {
        const real* F_ = P.R[f];
        const real* tf = P.p[f];
        px += R0 * tf[0] + R1 * tf[1] + R2 * tf[2];
        py += R3 * tf[0] + R4 * tf[1] + R5 * tf[2];
        pz += R6 * tf[0] + R7 * tf[1] + R8 * tf[2];
        const real a0 = R0 * F_[0] + R1 * F_[3] + R2 * F_[6], a1 = R0 * F_[1] + R1 * F_[4] + R2 * F_[7], a2 = R0 * F_[2] + R1 * F_[5] + R2 * F_[8];
        const real a3 = R3 * F_[0] + R4 * F_[3] + R5 * F_[6], a4 = R3 * F_[1] + R4 * F_[4] + R5 * F_[7], a5 = R3 * F_[2] + R4 * F_[5] + R5 * F_[8];
        const real a6 = R6 * F_[0] + R7 * F_[3] + R8 * F_[6], a7 = R6 * F_[1] + R7 * F_[4] + R8 * F_[7], a8 = R6 * F_[2] + R7 * F_[5] + R8 * F_[8];
        R0 = a0; R1 = a1; R2 = a2; R3 = a3; R4 = a4; R5 = a5; R6 = a6; R7 = a7; R8 = a8;
        visit(px, py, pz);
    }
}

// ---- structured chains ------------------------------------------------------------------------------
// CHAIN = 0: generic serial arm (fk_visit_links, runtime constants).
// CHAIN = 1 (spheres only) / 2 (+ self-collision code): "Panda structure" (7 joints): fixed rotations are identity / Rx(+-90 deg) / about-z, and most
// translation components are zero (panda_arm_no_gripper.urdf).  The STRUCTURE is compile-time — Rx(+-90)
// becomes a signed relabelling of columns, zero translation components vanish — while the translation VALUES
// still come from the descriptor.  The host selects it only when the descriptor matches this structure
// (cos(1.57079632679) = 4.9e-12 is below fp32 resolution; fp64 always uses the generic path).
// Further structure used: frames with zero translation share their parent's origin (evaluated once, weight
// 2); origins that do not depend on q (base, link1, link2) are folded into CostSmem::coll_const; joint 7 only
// spins the z axis along which the remaining translations point, so q[6] is never needed.
constexpr int PANDA_EVAL_LINKS = 6;   // link3, link4, link5(=link6), link7, link8(=hand), ee

template <typename real>
struct Cols {   // rotation columns
    real ax, ay, az, bx, by, bz, cx, cy, cz;
    __device__ __forceinline__ void rot_xp90() {   // R <- R Rx(+90): (a, b, c) -> (a, c, -b)
        const real tx = bx, ty = by, tz = bz;
        bx = cx; by = cy; bz = cz;
        cx = -tx; cy = -ty; cz = -tz;
    }
    __device__ __forceinline__ void rot_xm90() {   // R <- R Rx(-90): (a, b, c) -> (a, -c, b)
        const real tx = bx, ty = by, tz = bz;
        bx = -cx; by = -cy; bz = -cz;
        cx = tx; cy = ty; cz = tz;
    }
    __device__ __forceinline__ void rot_z(real q) {  // R <- R Rz(q)
        real s, c;
        sg_sincos(q, &s, &c);
        const real nax = c * ax + s * bx, nay = c * ay + s * by, naz = c * az + s * bz;
        bx = c * bx - s * ax; by = c * by - s * ay; bz = c * bz - s * az;
        ax = nax; ay = nay; az = naz;
    }
};

// Origins of the 6 q-dependent, distinct link frames of the Panda structure; weights {1,1,2,1,2,1}.
template <typename real>
__device__ __forceinline__ void fk_panda_origins(const CostParams<real>& P, const real* q, real (&X)[PANDA_EVAL_LINKS],
                                                 real (&Y)[PANDA_EVAL_LINKS], real (&Z)[PANDA_EVAL_LINKS]) {
    Cols<real> R{1, 0, 0, 0, 1, 0, 0, 0, 1};
    real px = 0, py = 0, pz = P.p[0][2];           // frame 0: t = (0,0,z), rot = I
    R.rot_z(q[0]);
    R.rot_xm90();                                  // frame 1: t = 0, Rx(-90)
    R.rot_z(q[1]);
    px += P.p[2][1] * R.bx; py += P.p[2][1] * R.by; pz += P.p[2][1] * R.bz;   // frame 2: t = (0,y,0), Rx(+90)
    X[0] = px; Y[0] = py; Z[0] = pz;               // link3
    R.rot_xp90();
    R.rot_z(q[2]);
    px += P.p[3][0] * R.ax; py += P.p[3][0] * R.ay; pz += P.p[3][0] * R.az;   // frame 3: t = (x,0,0), Rx(+90)
    X[1] = px; Y[1] = py; Z[1] = pz;               // link4
    R.rot_xp90();
    R.rot_z(q[3]);
    px += P.p[4][0] * R.ax + P.p[4][1] * R.bx;                                 // frame 4: t = (x,y,0), Rx(-90)
    py += P.p[4][0] * R.ay + P.p[4][1] * R.by;
    pz += P.p[4][0] * R.az + P.p[4][1] * R.bz;
    X[2] = px; Y[2] = py; Z[2] = pz;               // link5 (= link6: frame 5 has t = 0)
    R.rot_xm90();
    R.rot_z(q[4]);
    R.rot_xp90();                                  // frame 5: t = 0, Rx(+90)
    R.rot_z(q[5]);
    px += P.p[6][0] * R.ax; py += P.p[6][0] * R.ay; pz += P.p[6][0] * R.az;   // frame 6: t = (x,0,0), Rx(+90)
    X[3] = px; Y[3] = py; Z[3] = pz;               // link7
    R.rot_xp90();                                  // joint 7 spins about c: c unchanged, q[6] not needed
    px += P.p[7][2] * R.cx; py += P.p[7][2] * R.cy; pz += P.p[7][2] * R.cz;   // frame 7: t = (0,0,z)
    X[4] = px; Y[4] = py; Z[4] = pz;               // link8 (= hand: frame 8 has t = 0, rotation about z)
    px += P.p[9][2] * R.cx; py += P.p[9][2] * R.cy; pz += P.p[9][2] * R.cz;   // frame 9: t = (0,0,z), about z
    X[5] = px; Y[5] = py; Z[5] = pz;               // ee_link
}

template <typename real, int N, int CHAIN = 0>
struct TrajCost {
    real c_start, c_gp, c_goal, c_coll, c_is, c_self;
    real xp[2 * N];   // previous state

    __device__ __forceinline__ void begin() { c_start = c_gp = c_goal = c_coll = c_is = c_self = 0; }

    // Link fields of configuration q (one FK evaluation shared by both):
    //   spheres  sum_l sum_o exp(-0.5 |p_l - c_o|^2 / r_o^2)            LinkDistanceField 'rbf'   costs/fields.py:63-79
    //   self     sum_{i,j} exp(-|p_i - p_j|^2 / (2 margin^2)), all ordered pairs incl. i == j
    //                                                                    LinkSelfDistanceField     costs/fields.py:114-124
    __device__ __forceinline__ void link_fields(const CostParams<real>& P, const CostSmem<real>& sm, const real* q) {
        const int O = P.n_spheres;
        if constexpr (CHAIN >= 1) {
            static_assert(N == 7, "Panda structure has 7 joints");
            real X[PANDA_EVAL_LINKS], Y[PANDA_EVAL_LINKS], Z[PANDA_EVAL_LINKS], PP[PANDA_EVAL_LINKS];
            fk_panda_origins<real>(P, q, X, Y, Z);
#pragma unroll
            for (int l = 0; l < PANDA_EVAL_LINKS; ++l) PP[l] = X[l] * X[l] + Y[l] * Y[l] + Z[l] * Z[l];
            if (P.has_spheres) {
                real acc = 0, acc2 = 0;   // acc2: links with weight 2
                for (int o = 0; o < O; ++o) {
                    const real* s = sm.sph + SPH_STRIDE * o;
                    const real k = s[3];
                    real ax, ay, az, b;
                    load4(s + 4, ax, ay, az, b);
#pragma unroll
                    for (int l = 0; l < PANDA_EVAL_LINKS; ++l) {
                        const real e = rbf_exp((real)1, X[l] * ax + (Y[l] * ay + (Z[l] * az + (k * PP[l] + b))));
                        if (l == 2 || l == 4) acc2 += e; else acc += e;
                    }
                }
                c_coll += acc + (acc2 + acc2);
            }
            if (CHAIN == 2 && P.has_self) {   // CHAIN 2 = Panda structure with the self-collision code compiled in
                // weights of the 6 evaluated origins: {1,1,2,1,2,1}; constants: base (w = include_base) and
                // link1 = link2 at (0,0,z0) (w = 2).  Ordered pairs => every unordered pair counts twice.
                const real ks = P.self_k, z0 = P.p[0][2];
                real a1 = 0, a2 = 0, a4 = 0;     // sums of E over unordered pairs with weight product 1, 2, 4
#pragma unroll
                for (int l = 0; l < PANDA_EVAL_LINKS; ++l) {
                    const bool w2 = (l == 2 || l == 4);
                    const real e12 = rbf_exp(ks, PP[l] + z0 * (z0 - (real)2 * Z[l]));      // vs link1/link2 (w = 2)
                    if (w2) a4 += e12; else a2 += e12;
                    if (P.include_base) {
                        const real eb = rbf_exp(ks, PP[l]);                                // vs base (w = 1)
                        if (w2) a2 += eb; else a1 += eb;
                    }
#pragma unroll
                    for (int m = l + 1; m < PANDA_EVAL_LINKS; ++m) {
                        const real dx = X[l] - X[m], dy = Y[l] - Y[m], dz = Z[l] - Z[m];
                        const real e = rbf_exp(ks, dx * dx + dy * dy + dz * dz);
                        const int wp = (w2 ? 2 : 1) * ((m == 2 || m == 4) ? 2 : 1);
                        if (wp == 4) a4 += e; else if (wp == 2) a2 += e; else a1 += e;
                    }
                }
                c_self += (real)2 * (a1 + (real)2 * a2 + (real)4 * a4);
            }
        } else {
            if (P.has_self) {
                // generic chain: keep every origin, then the upper triangle of the pair matrix
                real PX[SGPMP_MAX_FRAMES + 1], PY[SGPMP_MAX_FRAMES + 1], PZ[SGPMP_MAX_FRAMES + 1];
                int L = 0;
                real acc = 0;
                fk_visit_links<real, N>(P, q, [&](real x, real y, real z) {
                    PX[L] = x; PY[L] = y; PZ[L] = z; ++L;
                    if (P.has_spheres)
                        for (int o = 0; o < O; ++o) {
                            const real* s = sm.sph + SPH_STRIDE * o;
                            const real dx = x - s[0], dy = y - s[1], dz = z - s[2];
                            acc += rbf_exp(s[3], dx * dx + dy * dy + dz * dz);
                        }
                });
                c_coll += acc;
                real sa = 0;
                for (int l = 0; l < L; ++l)
                    for (int m = l + 1; m < L; ++m) {
                        const real dx = PX[l] - PX[m], dy = PY[l] - PY[m], dz = PZ[l] - PZ[m];
                        sa += rbf_exp(P.self_k, dx * dx + dy * dy + dz * dz);
                    }
                c_self += (real)2 * sa + (real)L;
            } else {
                real acc = 0;
                fk_visit_links<real, N>(P, q, [&](real x, real y, real z) {
                    for (int o = 0; o < O; ++o) {
                        const real* s = sm.sph + SPH_STRIDE * o;
                        const real dx = x - s[0], dy = y - s[1], dz = z - s[2];
                        acc += rbf_exp(s[3], dx * dx + dy * dy + dz * dz);
                    }
                });
                c_coll += acc;
            }
        }
    }

    __device__ __forceinline__ real map_value(const CostParams<real>& P, const CostSmem<real>& sm, real x, real y) const {
        // X*(1/cell) + offset with two roundings, floor, int, clamp; value = map[iy][ix].  The reference
        // clamps ix with shape[0] and iy with shape[1] (obst_map.py:177-178); maps are square here.
        const real xo = sg_mul_add_2r(x, P.map_inv_cell, P.map_origin_x);
        const real yo = sg_mul_add_2r(y, P.map_inv_cell, P.map_origin_y);
        int ix = (int)sg_floor(xo), iy = (int)sg_floor(yo);
        ix = min(max(ix, 0), P.map_h - 1);
        iy = min(max(iy, 0), P.map_w - 1);
        return __ldg(sm.map + (size_t)iy * P.map_w + ix);
    }

    // feed state x_t (t = 0..T-1 in order)
    // brow: b_t = (Sigma^-1 mu)_t of this particle (2N reals, 16-byte aligned when 2N % 4 == 0 rows are padded), or null
    __device__ __forceinline__ void step(const CostParams<real>& P, const CostSmem<real>& sm, int t, int T,
                                         const real (&x)[2 * N], const real* brow) {
        if (t == 0) {
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) {
                const real e = sm.start[j] - x[j];
                c_start += e * e;
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const real ep = x[i] - xp[i] - P.dt * xp[N + i];
                const real ev = x[N + i] - xp[N + i];
                c_gp += P.q11 * ep * ep + P.q12x2 * ep * ev + P.q22 * ev * ev;
            }
            if (P.has_map) c_coll += map_value(P, sm, x[0], x[1]);
            if (P.has_spheres || P.has_self) link_fields(P, sm, x);
        }
        if (t == T - 1 && P.has_goal) {
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) {
                const real e = sm.goal[j] - x[j];
                c_goal += e * e;
            }
        }
        if (brow) {
            constexpr int DP4 = (2 * N + 3) / 4;
            real b[4 * DP4];
#pragma unroll
            for (int k = 0; k < DP4; ++k) load4(brow + 4 * k, b[4 * k], b[4 * k + 1], b[4 * k + 2], b[4 * k + 3]);
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) c_is += x[j] * b[j];
        }
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) xp[j] = x[j];
    }

    __device__ __forceinline__ void finish(const CostParams<real>& P, const CostSmem<real>& sm, int T) {
        c_start *= P.inv_sig_start2;
        c_goal *= P.inv_sig_goal2;
        if (CHAIN >= 1) {
            c_coll += (real)(T - 1) * sm.coll_const;
            c_self += (real)(T - 1) * sm.self_const;
        }
        c_coll *= (P.has_map ? P.map_w_coll : P.sphere_w_coll);
        c_self *= P.self_w_coll;
        c_is *= P.temperature;
    }
    // summation order of the shipped examples' cost lists: CostGP (start + gp), CostGoalPrior, self-collision,
    // obstacle collision (examples/panda_environment.py:90), then += IS (planner.py:236)
    __device__ __forceinline__ real total() const { return ((((c_start + c_gp) + c_goal) + c_self) + c_coll) + c_is; }
};

// Per-CTA staging of the problem constants into shared memory: start [d], goal [d] of goal index g, the sphere
// table [MAX_SPHERES][8] followed by one slot for coll_const.  Ends with a __syncthreads().
constexpr int SPH_SMEM = SPH_STRIDE * SGPMP_MAX_SPHERES + 4;   // keeps what follows 16-byte aligned

template <typename real, int N, int CHAIN>
__device__ __forceinline__ void stage_cta_constants(const CostParams<real>& P, int b, int g, int G, real* start, real* goal,
                                                    real* sph) {
    constexpr int d = 2 * N;
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        start[k] = P.start[(size_t)b * d + k];
        goal[k] = P.has_goal ? P.goals[((size_t)b * G + g) * d + k] : (real)0;
    }
    if (P.has_spheres)
        for (int k = threadIdx.x; k < P.n_spheres; k += blockDim.x)
            stage_sphere<real>(P.spheres + ((size_t)(P.spheres_per_problem ? b : 0) * P.n_spheres + k) * 4, sph + SPH_STRIDE * k);
    __syncthreads();
    if (threadIdx.x == 0) {
        real cc = 0;
        if (CHAIN >= 1 && P.has_spheres) {
            // q-independent origins of the Panda structure: base (if counted), link1 and link2 at (0, 0, z0)
            const real z0 = P.p[0][2];
            for (int o = 0; o < P.n_spheres; ++o) {
                const real* s = sph + SPH_STRIDE * o;
                if (P.include_base) cc += rbf_exp((real)1, s[7]);
                cc += (real)2 * rbf_exp((real)1, z0 * s[6] + (s[3] * z0 * z0 + s[7]));
            }
        }
        sph[SPH_STRIDE * SGPMP_MAX_SPHERES] = cc;
        real sc = 0;
        if (CHAIN >= 1 && P.has_self) {
            // constant-constant pairs + the diagonal of the 6 evaluated origins (weights 1,1,2,1,2,1 -> 12)
            const real z0 = P.p[0][2], wb = P.include_base ? (real)1 : (real)0;
            sc = wb * wb + (real)4 + (real)4 * wb * rbf_exp(P.self_k, z0 * z0) + (real)12;
        }
        sph[SPH_STRIDE * SGPMP_MAX_SPHERES + 1] = sc;
    }
    __syncthreads();
}

// 1 if the descriptor's chain has the Panda STRUCTURE (see fk_panda_origins); host-side check.
inline int chain_is_panda_structure(const sgpmp_cost_desc_t& d, int n_dof) {
    if (n_dof != 7 || d.n_frames != 10) return 0;
    static const int rot[10] = {0, -1, 1, 1, -1, 1, 1, 0, 2, 2};          // 0: I, +-1: Rx(+-90), 2: about z
    static const int tmask[10] = {4, 0, 2, 1, 3, 0, 1, 4, 0, 4};          // bit0 x, bit1 y, bit2 z may be non-zero
    const double tol = 1e-9;
    for (int f = 0; f < 10; ++f) {
        const double* R = d.chain_R[f];
        for (int k = 0; k < 3; ++k)
            if (!((tmask[f] >> k) & 1) && d.chain_p[f][k] != 0.0) return 0;
        double want[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (rot[f] == 1) { double w[9] = {1, 0, 0, 0, 0, -1, 0, 1, 0}; memcpy(want, w, sizeof(w)); }
        if (rot[f] == -1) { double w[9] = {1, 0, 0, 0, 0, 1, 0, -1, 0}; memcpy(want, w, sizeof(w)); }
        if (rot[f] == 2) {   // rotation about z: third row/column are e_z
            if (fabs(R[2]) > tol || fabs(R[5]) > tol || fabs(R[6]) > tol || fabs(R[7]) > tol || fabs(R[8] - 1) > tol) return 0;
            continue;
        }
        for (int k = 0; k < 9; ++k)
            if (fabs(R[k] - want[k]) > tol) return 0;
    }
    return 1;
}

// b = P mu for one particle, computed in fp64 from the D/O blocks (the fp32 reference evaluates this
// contraction with catastrophic cancellation), stored as `real`.
// tabDO: [T][7] doubles (d11,d12,d22,o11,o12,o21,o22), O_t = P[t+1,t].
// MU_STRIDE: row stride of mu in reals (0 = dense rows of 2n).
template <typename real, int MU_STRIDE = 0>
__device__ __forceinline__ void precision_times_row(const double* tabDO, const real* mu, int T, int n, int t, int i,
                                                    real* bp, real* bv) {
    const int d = MU_STRIDE ? MU_STRIDE : 2 * n;
    const double* r = tabDO + t * 7;
    const double mp = mu[t * d + i], mv = mu[t * d + n + i];
    double p = r[0] * mp + r[1] * mv;
    double v = r[1] * mp + r[2] * mv;
    if (t > 0) {
        const double* q = tabDO + (t - 1) * 7;
        const double ap = mu[(t - 1) * d + i], av = mu[(t - 1) * d + n + i];
        p += q[3] * ap + q[4] * av;
        v += q[5] * ap + q[6] * av;
    }
    if (t < T - 1) {
        const double ap = mu[(t + 1) * d + i], av = mu[(t + 1) * d + n + i];
        p += r[3] * ap + r[5] * av;   // O_t^T
        v += r[4] * ap + r[6] * av;
    }
    *bp = (real)p;
    *bv = (real)v;
}

}  // namespace sgpmp
