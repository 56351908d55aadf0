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
    const real* sph;     // [O][4]   (cx, cy, cz, kexp) kexp = -0.5/r^2 (* log2 e in fp32)
    const real* map;     // occupancy map of this problem (global memory)
};

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

template <typename real, int N>
struct TrajCost {
    real c_start, c_gp, c_goal, c_coll, c_is;
    real xp[2 * N];   // previous state

    __device__ __forceinline__ void begin() { c_start = c_gp = c_goal = c_coll = c_is = 0; }

    // sum_l sum_o exp(-0.5 |p_l - c_o|^2 / r_o^2) over the link frames of configuration q
    __device__ __forceinline__ real link_sphere_rbf(const CostParams<real>& P, const CostSmem<real>& sm,
                                                    const real* q) const {
        real acc = 0;
        const int O = P.n_spheres;
        fk_visit_links<real, N>(P, q, [&](real x, real y, real z) {
            for (int o = 0; o < O; ++o) {
                const real dx = x - sm.sph[4 * o + 0], dy = y - sm.sph[4 * o + 1], dz = z - sm.sph[4 * o + 2];
                acc += rbf_exp(sm.sph[4 * o + 3], dx * dx + dy * dy + dz * dz);
            }
        });
        return acc;
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
    __device__ __forceinline__ void step(const CostParams<real>& P, const CostSmem<real>& sm, int t, int T,
                                         const real (&x)[2 * N]) {
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
            if (P.has_spheres) c_coll += link_sphere_rbf(P, sm, x);
        }
        if (t == T - 1 && P.has_goal) {
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) {
                const real e = sm.goal[j] - x[j];
                c_goal += e * e;
            }
        }
        if (sm.bvec) {
            const real* b = sm.bvec + (size_t)t * 2 * N;
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) c_is += x[j] * b[j];
        }
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) xp[j] = x[j];
    }

    __device__ __forceinline__ void finish(const CostParams<real>& P) {
        c_start *= P.inv_sig_start2;
        c_goal *= P.inv_sig_goal2;
        c_coll *= (P.has_map ? P.map_w_coll : P.sphere_w_coll);
        c_is *= P.temperature;
    }
    // reference summation order: CostGP (start + gp), CostGoalPrior, CostCollision, then += IS
    __device__ __forceinline__ real total() const { return (((c_start + c_gp) + c_goal) + c_coll) + c_is; }
};

// b = P mu for one particle, computed in fp64 from the D/O blocks (the fp32 reference evaluates this
// contraction with catastrophic cancellation), stored as `real`.
// tabDO: [T][7] doubles (d11,d12,d22,o11,o12,o21,o22), O_t = P[t+1,t].
template <typename real>
__device__ __forceinline__ void precision_times_row(const double* tabDO, const real* mu, int T, int n, int t, int i,
                                                    real* bp, real* bv) {
    const int d = 2 * n;
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
