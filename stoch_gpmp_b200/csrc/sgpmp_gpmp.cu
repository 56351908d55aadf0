// Gauss-Newton GPMP step (the reference's second planner, stoch_gpmp/planner.py:352-661) on the block-tridiagonal
// structure of its normal equations.  SURVEY §8(f) rank 4; reuses the cost descriptor and the FK chain of the StochGPMP path.
//
// Reference, per iteration (planner.py:575-600):
//   A, b, K = cost.get_linear_system(means)          dense [NP, rows, M] / [NP, rows, rows]   cost_functions.py:60-85
//   J^T J = A^T K A + damping, g = A^T K b           dense [NP, M, M] GEMMs                   planner.py:602-617
//   d_theta = solve(J^T J, g)                        dense LU / Cholesky per particle         planner.py:619-633
//   means += step_size d_theta                                                                planner.py:592
// Here: A^T K A is never formed.  Its block (t, t) is the per-DoF 2x2 block of the GP/start/goal precision (the same
// closed form as the sampling prior, with the COST sigmas) plus one rank-1 term  w h h^T  per link field on the
// position rows (h = -d field / d q_t, computed analytically by the Jacobian-transpose rule instead of autograd), its
// block (t+1, t) is -Q Phi.  Two kernels per iteration:
//   gpmp_assemble_kernel   CTA per particle, thread per time step: residuals, g, field values + gradients, b^T K b
//   gpmp_solve_kernel      warp per particle: block Cholesky (d x d blocks, fp64, shared memory), forward / backward
//                          substitution, means += step d_theta
// All arithmetic is fp64 whatever the storage dtype (cond(J^T J) ~ 1e9: the reference's fp32 runs solve in fp32 and
// are only matched loosely).
//
// solver_params (planner.py:375,581-586):
//   delta, trust_region   J^T J = A^T K A + delta I,  or  + delta diag(mean over the particles of one problem of A^T K A)
//   method 'inverse'      d = (J^T J)^-1 g
//   method 'cholesky'     AS WRITTEN in the reference (planner.py:626-629): l = chol(J^T J); z = l^-1 g; the second
//                         triangular solve is called with l^T and upper=False, so it only reads the diagonal:
//                         d = diag(l)^-1 z.  Reproduced as is (oracle/gpmp.py::solve, pinned by tests/golden/gpmp_*).
#include <math.h>

#include <algorithm>

#include "sgpmp_common.cuh"
#include "sgpmp_cost.cuh"

namespace sgpmp {

enum { GPMP_METHOD_INVERSE = 0, GPMP_METHOD_CHOLESKY = 1 };

struct GpmpArgs {
    int NP, K, G, T;
    const double* D;       // [T][3]   per-DoF diagonal blocks of the GP/start/goal normal equations (d11, d12, d22)
    const double* O;       // [T-1][4] per-DoF sub-diagonal blocks P[t+1,t] (o11, o12, o21, o22)
    double* gvec;          // [BP][T][d]     g = A^T K b
    double* hvec;          // [BP][T][3][N]  sqrt(w) h of the self / sphere / EE-goal rows (zero where absent)
    double* diagv;         // [BP][T][d]     diagonal of A^T K A (trust region)
    double* Lws;           // [BP][T][d][d]  Cholesky factors of the pivot blocks   (method inverse)
    double* Wws;           // [BP][T][d][d]  sub-diagonal factor blocks             (method inverse)
    double* zws;           // [BP][T][d]     forward-substituted right-hand side    (method inverse)
    int32_t* not_pd;       // [BP] 0, or 1 + index of the failing pivot block
    double delta, step;
    int trust_region, method;
    const void* means_in;  // [BP][T][d] real
    void* means_out;       // [BP][T][d] real (may alias means_in)
    void* d_theta;         // optional [BP][T][d] real
    void* costs;           // [BP] real: b^T K b
};

// ---- link fields with gradients ------------------------------------------------------------------------------------
// Value and gradient wrt q of
//   spheres  sum_p sum_o exp(-0.5 |x_p - c_o|^2 / r_o^2)        LinkDistanceField 'rbf'    costs/fields.py:63-79
//   self     sum_{p,p'} exp(-|x_p - x_p'|^2 / (2 margin^2))      LinkSelfDistanceField      costs/fields.py:114-124
// over the link-frame origins plus the interpolated points of each field (fields.py:68-74).  The gradient is
// sum_p F_p . d x_p / d q_j with F_p = d c / d x_p; an interpolated point X_i + (X_{i+1} - X_i) a hands (1 - a) F to
// link i and a F to link i+1, and for link frames  d X_l / d q_j = z_j x (X_l - o_j)  (l at or below joint j), so
//   d c / d q_j = z_j . sum_{l below j} (X_l - o_j) x F_l     — one backward sweep over the chain.
template <int N>
__device__ void link_fields_grad(const CostParams<double>& P, const double* sph, const double* q, bool want_ee,
                                 double& c_sph, double* g_sph, double& c_self, double* g_self, double& c_ee, double* g_ee) {
    constexpr int MAXL = SGPMP_MAX_FRAMES + 1;
    double X[MAXL][3], Zax[N][3];
    const int off = P.include_base ? 1 : 0;
    const int L = P.n_frames + off;
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};          // ends as the rotation of the LAST frame (the end-effector)
    {
        double p[3] = {0, 0, 0};
        if (off) { X[0][0] = X[0][1] = X[0][2] = 0; }
        for (int f = 0; f < P.n_frames; ++f) {
            const double* F_ = P.R[f];
            const double* tf = P.p[f];
            double Rn[9];
            for (int r = 0; r < 3; ++r) {
                p[r] = fma(R[3 * r], tf[0], fma(R[3 * r + 1], tf[1], fma(R[3 * r + 2], tf[2], p[r])));
                for (int c = 0; c < 3; ++c)
                    Rn[3 * r + c] = fma(R[3 * r], F_[c], fma(R[3 * r + 1], F_[3 + c], R[3 * r + 2] * F_[6 + c]));
            }
            if (f < N) {
                double sn, cs;
                sincos(q[f], &sn, &cs);
                for (int r = 0; r < 3; ++r) {
                    const double a = Rn[3 * r], b = Rn[3 * r + 1];
                    Rn[3 * r] = fma(cs, a, sn * b);
                    Rn[3 * r + 1] = fma(cs, b, -sn * a);
                    Zax[f][r] = Rn[3 * r + 2];                 // joint axis = z column of the joint's frame
                }
            }
            for (int k = 0; k < 9; ++k) R[k] = Rn[k];
            X[f + off][0] = p[0]; X[f + off][1] = p[1]; X[f + off][2] = p[2];
        }
    }
    // point m of a field with (n_interp, lo, alpha): m < L a link origin, else interpolated between links i and i+1
    auto point = [&](int m, int ni, int lo, const double* alpha, double* x, int& i0, double& a) {
        if (m < L) { x[0] = X[m][0]; x[1] = X[m][1]; x[2] = X[m][2]; i0 = m; a = 0.0; return; }
        const int k = m - L;
        i0 = lo + k / ni;
        a = alpha[k % ni];
        for (int c = 0; c < 3; ++c) x[c] = X[i0][c] + (X[i0 + 1][c] - X[i0][c]) * a;
    };
    auto sweep = [&](const double (*F)[3], double* g) {     // Jacobian-transpose: torque of the link forces about every joint axis
        double Fa[3] = {0, 0, 0}, Ma[3] = {0, 0, 0};
        for (int l = L - 1; l >= off; --l) {
            Fa[0] += F[l][0]; Fa[1] += F[l][1]; Fa[2] += F[l][2];
            Ma[0] += X[l][1] * F[l][2] - X[l][2] * F[l][1];
            Ma[1] += X[l][2] * F[l][0] - X[l][0] * F[l][2];
            Ma[2] += X[l][0] * F[l][1] - X[l][1] * F[l][0];
            const int j = l - off;
            if (j < N) {
                const double* o = X[l];
                const double mx = Ma[0] - (o[1] * Fa[2] - o[2] * Fa[1]);
                const double my = Ma[1] - (o[2] * Fa[0] - o[0] * Fa[2]);
                const double mz = Ma[2] - (o[0] * Fa[1] - o[1] * Fa[0]);
                g[j] = Zax[j][0] * mx + Zax[j][1] * my + Zax[j][2] * mz;
            }
        }
    };
    double F[MAXL][3];
    c_sph = 0;
    c_self = 0;
    c_ee = 0;
    for (int j = 0; j < N; ++j) { g_sph[j] = 0; g_self[j] = 0; g_ee[j] = 0; }
    if (want_ee && P.has_ee) {
        // EE SE(3) goal (CostGoal + EESE3DistanceField, cost_functions.py:282-337, fields.py:130-153; SE3_distance as in
        // oracle/se3.py):  dist = w_pos |p - p*| + w_rot acos(c),  c = (tr(R^T R*) - 1)/2 clamped to [-1, 1].
        //   d|p - p*|/dq_j = u . (z_j x (p - o_j)),  u = (p - p*)/|p - p*|
        //   d tr(R^T R*)/dq_j = z_j . v,  v = vee(M - M^T),  M = R* R^T      (dR/dq_j = [z_j]x R)
        //   d acos(c)/dq_j = -(1/sqrt(1 - c^2)) (1/2) z_j . v                 (zero where the clamp is active, as autograd)
        const double* pe = X[L - 1];
        const double dx = pe[0] - P.ee_p[0], dy = pe[1] - P.ee_p[1], dz = pe[2] - P.ee_p[2];
        const double dpos = sqrt(dx * dx + dy * dy + dz * dz);
        double tr = 0;
        for (int k = 0; k < 9; ++k) tr += R[k] * P.ee_R[k];
        double c = (tr - 1.0) * 0.5;
        const bool clamped = !(c > -1.0 && c < 1.0);
        c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
        const double dist = P.ee_w_pos * dpos + P.ee_w_rot * acos(c);
        // M = R* R^T ; v = (M21 - M12, M02 - M20, M10 - M01)
        double M[9];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                M[3 * r + cc] = P.ee_R[3 * r] * R[3 * cc] + P.ee_R[3 * r + 1] * R[3 * cc + 1] + P.ee_R[3 * r + 2] * R[3 * cc + 2];
        const double v[3] = {M[7] - M[5], M[2] - M[6], M[3] - M[1]};
        const double wr = clamped ? 0.0 : -P.ee_w_rot * 0.5 / sqrt(1.0 - c * c);
        const double wp = dpos > 0 ? P.ee_w_pos / dpos : 0.0;
        const double outer = P.ee_square ? 2.0 * dist : 1.0;
        c_ee = P.ee_square ? dist * dist : dist;
        for (int j = 0; j < N; ++j) {
            const double* z = Zax[j];
            const double* o = X[j + off];
            const double rx = pe[0] - o[0], ry = pe[1] - o[1], rz = pe[2] - o[2];
            const double cx = z[1] * rz - z[2] * ry, cy = z[2] * rx - z[0] * rz, cz = z[0] * ry - z[1] * rx;   // z x r
            g_ee[j] = outer * (wp * (dx * cx + dy * cy + dz * cz) + wr * (z[0] * v[0] + z[1] * v[1] + z[2] * v[2]));
        }
    }
    if (P.has_spheres) {
        for (int l = 0; l < L; ++l) F[l][0] = F[l][1] = F[l][2] = 0;
        const int ni = P.sphere_interp_n, Lp = L + ni * (P.sphere_interp_hi - P.sphere_interp_lo);
        for (int m = 0; m < Lp; ++m) {
            double x[3], a, f[3] = {0, 0, 0};
            int i0;
            point(m, ni, P.sphere_interp_lo, P.sphere_alpha, x, i0, a);
            for (int o = 0; o < P.n_spheres; ++o) {
                const double* s = sph + 4 * o;
                const double dx = x[0] - s[0], dy = x[1] - s[1], dz = x[2] - s[2], ir2 = 1.0 / (s[3] * s[3]);
                const double E = exp(-0.5 * (dx * dx + dy * dy + dz * dz) * ir2);
                c_sph += E;
                f[0] -= E * dx * ir2; f[1] -= E * dy * ir2; f[2] -= E * dz * ir2;
            }
            for (int c = 0; c < 3; ++c) {
                F[i0][c] += (1.0 - a) * f[c];
                if (m >= L) F[i0 + 1][c] += a * f[c];
            }
        }
        sweep(F, g_sph);
    }
    if (P.has_self) {
        for (int l = 0; l < L; ++l) F[l][0] = F[l][1] = F[l][2] = 0;
        const int ni = P.self_interp_n, Lp = L + ni * (P.self_interp_hi - P.self_interp_lo);
        const double k2 = P.self_k;                                    // -0.5 / margin^2 (fp64 lowering)
        for (int m = 0; m < Lp; ++m) {
            double xa[3], aa;
            int ia;
            point(m, ni, P.self_interp_lo, P.self_alpha, xa, ia, aa);
            for (int m2 = m + 1; m2 < Lp; ++m2) {
                double xb[3], ab;
                int ib;
                point(m2, ni, P.self_interp_lo, P.self_alpha, xb, ib, ab);
                const double dx = xa[0] - xb[0], dy = xa[1] - xb[1], dz = xa[2] - xb[2];
                const double E = exp(k2 * (dx * dx + dy * dy + dz * dz));
                c_self += 2.0 * E;                                     // ordered pairs (a,b) and (b,a)
                const double w = 2.0 * E * 2.0 * k2;                   // d/dx_a of 2 exp(k2 |x_a - x_b|^2) = w (x_a - x_b)
                const double f[3] = {w * dx, w * dy, w * dz};
                for (int c = 0; c < 3; ++c) {
                    F[ia][c] += (1.0 - aa) * f[c];
                    if (m >= L) F[ia + 1][c] += aa * f[c];
                    F[ib][c] -= (1.0 - ab) * f[c];
                    if (m2 >= L) F[ib + 1][c] -= ab * f[c];
                }
            }
        }
        c_self += (double)Lp;                                          // the diagonal pairs
        sweep(F, g_self);
    }
}

template <typename real, int N>
__global__ void __launch_bounds__(128)
gpmp_assemble_kernel(const __grid_constant__ CostParams<double> P, const __grid_constant__ GpmpArgs A) {
    constexpr int d = 2 * N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* x = reinterpret_cast<double*>(smem_raw);       // [T][d]
    double* sph = x + (size_t)A.T * d;                     // [MAX_SPHERES][4]
    double* start = sph + 4 * SGPMP_MAX_SPHERES;           // [d]
    double* goal = start + d;                              // [d]
    __shared__ double red[32];
    const int T = A.T;
    const int bp = blockIdx.x, b = bp / A.NP, p = bp - b * A.NP, gidx = p / A.K;
    const real* means = reinterpret_cast<const real*>(A.means_in) + (size_t)bp * T * d;
    for (int k = threadIdx.x; k < T * d; k += blockDim.x) x[k] = (double)means[k];
    const real* st = reinterpret_cast<const real*>(P.start);
    const real* gl = reinterpret_cast<const real*>(P.goals);
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        start[k] = (double)st[(size_t)b * d + k];
        goal[k] = P.has_goal ? (double)gl[((size_t)b * A.G + gidx) * d + k] : 0.0;
    }
    if (P.has_spheres) {
        const real* sp = reinterpret_cast<const real*>(P.spheres) + (size_t)(P.spheres_per_problem ? b : 0) * P.n_spheres * 4;
        for (int k = threadIdx.x; k < 4 * P.n_spheres; k += blockDim.x) sph[k] = (double)sp[k];
    }
    __syncthreads();
    const double q11 = P.q11, q12 = 0.5 * P.q12x2, q22 = P.q22, dt = P.dt;
    double cost = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double* g = A.gvec + ((size_t)bp * T + t) * d;
        double* hv = A.hvec + ((size_t)bp * T + t) * 3 * N;
        double* dg = A.diagv + ((size_t)bp * T + t) * d;
        const double* xt = x + t * d;
        double gp[N], gv[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { gp[i] = 0; gv[i] = 0; }
        if (t < T - 1) {          // GP factor t: rows H1 = Phi on x_t:  Phi^T Q e_t
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double ep = xt[d + i] - xt[i] - dt * xt[N + i], ev = xt[d + N + i] - xt[N + i];
                const double Qp = q11 * ep + q12 * ev, Qv = q12 * ep + q22 * ev;
                gp[i] += Qp;
                gv[i] += dt * Qp + Qv;
                cost += ep * Qp + ev * Qv;
            }
        }
        if (t > 0) {              // GP factor t-1: rows H2 = -I on x_t:  -Q e_{t-1}
            const double* xm = xt - d;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double ep = xt[i] - xm[i] - dt * xm[N + i], ev = xt[N + i] - xm[N + i];
                gp[i] -= q11 * ep + q12 * ev;
                gv[i] -= q12 * ep + q22 * ev;
            }
        }
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double e0 = start[i] - xt[i], e1 = start[N + i] - xt[N + i];
                gp[i] += P.inv_sig_start2 * e0;
                gv[i] += P.inv_sig_start2 * e1;
                cost += P.inv_sig_start2 * (e0 * e0 + e1 * e1);
            }
        }
        if (t == T - 1 && P.has_goal) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double e0 = goal[i] - xt[i], e1 = goal[N + i] - xt[N + i];
                gp[i] += P.inv_sig_goal2 * e0;
                gv[i] += P.inv_sig_goal2 * e1;
                cost += P.inv_sig_goal2 * (e0 * e0 + e1 * e1);
            }
        }
        double hs[N], hc[N], he[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { hs[i] = 0; hc[i] = 0; he[i] = 0; }
        if (t > 0 && (P.has_spheres || P.has_self || (P.has_ee && t == T - 1))) {
            // field rows exist for steps 1..T-1 (cost_functions.py:241-245), the EE-goal row for T-1 only (:300-304)
            double c_sph, c_self, c_ee, g_sph[N], g_self[N], g_ee[N];
            link_fields_grad<N>(P, sph, xt, t == T - 1, c_sph, g_sph, c_self, g_self, c_ee, g_ee);
            if (P.has_ee && t == T - 1) {
                const double w = P.ee_w;
#pragma unroll
                for (int i = 0; i < N; ++i) { gp[i] += w * (-g_ee[i]) * c_ee; he[i] = sqrt(w) * (-g_ee[i]); }
                cost += w * c_ee * c_ee;
            }
            if (P.has_spheres) {
                const double w = P.sphere_w_coll;
#pragma unroll
                for (int i = 0; i < N; ++i) { gp[i] += w * (-g_sph[i]) * c_sph; hc[i] = sqrt(w) * (-g_sph[i]); }   // H = -d err / d q
                cost += w * c_sph * c_sph;
            }
            if (P.has_self) {
                const double w = P.self_w_coll;
#pragma unroll
                for (int i = 0; i < N; ++i) { gp[i] += w * (-g_self[i]) * c_self; hs[i] = sqrt(w) * (-g_self[i]); }
                cost += w * c_self * c_self;
            }
        }
        const double* Dt = A.D + 3 * t;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            g[i] = gp[i]; g[N + i] = gv[i];
            hv[i] = hs[i]; hv[N + i] = hc[i]; hv[2 * N + i] = he[i];
            dg[i] = Dt[0] + hs[i] * hs[i] + hc[i] * hc[i] + he[i] * he[i];
            dg[N + i] = Dt[2];
        }
    }
    // block sum of the cost
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cost;
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0;
        for (int k = 0; k < (int)(blockDim.x + 31) / 32; ++k) c += red[k];
        reinterpret_cast<real*>(A.costs)[bp] = (real)c;
    }
}

// One warp per particle — or per TWO particles when d <= 16 (SUB = 2: lanes 0-15 and 16-31 each own one particle; with the
// Panda's d = 14 that fills 28 of 32 lanes instead of 14).  Lane i of a half owns row i of the current d x d block.
template <typename real, int N, int WPB>
__global__ void __launch_bounds__(32 * WPB)
gpmp_solve_kernel(const __grid_constant__ GpmpArgs A, int n_particles) {
    constexpr int d = 2 * N, LD = d + 1;
    static_assert(d <= 32, "one lane per row of a pivot block");
    constexpr int SUB = (d <= 16) ? 2 : 1;
    constexpr int HALF = 32 / SUB;
    __shared__ double Ssm[WPB * SUB][d * LD], Wsm[WPB * SUB][d * LD], zsm[WPB * SUB][d], xsm[WPB * SUB][d];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / HALF, src0 = sub * HALF;            // shuffles read lane (src0 + j) of the own half
    if ((blockIdx.x * WPB + warp) * SUB >= n_particles) return;
    const int slot = warp * SUB + sub;
    const int bp_raw = (blockIdx.x * WPB + warp) * SUB + sub;
    const bool live = bp_raw < n_particles;                    // an odd particle count leaves the last upper half idle
    const int bp = live ? bp_raw : n_particles - 1;
    double* S = Ssm[slot];
    double* W = Wsm[slot];
    double* z = zsm[slot];
    double* xn = xsm[slot];
    const int T = A.T, b = bp / A.NP;
    const int i = lane - src0, ai = i / N, ii = i - ai * N;    // row i = (pos|vel, dof)
    const bool row = live && i < d;
    const real* means = reinterpret_cast<const real*>(A.means_in) + (size_t)bp * T * d;
    real* mout = reinterpret_cast<real*>(A.means_out) + (size_t)bp * T * d;
    real* dth = A.d_theta ? reinterpret_cast<real*>(A.d_theta) + (size_t)bp * T * d : nullptr;
    int bad = 0;
    for (int t = 0; t < T; ++t) {
        const double* Dt = A.D + 3 * t;
        const double* hv = A.hvec + ((size_t)bp * T + t) * 3 * N;
        // ---- pivot block S = D_t + sum_f h_f h_f^T + damping - W_{t-1} W_{t-1}^T -----------------------------------
        if (row) {
            double damp = A.delta;
            if (A.trust_region) {
                double m = 0;
                for (int p2 = 0; p2 < A.NP; ++p2) m += A.diagv[(((size_t)b * A.NP + p2) * T + t) * d + i];
                damp = A.delta * (m / (double)A.NP);
            }
            for (int k = 0; k < d; ++k) {
                const int ak = k / N, kk = k - ak * N;
                double v = 0;
                if (ii == kk) v = (ai == 0 && ak == 0) ? Dt[0] : ((ai == 1 && ak == 1) ? Dt[2] : Dt[1]);
                if (ai == 0 && ak == 0) v += hv[ii] * hv[kk] + hv[N + ii] * hv[N + kk] + hv[2 * N + ii] * hv[2 * N + kk];
                if (k == i) v += damp;
                if (t > 0) {
                    double acc = 0;
                    for (int m = 0; m < d; ++m) acc += W[i * LD + m] * W[k * LD + m];
                    v -= acc;
                }
                S[i * LD + k] = v;
            }
        }
        __syncwarp();
        // rhs_t = g_t - W_{t-1} z_{t-1}
        double rhs = 0;
        if (row) {
            rhs = A.gvec[((size_t)bp * T + t) * d + i];
            if (t > 0)
                for (int m = 0; m < d; ++m) rhs -= W[i * LD + m] * z[m];
        }
        __syncwarp();
        // ---- Cholesky S = L L^T in place (lower), column by column ---------------------------------------------------
        for (int j = 0; j < d; ++j) {
            double v = 0;
            if (row && i >= j) {
                v = S[i * LD + j];
                for (int k = 0; k < j; ++k) v -= S[i * LD + k] * S[j * LD + k];
            }
            const double piv = __shfl_sync(0xffffffffu, v, src0 + j);
            if (live && !(piv > 0.0) && !bad) bad = 1 + t;
            const double lj = sqrt(piv);
            if (row && i >= j) S[i * LD + j] = (i == j) ? lj : v / lj;
            __syncwarp();
        }
        // ---- forward substitution z_t = L^-1 rhs ---------------------------------------------------------------------
        for (int j = 0; j < d; ++j) {
            const double zj = __shfl_sync(0xffffffffu, rhs, src0 + j) / S[j * LD + j];
            if (row && i > j) rhs -= S[i * LD + j] * zj;
            if (i == j) rhs = zj;
        }
        if (row) z[i] = rhs;
        if (A.method == GPMP_METHOD_CHOLESKY) {
            // the reference's 'cholesky' branch as written: d_theta = diag(l)^-1 l^-1 g  (see the header)
            if (row) {
                const double dd = rhs / S[i * LD + i];
                mout[t * d + i] = (real)((double)means[t * d + i] + A.step * dd);
                if (dth) dth[t * d + i] = (real)dd;
            }
        } else if (row) {
            double* Lg = A.Lws + ((size_t)bp * T + t) * d * d;
            for (int k = 0; k < d; ++k) Lg[i * d + k] = (k <= i) ? S[i * LD + k] : 0.0;
            A.zws[((size_t)bp * T + t) * d + i] = rhs;
        }
        __syncwarp();
        // ---- W_t = O_t L^-T:  lane c solves L w = (row c of O_t)^T;  W[c][j] = w_j ---------------------------------
        if (t < T - 1) {
            const double* Ot = A.O + 4 * t;
            if (row) {
                for (int j = 0; j < d; ++j) {
                    const int aj = j / N, jj = j - aj * N;
                    double v = (jj == ii) ? Ot[2 * ai + aj] : 0.0;          // O_t[i][j]
                    for (int k = 0; k < j; ++k) v -= S[j * LD + k] * W[i * LD + k];
                    W[i * LD + j] = v / S[j * LD + j];
                }
                if (A.method == GPMP_METHOD_INVERSE) {
                    double* Wg = A.Wws + ((size_t)bp * T + t) * d * d;
                    for (int k = 0; k < d; ++k) Wg[i * d + k] = W[i * LD + k];
                }
            }
        }
        __syncwarp();
    }
    if (A.method == GPMP_METHOD_INVERSE) {
        // ---- backward substitution: L_t^T x_t = z_t - W_t^T x_{t+1} ----------------------------------------------------
        for (int t = T - 1; t >= 0; --t) {
            const double* Lg = A.Lws + ((size_t)bp * T + t) * d * d;
            double rhs = 0;
            if (row) {
                for (int k = 0; k < d; ++k) S[i * LD + k] = Lg[i * d + k];
                rhs = A.zws[((size_t)bp * T + t) * d + i];
                if (t < T - 1) {
                    const double* Wg = A.Wws + ((size_t)bp * T + t) * d * d;
                    for (int c = 0; c < d; ++c) rhs -= Wg[c * d + i] * xn[c];
                }
            }
            __syncwarp();
            for (int j = d - 1; j >= 0; --j) {
                const double xj = __shfl_sync(0xffffffffu, rhs, src0 + j) / S[j * LD + j];
                if (row && i < j) rhs -= S[j * LD + i] * xj;
                if (i == j) rhs = xj;
            }
            __syncwarp();
            if (row) {
                xn[i] = rhs;
                mout[t * d + i] = (real)((double)means[t * d + i] + A.step * rhs);
                if (dth) dth[t * d + i] = (real)rhs;
            }
            __syncwarp();
        }
    }
    if (live && i == 0) A.not_pd[bp] = bad;
}

template <typename real, int N>
static int launch_gpmp_n(const sgpmp_shape_t& sh, const CostParams<double>& P, GpmpArgs& A, int n_iters, cudaStream_t st) {
    constexpr int d = 2 * N;
    if constexpr (d > 32) {
        set_error("sgpmp_gpmp_step: n_dof=%d needs %d-row pivot blocks (> 32 lanes)", N, d);
        return SGPMP_ERR_UNSUPPORTED;
    } else {
        const int BP = sh.B * sh.G * sh.K;
        const size_t smem = ((size_t)sh.T * d + 4 * SGPMP_MAX_SPHERES + 2 * d) * sizeof(double);
        if (smem > 227 * 1024) { set_error("sgpmp_gpmp_step: T=%d too large for shared memory", sh.T); return SGPMP_ERR_UNSUPPORTED; }
        if (smem > SGPMP_SMEM_OPTIN) cudaFuncSetAttribute(gpmp_assemble_kernel<real, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        constexpr int WPB = (d <= 16) ? 4 : 1;       // static shared memory: WPB * (2 d (d+1) + 2 d) doubles
        for (int it = 0; it < n_iters; ++it) {
            // one thread per time step: no more threads than steps (T = 64: 64-thread CTAs, twice the resident CTAs per SM)
            const int abs_ = std::min(128, ((sh.T + 31) / 32) * 32);
            gpmp_assemble_kernel<real, N><<<BP, abs_, smem, st>>>(P, A);
            SGPMP_CHECK_LAUNCH("sgpmp_gpmp_step(assemble)");
            constexpr int PPC = WPB * ((d <= 16) ? 2 : 1);       // particles per CTA
            gpmp_solve_kernel<real, N, WPB><<<(BP + PPC - 1) / PPC, 32 * WPB, 0, st>>>(A, BP);
            SGPMP_CHECK_LAUNCH("sgpmp_gpmp_step(solve)");
            A.means_in = A.means_out;
        }
        return SGPMP_OK;
    }
}

static size_t gpmp_ws_doubles(const sgpmp_shape_t& sh, int method, size_t* offs /*[6]*/) {
    const size_t BP = (size_t)sh.B * sh.G * sh.K, T = sh.T, d = 2 * (size_t)sh.n_dof;
    size_t o = 0;
    offs[0] = o; o += BP * T * d;              // gvec
    offs[1] = o; o += BP * T * (3 * d / 2);    // hvec (3 N)
    offs[2] = o; o += BP * T * d;              // diagv
    offs[3] = o; if (method == GPMP_METHOD_INVERSE) o += BP * T * d * d;   // Lws
    offs[4] = o; if (method == GPMP_METHOD_INVERSE) o += BP * T * d * d;   // Wws
    offs[5] = o; if (method == GPMP_METHOD_INVERSE) o += BP * T * d;       // zws
    return o;
}

template <typename real>
static int launch_gpmp(const sgpmp_shape_t& sh, const sgpmp_cost_desc_t& desc, const double* D, const double* O, double delta,
                       int trust_region, int method, double step, int n_iters, void* means, void* d_theta, void* costs,
                       int32_t* not_pd, double* ws, cudaStream_t st) {
    CostParams<double> P;
    int rc = lower_cost_desc<double>(sh, desc, P);
    if (rc != SGPMP_OK) return rc;
    if (P.has_map) {
        set_error("sgpmp_gpmp_step: the occupancy-map field has no gradient (a floor() lookup, obst_map.py:164-182); the "
                  "reference's GPMP cannot differentiate it either (field_factor.py:35)");
        return SGPMP_ERR_UNSUPPORTED;
    }
    if (P.has_spheres && P.sphere_mode != SGPMP_FIELD_RBF) {
        set_error("sgpmp_gpmp_step: only the 'rbf' sphere field is differentiated");
        return SGPMP_ERR_UNSUPPORTED;
    }
    size_t offs[6];
    gpmp_ws_doubles(sh, method, offs);
    GpmpArgs A;
    A.NP = sh.G * sh.K; A.K = sh.K; A.G = sh.G; A.T = sh.T;
    A.D = D; A.O = O;
    A.gvec = ws + offs[0]; A.hvec = ws + offs[1]; A.diagv = ws + offs[2];
    A.Lws = ws + offs[3]; A.Wws = ws + offs[4]; A.zws = ws + offs[5];
    A.not_pd = not_pd;
    A.delta = delta; A.step = step; A.trust_region = trust_region; A.method = method;
    A.means_in = means; A.means_out = means; A.d_theta = d_theta; A.costs = costs;
    switch (sh.n_dof) {
#define SGPMP_DOF_CASE(N) case N: return launch_gpmp_n<real, N>(sh, P, A, n_iters, st);
#include "sgpmp_dof_list.inc"
#undef SGPMP_DOF_CASE
        default:
            set_error("sgpmp_gpmp_step: n_dof=%d is not instantiated (see sgpmp_dof_list.inc)", sh.n_dof);
            return SGPMP_ERR_UNSUPPORTED;
    }
}

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int64_t sgpmp_gpmp_workspace_bytes(const sgpmp_shape_t* shape, int32_t method) {
    if (!shape_ok(shape)) return -1;
    size_t offs[6];
    return (int64_t)(gpmp_ws_doubles(*shape, method, offs) * sizeof(double));
}

extern "C" int sgpmp_gpmp_step(const sgpmp_shape_t* shape, const sgpmp_cost_desc_t* desc, const double* D, const double* O,
                               double delta, int32_t trust_region, int32_t method, double step_size, int32_t n_iters,
                               void* means, void* d_theta, void* costs, void* workspace, int64_t workspace_bytes,
                               int32_t* not_pd, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_gpmp_step: invalid shape");
    SGPMP_REQUIRE(desc && D && O && means && costs && workspace && not_pd, "sgpmp_gpmp_step: null pointer");
    SGPMP_REQUIRE(n_iters >= 1, "sgpmp_gpmp_step: n_iters must be >= 1");
    SGPMP_REQUIRE(method == GPMP_METHOD_INVERSE || method == GPMP_METHOD_CHOLESKY, "sgpmp_gpmp_step: unknown method %d", method);
    SGPMP_REQUIRE(delta >= 0, "sgpmp_gpmp_step: delta must be >= 0");
    SGPMP_REQUIRE(workspace_bytes >= sgpmp_gpmp_workspace_bytes(shape, method), "sgpmp_gpmp_step: workspace too small");
    if (shape->dtype == SGPMP_F32)
        return launch_gpmp<float>(*shape, *desc, D, O, delta, trust_region, method, step_size, n_iters, means, d_theta, costs,
                                  not_pd, (double*)workspace, (cudaStream_t)stream);
    return launch_gpmp<double>(*shape, *desc, D, O, delta, trust_region, method, step_size, n_iters, means, d_theta, costs,
                               not_pd, (double*)workspace, (cudaStream_t)stream);
}
