// K1 — block-tridiagonal Cholesky of the constant-velocity GP precision (fp64), and the dense
// scale_tril expansion.
//
// Reference being replaced: MultiMPPrior.update_dist (mp_priors_multi.py:100-110) ->
// torch MultivariateNormal(precision_matrix=P) -> _precision_to_scale_tril
// (multivariate_normal.py:79-85): Lf = chol(flip P); L = (flip(Lf)^T)^-1, executed on NP dense copies of
// the M x M precision and re-executed on every set_mean().  Here P is kept as its per-DoF 2x2 blocks
// (the precision of mp_priors_multi.py:170-202 is (2x2) (x) I_n block-tridiagonal), factored once,
// O(T) work, by the reverse recursion  P = U U^T,  U[t,t] = A_t (upper), U[t,t+1] = C_t.
//
// The arithmetic mirrors oracle/prior.py::banded_factor operation for operation (explicit _rn
// intrinsics: no FMA contraction), because with cond(P) ~ 1e7..1e9 any re-association moves the
// factor by cond * 2^-53.
#include "sgpmp_common.cuh"

namespace sgpmp {

__global__ void prior_factor_kernel(int n_priors, int T, const double* __restrict__ Dg,
                                    const double* __restrict__ Og, double* __restrict__ tables,
                                    int32_t* __restrict__ not_pd) {
    const int pr = blockIdx.x * blockDim.x + threadIdx.x;
    if (pr >= n_priors) return;
    const double* D = Dg + (size_t)pr * T * 3;
    const double* O = Og + (size_t)pr * (T - 1) * 4;
    double* tab = tables + (size_t)pr * T * SGPMP_TABLE_STRIDE;
    int32_t bad = 0;
    double g11 = 0.0, g21 = 0.0, g22 = 0.0;      // G_{t+1}
    double c11 = 0.0, c12 = 0.0, c21 = 0.0, c22 = 0.0;
    // The blocks of step t - 1 are loaded while step t is factored: the recursion is one dependent chain of ~40 fp64 operations
    // (two square roots, three divisions) per step, and a just-in-time load would put an L2 round trip in front of every link of it
    // (T = 1024: 0.69 -> see profiles/r2/c5_sweep.jsonl).
    double nd11 = D[(T - 1) * 3 + 0], nd12 = D[(T - 1) * 3 + 1], nd22 = D[(T - 1) * 3 + 2];
    double no11 = 0.0, no12 = 0.0, no21 = 0.0, no22 = 0.0;
    for (int t = T - 1; t >= 0; --t) {
        const double d11 = nd11, d12 = nd12, d22 = nd22;
        double s11 = d11, s12 = d12, s22 = d22;
        double o11 = 0.0, o12 = 0.0, o21 = 0.0, o22 = 0.0;
        const bool inner = t < T - 1;
        if (inner) { o11 = no11; o12 = no12; o21 = no21; o22 = no22; }
        if (t > 0) {
            nd11 = D[(t - 1) * 3 + 0]; nd12 = D[(t - 1) * 3 + 1]; nd22 = D[(t - 1) * 3 + 2];
            no11 = O[(t - 1) * 4 + 0]; no12 = O[(t - 1) * 4 + 1]; no21 = O[(t - 1) * 4 + 2]; no22 = O[(t - 1) * 4 + 3];
        }
        if (inner) {
            // C_t = O_t^T G_{t+1}
            c11 = __dadd_rn(__dmul_rn(o11, g11), __dmul_rn(o21, g21));
            c12 = __dmul_rn(o21, g22);
            c21 = __dadd_rn(__dmul_rn(o12, g11), __dmul_rn(o22, g21));
            c22 = __dmul_rn(o22, g22);
            s11 = __dsub_rn(d11, __dadd_rn(__dmul_rn(c11, c11), __dmul_rn(c12, c12)));
            s12 = __dsub_rn(d12, __dadd_rn(__dmul_rn(c11, c21), __dmul_rn(c12, c22)));
            s22 = __dsub_rn(d22, __dadd_rn(__dmul_rn(c21, c21), __dmul_rn(c22, c22)));
            // H_{t+1} = G_{t+1} C_t^T   (uses G_{t+1} still held in g**)
            double* nx = tab + (size_t)(t + 1) * SGPMP_TABLE_STRIDE;
            nx[SGPMP_TAB_H11] = __dmul_rn(g11, c11);
            nx[SGPMP_TAB_H12] = __dmul_rn(g11, c21);
            nx[SGPMP_TAB_H21] = __dadd_rn(__dmul_rn(g21, c11), __dmul_rn(g22, c12));
            nx[SGPMP_TAB_H22] = __dadd_rn(__dmul_rn(g21, c21), __dmul_rn(g22, c22));
        }
        // S = A A^T with A upper-triangular
        if (!(s22 > 0.0) && !bad) bad = 1 + t;
        const double a22 = __dsqrt_rn(s22);
        const double a12 = __ddiv_rn(s12, a22);
        const double r = __dsub_rn(s11, __dmul_rn(a12, a12));
        if (!(r > 0.0) && !bad) bad = 1 + t;
        const double a11 = __dsqrt_rn(r);
        g11 = __ddiv_rn(1.0, a11);
        g22 = __ddiv_rn(1.0, a22);
        g21 = __dmul_rn(-__dmul_rn(a12, g11), g22);
        double* row = tab + (size_t)t * SGPMP_TABLE_STRIDE;
        row[SGPMP_TAB_G11] = g11; row[SGPMP_TAB_G21] = g21; row[SGPMP_TAB_G22] = g22;
        row[SGPMP_TAB_D11] = d11; row[SGPMP_TAB_D12] = d12; row[SGPMP_TAB_D22] = d22;
        row[SGPMP_TAB_O11] = o11; row[SGPMP_TAB_O12] = o12; row[SGPMP_TAB_O21] = o21; row[SGPMP_TAB_O22] = o22;
        row[14] = 0.0; row[15] = 0.0;
        if (t == 0) {
            row[SGPMP_TAB_H11] = 0.0; row[SGPMP_TAB_H12] = 0.0; row[SGPMP_TAB_H21] = 0.0; row[SGPMP_TAB_H22] = 0.0;
        }
    }
    not_pd[pr] = bad;
}

// Column j of the per-DoF 2T x 2T factor = recurrence applied to e_j; scattered to the n DoF copies.
template <typename real>
__global__ void dense_L_kernel(int T, int n, const double* __restrict__ tab, real* __restrict__ L) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // per-DoF column index (tj, aj)
    if (j >= 2 * T) return;
    const int tj = j >> 1, aj = j & 1;
    const int d = 2 * n;
    const size_t M = (size_t)T * d;
    double yp = 0.0, yv = 0.0;
    for (int t = tj; t < T; ++t) {
        const double* r = tab + (size_t)t * SGPMP_TABLE_STRIDE;
        const double ep = (t == tj && aj == 0) ? 1.0 : 0.0;
        const double ev = (t == tj && aj == 1) ? 1.0 : 0.0;
        const double np_ = r[SGPMP_TAB_G11] * ep - (r[SGPMP_TAB_H11] * yp + r[SGPMP_TAB_H12] * yv);
        const double nv_ = r[SGPMP_TAB_G21] * ep + r[SGPMP_TAB_G22] * ev - (r[SGPMP_TAB_H21] * yp + r[SGPMP_TAB_H22] * yv);
        yp = np_; yv = nv_;
        for (int i = 0; i < n; ++i) {
            const size_t col = (size_t)tj * d + (size_t)aj * n + i;
            L[((size_t)t * d + i) * M + col] = (real)yp;
            L[((size_t)t * d + n + i) * M + col] = (real)yv;
        }
    }
}

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int sgpmp_prior_factor(int32_t n_priors, int32_t T, const double* D, const double* O,
                                  double* tables, int32_t* not_pd, void* stream) {
    SGPMP_REQUIRE(n_priors > 0 && T >= 2, "sgpmp_prior_factor: need n_priors > 0 and T >= 2 (got %d, %d)", n_priors, T);
    SGPMP_REQUIRE(D && O && tables && not_pd, "sgpmp_prior_factor: null pointer");
    const int bs = 32;
    prior_factor_kernel<<<(n_priors + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(n_priors, T, D, O, tables, not_pd);
    SGPMP_CHECK_LAUNCH("sgpmp_prior_factor");
    return SGPMP_OK;
}

extern "C" int sgpmp_prior_dense_L(int32_t T, int32_t n_dof, const double* tables, int32_t dtype, void* L,
                                   void* stream) {
    SGPMP_REQUIRE(T >= 2 && n_dof > 0 && tables && L, "sgpmp_prior_dense_L: bad arguments");
    SGPMP_REQUIRE(dtype == SGPMP_F32 || dtype == SGPMP_F64, "sgpmp_prior_dense_L: bad dtype %d", dtype);
    const size_t M = (size_t)T * 2 * n_dof;
    const size_t bytes = M * M * (dtype == SGPMP_F32 ? 4 : 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(L, 0, bytes, st) != cudaSuccess) {
        set_error("sgpmp_prior_dense_L: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
        return SGPMP_ERR_CUDA;
    }
    const int bs = 64;
    if (dtype == SGPMP_F32)
        dense_L_kernel<float><<<(2 * T + bs - 1) / bs, bs, 0, st>>>(T, n_dof, tables, (float*)L);
    else
        dense_L_kernel<double><<<(2 * T + bs - 1) / bs, bs, 0, st>>>(T, n_dof, tables, (double*)L);
    SGPMP_CHECK_LAUNCH("sgpmp_prior_dense_L");
    return SGPMP_OK;
}
