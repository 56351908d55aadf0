// Counter-based normal stream of the sampler (stands in for torch's `normal_()` at
// torch/distributions/multivariate_normal.py:253; the reference's generator is not reproducible on a
// GPU, parity with it uses injected eps).  The stream definition is restated in oracle/philox.py.
//
//   Philox4x32-10, key = (seed_lo, seed_hi),
//   counter = ((tpair << 8) | dof, sample, global particle id, draw index)
//   4 words -> two Box-Muller pairs -> eps[2*tpair + {0,1}][{pos,vel}][dof]
#pragma once
#include "sgpmp_common.cuh"

namespace sgpmp {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// u1 = (w0 + 0.5) 2^-32, u2 = (w1 + 0.5) 2^-32; r = sqrt(-2 ln u1); th = pi (2 u2 - 1)
// fp32: MUFU forms — lg2.approx (|err| <= 2^-22 abs), sqrt.approx, sin/cos.approx on [-pi, pi) (|err| <= 2^-20.9
// abs): ~13 issue slots per pair instead of ~66 for logf + sqrtf + sincospif, with |d eps| <~ 1e-6 except in the
// vanishing-radius corner u1 -> 1 (the variate stays N(0,1) to that accuracy; parity uses injected eps).
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& z0, float& z1) {
    const float u1 = fmaf((float)w0, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float th = fmaf((float)w1, 1.4629180792671596e-09f, 7.314590396335798e-10f - 3.14159265358979f);
    float l2, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));      // -2 ln2 * log2(u1)
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    z0 = r * c;
    z1 = r * s;
}
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, double& z0, double& z1) {
    const double u1 = ((double)w0 + 0.5) * 2.3283064365386963e-10;
    const double v = ((double)w1 + 0.5) * 4.6566128730773926e-10 - 1.0;
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(v, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

struct RngKey {
    uint32_t k0, k1, draw;
};

// Four normals for (tpair, dof, sample, particle): (pos,vel) of t = 2*tpair and of t = 2*tpair + 1.
template <typename real>
__device__ __forceinline__ void normal4(const RngKey& key, uint32_t tpair, uint32_t dof, uint32_t sample,
                                        uint32_t particle_gid, real& p0, real& v0, real& p1, real& v1) {
    const uint4 w = philox4x32_10(make_uint4((tpair << 8) | dof, sample, particle_gid, key.draw), key.k0, key.k1);
    box_muller(w.x, w.y, p0, v0);
    box_muller(w.z, w.w, p1, v1);
}

}  // namespace sgpmp
