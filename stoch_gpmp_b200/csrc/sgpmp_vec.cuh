// Value types of the per-sample arithmetic: scalar `float` / `double`, or `F2` = two fp32 samples packed in one
// 64-bit register pair and processed with sm_100's packed-FP32 instructions (PTX add/sub/mul/fma .f32x2 ->
// SASS FADD2 / FMUL2 / FFMA2).
//
// Why: the fused kernel is bound by the single issue port of each SM sub-partition (1 warp instruction per
// clock), not by the FP32 lanes: FFMA2 performs two FMAs per lane per ISSUED instruction (it occupies the FMA
// pipe for two passes, so peak FLOP/s is unchanged — measured 70.6 TFLOP/s for both forms) and frees every
// second FP issue slot for the integer (Philox), MUFU and LDS instructions of the same warp
// (profiles/r1/ffma2_probe.txt: 128 FFMA + 64 LOP3 = 4.25 ms, 64 FFMA2 + 64 LOP3 = 3.37 ms).
// One thread therefore owns TWO trajectory samples (lanes .x / .y).  Scalars broadcast for free: ptxas turns
// pack(s, s) into the `R.F32` broadcast operand form, and lane extraction is register-pair aliasing.
#pragma once
#include "sgpmp_common.cuh"

namespace sgpmp {

struct F2 {
    unsigned long long v;
};

__device__ __forceinline__ F2 f2(float a, float b) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lane0(F2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
__device__ __forceinline__ float lane1(F2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }

__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator+(F2 a, float s) { return a + f2(s, s); }
__device__ __forceinline__ F2 operator+(float s, F2 a) { return f2(s, s) + a; }
__device__ __forceinline__ F2 operator-(F2 a, float s) { return a - f2(s, s); }
__device__ __forceinline__ F2 operator-(float s, F2 a) { return f2(s, s) - a; }
__device__ __forceinline__ F2 operator*(F2 a, float s) { return a * f2(s, s); }
__device__ __forceinline__ F2 operator*(float s, F2 a) { return f2(s, s) * a; }
__device__ __forceinline__ F2& operator+=(F2& a, F2 b) { a = a + b; return a; }

// ---- uniform interface over scalar and packed values -------------------------------------------------
template <typename V> struct VT;
template <> struct VT<float>  { using real = float;  static constexpr int W = 1; };
template <> struct VT<double> { using real = double; static constexpr int W = 1; };
template <> struct VT<F2>     { using real = float;  static constexpr int W = 2; };

// a * b + c in one rounding
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double vfma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ F2 vfma(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 vfma(float a, F2 b, F2 c) { return vfma(f2(a, a), b, c); }
__device__ __forceinline__ F2 vfma(F2 a, float b, F2 c) { return vfma(a, f2(b, b), c); }
__device__ __forceinline__ F2 vfma(F2 a, F2 b, float c) { return vfma(a, b, f2(c, c)); }
__device__ __forceinline__ F2 vfma(float a, F2 b, float c) { return vfma(f2(a, a), b, f2(c, c)); }
__device__ __forceinline__ F2 vfma(F2 a, float b, float c) { return vfma(a, f2(b, b), f2(c, c)); }

template <typename V> __device__ __forceinline__ V vbroadcast(typename VT<V>::real s);
template <> __device__ __forceinline__ float vbroadcast<float>(float s) { return s; }
template <> __device__ __forceinline__ double vbroadcast<double>(double s) { return s; }
template <> __device__ __forceinline__ F2 vbroadcast<F2>(float s) { return f2(s, s); }

__device__ __forceinline__ float vlane(float a, int) { return a; }
__device__ __forceinline__ double vlane(double a, int) { return a; }
__device__ __forceinline__ float vlane(F2 a, int k) { return k ? lane1(a) : lane0(a); }

// 2^x (fp32: ex2.approx) / e^x (fp64) — the RBF kernels pre-scale their exponent accordingly
__device__ __forceinline__ float vexp2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double vexp2_fast(double x) { return exp(x); }
__device__ __forceinline__ F2 vexp2_fast(F2 x) { return f2(vexp2_fast(lane0(x)), vexp2_fast(lane1(x))); }

// sin/cos of joint angles.  fp32: branch-free Cody-Waite + cephes minimax polynomials on the FMA pipe (max abs
// error 7e-8 on |x| <= 12; libdevice sincosf costs ~28 issue slots incl. a divergent slow-path guard, this ~21,
// and the packed form ~13 per angle).  fp64: libdevice.
__device__ __forceinline__ void vsincos(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ void vsincos(float x, float* sp, float* cp) {
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
__device__ __forceinline__ void vsincos(F2 x, F2* sp, F2* cp) {
    const F2 t = vfma(x, 0.636619772367581f, 12582912.0f);
    const int q0 = __float_as_int(lane0(t)), q1 = __float_as_int(lane1(t));
    const F2 j = t - 12582912.0f;
    F2 r = vfma(j, -1.5707963705062866f, x);
    r = vfma(j, 4.371138828673793e-08f, r);
    const F2 z = r * r;
    const F2 ps = vfma(vfma(f2(-1.9515295891e-4f, -1.9515295891e-4f), z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z;
    const F2 s = vfma(ps, r, r);
    const F2 c = vfma(vfma(vfma(vfma(f2(2.443315711809948e-5f, 2.443315711809948e-5f), z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z, -0.5f), z, 1.0f);
    const float s0 = lane0(s), s1 = lane1(s), c0 = lane0(c), c1 = lane1(c);
    const float ss0 = (q0 & 1) ? c0 : s0, cc0 = (q0 & 1) ? s0 : c0;
    const float ss1 = (q1 & 1) ? c1 : s1, cc1 = (q1 & 1) ? s1 : c1;
    *sp = f2(__int_as_float(__float_as_int(ss0) ^ ((q0 << 30) & 0x80000000)), __int_as_float(__float_as_int(ss1) ^ ((q1 << 30) & 0x80000000)));
    *cp = f2(__int_as_float(__float_as_int(cc0) ^ (((q0 + 1) << 30) & 0x80000000)), __int_as_float(__float_as_int(cc1) ^ (((q1 + 1) << 30) & 0x80000000)));
}

}  // namespace sgpmp
