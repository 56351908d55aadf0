// Roofline denominators that MEASURED_PEAKS.json does not carry: FP32 FMA pipe and MUFU (SFU) pipe peaks.
// bench.py times these with CUDA events in the same run as the kernels it reports on.
#include "sgpmp_common.cuh"

namespace sgpmp {

// 8 independent FMA chains per thread, fully unrolled inner block: issue-bound on the FP32 pipe.
__global__ void __launch_bounds__(256) probe_fma_kernel(int iters, float seed, float* __restrict__ out) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456f) out[0] = r;   // never true; keeps the chains alive
}

// 8 independent ex2.approx chains per thread: issue-bound on the MUFU pipe.
__global__ void __launch_bounds__(256) probe_mufu_kernel(int iters, float seed, float* __restrict__ out) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + 0.01f * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
        }
    }
    float r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += a[k];
    if (r == 123.456f) out[0] = r;
}

__device__ __forceinline__ unsigned long long ffma2_raw(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Packed-FP32 / issue-port experiments (64 "FP slots" per thread per iteration):
//   mode 2: 64 FFMA2 (packed fp32x2: 128 FMA)              mode 5: 128 FFMA (same FP work, scalar)
//   mode 3: 128 FFMA + 64 independent LOP3                 mode 4: 64 FFMA2 + 64 independent LOP3
// If FFMA2 frees issue slots, mode 4 runs in the time of mode 2 while mode 3 needs 1.5x mode 5.
template <int mode>
__global__ void __launch_bounds__(256) probe_mix_kernel(int iters, float seed, float* __restrict__ out) {
    float2 f[8];
    unsigned u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { f[k] = make_float2(seed + threadIdx.x + k, seed - k); u[k] = threadIdx.x * 2654435761u + k; }
    const float2 m = make_float2(0.999f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    const unsigned long long mm = *reinterpret_cast<const unsigned long long*>(&m), cc = *reinterpret_cast<const unsigned long long*>(&c);
    const unsigned ka = 0x9E3779B9u + threadIdx.x, kb = 0x85EBCA6Bu;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (mode == 2 || mode == 4) {
                    unsigned long long v = *reinterpret_cast<unsigned long long*>(&f[k]);
                    v = ffma2_raw(v, mm, cc);
                    f[k] = *reinterpret_cast<float2*>(&v);
                } else {
                    f[k].x = fmaf(f[k].x, m.x, c.x);
                    f[k].y = fmaf(f[k].y, m.y, c.y);
                }
                if (mode == 3 || mode == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(ka), "r"(kb));
            }
        }
    }
    float r = 0;
    unsigned q = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { r += f[k].x + f[k].y; q ^= u[k]; }
    if (r == 123.456f || q == 0x12345u) out[0] = r + q;
}

}  // namespace sgpmp

using namespace sgpmp;

/* mode 0: FP32 FMA (2 flop each; launches `blocks` x 256 threads, 128 FMA per thread per iteration);
 * mode 1: MUFU ex2 (64 per thread per iteration);
 * modes 2-4: packed-FP32 / mixed-issue experiments (64 FP instructions per thread per iteration, see probe_mix_kernel). */
extern "C" int sgpmp_probe(int32_t mode, int32_t blocks, int32_t iters, void* scratch, void* stream) {
    SGPMP_REQUIRE(blocks > 0 && iters > 0 && scratch, "sgpmp_probe: bad arguments");
    if (mode == 0)
        probe_fma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)scratch);
    else if (mode == 1)
        probe_mufu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5f, (float*)scratch);
    else if (mode == 2)
        probe_mix_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)scratch);   // 64 FP instr / thread / iter
    else if (mode == 3)
        probe_mix_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)scratch);
    else if (mode == 4)
        probe_mix_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)scratch);
    else
        probe_mix_kernel<5><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)scratch);
    SGPMP_CHECK_LAUNCH("sgpmp_probe");
    return SGPMP_OK;
}
