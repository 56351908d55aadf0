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

}  // namespace sgpmp

using namespace sgpmp;

/* mode 0: FP32 FMA (2 flop each; launches `blocks` x 256 threads, 128 FMA per thread per iteration);
 * mode 1: MUFU ex2 (64 per thread per iteration). */
extern "C" int sgpmp_probe(int32_t mode, int32_t blocks, int32_t iters, void* scratch, void* stream) {
    SGPMP_REQUIRE(blocks > 0 && iters > 0 && scratch, "sgpmp_probe: bad arguments");
    if (mode == 0)
        probe_fma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)scratch);
    else
        probe_mufu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5f, (float*)scratch);
    SGPMP_CHECK_LAUNCH("sgpmp_probe");
    return SGPMP_OK;
}
