// Microbenchmark: does packed fp32 (fma.rn.f32x2 -> FFMA2) free issue slots on sm_100a?
// Each round issues NF scalar FFMAs (or NF/2 FFMA2) on 8 independent chains plus NI integer LOP3s on 4 independent
// chains (asm volatile: nothing is folded).  cycles/round = time * clock / rounds / warps per scheduler.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}

template <bool PACKED, int NI>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s, uint32_t m) {
    float a[8];
    unsigned long long p[4];
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { p[i] = pk(a[2 * i], a[2 * i + 1]); u[i] = threadIdx.x + i; }
    const unsigned long long s2 = pk(s, s), c2 = pk(1e-3f, 2e-3f);
    const float c1 = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (!PACKED) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(c1));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(s2), "l"(c2));
            }
#pragma unroll
            for (int i = 0; i < NI; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i & 3]) : "r"(m), "r"(it));
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)) + __uint_as_float(u[i] & 0x3f800000u);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <bool PACKED, int NI>
void run(const char *name, float *d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;  // 8 CTAs x 8 warps per SM = 16 warps per scheduler
    k<PACKED, NI><<<blocks, 256>>>(d, iters, 0.999f, 0x5bd1e995u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<PACKED, NI><<<blocks, 256>>>(d, iters, 0.999f, 0x5bd1e995u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 8 * (double)iters * blocks * 256;
    const double cycles_per_round = ms * 1e-3 * 1.92e9 / ((double)iters * 8 * 16);
    printf("%-34s %8.3f ms  %6.2f TFLOP/s  %5.2f cycles/round/warp at 1.92 GHz (%s)\n", name, ms, flops / ms * 1e-9,
           cycles_per_round, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    run<false, 0>("8 FFMA", d, iters);
    run<true, 0>("4 FFMA2", d, iters);
    run<false, 2>("8 FFMA + 2 LOP3", d, iters);
    run<true, 2>("4 FFMA2 + 2 LOP3", d, iters);
    run<false, 4>("8 FFMA + 4 LOP3", d, iters);
    run<true, 4>("4 FFMA2 + 4 LOP3", d, iters);
    run<false, 8>("8 FFMA + 8 LOP3", d, iters);
    run<true, 8>("4 FFMA2 + 8 LOP3", d, iters);
    return 0;
}
