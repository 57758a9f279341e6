// Issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100a) with 8 independent chains per thread.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_probe ffma2_probe.cu     Run: ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
    float acc = 0.f;
    if (MODE == 0) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = seed + i + threadIdx.x;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fma1(v[i], 0.999f, 0.001f);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i];
    } else {
        u64 v[8];
        const float2 m = make_float2(0.999f, 0.999f), c = make_float2(0.001f, 0.001f);
        const u64 mm = *reinterpret_cast<const u64*>(&m), cc = *reinterpret_cast<const u64*>(&c);
#pragma unroll
        for (int i = 0; i < 8; ++i) { float2 t = make_float2(seed + 2 * i + threadIdx.x, seed + 2 * i + 1 + threadIdx.x); v[i] = *reinterpret_cast<u64*>(&t); }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma2(v[i], mm, cc);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&v[i]); acc += t.x + t.y; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int blocks_per_sm : {1, 2, 4, 8}) {
            const int grid = 148 * blocks_per_sm;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) probe<0><<<grid, 256>>>(out, iters, 1.f); else probe<1><<<grid, 256>>>(out, iters, 1.f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fmas = double(grid) * 256 * 16 * iters;          // scalar-equivalent FMAs (both modes do 16 per thread per iteration)
            printf("%s  %d blocks/SM (%d warps/SM): %.3f ms  %.1f TFLOP/s fp32  (%.1f FMA lanes / clk / SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ",
                   blocks_per_sm, blocks_per_sm * 8, ms, 2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / 148 / 1.965e9);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
