// Throughput of scalar FFMA against packed FFMA2 (fma.rn.f32x2) on sm_100a: same FLOPs, half the issue slots.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

template <bool PACKED>
__global__ void __launch_bounds__(256) k(float *out, float x, int iters) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 m = make_float2(x, x * 0.999f), a = make_float2(1e-6f, 2e-6f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) {
                acc[i] = fma2(acc[i], m, a);
            } else {
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(acc[i].x) : "f"(acc[i].x), "f"(m.x), "f"(a.x));
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(acc[i].y) : "f"(acc[i].y), "f"(m.y), "f"(a.y));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    const int ctas = 148 * 8, threads = 256, iters = 1 << 14;
    float *out;
    cudaMalloc(&out, sizeof(float) * ctas * threads);
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(s);
            if (packed) k<true><<<ctas, threads>>>(out, 0.9999f, iters);
            else k<false><<<ctas, threads>>>(out, 0.9999f, iters);
            cudaEventRecord(e);
            cudaEventSynchronize(e);
            float ms;
            cudaEventElapsedTime(&ms, s, e);
            const double fma = double(ctas) * threads * iters * 16.0;
            if (rep == 2) printf("%s: %.3f ms, %.1f TFMA/s (%.1f TFLOP/s)\n", packed ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
