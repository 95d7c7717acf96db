// Micro-benchmark: CUDA-core FP32 issue rates on sm_100a (FFMA 3-reg, FFMA2 / FADD2 packed, mixed FADD+FFMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dev_fp32_rate tests/devtools/dev_fp32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[16]; u64 p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 1.f); }
    const u64 ss = pk(s, s * 1.0001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], s, a[(i + 1) & 15]);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(ss), "l"(p[(i + 1) & 15]));
            if (MODE == 2) asm volatile("add.f32x2 %0, %1, %2;" : "=l"(p[i]) : "l"(p[i]), "l"(ss));
            if (MODE == 3) { float d = a[i] - s; a[(i + 1) & 15] = fmaf(d, d, a[(i + 1) & 15]); }
            if (MODE == 4) { u64 d; asm volatile("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(p[i]), "l"(ss));
                             asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(p[(i + 1) & 15]) : "l"(d), "l"(p[(i + 1) & 15])); }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i] + lo(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int ops_per_inner, int flop_per_op) {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 4, 512>>>(out, 100, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 148.0 * 4 * 512 * (double)iters * 16 * ops_per_inner;   // thread-level instructions
    printf("%-28s %8.3f ms  %7.2f Tinst-lanes/s  %7.2f TFLOP/s\n", name, ms, inst / ms * 1e-9, inst * flop_per_op / ms * 1e-9);
    cudaFree(out);
}
int main() {
    run<0>("FFMA (3-reg)", 1, 2);
    run<1>("FFMA2 (packed)", 1, 4);
    run<2>("FADD2 (packed)", 1, 2);
    run<3>("FADD + FFMA (diff^2)", 2, 1);     // flop column = pairs-of-(sub,fma) x 1.5 ... report inst rate
    run<4>("FADD2 + FFMA2 (diff^2 x2)", 2, 2);
    return 0;
}
