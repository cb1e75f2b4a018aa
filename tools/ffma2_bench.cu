// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
    float a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    float m = 1.0001f, c = 0.5f;
    unsigned long long m2, c2;
    { float2 t = make_float2(m, m); m2 = *(unsigned long long*)&t; t = make_float2(c, c); c2 = *(unsigned long long*)&t; }
    unsigned long long* p = (unsigned long long*)a;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], m, c);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
        }
    }
    float s = 0; for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters); else k<1><<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * 16.0 * iters;
            if (rep) printf("%s: %.3f ms, %.1f TFLOP/s fp32 (%.1f G fma-lanes/s)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9, fma / ms / 1e6);
        }
    }
    return 0;
}
