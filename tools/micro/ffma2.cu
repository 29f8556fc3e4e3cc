// FFMA vs FFMA2 issue throughput on sm_100a: nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3f + i;
    float m = s, c = s * 0.5f;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], m, c);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long X, Y, Z, R;
                float2 x = make_float2(a[i], a[i + 1]), y = make_float2(m, m), z = make_float2(c, c);
                X = *(unsigned long long*)&x; Y = *(unsigned long long*)&y; Z = *(unsigned long long*)&z;
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(R) : "l"(X), "l"(Y), "l"(Z));
                float2 r = *(float2*)&R; a[i] = r.x; a[i + 1] = r.y;
            }
        }
    }
    float t = 0; for (int i = 0; i < 16; i++) t += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) for (int rep = 0; rep < 2; rep++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f); else k<1><<<148 * 8, 256>>>(d, iters, 1.0001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = 148.0 * 8 * 256 * 16.0 * iters;
        printf("{\"mode\": \"%s\", \"ms\": %.3f, \"TFMA_per_s\": %.2f, \"fma_per_clk_per_sm_at_1.965GHz\": %.1f}\n", mode ? "FFMA2" : "FFMA", ms, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
