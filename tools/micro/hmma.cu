// legacy warp-level mma.sync throughput on sm_100a (HMMA m16n8k16 f16 -> f32, m16n8k8 tf32 -> f32)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters) {
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x + 1u, 5u};
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float t = 0; for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) t += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 4000;
    for (int mode = 0; mode < 2; mode++) for (int rep = 0; rep < 2; rep++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters); else k<1><<<148 * 8, 256>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = 148.0 * 8 * 8 * 8.0 * iters;   // warps x 8 per iteration
        double mac = mmas * (mode == 0 ? 16 * 8 * 16 : 16 * 8 * 8);
        printf("{\"mode\": \"%s\", \"ms\": %.3f, \"mma_per_clk_per_sm\": %.3f, \"mac_per_clk_per_sm\": %.1f, \"dense_TFLOPs\": %.1f}\n", mode ? "m16n8k8.tf32" : "m16n8k16.f16", ms,
               mmas / (ms * 1e-3) / 148 / 1.965e9, mac / (ms * 1e-3) / 148 / 1.965e9, 2 * mac / (ms * 1e-3) / 1e12);
    }
    return 0;
}
