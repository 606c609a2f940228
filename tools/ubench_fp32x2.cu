// Micro-benchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 (fma.rn.f32x2 / add.rn.f32x2) issue throughput on sm_100a.
// Decides whether the radix-16 butterflies should be written with float2-packed arithmetic.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void k(float* out, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, aa.x, bb.x); x[i].y = fmaf(x[i].y, aa.y, bb.y); }
            if (MODE == 1) { x[i] = __ffma2_rn(x[i], aa, bb); }
            if (MODE == 2) { x[i].x = x[i].x + bb.x; x[i].y = x[i].y + bb.y; }
            if (MODE == 3) { x[i] = __fadd2_rn(x[i], bb); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, 1.0001f, 1e-6f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(d, 1.0001f, 1e-6f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 5.0 * grid * block * (double)ITERS * 16;   // 16 scalar results per iteration per thread
    printf("%-8s %8.3f ms  %8.2f T scalar-results/s  (%.1f results/clk/SM at 1.965 GHz)\n", name, ms, lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("FADD", d); run<3>("FADD2", d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
