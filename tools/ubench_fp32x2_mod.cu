// Micro-benchmark: do the operand modifiers of the packed FP32 pipe (F32x2.LO_HI swap, .F32 broadcast, .NP per-half
// sign) cost throughput?  Complex multiply as 4 scalar ops vs FMUL2+FFMA2, radix-4 butterfly in both forms.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
template <int MODE> __global__ void k(float* out, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
    for (int it = 0; it < ITERS; ++it) {
        if (MODE < 6) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = __ffma2_rn(x[i], aa, bb);                               // plain
                if (MODE == 1) x[i] = __ffma2_rn(swp(x[i]), aa, bb);                          // LO_HI
                if (MODE == 2) x[i] = __ffma2_rn(x[i], make_float2(aa.x, aa.x), bb);          // .F32 broadcast
                if (MODE == 3) x[i] = __ffma2_rn(swp(x[i]), make_float2(-aa.y, aa.y), bb);    // LO_HI.NP + .F32
                if (MODE == 4) x[i] = make_float2(x[i].x * aa.x - x[i].y * aa.y, x[i].x * aa.y + x[i].y * aa.x);   // scalar cmul
                if (MODE == 5) { const float2 t = __fmul2_rn(x[i], make_float2(aa.x, aa.x)); x[i] = __ffma2_rn(swp(x[i]), make_float2(-aa.y, aa.y), t); }
            }
        } else {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float2 &a0 = x[4 * g], &a1 = x[4 * g + 1], &a2 = x[4 * g + 2], &a3 = x[4 * g + 3];
                const float2 t0 = add2(a0, a2), t1 = sub2(a0, a2), t2 = add2(a1, a3);
                if (MODE == 6) {
                    const float2 t3s = make_float2(a1.y - a3.y, a1.x - a3.x);
                    a0 = add2(t0, t2); a2 = sub2(t0, t2);
                    a1 = __ffma2_rn(t3s, make_float2(1.f, -1.f), t1);
                    a3 = __ffma2_rn(t3s, make_float2(-1.f, 1.f), t1);
                } else {
                    const float2 d = sub2(a1, a3);
                    a0 = add2(t0, t2); a2 = sub2(t0, t2);
                    a1 = __ffma2_rn(swp(d), make_float2(1.f, -1.f), t1);
                    a3 = __ffma2_rn(swp(d), make_float2(-1.f, 1.f), t1);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) x[4 * g + j] = __fmul2_rn(x[4 * g + j], make_float2(0.5f, 0.5f));
            }
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
    double items = 5.0 * grid * block * (double)ITERS * 8;   // 8 complex items per iteration per thread
    printf("%-22s %8.3f ms  %7.2f complex items/clk/SM at 1.965 GHz\n", name, ms, items / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FFMA2 plain", d); run<1>("FFMA2 LO_HI", d); run<2>("FFMA2 .F32", d); run<3>("FFMA2 LO_HI.NP+.F32", d);
    run<4>("cmul scalar (4 ops)", d); run<5>("cmul FMUL2+FFMA2", d); run<6>("fft4 scalar-diff", d); run<7>("fft4 packed-diff", d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
