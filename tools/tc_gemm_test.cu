// Standalone bring-up of the tcgen05 (TF32 -> FP32 in TMEM) GEMM used by the batched OMP correlation.
//   D[M x N] = A[M x K] * B[N x K]^T, A and B K-major (row = M or N index, K contiguous), FP32 storage read as TF32.
// TMA (cp.async.bulk.tensor.2d, 128B swizzle) feeds a 4-stage shared-memory ring; one thread issues tcgen05.mma
// (cta_group::1, M=128, N=128, K=8 per instruction); four warps drain the TMEM accumulator with tcgen05.ld.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tc_gemm_test tools/tc_gemm_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ofdm-course_b200/csrc/tc_gemm.cuh"

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 256, N = argc > 2 ? atoi(argv[2]) : 384, K = argc > 3 ? atoi(argv[3]) : 512;
    printf("tcgen05 TF32 GEMM test: M=%d N=%d K=%d\n", M, N, K);
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, (size_t)M * N * 4);
    CUtensorMap mapA, mapB;
    if (!tc_make_kmajor_map(&mapA, dA, M, K) || !tc_make_kmajor_map(&mapB, dB, N, K)) { printf("tensor map creation failed\n"); return 2; }
    cudaFuncSetAttribute(tc_gemm_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    dim3 grid(M / TC_BM, N / TC_BN);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tc_gemm_store_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES>>>(mapA, mapB, dD, N, K, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("first launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) tc_gemm_store_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES>>>(mapA, mapB, dD, N, K, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    std::vector<float> D((size_t)M * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M; i += 7)
        for (int j = 0; j < N; j += 5) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[(size_t)i * K + k] * B[(size_t)j * K + k];
            maxerr = fmax(maxerr, fabs(s - D[(size_t)i * N + j]));
            maxref = fmax(maxref, fabs(s));
        }
    printf("max |err| = %.3e (max |ref| = %.3f) -> relative %.2e [TF32 expects ~1e-3]; %.3f ms, %.1f TFLOP/s\n", maxerr, maxref, maxerr / maxref, ms,
           2.0 * M * N * K / ms / 1e9);
    printf(maxerr / maxref < 5e-3 ? "PASS\n" : "FAIL\n");
    return maxerr / maxref < 5e-3 ? 0 : 1;
}
