// Bring-up of tc_corr_top_kernel: random complex dictionary and residuals, top-4 screened indices vs a CPU argmax.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ofdm-course_b200/csrc/tc_gemm.cuh"
int main(int argc, char** argv) {
    const int F = argc > 1 ? atoi(argv[1]) : 256, L = argc > 2 ? atoi(argv[2]) : 1024, Np = argc > 3 ? atoi(argv[3]) : 256;
    const int K2 = 2 * Np;
    printf("corr-top test: frames=%d Ldict=%d Np=%d\n", F, L, Np);
    std::vector<float> Rt((size_t)F * K2), Bt((size_t)2 * L * K2);
    srand(2);
    for (auto& x : Rt) x = rand() / (float)RAND_MAX - 0.5f;
    std::vector<float> Ar((size_t)L * Np), Ai((size_t)L * Np);
    for (size_t i = 0; i < Ar.size(); ++i) { Ar[i] = rand() / (float)RAND_MAX - 0.5f; Ai[i] = rand() / (float)RAND_MAX - 0.5f; }
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < Np; ++i) {
            Bt[(size_t)(2 * l) * K2 + i] = Ar[(size_t)l * Np + i]; Bt[(size_t)(2 * l) * K2 + Np + i] = Ai[(size_t)l * Np + i];
            Bt[(size_t)(2 * l + 1) * K2 + i] = -Ai[(size_t)l * Np + i]; Bt[(size_t)(2 * l + 1) * K2 + Np + i] = Ar[(size_t)l * Np + i];
        }
    float *dR, *dB, *dS; int32_t* dC;
    cudaMalloc(&dR, Rt.size() * 4); cudaMalloc(&dB, Bt.size() * 4); cudaMalloc(&dC, (size_t)F * TC_TOP * 4); cudaMalloc(&dS, (size_t)F * TC_TOP * 4);
    cudaMemcpy(dR, Rt.data(), Rt.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap mapA, mapB;
    if (!tc_make_kmajor_map(&mapA, dR, F, K2) || !tc_make_kmajor_map(&mapB, dB, 2 * L, K2)) { printf("map failed\n"); return 2; }
    cudaFuncSetAttribute(tc_corr_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_TOP_BYTES);
    tc_corr_top_kernel<<<F / TC_BM, TC_THREADS, TC_SMEM_TOP_BYTES>>>(mapA, mapB, 2 * L / TC_BN, K2, L, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) tc_corr_top_kernel<<<F / TC_BM, TC_THREADS, TC_SMEM_TOP_BYTES>>>(mapA, mapB, 2 * L / TC_BN, K2, L, dC, dS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    std::vector<int32_t> C((size_t)F * TC_TOP); std::vector<float> S((size_t)F * TC_TOP);
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxrel = 0;
    for (int f = 0; f < F; f += 3) {
        double best = -1; int bl = -1;
        for (int l = 0; l < L; ++l) {
            double cr = 0, ci = 0;
            for (int i = 0; i < Np; ++i) {
                double rr = Rt[(size_t)f * K2 + i], ri = Rt[(size_t)f * K2 + Np + i], ar = Ar[(size_t)l * Np + i], ai = Ai[(size_t)l * Np + i];
                cr += ar * rr + ai * ri; ci += ar * ri - ai * rr;
            }
            double s = cr * cr + ci * ci;
            if (s > best) { best = s; bl = l; }
        }
        bool found = false;
        for (int i = 0; i < TC_TOP; ++i) if (C[(size_t)f * TC_TOP + i] == bl) found = true;
        if (!found) ++bad;
        maxrel = fmax(maxrel, fabs(S[(size_t)f * TC_TOP] - best) / best);
    }
    printf("frames whose exact argmax is missing from the top-%d: %d; top score rel. error %.2e; %.3f ms, %.1f TFLOP/s\n", TC_TOP, bad, maxrel, ms,
           2.0 * F * 2 * L * K2 / ms / 1e9);
    printf(bad == 0 ? "PASS\n" : "FAIL\n");
    return bad != 0;
}
