// Counter-based RNG for throughput runs: Philox4x32-10 keyed by (seed), counter = (stream id, sample
// index), so a stream's noise is identical for any batch split or GPU count.
// The reference's `normrnd` stream (`Task 5/Noise.m:7-8`) is MATLAB-only; parity runs import normals.
#pragma once
#include <stdint.h>

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}
// Unit normals for a PAIR of samples (2*pair, 2*pair + 1) of a stream from ONE Philox call: the four 32-bit outputs
// feed two Box-Muller transforms, (g[0], g[1]) = real/imaginary normal of the even sample, (g[2], g[3]) of the odd one.
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void philox_normal_quad(uint64_t seed, uint64_t stream, uint64_t pair, float g[4]) {
    uint32_t c[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        // u1 = (c+1) 2^-32 in (0,1] is never denormal, so the bare MUFU forms do: r = sqrt(-2 ln u1) = sqrt(-2 ln2 * log2 u1)
        const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;
        const float r = sqrt_approx(lg2_approx(u1) * -1.3862943611198906f);
        float s, co;
        __sincosf((float)c[2 * h + 1] * 1.4629180792671596e-9f, &s, &co);     // 2 pi u2, u2 = c 2^-32 in [0,1)
        g[2 * h] = r * co; g[2 * h + 1] = r * s;
    }
}
// two independent unit normals for (stream, sample n): the half of the pair's quad that belongs to n
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t stream, uint64_t n, float& g1, float& g2) {
    float g[4];
    philox_normal_quad(seed, stream, n >> 1, g);
    g1 = (n & 1) ? g[2] : g[0];
    g2 = (n & 1) ? g[3] : g[1];
}
