// Register-resident FFT building blocks (FP32, forward transform) shared by the fused RX kernels, and the
// separable 16QAM hard decision.  fft16 is two radix-4 stages on packed FP32x2 arithmetic; fft32 is a radix-2
// combination of two fft16.  Everything is meant to be fully unrolled on register arrays, so unused outputs
// are pruned by dead-code elimination (the callers only consume the N_carrier lowest bins,
// `Task 5/equalize_signal.m:6`).
#pragma once
#include "common.cuh"

// Packed FP32x2 arithmetic (FADD2 / FFMA2, new on sm_100): one issue slot per complex add.  Measured on
// B200 (tools/ubench_fp32x2.cu): FADD2 127, FFMA2 117, scalar FADD 117, scalar 3-register FFMA 71
// results/clk/SM -- the packed forms halve the issue slots of the butterflies.
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = add2(a0, a2), t1 = sub2(a0, a2), t2 = add2(a1, a3);
    const float2 t3s = make_float2(a1.y - a3.y, a1.x - a3.x);          // (a1 - a3) with the halves swapped
    a0 = add2(t0, t2);
    a2 = sub2(t0, t2);
    a1 = __ffma2_rn(t3s, make_float2(1.f, -1.f), t1);                   // t1 + (-i)(a1 - a3)
    a3 = __ffma2_rn(t3s, make_float2(-1.f, 1.f), t1);                   // t1 - (-i)(a1 - a3)
}
#define C16_1 0.92387953251128674f
#define S16_1 0.38268343236508977f
#define RSQ2 0.70710678118654752f
// v[4a+b] in  ->  X[c+4d] at v[4c+d]   (forward 16-point DFT, radix 4x4)
__device__ __forceinline__ void fft16_steps12(float2* v) {
#pragma unroll
    for (int b = 0; b < 4; ++b) fft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // y_b[c] sits at v[4c+b]; multiply by W16^{bc}
    float2 t;
    t = v[5];  v[5]  = make_float2(t.x * C16_1 + t.y * S16_1, t.y * C16_1 - t.x * S16_1);       // W16^1
    t = v[6];  v[6]  = make_float2((t.x + t.y) * RSQ2, (t.y - t.x) * RSQ2);                       // W16^2
    t = v[7];  v[7]  = make_float2(t.x * S16_1 + t.y * C16_1, t.y * S16_1 - t.x * C16_1);       // W16^3
    t = v[9];  v[9]  = make_float2((t.x + t.y) * RSQ2, (t.y - t.x) * RSQ2);                       // W16^2
    t = v[10]; v[10] = make_float2(t.y, -t.x);                                                      // W16^4 = -i
    t = v[11]; v[11] = make_float2((t.y - t.x) * RSQ2, -(t.x + t.y) * RSQ2);                      // W16^6
    t = v[13]; v[13] = make_float2(t.x * S16_1 + t.y * C16_1, t.y * S16_1 - t.x * C16_1);       // W16^3
    t = v[14]; v[14] = make_float2((t.y - t.x) * RSQ2, -(t.x + t.y) * RSQ2);                      // W16^6
    t = v[15]; v[15] = make_float2(-t.x * C16_1 - t.y * S16_1, t.x * S16_1 - t.y * C16_1);      // W16^9
}
// Same transform when only v[0..3] are non-zero (the TX side: rows 1..1024 of 4096 carry carriers): the first radix-4
// stage degenerates to copies, so it is the twiddles and the second stage only.
__device__ __forceinline__ void fft16_in4(float2* v) {
    const float2 x1 = v[1], x2 = v[2], x3 = v[3];
    v[4] = v[0]; v[8] = v[0]; v[12] = v[0];
    v[5]  = make_float2(x1.x * C16_1 + x1.y * S16_1, x1.y * C16_1 - x1.x * S16_1);            // W16^1
    v[6]  = make_float2((x2.x + x2.y) * RSQ2, (x2.y - x2.x) * RSQ2);                            // W16^2
    v[7]  = make_float2(x3.x * S16_1 + x3.y * C16_1, x3.y * S16_1 - x3.x * C16_1);            // W16^3
    v[9]  = make_float2((x1.x + x1.y) * RSQ2, (x1.y - x1.x) * RSQ2);                            // W16^2
    v[10] = make_float2(x2.y, -x2.x);                                                            // W16^4 = -i
    v[11] = make_float2((x3.y - x3.x) * RSQ2, -(x3.x + x3.y) * RSQ2);                           // W16^6
    v[13] = make_float2(x1.x * S16_1 + x1.y * C16_1, x1.y * S16_1 - x1.x * C16_1);            // W16^3
    v[14] = make_float2((x2.y - x2.x) * RSQ2, -(x2.x + x2.y) * RSQ2);                           // W16^6
    v[15] = make_float2(-x3.x * C16_1 - x3.y * S16_1, x3.x * S16_1 - x3.y * C16_1);           // W16^9
#pragma unroll
    for (int c = 0; c < 4; ++c) fft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void fft16(float2* v) {
    fft16_steps12(v);
#pragma unroll
    for (int c = 0; c < 4; ++c) fft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}


// 32-point forward DFT of v[n], n = 0..31, result X[k] in v[k] (natural order).  Only outputs k < KOUT are
// produced (KOUT <= 16 needs the "+" half of the last radix-2 stage only).
template <int KOUT>
__device__ __forceinline__ void fft32(float2* v) {
    float2 e[16], o[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) { e[m] = v[2 * m]; o[m] = v[2 * m + 1]; }
    fft16(e);
    fft16(o);
    // W32^k = exp(-2*pi*i*k/32), k = 0..15
    const float wc[16] = {1.f, 0.98078528040323044f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960222f,
                          0.38268343236508977f, 0.19509032201612827f, 0.f, -0.19509032201612827f, -0.38268343236508977f, -0.55557023301960222f,
                          -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323044f};
    const float ws[16] = {0.f, -0.19509032201612827f, -0.38268343236508977f, -0.55557023301960222f, -0.70710678118654752f, -0.83146961230254524f,
                          -0.92387953251128674f, -0.98078528040323044f, -1.f, -0.98078528040323044f, -0.92387953251128674f, -0.83146961230254524f,
                          -0.70710678118654752f, -0.55557023301960222f, -0.38268343236508977f, -0.19509032201612827f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < KOUT) {
            const int pos = 4 * (k & 3) + (k >> 2);       // fft16 leaves X[c + 4d] at [4c + d]
            const float2 E = e[pos], O = o[pos];
            float2 t;
            if (k == 0) t = O;
            else if (k == 8) t = make_float2(O.y, -O.x);
            else t = make_float2(O.x * wc[k] - O.y * ws[k], O.x * ws[k] + O.y * wc[k]);
            v[k] = add2(E, t);
            if (k + 16 < KOUT) v[k + 16] = sub2(E, t);
        }
    }
}

// 16QAM hard decision, separable, with the reference's first-minimum tie rule (`demapping.m:12`).
// Table index = 8*(x>0) + 4*(|x|<2a) + 2*(y<0) + (|y|<2a): I levels {-3,-1,+3,+1} -> codes {00,01,10,11},
// Q levels {+3,+1,-3,-1} -> {00,01,10,11}; ties (x = 0, |x| = 2a, ...) fall to the lower table index
// exactly as `min` does, and a NaN never wins a '<' so it decodes to index 1 (bits 0000).
// Returned value is the index bit-reversed (MSB-first symbol bits inside LSB-first packing).
template <bool NEAR>
__device__ __forceinline__ uint32_t demap16_nib(float x, float y, float two_a, float* margin) {
    const float ax = fabsf(x), ay = fabsf(y);
    uint32_t nib = (x > 0.f ? 1u : 0u) | (ax < two_a ? 2u : 0u) | (y < 0.f ? 4u : 0u) | (ay < two_a ? 8u : 0u);
    if (!(ax + ay <= CUDART_INF_F)) nib = 0u;                             // NaN in either part
    if (NEAR) {
        float dx = fminf(ax, fabsf(ax - two_a)), dy = fminf(ay, fabsf(ay - two_a));
        *margin = 2.f * two_a * fminf(dx, dy);  // second-best minus best squared distance
    }
    return nib;
}

