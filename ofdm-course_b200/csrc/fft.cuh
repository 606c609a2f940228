// Block-cooperative shared-memory Stockham FFT (radix-4 passes, one leading radix-2 pass when
// log2 N is odd).  Generic size path, N = 2^k; the fused M1 chain has its own register-resident
// radix-16 kernel in chain_rx.cu.  Replaces MATLAB fft/ifft as called from
// `Task 5/OFDM_modulator.m:5`, `OFDM_demodulator.m:8`, `remove_IFO.m:5`, `OMP_estimate.m:36`.
#pragma once
#include "common.cuh"

// a, b: two shared-memory buffers of N complex each; input in `a` (natural order).  `tw` is the
// table W_N^k = exp(-2*pi*i*k/N), k = 0..N-1 (global memory, context precision).  All threads of
// the block must call; returns the buffer that holds the natural-order result.  The caller must
// __syncthreads() between filling `a` and calling; the result is synchronised on return.
template <typename T, bool INV>
__device__ __forceinline__ cx<T>* block_fft(cx<T>* a, cx<T>* b, int N, int logN, const cx<T>* __restrict__ tw) {
    using C = cx<T>;
    const int tid = threadIdx.x, nt = blockDim.x;
    int Ns = 1;
    if (logN & 1) {
        for (int j = tid; j < (N >> 1); j += nt) {
            C v0 = a[j], v1 = a[j + (N >> 1)];
            b[2 * j] = v0 + v1;
            b[2 * j + 1] = v0 - v1;
        }
        __syncthreads();
        C* t = a; a = b; b = t;
        Ns = 2;
    }
    const int Q = N >> 2;
    while (Ns < N) {
        const int tmul = N / (4 * Ns);
        for (int j = tid; j < Q; j += nt) {
            const int k = j & (Ns - 1);
            const int ts = k * tmul;
            C w1 = tw[ts], w2 = tw[2 * ts], w3 = tw[3 * ts];
            if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            C v0 = a[j];
            C v1 = cmul(a[j + Q], w1);
            C v2 = cmul(a[j + 2 * Q], w2);
            C v3 = cmul(a[j + 3 * Q], w3);
            C t0 = v0 + v2, t1 = v0 - v2, t2 = v1 + v3;
            C t3 = INV ? mul_pi(v1 - v3) : mul_mi(v1 - v3);
            const int d = ((j - k) << 2) + k;
            b[d] = t0 + t2;
            b[d + Ns] = t1 + t3;
            b[d + 2 * Ns] = t0 - t2;
            b[d + 3 * Ns] = t1 - t3;
        }
        __syncthreads();
        C* t = a; a = b; b = t;
        Ns <<= 2;
    }
    return a;
}

static inline int fft_threads(int N) { int t = N / 4; if (t < 32) t = 32; if (t > 256) t = 256; return t; }
