// Helpers shared by the SIMT pursuit kernel (sparse.cu) and the tensor-core OMP path (sparse_tc.cu).
#pragma once
#include "fft.cuh"

#define PU_THREADS 256
#define PU_MAXK 32

template <typename T> __device__ __forceinline__ void block_argmax(T& val, int& idx, T* sval, int* sidx) {
    // first maximum wins (MATLAB max): larger value, or equal value with lower index
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T ov = __shfl_xor_sync(0xffffffffu, val, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
    __syncthreads();
    if (lane == 0) { sval[w] = val; sidx[w] = idx; }
    __syncthreads();
    T v = (lane < nw) ? sval[lane] : (T)-CUDART_INF;
    int i = (lane < nw) ? sidx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    val = v; idx = i;
}

__device__ __forceinline__ double2 block_csum(double2 v, double* red) {
    double a = block_sum(v.x, red);
    double b = block_sum(v.y, red);
    return make_double2(a, b);
}

// Solve the Hermitian positive (semi-)definite k x k system G x = g by Cholesky in double (thread 0).
static __device__ void chol_solve(int k, const double2 (*G)[PU_MAXK], const double2* g, double2* x, double2 (*Lm)[PU_MAXK]) {
    for (int j = 0; j < k; ++j) {
        double s = G[j][j].x;
        for (int q = 0; q < j; ++q) s -= Lm[j][q].x * Lm[j][q].x + Lm[j][q].y * Lm[j][q].y;
        double dj = sqrt(fmax(s, 0.0));
        Lm[j][j] = make_double2(dj, 0.0);
        for (int i = j + 1; i < k; ++i) {
            double2 a = G[i][j];
            for (int q = 0; q < j; ++q) a = a - cmulc(Lm[i][q], Lm[j][q]);
            Lm[i][j] = dj > 0 ? cscale(a, 1.0 / dj) : make_double2(0, 0);
        }
    }
    double2 z[PU_MAXK];
    for (int i = 0; i < k; ++i) {          // L z = g
        double2 a = g[i];
        for (int q = 0; q < i; ++q) a = a - cmul(Lm[i][q], z[q]);
        z[i] = Lm[i][i].x > 0 ? cscale(a, 1.0 / Lm[i][i].x) : make_double2(0, 0);
    }
    for (int i = k - 1; i >= 0; --i) {     // L^H x = z
        double2 a = z[i];
        for (int q = i + 1; q < k; ++q) a = a - cmul(cconj(Lm[q][i]), x[q]);
        x[i] = Lm[i][i].x > 0 ? cscale(a, 1.0 / Lm[i][i].x) : make_double2(0, 0);
    }
}

