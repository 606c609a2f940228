// Device side of the interpolation plan (see ctx_plan in api.cu): edge extension exactly as
// `Task 5/interpolate.m:7-16`, banded derivative operator, Hermite evaluation per query carrier.
#pragma once
#include "common.cuh"

template <typename T> struct PlanDev {
    int n_knots, n_src, ext_lo, ext_hi, nq, hb, folded;
    T lo_den, lo_mul, hi_den, hi_mul;
    const T* band;
    const T* qw;
    const int32_t* qk;
};
template <typename T> static inline PlanDev<T> plan_dev(const InterpPlan* p) {
    PlanDev<T> d;
    d.n_knots = p->n_knots; d.n_src = p->n_src; d.ext_lo = p->ext_lo; d.ext_hi = p->ext_hi; d.nq = p->nq; d.hb = p->hb; d.folded = p->folded;
    d.lo_den = (T)p->lo_den; d.lo_mul = (T)p->lo_mul; d.hi_den = (T)p->hi_den; d.hi_mul = (T)p->hi_mul;
    d.band = (const T*)p->band; d.qw = (const T*)p->qw; d.qk = p->qk;
    return d;
}

// y: shared, n_knots entries, caller has written the n_src pilot values at y[ext_lo ...] and synchronised.
// d: shared scratch, n_knots entries.  emit(q, value) receives the nq interpolated values.  Block-cooperative,
// two barriers: the extrapolated end knots (`interpolate.m:7-16`) are produced by two threads *during* the
// derivative stage, whose operator was folded on the host so that it never reads them.
// Stage 1 of plan_apply_fn: edge extension + derivative operator (d = Band * y).  A block barrier must separate it from stage 2.
template <typename T>
__device__ __forceinline__ void plan_band_stage(const PlanDev<T>& p, cx<T>* y, cx<T>* d) {
    using C = cx<T>;
    const int n = p.n_knots;
    if (!p.folded && threadIdx.x == 0) {   // unfolded plans (tiny knot sets): extension first, then a barrier
        if (p.ext_lo) { C s = mk<T>((y[2].x - y[1].x) / p.lo_den, (y[2].y - y[1].y) / p.lo_den); y[0] = mk<T>(y[1].x - s.x * p.lo_mul, y[1].y - s.y * p.lo_mul); }
        if (p.ext_hi) { C s = mk<T>((y[n - 2].x - y[n - 3].x) / p.hi_den, (y[n - 2].y - y[n - 3].y) / p.hi_den); y[n - 1] = mk<T>(y[n - 2].x + s.x * p.hi_mul, y[n - 2].y + s.y * p.hi_mul); }
    }
    if (!p.folded) __syncthreads();
    C e_lo = mk<T>(0, 0), e_hi = mk<T>(0, 0);
    if (p.folded) {
        // slope = (H(2)-H(1))/(loc(2)-loc(1)); H0 = H(1) - slope*(loc(1)-1)      (`interpolate.m:8-9`)
        if (p.ext_lo && threadIdx.x == 0) { C s = mk<T>((y[2].x - y[1].x) / p.lo_den, (y[2].y - y[1].y) / p.lo_den); e_lo = mk<T>(y[1].x - s.x * p.lo_mul, y[1].y - s.y * p.lo_mul); }
        // slope = (H(end)-H(end-1))/(...); Hend = H(end) + slope*(N-loc(end))   (`interpolate.m:13-14`)
        if (p.ext_hi && threadIdx.x == 32 % blockDim.x) { C s = mk<T>((y[n - 2].x - y[n - 3].x) / p.hi_den, (y[n - 2].y - y[n - 3].y) / p.hi_den); e_hi = mk<T>(y[n - 2].x + s.x * p.hi_mul, y[n - 2].y + s.y * p.hi_mul); }
    }
    const int bw = 2 * p.hb + 1;
    const int jlo = p.folded ? p.ext_lo : 0, jhi = p.folded ? n - p.ext_hi : n;   // columns the operator may read
    // A few rows beyond a multiple of the block size (n = Np + 2 = 258 with 256 threads) would make two threads
    // walk a second row while everybody else waits at the barrier: give each of them to a warp, one lane per tap.
    int n_rows = n;
    {
        const int left = n % (int)blockDim.x, nw = (int)blockDim.x >> 5;
        if (p.hb > 0 && n > (int)blockDim.x && left > 0 && left <= nw) {
            n_rows = n - left;
            const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (w < left) {
                const int i = n_rows + w;
                const int c0 = max(0, p.hb - i + jlo), c1 = min(bw, jhi - i + p.hb);
                T ar = 0, ai = 0;
                for (int c = c0 + lane; c < c1; c += 32) { T wt = p.band[(size_t)c * n + i]; C v = y[i - p.hb + c]; ar += wt * v.x; ai += wt * v.y; }
                ar = warp_sum(ar); ai = warp_sum(ai);
                if (lane == 0) d[i] = mk<T>(ar, ai);
            }
        }
    }
    for (int i = threadIdx.x; i < n_rows; i += blockDim.x) {
        T ar = 0, ai = 0;
        if (p.hb > 0) {
            const T* col = p.band + i;            // band is stored tap-major [bw][n]: coalesced across threads
            const int c0 = max(0, p.hb - i + jlo), c1 = min(bw, jhi - i + p.hb);
#pragma unroll 4
            for (int c = c0; c < c1; ++c) { T w = col[(size_t)c * n]; C v = y[i - p.hb + c]; ar += w * v.x; ai += w * v.y; }
        }
        d[i] = mk<T>(ar, ai);
    }
    if (p.folded) {   // nobody reads y[0] / y[n-1] before the barrier below
        if (p.ext_lo && threadIdx.x == 0) y[0] = e_lo;
        if (p.ext_hi && threadIdx.x == 32 % blockDim.x) y[n - 1] = e_hi;
    }
}

// Stage 2 of plan_apply_fn: Hermite evaluation at the nq query carriers; emit(q, value).
template <typename T, typename Emit>
__device__ __forceinline__ void plan_hermite_stage(const PlanDev<T>& p, const cx<T>* y, const cx<T>* d, Emit emit) {
    using C = cx<T>;
    for (int q = threadIdx.x; q < p.nq; q += blockDim.x) {
        int k = p.qk[q];
        T w0 = p.qw[q], w1 = p.qw[p.nq + q], w2 = p.qw[2 * p.nq + q], w3 = p.qw[3 * p.nq + q];   // four planes [4][nq]
        C y0 = y[k], y1 = y[k + 1], d0 = d[k], d1 = d[k + 1];
        emit(q, mk<T>(w0 * y0.x + w1 * d0.x + w2 * y1.x + w3 * d1.x, w0 * y0.y + w1 * d0.y + w2 * y1.y + w3 * d1.y));
    }
}

template <typename T, typename Emit>
__device__ __forceinline__ void plan_apply_fn(const PlanDev<T>& p, cx<T>* y, cx<T>* d, Emit emit) {
    plan_band_stage<T>(p, y, d);
    __syncthreads();
    plan_hermite_stage<T>(p, y, d, emit);
}

template <typename T>
__device__ __forceinline__ void plan_apply(const PlanDev<T>& p, cx<T>* y, cx<T>* d, cx<T>* out) {
    plan_apply_fn<T>(p, y, d, [out](int q, cx<T> v) { out[q] = v; });
}
