// AutoCorrFunction / remove_IFO / fine_sync -- Task-4 synchronisation on B independent streams.
#include "fft.cuh"

// =====================================================================================
// AutoCorrFunction (`Task 5/AutoCorrFunction.m:1-28`)
// Pass 1: per (stream, tile of outputs) the three sliding-window sums come from one block-wide
// prefix scan held in double (no FP32 cancellation), rho(n) and the "> 0.77 and index > W" flag
// are produced; flags go out as one ballot word per 32 positions.
// Pass 2: one warp per stream finds the first run of consecutive flagged indices, requires a
// second run to exist (else the reference's catch branch: TgPosition = 65), and evaluates
// FreqOffset = -angle(rho(TgPosition))/(2*pi) from a directly summed window.
// =====================================================================================
#define AC_TILE 1024
#define AC_THREADS 256

template <typename T>
__global__ void __launch_bounds__(AC_THREADS) autocorr_kernel(const cx<T>* __restrict__ rx, int64_t L, int W, int Nfft, int64_t n_out,
                                                              cx<T>* __restrict__ ac_out, uint32_t* __restrict__ flags, int64_t flag_words) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int64_t b = blockIdx.x;
    const int64_t n0 = (int64_t)blockIdx.y * AC_TILE;
    const int tile = (int)min((int64_t)AC_TILE, n_out - n0);
    const int M = tile + W - 1;
    double* Sre = (double*)smem_raw;
    double* Sim = Sre + (AC_TILE + W);
    double* Sa = Sim + (AC_TILE + W);
    double* Sb = Sa + (AC_TILE + W);
    __shared__ double wtot[4][AC_THREADS / 32];
    __shared__ double tot[4][AC_THREADS];
    const cx<T>* r = rx + b * L;
    const int tid = threadIdx.x;
    for (int i = tid; i < M; i += AC_THREADS) {
        cx<T> x = r[n0 + i], y = r[n0 + i + Nfft];
        double xr = x.x, xi = x.y, yr = y.x, yi = y.y;
        Sre[i] = xr * yr + xi * yi;     // x * conj(y)
        Sim[i] = xi * yr - xr * yi;
        Sa[i] = xr * xr + xi * xi;
        Sb[i] = yr * yr + yi * yi;
    }
    __syncthreads();
    // chunked inclusive scan: odd chunk length keeps the strided shared accesses conflict-free
    int CH = (M + AC_THREADS - 1) / AC_THREADS;
    CH |= 1;
    const int lo = min(tid * CH, M), hi = min(lo + CH, M);
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int i = lo; i < hi; ++i) {
        t0 += Sre[i]; Sre[i] = t0;
        t1 += Sim[i]; Sim[i] = t1;
        t2 += Sa[i]; Sa[i] = t2;
        t3 += Sb[i]; Sb[i] = t3;
    }
    // Offsets of the chunks.  They must come from ONE fixed summation order: a run of exactly-zero
    // samples (STO zero fill) then leaves the prefix bit-identical on both sides of a window, so the
    // window sums are exactly 0 and 0/0 gives the reference's NaN (`AutoCorrFunction.m:6`).
    const int lane = tid & 31, w = tid >> 5;
    tot[0][tid] = t0; tot[1][tid] = t1; tot[2][tid] = t2; tot[3][tid] = t3;
    __syncthreads();
    if (tid < 4 * (AC_THREADS / 32)) {            // one thread per (quantity, warp): sequential lane scan
        const int q = tid & 3, ww = tid >> 2;
        double run = 0;
        for (int l = 0; l < 32; ++l) { double x = tot[q][ww * 32 + l]; tot[q][ww * 32 + l] = run; run += x; }
        wtot[q][ww] = run;
    }
    __syncthreads();
    double off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double base = 0;
        for (int k = 0; k < w; ++k) base += wtot[q][k];
        off[q] = base + tot[q][tid];
    }
    for (int i = lo; i < hi; ++i) { Sre[i] += off[0]; Sim[i] += off[1]; Sa[i] += off[2]; Sb[i] += off[3]; }
    __syncthreads();
    for (int j0 = 0; j0 < AC_TILE; j0 += AC_THREADS) {
        const int j = j0 + tid;
        bool flag = false;
        if (j < tile) {
            double nr = Sre[j + W - 1], ni = Sim[j + W - 1], pa = Sa[j + W - 1], pb = Sb[j + W - 1];
            if (j > 0) { nr -= Sre[j - 1]; ni -= Sim[j - 1]; pa -= Sa[j - 1]; pb -= Sb[j - 1]; }
            double den = sqrt(pa * pb);
            double ar = nr / den, ai = ni / den;   // 0/0 -> NaN as in MATLAB
            const int64_t n = n0 + j;
            if (ac_out) ac_out[b * n_out + n] = mk<T>((T)ar, (T)ai);
            double amp = sqrt(ar * ar + ai * ai);
            flag = (amp > 0.77) && (n + 1 > (int64_t)W);   // :10-13 (NaN compares false)
        }
        unsigned bal = __ballot_sync(0xffffffffu, flag);
        const int base = j0 + (tid & ~31);   // first output of this warp's 32-wide word (n0 is a multiple of 32)
        if (lane == 0 && base < tile) flags[b * flag_words + ((n0 + base) >> 5)] = bal;
    }
}

template <typename T>
__global__ void autocorr_detect_kernel(const cx<T>* __restrict__ rx, int64_t B, int64_t L, int W, int Nfft, int64_t n_out,
                                       const uint32_t* __restrict__ flags, int64_t flag_words, int32_t* __restrict__ tg_pos,
                                       double* __restrict__ freq_off, int32_t* __restrict__ fail) {
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const uint32_t* fw = flags + b * flag_words;
    const int BIG = 0x7fffffff;
    // first flagged index
    int f = BIG;
    for (int64_t wi = lane; wi < flag_words; wi += 32) { uint32_t x = fw[wi]; if (x) { f = (int)(wi * 32 + __ffs(x) - 1); break; } }
    f = warp_min(f);
    int e = BIG, g = BIG;
    if (f != BIG) {
        // first un-flagged index after f (positions >= n_out count as un-flagged)
        for (int64_t wi = (f >> 5) + ((lane - ((f >> 5) & 31)) & 31); wi < flag_words; wi += 32) {
            uint32_t x = ~fw[wi];
            if (wi == (f >> 5)) x &= ~((2u << (f & 31)) - 1u);   // only positions > f
            if (x) { e = (int)(wi * 32 + __ffs(x) - 1); break; }
        }
        e = warp_min(e);
        if (e == BIG || e > n_out) e = (int)n_out;
        // any flagged index after e -> a second run exists (`result(2)` is defined)
        for (int64_t wi = (e >> 5) + ((lane - ((e >> 5) & 31)) & 31); wi < flag_words; wi += 32) {
            uint32_t x = fw[wi];
            if (wi == (e >> 5)) x &= ~((1u << (e & 31)) - 1u);   // positions >= e (e itself is clear)
            if (x) { g = (int)(wi * 32 + __ffs(x) - 1); break; }
        }
        g = warp_min(g);
    }
    int tg = 65, failed = 1;
    if (f != BIG && g != BIG) { tg = (f + 1 + e) / 2; failed = 0; }   // floor((first+last)/2), 1-based
    // rho(TgPosition) summed directly in double
    double nr = 0, ni = 0, pa = 0, pb = 0;
    const int64_t n = tg - 1;
    if (n < n_out) {
        const cx<T>* r = rx + b * L;
        for (int k = lane; k < W; k += 32) {
            cx<T> x = r[n + k], y = r[n + k + Nfft];
            double xr = x.x, xi = x.y, yr = y.x, yi = y.y;
            nr += xr * yr + xi * yi; ni += xi * yr - xr * yi; pa += xr * xr + xi * xi; pb += yr * yr + yi * yi;
        }
        nr = warp_sum(nr); ni = warp_sum(ni); pa = warp_sum(pa); pb = warp_sum(pb);
    } else { nr = ni = CUDART_NAN; }
    if (lane == 0) {
        double den = sqrt(pa * pb);
        tg_pos[b] = tg;
        freq_off[b] = -atan2(ni / den, nr / den) / (2.0 * CUDART_PI);
        if (fail) fail[b] = failed;
    }
}

extern "C" int ofdm_cp_autocorr(ofdm_ctx* ctx, const void* rx, int64_t B, int64_t L, int W, int Nfft, void* autocorr,
                                int32_t* tg_pos, double* freq_off, int32_t* fail) {
    if (!ctx) return OFDM_ERR_INVALID;
    REQUIRE(ctx, rx && tg_pos && freq_off && B >= 0 && W > 0 && Nfft > 0, "bad argument");
    const int64_t n_out = L - W - Nfft;
    REQUIRE(ctx, n_out >= 65, "stream too short for AutoCorrFunction (needs L-W-Nfft >= 65)");
    REQUIRE(ctx, W <= 2048, "window wider than 2048 samples");
    if (B == 0) return OFDM_OK;
    const int64_t flag_words = (n_out + 31) / 32;
    uint32_t* flags = (uint32_t*)ctx_scratch(ctx, sizeof(uint32_t) * B * flag_words);
    REQUIRE(ctx, flags != nullptr, "scratch allocation failed");
    const int tiles = (int)cdiv64(n_out, AC_TILE);
    size_t smem = 4 * sizeof(double) * (AC_TILE + W);
    DISPATCH_T(ctx, {
        auto k1 = autocorr_kernel<T>;
        if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1<<<dim3((unsigned)B, tiles), AC_THREADS, smem, ctx->stream>>>((const cx<T>*)rx, L, W, Nfft, n_out, (cx<T>*)autocorr, flags, flag_words);
        ctx->launches++;
        autocorr_detect_kernel<T><<<(unsigned)cdiv64(B * 32, 128), 128, 0, ctx->stream>>>((const cx<T>*)rx, B, L, W, Nfft, n_out, flags, flag_words,
                                                                                            tg_pos, freq_off, fail);
    });
    LAUNCH_CHECK(ctx);
    return OFDM_OK;
}

// =====================================================================================
// remove_IFO (`Task 5/remove_IFO.m:1-11`): first bin of abs(fft(rx(Nfft+1:2*Nfft))) above 0.77,
// then add_CFO(rx, -IFO, Nfft) with the exact integer phase (IFO*n mod Nfft) from the twiddle table.
// =====================================================================================
template <typename T>
__global__ void ifo_detect_kernel(const cx<T>* __restrict__ rx, int64_t L, int N, int logN, const cx<T>* __restrict__ tw,
                                  int32_t* __restrict__ ifo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int red[32];
    using C = cx<T>;
    C* a = (C*)smem_raw;
    C* bb = a + N;
    const int64_t b = blockIdx.x;
    const C* src = rx + b * L + N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) a[i] = src[i];
    __syncthreads();
    C* r = block_fft<T, false>(a, bb, N, logN, tw);
    int first = 0x7fffffff;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double re = r[i].x, im = r[i].y;
        if (sqrt(re * re + im * im) > 0.77) { first = i; break; }   // strided ascending: first hit per thread is its minimum
    }
    first = block_min(first, red);
    if (threadIdx.x == 0) ifo[b] = (first == 0x7fffffff) ? -1 : first;
}
template <typename T>
__global__ void ifo_derotate_kernel(const cx<T>* __restrict__ rx, int64_t B, int64_t L, int N, const cx<T>* __restrict__ tw,
                                    const int32_t* __restrict__ ifo, cx<T>* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * L) return;
    int64_t b = i / L, n = i - b * L;
    int k = ifo[b];
    cx<T> v = rx[i];
    if (k > 0) v = cmul(v, tw[(int)(((int64_t)k * n) & (N - 1))]);   // exp(-2j*pi*IFO*n/Nfft)
    out[i] = v;
}
extern "C" int ofdm_remove_ifo(ofdm_ctx* ctx, const void* rx, int64_t B, int64_t L, int Nfft, void* out, int32_t* ifo) {
    if (!ctx) return OFDM_ERR_INVALID;
    REQUIRE(ctx, rx && out && ifo && B >= 0, "bad argument");
    REQUIRE(ctx, is_pow2(Nfft) && Nfft >= 8 && Nfft <= (ctx->precision == OFDM_PREC_F64 ? 4096 : 8192), "unsupported Nfft");
    REQUIRE(ctx, L >= 2 * (int64_t)Nfft, "stream shorter than 2*Nfft");
    if (B == 0) return OFDM_OK;
    const void* tw = ctx_twiddles(ctx, Nfft);
    REQUIRE(ctx, tw != nullptr, "twiddle allocation failed");
    DISPATCH_T(ctx, {
        size_t smem = 2 * (size_t)Nfft * sizeof(cx<T>);
        auto k1 = ifo_detect_kernel<T>;
        if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1<<<(unsigned)B, fft_threads(Nfft), smem, ctx->stream>>>((const cx<T>*)rx, L, Nfft, ilog2(Nfft), (const cx<T>*)tw, ifo);
        ctx->launches++;
        ifo_derotate_kernel<T><<<(unsigned)cdiv64(B * L, 256), 256, 0, ctx->stream>>>((const cx<T>*)rx, B, L, Nfft, (const cx<T>*)tw, ifo, (cx<T>*)out);
    });
    LAUNCH_CHECK(ctx);
    return OFDM_OK;
}

// =====================================================================================
// fine_sync (`Task 4/fine_sync.m:1-60`).  Estimators run in double on the (T-typed) pilots.
// =====================================================================================
#define FS_THREADS 256
template <typename T>
__global__ void __launch_bounds__(FS_THREADS) fine_sync_est_kernel(const cx<T>* __restrict__ grid, int S, int Nfft, const int32_t* __restrict__ pc0,
                                                                   int Np, const double2* __restrict__ txp /* Np x S col-major */, int time_desync,
                                                                   double* __restrict__ tau_out, double* __restrict__ phase_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* taus = (double*)smem_raw;       // M-1 entries
    __shared__ int cnt[FS_THREADS];
    __shared__ double red[32];
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const int M = Np * S;
    const cx<T>* g = grid + b * (int64_t)S * Nfft;
    const double deltak = (double)(pc0[1] - pc0[0]);                     // :6
    auto q_at = [&](int i) -> double2 {                                   // tx * conj(rx), column-major flat index
        int s = i / Np, p = i - s * Np;
        double2 rxv = to_d(g[(int64_t)s * Nfft + pc0[p]]);
        return cmulc(txp[i], rxv);
    };
    for (int j = tid; j < M - 1; j += FS_THREADS) {                        // taus(j+1) in MATLAB terms (:25-30)
        double2 q1 = q_at(j + 1), q0 = q_at(j);
        double2 d = cmulc(q1, q0);
        taus[j] = atan2(d.y, d.x) / (2.0 * CUDART_PI * deltak);
    }
    __syncthreads();
    // mask = [false, abs(diffs)<1e-3 & abs(diffs)~=0]; taus_result = taus(mask); mean(taus_result(Np+1:end))
    const int n = M - 1;
    const int CH = (n + FS_THREADS - 1) / FS_THREADS;
    const int lo = min(tid * CH, n), hi = min(lo + CH, n);
    int c = 0;
    for (int j = max(lo, 1); j < hi; ++j) { double d = fabs(taus[j] - taus[j - 1]); c += (d < 1e-3 && d != 0.0); }
    cnt[tid] = c;
    __syncthreads();
    int rank = 0;
    for (int k = 0; k < tid; ++k) rank += cnt[k];
    double sum = 0; int kept = 0;
    for (int j = max(lo, 1); j < hi; ++j) {
        double d = fabs(taus[j] - taus[j - 1]);
        if (d < 1e-3 && d != 0.0) { if (rank >= Np) { sum += taus[j]; ++kept; } ++rank; }
    }
    sum = block_sum(sum, red);
    double nk = block_sum((double)kept, red);
    const double tau = sum / nk;                                            // empty -> NaN like mean([])
    // common phase after the (optional) timing correction (:47-52)
    double ps = 0; int pn = 0;
    for (int i = tid; i < M; i += FS_THREADS) {
        int s = i / Np, p = i - s * Np;
        double2 rxv = to_d(g[(int64_t)s * Nfft + pc0[p]]);
        if (time_desync) { double sn, cs; sincospi(2.0 * tau * (double)pc0[p], &sn, &cs); rxv = cmul(rxv, make_double2(cs, sn)); }
        double2 qq = cmulc(txp[i], rxv);
        double a = atan2(qq.y, qq.x);
        if (fabs(a) > 1e-3) { ps += a; ++pn; }
    }
    ps = block_sum(ps, red);
    double pk = block_sum((double)pn, red);
    if (tid == 0) { tau_out[b] = tau; phase_out[b] = ps / pk; }
}
template <typename T>
__global__ void fine_sync_apply_kernel(const cx<T>* __restrict__ grid, int64_t B, int S, int Nfft, const double* __restrict__ tau,
                                       const double* __restrict__ phase, int time_desync, int freq_desync, cx<T>* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * S * (int64_t)Nfft) return;
    int n = (int)(i % Nfft);
    int64_t b = i / ((int64_t)S * Nfft);
    double2 v = to_d(grid[i]);
    if (time_desync) { double sn, cs; sincospi(2.0 * tau[b] * (double)n, &sn, &cs); v = cmul(v, make_double2(cs, sn)); }  // nn_exp' = exp(+2j*pi*tau*n)
    if (freq_desync) { double sn, cs; sincos(phase[b], &sn, &cs); v = cmul(v, make_double2(cs, sn)); }
    out[i] = from_d<T>(v);
}

const void* ofdm_upload_pilots(ofdm_ctx* ctx, const double* pv, size_t n_complex);

extern "C" int ofdm_fine_sync(ofdm_ctx* ctx, const void* grid, int64_t B, int S, int Nfft, const int32_t* pc, int Np, const double* pv,
                              int time_desync, int freq_desync, void* out, double* tau_dev, double* phase_dev) {
    if (!ctx) return OFDM_ERR_INVALID;
    REQUIRE(ctx, grid && out && pc && pv && B >= 0 && S > 0 && Np >= 2, "bad argument");
    if (B == 0) return OFDM_OK;
    std::vector<int32_t> p0(Np);
    for (int i = 0; i < Np; ++i) { REQUIRE(ctx, pc[i] >= 1 && pc[i] <= Nfft, "pilot index out of range"); p0[i] = pc[i] - 1; }
    const int32_t* pc0 = (const int32_t*)ctx_blob(ctx, p0.data(), sizeof(int32_t) * Np);
    const double2* txp = (const double2*)ctx_blob(ctx, pv, sizeof(double) * 2 * (size_t)Np * S);
    REQUIRE(ctx, pc0 && txp, "device upload failed");
    double* est = nullptr;
    if (!tau_dev || !phase_dev) {
        est = (double*)ctx_scratch(ctx, sizeof(double) * 2 * B);
        REQUIRE(ctx, est != nullptr, "scratch allocation failed");
        if (!tau_dev) tau_dev = est;
        if (!phase_dev) phase_dev = est + B;
    }
    size_t smem = sizeof(double) * (size_t)(Np * S);
    DISPATCH_T(ctx, {
        auto k1 = fine_sync_est_kernel<T>;
        if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1<<<(unsigned)B, FS_THREADS, smem, ctx->stream>>>((const cx<T>*)grid, S, Nfft, pc0, Np, txp, time_desync, tau_dev, phase_dev);
        ctx->launches++;
        int64_t n = B * S * (int64_t)Nfft;
        fine_sync_apply_kernel<T><<<(unsigned)cdiv64(n, 256), 256, 0, ctx->stream>>>((const cx<T>*)grid, B, S, Nfft, tau_dev, phase_dev, time_desync,
                                                                                       freq_desync, (cx<T>*)out);
    });
    LAUNCH_CHECK(ctx);
    return OFDM_OK;
}
