// OMP_estimate for partial-DFT dictionaries (`Task 5/OMP_estimate.m:2-37` with the sensing matrix of
// `Task 5/Main_model_Task_5.m:182-190` / `Task5_part2.m:181-190`: A(i,l) = exp(-2*pi*1j*p_i*l/Nfft), p_i = pilot bin, 0-based).
//
// Batch-OMP on the structure of that dictionary -- one correlation per frame instead of one per iteration:
//   * alpha0 = A^H y is one inverse FFT of the measurement scattered onto the pilot bins (SURVEY KAT 6);
//   * the Gram matrix of the dictionary is Toeplitz, (A^H A)[l, c] = g[(l - c) mod Nfft] with
//     g[d] = sum_i exp(+2*pi*1j*p_i*d/Nfft): ONE Nfft-long vector replaces the Ldict x Ldict matrix, so the correlation
//     with the residual r = y - A_S x is  A^H r = alpha0 - sum_q x_q g[(l - c_q) mod Nfft]   (|S| terms per column);
//   * the normal equations of the re-fit (`pinv(A_sel)*y`, :9,17) need only g at the pairwise tap distances and
//     alpha0-like right-hand sides; the stopping rule ||r_i - r_{i-1}|| / ||r_{i-1}|| < 1e-2 (:20) follows from the same
//     Gram quantities (all of that in double, g tabulated in double).
// One CTA per frame; alpha0 lives in registers (16 columns per thread), g in shared memory.  Per frame that is one
// 4096-point FFT + 36 x Ldict complex multiply-adds for K = 9, against 9 x Np x Ldict for the explicit correlations.
// The argmax is taken on FP32 values of alpha; frames whose best two |alpha|^2 differ by less than tie_eps (relative) are
// counted per frame in near_ties (their tap ORDER may differ from a float64 evaluation).
#include "pursuit_common.cuh"

#define OD_THREADS 256
#define OD_EPT 16            // dictionary columns per thread: Ldict <= 4096
#define OD_MAXK 16

// g[d] = sum_i exp(+2*pi*1j*p_i*d/N) = sum_i conj(W_N^{(p_i d) mod N}) from the double twiddle table (exact argument reduction)
__global__ void omp_dft_gram_kernel(const int32_t* __restrict__ p0, int Np, int N, const double2* __restrict__ tw_d, double2* __restrict__ g_d,
                                    float2* __restrict__ g_f) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= N) return;
    double sr = 0, si = 0;
    for (int i = 0; i < Np; ++i) { const double2 w = tw_d[(p0[i] * d) & (N - 1)]; sr += w.x; si -= w.y; }
    g_d[d] = make_double2(sr, si);
    g_f[d] = make_float2((float)sr, (float)si);
}

// Is a dense (Np x Ldict, column-major) dictionary the partial-DFT matrix exp(-2*pi*1j*p_i*l/N) for integer p_i?
// Row i: p_i from the phase of column 1, then every entry is compared.  flag[0] is cleared on the first mismatch.
__global__ void omp_dft_probe_kernel(const float2* __restrict__ A, int Np, int Ldict, int N, int32_t* __restrict__ p0, int32_t* __restrict__ flag) {
    const int i = blockIdx.x;
    __shared__ int sp;
    if (threadIdx.x == 0) {
        const float2 a1 = A[(size_t)Np + i];
        const double turns = -atan2((double)a1.y, (double)a1.x) / (2.0 * CUDART_PI);     // p_i / N mod 1
        int p = (int)llrint(turns * N);
        p &= (N - 1);
        sp = p;
        p0[i] = p;
    }
    __syncthreads();
    const int p = sp;
    bool ok = true;
    for (int l = threadIdx.x; l < Ldict; l += blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)((p * l) & (N - 1)) / (double)N, &s, &c);
        const float2 a = A[(size_t)l * Np + i];
        if (!(fabs((double)a.x - c) < 2e-6 && fabs((double)a.y - s) < 2e-6)) ok = false;
    }
    if (!ok) atomicAnd(flag, 0);
}

// (best, index, runner-up) merge for the fused argmax + near-tie reduction: first maximum wins (MATLAB max)
struct Top2 { float b1; int i1; float b2; };
__device__ __forceinline__ Top2 top2_merge(Top2 a, Top2 b) {
    Top2 r;
    const bool a_wins = a.b1 > b.b1 || (a.b1 == b.b1 && a.i1 < b.i1);
    if (a_wins) { r.b1 = a.b1; r.i1 = a.i1; r.b2 = fmaxf(a.b2, b.b1); }
    else { r.b1 = b.b1; r.i1 = b.i1; r.b2 = fmaxf(b.b2, a.b1); }
    return r;
}

// NG = groups of 256 dictionary columns a thread works on (Ldict <= 256 NG).
//
// The re-fit is carried in ORTHOGONALISED form, so that an iteration needs no triangular solve and the correlation is
// updated in place (only the running alpha lives in registers):
//   u_n = a_{c_n} - sum_{j<n} gamma_jn u_j  = sum_{i<=n} T_in a_{c_i}        (Gram-Schmidt on the selected columns)
//   gamma_jn = u_j^H a_{c_n} / ||u_j||^2 = sum_{i<=j} conj(T_ij) G[c_i][c_n] / ||u_j||^2,   G[c_i][c_n] = g[(c_i - c_n) mod N]
//   beta_n = u_n^H y / ||u_n||^2 = sum_{i<=n} conj(T_in) b_i / ||u_n||^2,     b_i = conj(a_{c_i}) . y
//   r_n = r_{n-1} - u_n beta_n   =>   alpha_n[l] = alpha_{n-1}[l] - sum_{i<=n} (beta_n T_in) g[(l - c_i) mod N]
//   ||r_n - r_{n-1}|| = |beta_n| ||u_n||,  ||r_n||^2 = ||r_{n-1}||^2 - |beta_n|^2 ||u_n||^2      (the stopping rule, :20)
//   x = T beta  is exactly pinv(A_sel) * y  (`OMP_estimate.m:9,17`), formed once at the end.
// All of that is O(k) per lane and iteration, in double; the k+1 coefficients beta_n T_in go to the FP32 update.
template <int NG>
__global__ void __launch_bounds__(OD_THREADS, 3) omp_dft_kernel(const float2* __restrict__ Y, int Np, const int32_t* __restrict__ p0, int Ldict, int Nfft,
                                                                int logN, const float2* __restrict__ tw, const double2* __restrict__ tw_d,
                                                                const float2* __restrict__ g_f, const double2* __restrict__ g_d, int K,
                                                                float2* __restrict__ Hout, float2* __restrict__ hout, int32_t* __restrict__ index_out,
                                                                int32_t* __restrict__ iters_out, int32_t* __restrict__ near_out, float tie_eps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[2 * (OD_THREADS / 32)];
    __shared__ Top2 stop2[OD_THREADS / 32];
    __shared__ int sel[OD_MAXK], ucol[OD_MAXK];
    __shared__ double2 Tm[OD_MAXK][OD_MAXK + 1];      // T[i][n], i <= n
    __shared__ double2 bvec[OD_MAXK], beta[OD_MAXK];
    __shared__ double un2[OD_MAXK];                    // ||u_j||^2
    __shared__ float2 wf[OD_MAXK], xs[OD_MAXK];
    __shared__ int s_nu, s_nsel, s_stop, s_near, s_col, s_new;
    __shared__ double s_rr;                            // ||r_{n-1}||^2
    float2* fa = (float2*)smem_raw;
    float2* fb = fa + Nfft;
    const int64_t f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Nmask = Nfft - 1;
    const float2* y = Y + f * Np;

    // ---- alpha0 = A^H y: scatter onto the pilot bins + unnormalised inverse FFT; ||y||^2 in double
    for (int i = tid; i < Nfft; i += OD_THREADS) fa[i] = make_float2(0.f, 0.f);
    __syncthreads();
    double yy = 0;
    for (int i = tid; i < Np; i += OD_THREADS) { const float2 v = y[i]; fa[p0[i]] = v; yy += (double)v.x * v.x + (double)v.y * v.y; }
    if (tid == 0) { s_nu = 0; s_nsel = 0; s_stop = 0; s_near = 0; }
    yy = block_sum(yy, red);          // (barriers inside: the scatter is complete)
    if (tid == 0) s_rr = yy;
    __syncthreads();
    float2* c = block_fft<float, true>(fa, fb, Nfft, logN, tw);
    float2 a[NG];                     // the running correlation A^H r
#pragma unroll
    for (int j = 0; j < NG; ++j) { const int l = tid + OD_THREADS * j; a[j] = l < Ldict ? c[l] : make_float2(0.f, 0.f); }
    __syncthreads();
    // g twice in a row (both FFT buffers are free now): entry (l - c) mod Nfft is read as gS[base + 256 j] with ONE base per
    // selected tap and immediate offsets, no wrap-around arithmetic in the inner loop
    float2* gS = fa;
    for (int i = tid; i < 2 * Nfft; i += OD_THREADS) gS[i] = g_f[i & Nmask];
    __syncthreads();

    for (int it = 0; it < K; ++it) {
        const int nu = s_nu, nsel = s_nsel;
        // ---- first maximum of |A^H r|^2 and the runner-up, one fused block reduction (`OMP_estimate.m:7,14`)
        Top2 t; t.b1 = -CUDART_INF_F; t.i1 = 0x7fffffff; t.b2 = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            const int l = tid + OD_THREADS * j;
            const float m = l < Ldict ? a[j].x * a[j].x + a[j].y * a[j].y : -CUDART_INF_F;
            if (m > t.b1) { t.b2 = t.b1; t.b1 = m; t.i1 = l; } else t.b2 = fmaxf(t.b2, m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Top2 u;
            u.b1 = __shfl_xor_sync(0xffffffffu, t.b1, o); u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o); u.b2 = __shfl_xor_sync(0xffffffffu, t.b2, o);
            t = top2_merge(t, u);
        }
        if (lane == 0) stop2[warp] = t;
        __syncthreads();
        if (warp == 0) {
            t = lane < OD_THREADS / 32 ? stop2[lane] : Top2{-CUDART_INF_F, 0x7fffffff, -CUDART_INF_F};
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                Top2 u;
                u.b1 = __shfl_xor_sync(0xffffffffu, t.b1, o); u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o); u.b2 = __shfl_xor_sync(0xffffffffu, t.b2, o);
                t = top2_merge(t, u);
            }
            if (lane == 0) {
                const int col = (t.i1 == 0x7fffffff) ? 0 : t.i1;          // all-NaN correlation: MATLAB's max returns index 1
                if (!(t.b1 - t.b2 > tie_eps * t.b1)) s_near += 1;
                int slot = -1;
                for (int q = 0; q < nu; ++q) if (ucol[q] == col) slot = q;
                sel[nsel] = col;
                if (slot < 0) { ucol[nu] = col; s_new = 1; s_nu = nu + 1; } else s_new = 0;
                s_nsel = nsel + 1;
                s_col = col;
            }
        }
        __syncthreads();
        const int col = s_col, is_new = s_new;
        if (!is_new) break;               // a column selected twice leaves the residual unchanged: ||r_i - r_{i-1}|| = 0 < 1e-2 (`:20`), it >= 1 always here
        {
            // b_n = conj(a_new) . y in double: one table twiddle per pilot, block-wide
            double ar = 0, ai = 0;
            for (int i = tid; i < Np; i += OD_THREADS) {
                const double2 w = tw_d[(p0[i] * col) & Nmask];             // A(i, col)
                const float2 v = y[i];
                ar += w.x * v.x + w.y * v.y; ai += w.x * v.y - w.y * v.x;  // conj(w) * v
            }
            ar = warp_sum(ar); ai = warp_sum(ai);
            if (lane == 0) { red[2 * warp] = ar; red[2 * warp + 1] = ai; }
            __syncthreads();
        }
        if (warp == 0) {
            const int n = nu;                                              // index of the new column
            double2 bn = make_double2(0, 0);
            for (int w = 0; w < OD_THREADS / 32; ++w) { bn.x += red[2 * w]; bn.y += red[2 * w + 1]; }
            if (lane == 0) bvec[n] = bn;
            // gamma_j (lane j < n) = sum_{i<=j} conj(T_ij) G[c_i][c_n] / ||u_j||^2
            const double2 gcol = lane < n ? g_d[(ucol[lane] - col) & Nmask] : make_double2(0, 0);     // G[c_lane][c_n]
            double2 gam = make_double2(0, 0);
            for (int i = 0; i < n; ++i) {
                const double2 gi = make_double2(__shfl_sync(0xffffffffu, gcol.x, i), __shfl_sync(0xffffffffu, gcol.y, i));
                if (lane < n && i <= lane) gam = gam + cmul(cconj(Tm[i][lane]), gi);
            }
            const double u2 = lane < n ? un2[lane] : 1.0;
            gam = (lane < n && u2 > 0) ? cscale(gam, 1.0 / u2) : make_double2(0, 0);
            // ||u_n||^2 = G_nn - sum_j |gamma_j|^2 ||u_j||^2
            double nn = lane < n ? (gam.x * gam.x + gam.y * gam.y) * u2 : 0.0;
            nn = warp_sum(nn);
            const double un = fmax(g_d[0].x - nn, 0.0);
            // T_in (lane i < n) = - sum_{j=i}^{n-1} gamma_j T_ij ; T_nn = 1
            double2 tin = make_double2(0, 0);
            for (int j = 0; j < n; ++j) {
                const double2 gj = make_double2(__shfl_sync(0xffffffffu, gam.x, j), __shfl_sync(0xffffffffu, gam.y, j));
                if (lane < n && j >= lane) tin = tin - cmul(gj, Tm[lane][j]);
            }
            if (lane == n) tin = make_double2(1, 0);
            if (lane <= n) Tm[lane][n] = tin;
            // beta_n = sum_{i<=n} conj(T_in) b_i / ||u_n||^2
            double2 part = make_double2(0, 0);
            if (lane < n) part = cmul(cconj(tin), bvec[lane]);
            else if (lane == n) part = bn;
            part.x = warp_sum(part.x); part.y = warp_sum(part.y);
            const double2 bt = un > 0 ? cscale(part, 1.0 / un) : make_double2(0, 0);
            if (lane == 0) { un2[n] = un; beta[n] = bt; }
            if (lane <= n) { const double2 w = cmul(bt, tin); wf[lane] = make_float2((float)w.x, (float)w.y); }
            // stopping rule: ||r_n - r_{n-1}|| / ||r_{n-1}|| < 1e-2 from the second selection on
            if (lane == 0) {
                const double dn = (bt.x * bt.x + bt.y * bt.y) * un, on = s_rr;
                if (it >= 1 && sqrt(fmax(dn, 0.0)) / sqrt(fmax(on, 0.0)) < 1e-2) s_stop = 1;
                s_rr = on - dn;
            }
        }
        __syncthreads();
        if (s_stop || it == K - 1) break;
        // ---- alpha -= sum_{i<=n} (beta_n T_in) g[(l - c_i) mod N]
        // (two packed FFMA2 per column: the table value is the broadcast operand, -w and its rotation are hoisted per tap)
        for (int q = 0; q <= nu; ++q) {
            const float2 x = wf[q];
            const float2 w0 = make_float2(-x.x, -x.y), w1 = make_float2(x.y, -x.x);      // a -= g*x = g.x*(-x) + g.y*(-i x)... (x.y, -x.x) = -i*x negated
            const float2* gp = gS + ((tid - ucol[q]) & Nmask);
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                const float2 gv = gp[OD_THREADS * j];
                a[j] = __ffma2_rn(make_float2(gv.x, gv.x), w0, a[j]);
                a[j] = __ffma2_rn(make_float2(gv.y, gv.y), w1, a[j]);
            }
        }
    }
    __syncthreads();
    // ---- gains x = T beta on the unique columns; a repeated last column shares its unknown (pinv's minimum-norm split)
    const int nsel = s_nsel, nu = s_nu;
    __shared__ float2 hval[OD_MAXK];
    if (warp == 0) {
        double2 xv = make_double2(0, 0);
        if (lane < nu) for (int n = lane; n < nu; ++n) xv = xv + cmul(Tm[lane][n], beta[n]);
        const int last = sel[nsel - 1];
        const bool dup = nsel > nu;                      // the last selection repeated column `last`
        if (lane < nu) {
            hval[lane] = (dup && ucol[lane] == last) ? make_float2((float)(0.5 * xv.x), (float)(0.5 * xv.y)) : make_float2((float)xv.x, (float)xv.y);
        }
        __syncwarp();
        if (lane == 0) {
            if (index_out) for (int q = 0; q < K; ++q) index_out[f * K + q] = q < nsel ? sel[q] + 1 : 0;
            if (iters_out) iters_out[f] = nsel;
            if (near_out) near_out[f] = s_near;
        }
    }
    __syncthreads();
    if (hout) {
        float2* hb = hout + f * (int64_t)Nfft;
        for (int i = tid; i < Nfft; i += OD_THREADS) {
            float2 v = make_float2(0.f, 0.f);
            for (int u = 0; u < nu; ++u) if (ucol[u] == i) v = hval[u];
            hb[i] = v;
        }
    }
    if (Hout) {
        // H(m) = sum_u h_u W^{c_u m}; with m = 64 a + b the twiddle is W^{64 c a} * W^{c b}: two 64-entry rows per tap
        float2* t_hi = fb;                       // [nu][64]  h_u * W^{64 c a}
        float2* t_lo = fa;                       // [nu][64]  W^{c b}   (g is no longer needed)
        for (int e = tid; e < nu * 64; e += OD_THREADS) {
            const int u = e >> 6, j = e & 63;
            const float2 v = hval[u];
            const float2 w = tw[(64 * j * ucol[u]) & Nmask];
            t_hi[e] = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
            t_lo[e] = tw[(j * ucol[u]) & Nmask];
        }
        __syncthreads();
        for (int mm = tid; mm < Nfft; mm += OD_THREADS) {
            const int aa = mm >> 6, bq = mm & 63;
            float2 acc = make_float2(0.f, 0.f);
            for (int u = 0; u < nu; ++u) {
                const float2 x = t_hi[64 * u + aa], w = t_lo[64 * u + bq];
                acc = __ffma2_rn(make_float2(x.x, x.x), w, acc);                        // x * w, x the broadcast operand
                acc = __ffma2_rn(make_float2(x.y, x.y), make_float2(-w.y, w.x), acc);
            }
            Hout[f * (int64_t)Nfft + mm] = acc;
        }
    }
}

// Batch-OMP path.  p0_dev: Np 0-based pilot bins on the device.  Returns OFDM_OK and sets *handled when it ran.
int ofdm_omp_dft(ofdm_ctx* ctx, const void* y, int64_t B, int Np, const int32_t* p0_dev, int Ldict, int Nfft, int K, void* H, void* h, int32_t* index,
                 int32_t* iters, int32_t* near_ties, double tie_eps, bool* handled) {
    *handled = false;
    if (ctx->precision != OFDM_PREC_F32 || K > OD_MAXK || Ldict > OD_THREADS * OD_EPT || Nfft > 4096 || Nfft < 1024) return OFDM_OK;
    if (getenv("OFDM_B200_NO_BATCH_OMP")) return OFDM_OK;
    const float2* tw = (const float2*)ctx_twiddles(ctx, Nfft);
    const double2* tw_d = (const double2*)ctx_twiddles_prec(ctx, Nfft, OFDM_PREC_F64);
    REQUIRE(ctx, tw != nullptr && tw_d != nullptr, "twiddle allocation failed");
    unsigned char* gbuf = nullptr;
    CUDA_TRY(ctx, cudaMallocAsync((void**)&gbuf, (sizeof(double2) + sizeof(float2)) * (size_t)Nfft, ctx->stream));
    double2* g_d = (double2*)gbuf;
    float2* g_f = (float2*)(g_d + Nfft);
    omp_dft_gram_kernel<<<(Nfft + 63) / 64, 64, 0, ctx->stream>>>(p0_dev, Np, Nfft, tw_d, g_d, g_f);
    const size_t smem = sizeof(float2) * 2 * (size_t)Nfft;
    auto kern = Ldict <= 1024 ? omp_dft_kernel<4> : (Ldict <= 2048 ? omp_dft_kernel<8> : omp_dft_kernel<16>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)B, OD_THREADS, smem, ctx->stream>>>((const float2*)y, Np, p0_dev, Ldict, Nfft, ilog2(Nfft), tw, tw_d, g_f, g_d, K, (float2*)H, (float2*)h,
                                                         index, iters, near_ties, (float)tie_eps);
    ctx->launches += 2;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(gbuf, ctx->stream);
    if (e != cudaSuccess) return ctx_fail(ctx, OFDM_ERR_CUDA, "Batch-OMP launch failed: %s", cudaGetErrorString(e));
    *handled = true;
    return OFDM_OK;
}

// Dense dictionary: probe for the partial-DFT structure (one 4-byte read-back); on success p0_out (device, Np ints,
// stream-ordered allocation the caller frees) holds the pilot bins.
int ofdm_omp_probe_dft(ofdm_ctx* ctx, const void* A, int Np, int Ldict, int Nfft, int32_t** p0_out, bool* is_dft) {
    *is_dft = false;
    *p0_out = nullptr;
    if (ctx->precision != OFDM_PREC_F32 || Ldict < 2) return OFDM_OK;
    int32_t* buf = nullptr;
    CUDA_TRY(ctx, cudaMallocAsync((void**)&buf, sizeof(int32_t) * ((size_t)Np + 1), ctx->stream));
    int32_t flag = 0;
    CUDA_TRY(ctx, cudaMemsetAsync(buf + Np, 0xFF, sizeof(int32_t), ctx->stream));    // all ones: cleared by the first mismatch
    omp_dft_probe_kernel<<<Np, 128, 0, ctx->stream>>>((const float2*)A, Np, Ldict, Nfft, buf, buf + Np);
    ctx->launches++;
    CUDA_TRY(ctx, cudaMemcpyAsync(&flag, buf + Np, sizeof flag, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag != 0) { *is_dft = true; *p0_out = buf; }
    else cudaFreeAsync(buf, ctx->stream);
    return OFDM_OK;
}
