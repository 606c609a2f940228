// Scrambler / DeScrambler / mapping / demapping / BER_func / MER_func on packed bitstreams.
#include "common.cuh"

// Register cell m (1-based) holds s[-m]; as a "previous 32-bit chunk" s[-m] sits at bit 32-m.
static uint32_t reg_to_prev(const uint8_t* reg) {
    uint32_t p = 0;
    for (int m = 1; m <= 15; ++m) if (reg[m - 1] & 1) p |= 1u << (32 - m);
    return p;
}

// ---- Scrambler (`Task 5/Scrambler.m:7-15,20-21`): s[i] = in[i] ^ s[i-13] ^ s[i-14].
// Word-recurrence form: over a 32-bit chunk, s = t*(1+p)(1+p^2) mod x^32 with p = x^13 + x^14
// (p^4 has degree >= 32), t = in ^ carry-in from the previous chunk.  One thread walks one frame;
// a warp covers 32 frames so consecutive chunk loads of a frame hit the same L1 lines.
__global__ void scramble_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n_frames,
                                int64_t frame_bits, uint32_t prev0, uint8_t* __restrict__ final_regs) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const int64_t total = n_frames * frame_bits;
    const int64_t base = f * frame_bits;
    uint32_t prev = prev0, prev2 = 0;
    for (int64_t c = 0; c < frame_bits; c += 32) {
        int n = (int)min((int64_t)32, frame_bits - c);
        uint32_t t = bits_get32(in, base + c, total);
        t ^= (prev >> 19) ^ (prev >> 18);
        uint32_t u = t ^ (t << 13) ^ (t << 14);
        uint32_t s = u ^ (u << 26) ^ (u << 28);
        if (n < 32) s &= (1u << n) - 1u;
        bits_put(out, base + c, n, s);
        prev2 = prev;
        prev = s;
        if (n < 32) {  // keep `prev` meaning "the last 32 bits ending at the frame end"
            prev = (s << (32 - n)) | (prev2 >> n);
        }
    }
    if (final_regs) {  // Register(m) = s[L-m]
        for (int m = 1; m <= 15; ++m) final_regs[f * 15 + (m - 1)] = (prev >> (32 - m)) & 1u;
    }
}

// ---- DeScrambler (`Task 5/DeScrambler.m:7-14`): out[i] = in[i] ^ in[i-13] ^ in[i-14]; one
// thread per 32-bit chunk of a frame.
__global__ void descramble_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n_frames,
                                  int64_t frame_bits, uint32_t prev0, int64_t chunks_per_frame) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_frames * chunks_per_frame) return;
    int64_t f = gid / chunks_per_frame, c = (gid - f * chunks_per_frame) * 32;
    const int64_t total = n_frames * frame_bits;
    const int64_t base = f * frame_bits;
    int n = (int)min((int64_t)32, frame_bits - c);
    uint32_t cur = bits_get32(in, base + c, min(total, base + frame_bits));
    uint32_t prev = (c == 0) ? prev0 : bits_get32(in, base + c - 32, total);
    uint32_t o = cur ^ ((cur << 13) | (prev >> 19)) ^ ((cur << 14) | (prev >> 18));
    bits_put(out, base + c, n, o);
}

__global__ void descramble_final_regs_kernel(const uint32_t* __restrict__ in, int64_t n_frames, int64_t frame_bits,
                                             uint32_t prev0, uint8_t* __restrict__ final_regs) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const int64_t total = n_frames * frame_bits;
    for (int m = 1; m <= 15; ++m) {  // Register(m) = in[L-m] (feedback is the received bit)
        int64_t i = frame_bits - m;
        uint32_t b;
        if (i >= 0) b = bits_get32(in, f * frame_bits + i, total) & 1u;
        else b = (prev0 >> (32 + i)) & 1u;  // s[i], i<0, lives at bit 32+i of prev0
        final_regs[f * 15 + (m - 1)] = (uint8_t)b;
    }
}

static int scr_common(ofdm_ctx* ctx, bool desc, const uint32_t* in, uint32_t* out, int64_t n_frames, int64_t frame_bits,
                      const uint8_t* reg0, uint8_t* final_regs) {
    REQUIRE(ctx, in && out && reg0, "null pointer");
    REQUIRE(ctx, n_frames >= 0 && frame_bits >= 0, "negative size");
    if (n_frames == 0 || frame_bits == 0) return OFDM_OK;
    const int64_t total = n_frames * frame_bits;
    uint32_t prev0 = reg_to_prev(reg0);
    CUDA_TRY(ctx, cudaMemsetAsync(out, 0, sizeof(uint32_t) * OFDM_BIT_WORDS(total), ctx->stream));
    if (!desc) {
        int th = 128;
        scramble_kernel<<<(unsigned)cdiv64(n_frames, th), th, 0, ctx->stream>>>(in, out, n_frames, frame_bits, prev0, final_regs);
        LAUNCH_CHECK(ctx);
    } else {
        int64_t cpf = cdiv64(frame_bits, 32);
        int th = 256;
        descramble_kernel<<<(unsigned)cdiv64(n_frames * cpf, th), th, 0, ctx->stream>>>(in, out, n_frames, frame_bits, prev0, cpf);
        LAUNCH_CHECK(ctx);
        if (final_regs) {
            descramble_final_regs_kernel<<<(unsigned)cdiv64(n_frames, 128), 128, 0, ctx->stream>>>(in, n_frames, frame_bits, prev0, final_regs);
            LAUNCH_CHECK(ctx);
        }
    }
    return OFDM_OK;
}
extern "C" int ofdm_scramble(ofdm_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n_frames, int64_t frame_bits,
                             const uint8_t* reg0, uint8_t* final_regs) {
    if (!ctx) return OFDM_ERR_INVALID;
    return scr_common(ctx, false, in, out, n_frames, frame_bits, reg0, final_regs);
}
extern "C" int ofdm_descramble(ofdm_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n_frames, int64_t frame_bits,
                               const uint8_t* reg0, uint8_t* final_regs) {
    if (!ctx) return OFDM_ERR_INVALID;
    return scr_common(ctx, true, in, out, n_frames, frame_bits, reg0, final_regs);
}

// ---- mapping (`Task 5/mapping.m:14-21`): groups of bps bits, MSB first -> table lookup.
template <typename T>
__global__ void map_kernel(const uint32_t* __restrict__ bits, int64_t n_bits, int64_t n_sym, DevConst<T> c, cx<T>* __restrict__ iq) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sym) return;
    uint32_t v = bits_get32(bits, k * c.bps, n_bits);  // missing tail bits read as 0 == zero padding (:11)
    int idx = 0;
    for (int b = 0; b < c.bps; ++b) idx = (idx << 1) | ((v >> b) & 1u);
    iq[k] = mk<T>(c.re[idx], c.im[idx]);
}

extern "C" int ofdm_map(ofdm_ctx* ctx, const uint32_t* bits, int64_t n_bits, int constellation, void* iq, int* pad) {
    if (!ctx) return OFDM_ERR_INVALID;
    ConstTable ct = host_constellation(constellation);
    REQUIRE(ctx, ct.bps > 0, "unknown constellation");
    REQUIRE(ctx, bits && iq && n_bits >= 0, "bad argument");
    int64_t rem = n_bits % ct.bps;
    if (pad) *pad = rem ? (int)(ct.bps - rem) : -1;
    int64_t n_sym = cdiv64(n_bits, ct.bps);
    if (n_sym == 0) return OFDM_OK;
    DISPATCH_T(ctx, {
        map_kernel<T><<<(unsigned)cdiv64(n_sym, 256), 256, 0, ctx->stream>>>(bits, n_bits, n_sym, make_devconst<T>(constellation), (cx<T>*)iq);
    });
    LAUNCH_CHECK(ctx);
    return OFDM_OK;
}

// ---- demapping (`Task 5/demapping.m:7-18`): one thread per output word; symbols that straddle a
// word (8PSK) are decided by both neighbours, which agree because the decision is deterministic.
template <typename T>
__global__ void demap_kernel(const cx<T>* __restrict__ iq, int64_t n_sym, DevConst<T> c, uint32_t* __restrict__ bits,
                             int64_t n_words, T near_eps, unsigned long long* __restrict__ near) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int local_near = 0;
    if (w < n_words) {
        const int64_t n_bits = n_sym * c.bps;
        const int64_t b0 = w * 32, b1 = min(b0 + 32, n_bits);
        uint32_t word = 0;
        for (int64_t k = b0 / c.bps; k * c.bps < b1; ++k) {
            cx<T> v = iq[k];
            T margin;
            int idx = nearest_idx(c, v.x, v.y, &margin);
            if (near && k * c.bps >= b0 && margin < near_eps) ++local_near;  // count a symbol once (by its first bit)
            for (int b = 0; b < c.bps; ++b) {
                int64_t pos = k * c.bps + b;  // bit b of the group is the MSB-first bit (bps-1-b) of idx (`int2bit`)
                if (pos >= b0 && pos < b1 && ((idx >> (c.bps - 1 - b)) & 1)) word |= 1u << (int)(pos - b0);
            }
        }
        bits[w] = word;
    }
    if (near) {
        local_near = warp_sum(local_near);
        if ((threadIdx.x & 31) == 0 && local_near) atomicAdd(near, (unsigned long long)local_near);
    }
}

extern "C" int ofdm_demap(ofdm_ctx* ctx, const void* iq, int64_t n_sym, int constellation, uint32_t* bits, double near_eps,
                          int64_t* near_dev) {
    if (!ctx) return OFDM_ERR_INVALID;
    ConstTable ct = host_constellation(constellation);
    REQUIRE(ctx, ct.bps > 0, "unknown constellation");
    REQUIRE(ctx, iq && bits && n_sym >= 0, "bad argument");
    if (n_sym == 0) return OFDM_OK;
    int64_t n_words = OFDM_BIT_WORDS(n_sym * ct.bps);
    DISPATCH_T(ctx, {
        demap_kernel<T><<<(unsigned)cdiv64(n_words, 128), 128, 0, ctx->stream>>>((const cx<T>*)iq, n_sym, make_devconst<T>(constellation), bits,
                                                                                   n_words, (T)near_eps, (unsigned long long*)near_dev);
    });
    LAUNCH_CHECK(ctx);
    return OFDM_OK;
}

// ---- BER_func (`Task 5/BER_func.m:3`): popcount of the XOR, 128-bit loads where aligned.
__global__ void ber_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n_bits,
                           unsigned long long* __restrict__ counts) {
    const int64_t n_words = (n_bits + 31) >> 5;
    const int64_t n_vec = n_words >> 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int errs = 0;
    const uint4* a4 = (const uint4*)a;
    const uint4* b4 = (const uint4*)b;
    const bool aligned = ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0;
    int64_t done_words = 0;
    if (aligned) {
        int64_t full_vec = (n_bits >> 7);  // vectors made only of complete words
        for (int64_t v = i; v < full_vec && v < n_vec; v += stride) {
            uint4 x = a4[v], y = b4[v];
            errs += __popc(x.x ^ y.x) + __popc(x.y ^ y.y) + __popc(x.z ^ y.z) + __popc(x.w ^ y.w);
        }
        done_words = min(full_vec, n_vec) * 4;
    }
    for (int64_t w = done_words + i; w < n_words; w += stride) {
        uint32_t x = a[w] ^ b[w];
        int64_t rem = n_bits - w * 32;
        if (rem < 32) x &= (1u << (int)rem) - 1u;
        errs += __popc(x);
    }
    errs = warp_sum(errs);
    if ((threadIdx.x & 31) == 0 && errs) atomicAdd(&counts[0], (unsigned long long)errs);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)n_bits);
}

extern "C" int ofdm_ber_count(ofdm_ctx* ctx, const uint32_t* tx, const uint32_t* rx, int64_t n_bits, int64_t* counts) {
    if (!ctx) return OFDM_ERR_INVALID;
    REQUIRE(ctx, tx && rx && counts && n_bits >= 0, "bad argument");
    if (n_bits == 0) return OFDM_OK;
    int64_t n_words = OFDM_BIT_WORDS(n_bits);
    int blocks = (int)std::min<int64_t>(cdiv64(n_words, 256 * 4), (int64_t)ctx->sm_count * 8);
    if (blocks < 1) blocks = 1;
    ber_kernel<<<blocks, 256, 0, ctx->stream>>>(tx, rx, n_bits, (unsigned long long*)counts);
    LAUNCH_CHECK(ctx);
    return OFDM_OK;
}

// ---- MER_func (`Task 5/MER_func.m:7-25`): nearest by abs() with strict '<', double accumulators.
template <typename T>
__global__ void mer_kernel(const cx<T>* __restrict__ iq, int64_t n_sym, DevConst<T> c, double* __restrict__ sums) {
    __shared__ double red[32];
    double s1 = 0, s2 = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_sym; k += (int64_t)gridDim.x * blockDim.x) {
        cx<T> v = iq[k];
        // abs() is monotone in the squared distance; ties resolve to the first index either way
        int idx = nearest_idx(c, v.x, v.y, (T*)nullptr);
        double ir = c.re[idx], ii = c.im[idx];
        double er = ir - (double)v.x, ei = ii - (double)v.y;
        s1 += ir * ir + ii * ii;
        s2 += er * er + ei * ei;
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { atomicAdd(&sums[0], s1); atomicAdd(&sums[1], s2); }
}

extern "C" int ofdm_mer(ofdm_ctx* ctx, const void* iq, int64_t n_sym, int constellation, double* sums) {
    if (!ctx) return OFDM_ERR_INVALID;
    ConstTable ct = host_constellation(constellation);
    REQUIRE(ctx, ct.bps > 0, "unknown constellation");
    REQUIRE(ctx, iq && sums && n_sym >= 0, "bad argument");
    if (n_sym == 0) return OFDM_OK;
    int blocks = (int)std::min<int64_t>(cdiv64(n_sym, 256), (int64_t)ctx->sm_count * 4);
    DISPATCH_T(ctx, {
        mer_kernel<T><<<blocks, 256, 0, ctx->stream>>>((const cx<T>*)iq, n_sym, make_devconst<T>(constellation), sums);
    });
    LAUNCH_CHECK(ctx);
    return OFDM_OK;
}
