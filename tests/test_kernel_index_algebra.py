"""CPU checks of the index algebra the fused kernels rely on (constants are read from the CUDA sources, so an edit
that breaks an invariant fails here before it reaches a GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ofdm-course_b200", "csrc")


def _define(src, name):
    m = re.search(r"#define\s+%s\s+(\d+)" % name, src)
    assert m, name
    return int(m.group(1))


def _read(name):
    return open(os.path.join(CSRC, name)).read()


def test_rx4096_exchange_layout_and_pass_c_reads_are_conflict_free():
    src = _read("chain_rx4096.cu")
    XROW, XGRP, XBUF = _define(src, "XROW"), _define(src, "XGRP"), _define(src, "XBUF")
    assert XBUF == 16 * XROW and XROW >= 15 * XGRP + 16 and XGRP >= 16 and XGRP % 2 == 0      # 16-byte aligned groups of 16 samples
    # pass C: lane (k1c, k2c) reads eight float4 at samples k1c*XROW + k2c*XGRP + 2j; a 128-bit request is served per
    # quarter warp, whose eight lanes hold eight consecutive k2 of one k1: their 16-byte bank groups must differ
    for tid0 in range(0, 256, 8):
        lanes = range(tid0, tid0 + 8)
        for j in range(8):
            groups = set()
            for t in lanes:
                k1c = (t >> 6) + 4 * ((t >> 3) & 3)
                k2c = ((t >> 5) & 1) * 8 + (t & 7)
                byte = 8 * (k1c * XROW + k2c * XGRP + 2 * j)
                assert byte % 16 == 0
                groups.add((byte // 16) % 8)
            assert len(groups) == 8


def test_rx4096_inverse_estimate_slots_match_the_pass_c_owner():
    # Hinv is stored so that thread tid finds 1/H of its carriers kq + 256c at Hinv[tid + 256c] (chain_rx4096.cu, `keep`)
    owner = {}
    for tid in range(256):
        k1c = (tid >> 6) + 4 * ((tid >> 3) & 3)
        k2c = ((tid >> 5) & 1) * 8 + (tid & 7)
        for c in range(4):
            owner[k1c + 16 * k2c + 256 * c] = tid + 256 * c
    assert sorted(owner) == list(range(1024))
    for q in range(1024):
        k1, k2 = q & 15, (q >> 4) & 15
        slot = (((k1 & 3) << 6) | ((k2 >> 3) << 5) | ((k1 >> 2) << 3) | (k2 & 7)) + (q & ~255)
        assert slot == owner[q]
    # comb-4 pilots (carriers 0, 4, 8, ... 0-based) fall into whole warps: the ones the kernel lets skip pass C
    pilot_warps = {tid >> 5 for tid in range(256) for c in range(4)
                   if (((tid >> 6) + 4 * ((tid >> 3) & 3)) + 16 * (((tid >> 5) & 1) * 8 + (tid & 7)) + 256 * c) % 4 == 0}
    assert pilot_warps == {0, 1}


def test_tx_scrambler_zero_pad_covers_every_shifted_read():
    src = _read("chain.cu")
    m = re.search(r"#define\s+TXF_PAD\(fw\)\s+\(\(fw\) \+ \(\(fw\) >> (\d+)\) \+ (\d+)\)", src)
    assert m
    sh, add = int(m.group(1)), int(m.group(2))
    for fw in list(range(1, 200)) + [672, 1024, 2048, 3333]:
        pad = fw + (fw >> sh) + add
        for frame_bits in {32 * (fw - 1) + 1, 32 * fw - 7, 32 * fw} - {0}:
            if frame_bits <= 0:
                continue
            s13, s14 = 13, 14
            while s13 < frame_bits:                                   # the doubling loop of tx4096_kernel
                reach = (s14 >> 5) + 1                                # lowest word read: w - q14 - 1 with w = 0
                assert reach <= pad, (fw, frame_bits, s14)
                s13, s14 = 2 * s13, 2 * s14


def test_channel_history_regeneration_fills_exactly_the_history_slots():
    src = _read("channel.cu")
    tile, chunk = _define(src, "CH_TILE"), _define(src, "CH_CHUNK")
    assert tile % 512 == 0 and chunk >= 1 and _define(src, "CH_MAXD") - 1 <= tile
    c0 = 4 * tile
    for D in range(1, 60):
        slots = {}
        q = 0
        while 2 * q < D - 1:                                         # pair c0/2 - 1 - q covers samples c0 - 2q - 2, c0 - 2q - 1
            pr = (c0 >> 1) - 1 - q
            j1 = D - 2 - 2 * q
            slots[j1] = 2 * pr + 1
            if j1 >= 1:
                slots[j1 - 1] = 2 * pr
            q += 1
        assert sorted(slots) == list(range(D - 1))
        assert all(slots[j] == c0 - (D - 1) + j for j in slots)      # slot j holds sample c0 - (D-1) + j


# ---------------------------------------------------------------------------------------------------------------------
# Round-2 kernels: the same algebra restated in NumPy / Python integers and compared with the oracle, so that the
# formulations (not the CUDA) are checked on CPU.
import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
from oracle import chains as OC  # noqa: E402


def _prev0(reg):
    p = 0
    for m in range(1, 16):
        if int(reg[m - 1]) & 1:
            p |= 1 << (32 - m)
    return p


def test_windowed_stream_descrambler_equals_per_frame_descrambler():
    """chain_rx_t4.cu, word-aligned path: out word = ((X ^ X<<13 ^ X<<14) >> 32) on the 64-bit window (previous word : word), with
    the initial register's history spliced in where a frame starts inside the window -- against DeScrambler per frame."""
    rng = np.random.default_rng(5)
    M64 = (1 << 64) - 1
    for frame_bits, frames in ((6640, 10), (64, 7), (21504, 2), (96, 5), (1328, 50)):
        stream_bits = frame_bits * frames
        if stream_bits % 32:
            continue
        raw = rng.integers(0, 2, stream_bits).astype(np.uint8)
        for reg in (O.DEFAULT_REGISTER, rng.integers(0, 2, 15).astype(np.uint8)):
            ref = np.concatenate([O.DeScrambler_fast(reg, raw[f * frame_bits:(f + 1) * frame_bits])[0] for f in range(frames)])
            words = np.packbits(raw, bitorder="little").view(np.uint32)
            prev0 = _prev0(reg)
            out = np.zeros_like(words)
            for w in range(words.size):
                R, P = int(words[w]), int(words[w - 1]) if w else 0
                X = (R << 32) | P
                o = ((X ^ (X << 13) ^ (X << 14)) & M64) >> 32
                fl = (32 * w + 31) // frame_bits
                t = fl * frame_bits - 32 * w
                if t > -14:
                    sh = 32 + t
                    keep = (M64 << sh) & M64
                    hist = ((prev0 << (sh - 32)) if sh >= 32 else (prev0 >> (32 - sh))) & M64
                    Xf = (X & keep) | (hist & ~keep & M64)
                    of = ((Xf ^ (Xf << 13) ^ (Xf << 14)) & M64) >> 32
                    before = ((1 << t) - 1) if t > 0 else 0
                    o = (o & before) | (of & ~before & 0xFFFFFFFF)
                out[w] = o & 0xFFFFFFFF
            got = np.unpackbits(out.view(np.uint8), bitorder="little")[:stream_bits]
            assert np.array_equal(got, ref), (frame_bits, frames)


def test_batch_omp_recurrences_equal_the_reference_omp():
    """sparse_dft.cu: Batch-OMP on the Toeplitz Gram vector with the re-fit in orthogonalised form (T, beta), in float64 NumPy:
    same tap indices, gains and stopping iteration as the oracle's OMP_estimate (pinv re-fit on the explicit residual)."""
    rng = np.random.default_rng(9)
    N, Np = 1024, 96
    for trial in range(6):
        K = 8 if trial < 4 else 5          # noise-free trials: exactly as many iterations as taps (beyond that the residual is rounding noise)
        pil = np.sort(rng.permutation(N // 4)[:Np])                      # 0-based pilot bins
        Ld = N if trial % 2 else N // 4
        A = O.sensing_matrix_dft(pil + 1, N, Ld)
        h = np.zeros(N, dtype=complex)
        taps = rng.permutation(60)[:5]
        h[taps] = (rng.standard_normal(5) + 1j * rng.standard_normal(5)) * np.linspace(1, .3, 5)
        y = np.fft.fft(h)[pil] + (0.03 if trial < 4 else 0.0) * (rng.standard_normal(Np) + 1j * rng.standard_normal(Np))
        H_ref, h_ref, idx_ref = O.OMP_estimate(y, A, N, K, 20)
        # --- the kernel's algebra
        d = np.arange(N)
        g = np.exp(2j * np.pi * np.outer(d, pil) / N).sum(axis=1)        # g[d] = sum_i exp(+2 pi j p_i d / N)
        s = np.zeros(N, dtype=complex)
        s[pil] = y
        alpha = (N * np.fft.ifft(s))[:Ld]                                # A^H y
        cols, T, beta, un2, bvec = [], np.zeros((K, K), dtype=complex), np.zeros(K, dtype=complex), np.zeros(K), np.zeros(K, dtype=complex)
        rr, sel = float(np.vdot(y, y).real), []
        for it in range(K):
            col = int(np.argmax(np.abs(alpha) ** 2))
            sel.append(col)
            if col in cols:
                break
            n = len(cols)
            cols.append(col)
            bvec[n] = np.sum(np.exp(2j * np.pi * pil * col / N) * y)      # conj(a_col) . y
            gcol = np.array([g[(c - col) % N] for c in cols[:n]])         # G[c_i][c_n]
            gam = np.array([np.sum(np.conj(T[:j + 1, j]) * gcol[:j + 1]) / un2[j] for j in range(n)]) if n else np.zeros(0)
            un = g[0].real - np.sum(np.abs(gam) ** 2 * un2[:n])
            for i in range(n):
                T[i, n] = -np.sum(gam[i:n] * T[i, i:n])
            T[n, n] = 1
            bt = np.sum(np.conj(T[:n + 1, n]) * bvec[:n + 1]) / un
            un2[n], beta[n] = un, bt
            dn = abs(bt) ** 2 * un
            stop = it >= 1 and np.sqrt(max(dn, 0)) / np.sqrt(max(rr, 0)) < 1e-2
            rr -= dn
            if stop or it == K - 1:
                break
            for i in range(n + 1):
                alpha = alpha - (bt * T[i, n]) * g[(np.arange(Ld) - cols[i]) % N]
        x = np.array([np.sum(T[i, i:len(cols)] * beta[i:len(cols)]) for i in range(len(cols))])
        assert [c + 1 for c in sel] == list(idx_ref), (trial, sel, idx_ref)
        h_got = np.zeros(N, dtype=complex)
        h_got[cols] = x
        assert np.linalg.norm(h_got - h_ref) / np.linalg.norm(h_ref) < 1e-9


def test_rx4096_dead_row_mask_and_pass_b_order():
    """chain_rx4096.cu: a pass-A row k1 = k mod 16 is dead after symbol 0 iff none of its carriers is a data carrier; the pass-B
    order lists the dead rows first so that they fill whole warps (two rows per warp)."""
    def perm(mask):
        return [i for i in range(16) if (mask >> i) & 1] + [i for i in range(16) if not (mask >> i) & 1]
    for comb, want in ((4, 0x1111), (8, 0x0101), (16, 0x0001), (2, 0x5555), (5, 0), (7, 0), (32, 0)):
        pil, dat = O.pilot_layout_comb(1024, comb)
        dead = 0xFFFF
        for c in dat:
            dead &= ~(1 << ((int(c) - 1) & 15))
        assert dead == want, (comb, hex(dead))
        pm = perm(dead)
        assert sorted(pm) == list(range(16))
        n_dead = bin(dead).count("1")
        for w in range(8):                                   # warp w owns rows pm[2w], pm[2w+1]; it idles iff both are dead
            idle = 2 * w + 1 < n_dead
            assert idle == all((dead >> r) & 1 for r in pm[2 * w:2 * w + 2])
    src = _read("chain_rx4096.cu")
    assert "mask_perm(MASK, tid >> 4)" in src and "2 * (tid >> 5) + 1 < mask_popc(MASK)" in src


def test_batch_omp_radix16_stockham_passes_and_swizzle():
    """`od_ifft4096` (sparse_dft.cu): three radix-16 Stockham passes on the conjugate with the pass-1 stores XOR-swizzled
    inside groups of 16; a numpy model with the kernel's index expressions must give the unnormalised inverse DFT, the
    swizzled stores / loads must be bank-conflict free per half warp, and pass 3 must leave thread t with l = t + 256 r."""
    import numpy as np
    src = _read("sparse_dft.cu")
    assert "fb[16 * tid + (r ^ (tid & 15))]" in src and "fb[(tid + 256 * q) ^ sw]" in src and "fa + 16 * tid - 15 * k" in src
    N = 4096
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    tw = np.exp(-2j * np.pi * np.arange(N) / N)

    def twiddled(v, t):
        wb = [1, tw[t], tw[2 * t], tw[3 * t]]
        wa = [1, tw[4 * t], tw[8 * t], tw[12 * t]]
        return np.array([v[q] * wb[q & 3] * wa[q >> 2] for q in range(16)])

    fa, fb = np.conj(x).copy(), np.zeros(N, complex)
    for t in range(256):
        V = np.fft.fft(fa[t + 256 * np.arange(16)])
        for r in range(16):
            fb[16 * t + (r ^ (t & 15))] = V[r]
    fa2 = np.zeros(N, complex)
    for t in range(256):
        k, sw = t & 15, (t >> 4) & 15
        V = np.fft.fft(twiddled(np.array([fb[(t + 256 * q) ^ sw] for q in range(16)]), 16 * k))
        for r in range(16):
            fa2[16 * t - 15 * k + 16 * r] = V[r]
    out = np.zeros(N, complex)
    for t in range(256):
        V = np.fft.fft(twiddled(fa2[t + 256 * np.arange(16)], t))
        out[t + 256 * np.arange(16)] = np.conj(V)                   # thread t owns columns t + 256 r
    assert np.max(np.abs(out - np.fft.ifft(x) * N)) < 1e-9 * N
    # 8-byte accesses are served per half warp: sixteen lanes must hit sixteen different 8-byte bank pairs
    for half in range(0, 256, 16):
        lanes = range(half, half + 16)
        for r in range(16):
            assert len({(16 * t + (r ^ (t & 15))) % 16 for t in lanes}) == 16          # pass-1 stores
        for q in range(16):
            assert len({((t + 256 * q) ^ ((t >> 4) & 15)) % 16 for t in lanes}) == 16  # pass-2 loads
            assert len({(16 * t - 15 * (t & 15) + 16 * q) % 16 for t in lanes}) == 16  # pass-2 stores


def test_batch_omp_redux_keys_order_like_the_floats():
    """`od_top2_redux`: |.|^2 >= 0, so bits + 1 is an order-preserving unsigned key, 0 marks "nothing valid" (NaN or -inf
    never wins); the (max key, min index, max of the rest) triple must equal the shuffle-and-merge top-2 it replaced."""
    import numpy as np
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.random(200).astype(np.float32), np.float32([0.0, 0.0, 1e-38, 3.4e38, np.inf])])
    key = lambda v: 0 if not (v >= 0) else int(np.float32(v).view(np.uint32)) + 1
    order = np.argsort(vals, kind="stable")
    keys = np.array([key(v) for v in vals], dtype=np.uint64)
    assert np.all(np.diff(keys[order]) >= 0) and key(np.float32(np.nan)) == 0 and key(np.float32(-np.inf)) == 0
    for trial in range(200):
        m = rng.random(64).astype(np.float32)
        if trial % 3 == 0:
            m[rng.integers(0, 64, 3)] = m.max()                    # exact ties: the smallest index wins, the runner-up equals the best
        if trial % 5 == 0:
            m[rng.integers(0, 64, 5)] = np.nan
        k = np.array([key(v) for v in m], dtype=np.int64)
        K1 = k.max(); I1 = int(np.min(np.where(k == K1)[0])); K2 = int(np.max(np.where(np.arange(64) == I1, 0, k)))
        valid = ~np.isnan(m)
        best = np.nanmax(m); i_best = int(np.min(np.where(valid & (m == best))[0]))
        rest = np.where((np.arange(64) != i_best) & valid, m, -np.inf).max()
        assert I1 == i_best and K1 == key(best) and K2 == key(rest)


def test_task4_channel_pair_staging_covers_every_slot_once():
    """`channel_t5_kernel<T, true>`: Philox pairs are indexed by the SOURCE sample, so a tile of CH_TILE slots starting at
    source index qb = n0 + sto is covered by pairs P0 + j, j = 0 .. CH_TILE / 2, slot l = 2 j - par + e.  Every slot must be
    written exactly once, with the source index and the (m & 255) / (m >> 8) rotation parts the kernel assumes."""
    src = _read("channel.cu")
    T = _define(src, "CH_TILE")
    assert "rq_s[(l >> 8) + 1]" in src and "const int la = 2 * (int)threadIdx.x - par;" in src
    for n0 in (0, T, 5 * T):
        for sto in (0, 1, 2, 611, -3, -40, 1153):
            qb = n0 + sto
            par = qb & 1
            P0 = (qb - par) >> 1
            seen = {}
            for i in range(T // 512 + 1):
                for tid in range(256):
                    j = tid + 256 * i
                    if j > T // 2:
                        continue
                    la = 2 * tid - par
                    for e in (0, 1):
                        l = la + e + 512 * i
                        if 0 <= l < T:
                            assert l not in seen
                            seen[l] = 2 * (P0 + j) + e
                            m = n0 + l
                            assert (m & 255) == ((la + e) & 255) and (m >> 8) == (n0 >> 8) + (l >> 8)
            assert sorted(seen) == list(range(T))
            assert all(seen[l] == n0 + l + sto for l in seen)          # slot l holds source sample m + sto


def test_two_level_cfo_rotation_is_the_direct_one_to_double_rounding():
    """`cfo_rot_apply` (channel.cu): rot(n) = R(256 (n >> 8)) * R(n & 255) in double, R(k) = exp(2j pi frac(CFO k / Nfft))."""
    import numpy as np
    for cfo in (0.24, -0.5, 3.3, 30.4, 12.5):
        n = np.arange(0, 57600, 7)
        R = lambda k: np.exp(2j * np.pi * np.mod(cfo * k / 1024.0, 1.0))
        two = R(256 * (n >> 8)) * R(n & 255)
        assert np.max(np.abs(two - np.exp(2j * np.pi * cfo * n / 1024.0))) < 2e-11     # the reference's own phase at n = 57,600 carries 1e-12 of rounding


def test_task4_post_kernel_word_decisions_never_straddle_a_symbol():
    """`t4_post_kernel`: a thread decides the eight consecutive payload symbols of one 32-bit word as two groups of four;
    with Nd % 4 == 0 a group of four stays inside one OFDM symbol and its equaliser taps are Gd[dr .. dr + 3]."""
    src = _read("chain_rx_t4.cu")
    assert "(p.Nd & 3) == 0" in src and "const int dr1 = dr0 + 4 >= p.Nd ? 0 : dr0 + 4;" in src
    for Nd in (332, 4, 8, 100, 1024):
        S = 7
        for w in range(S * Nd // 8):
            i0 = 8 * w
            dr0 = i0 % Nd
            dr1 = 0 if dr0 + 4 >= Nd else dr0 + 4
            for u in range(8):
                dr = (dr0 if u < 4 else dr1) + (u & 3)
                assert dr == (i0 + u) % Nd and dr < Nd


def test_dynamic_stream_claims_cover_every_stream_once_and_never_clobber_a_live_slot():
    """Persistent kernels (`rx4096_kernel` SLIM, `tx4096_kernel`): CTA c starts on stream c; on symbol 0 of every stream its
    thread 0 claims the next one (grid + atomicAdd) into slot `parity of the stream count`, the other threads read that slot
    at the stream end.  A randomised interleaving of CTAs must process every stream exactly once, and a slot must not be
    rewritten between its write and the stream end that reads it."""
    import random
    for fn in ("chain_rx4096.cu", "chain.cu"):
        src = _read(fn)
        assert "atomicAdd(next_stream, 1ull)" in src and "kpar ^= 1" in src
    rnd = random.Random(3)
    for grid, B in ((444, 65536), (296, 300), (7, 7), (5, 23)):
        counter = 0
        done = []
        # per-CTA state: current stream, phase (0 = at symbol 0, 1 = at stream end), slots, parity, slot "live" flags
        ctas = [{"b": c, "phase": 0, "slot": [None, None], "live": [False, False], "k": 0} for c in range(min(grid, B))]
        active = list(range(len(ctas)))
        while active:
            c = rnd.choice(active)
            st = ctas[c]
            if st["phase"] == 0:                      # symbol 0: claim the next stream
                assert not st["live"][st["k"]]        # the slot written now is not awaited by a pending read
                st["slot"][st["k"]] = grid + counter
                st["live"][st["k"]] = True
                counter += 1
                st["phase"] = 1
            else:                                     # stream end: everybody reads the slot
                done.append(st["b"])
                st["b"] = st["slot"][st["k"]]
                st["live"][st["k"]] = False
                st["k"] ^= 1
                st["phase"] = 0
                if st["b"] >= B:
                    active.remove(c)
        assert sorted(done) == list(range(B))
