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
