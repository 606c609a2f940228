#!/usr/bin/env python
"""Round-2 feasibility study (CPU, NumPy): accuracy and flop count of a 4096-point DFT built from small DFT *matrix
products* -- the form tensor cores execute -- with split-precision operands, against the float64 FFT.

The M1 tolerance on post-FFT values is a relative L2 error of 2e-5.  Operands are rounded to the
tensor-core input type (TF32: 10 explicit mantissa bits, BF16: 7, FP16: 10) and split into 1-3 terms
(x = x0 + x1 + ..., each term representable); products of terms are accumulated in FP32, as the MMA does.
Cross terms of combined order >= `terms` are dropped (the usual 3xTF32 / 2xFP16 recipes).

    python tools/dft_gemm_accuracy.py
"""
import itertools

import numpy as np


def round_mantissa(x, bits):
    """Round float32 values to `bits` explicit mantissa bits (round to nearest even), keeping the FP32 exponent range."""
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = (u + half - np.uint64(1) + lsb) >> np.uint64(drop) << np.uint64(drop)
    return u.astype(np.uint32).view(np.float32)


def split(x, bits, terms):
    out, r = [], np.asarray(x, dtype=np.float32)
    for _ in range(terms):
        t = round_mantissa(r, bits)
        out.append(t)
        r = (r - t).astype(np.float32)
    return out


def mm_split(A, B, bits, terms):
    """Real matrix product with split operands, FP32 accumulation, cross terms of order i + j < terms."""
    As, Bs = split(A, bits, terms), split(B, bits, terms)
    acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.float32)
    for i, j in sorted(itertools.product(range(terms), repeat=2), key=lambda ij: -(ij[0] + ij[1])):   # small terms first
        if i + j < terms:
            acc = (acc + (As[i].astype(np.float32) @ Bs[j].astype(np.float32))).astype(np.float32)
    return acc


def cmm(Wr, Wi, Xr, Xi, bits, terms):
    """(Wr + i Wi)(Xr + i Xi) as four real products."""
    return (mm_split(Wr, Xr, bits, terms) - mm_split(Wi, Xi, bits, terms),
            mm_split(Wr, Xi, bits, terms) + mm_split(Wi, Xr, bits, terms))


def dft_by_gemm(x, radices, bits, terms):
    """Decimation-in-time over the digit list `radices` (product = len(x)); every stage is a DFT-matrix product followed
    by an FP32 twiddle multiplication (done on the SIMT side in the real kernel)."""
    N = len(x)
    xr, xi = np.real(x).astype(np.float32), np.imag(x).astype(np.float32)

    def rec(xr, xi, radices):
        n = len(xr)
        if len(radices) == 1:
            R = radices[0]
            k = np.arange(R)
            W = np.exp(-2j * np.pi * np.outer(k, k) / R)
            yr, yi = cmm(W.real.astype(np.float32), W.imag.astype(np.float32), xr.reshape(R, 1), xi.reshape(R, 1), bits, terms)
            return yr.ravel(), yi.ravel()
        R, M = radices[0], n // radices[0]
        # n = R*m + r  ->  X[k1 + M... ] : split input index n = r + R*m (r < R), output k = k2 + M*k1? use standard: x[n1*M + n2]
        Xr, Xi = xr.reshape(R, M), xi.reshape(R, M)                 # rows n1, columns n2 ; n = M*n1 + n2
        k = np.arange(R)
        W = np.exp(-2j * np.pi * np.outer(k, k) / R)
        Yr, Yi = cmm(W.real.astype(np.float32), W.imag.astype(np.float32), Xr, Xi, bits, terms)     # DFT over n1 -> k1 ; [k1, n2]
        tw = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(M)) / n)                          # W_n^{k1*n2}
        twr, twi = tw.real.astype(np.float32), tw.imag.astype(np.float32)
        Zr = (Yr * twr - Yi * twi).astype(np.float32)
        Zi = (Yr * twi + Yi * twr).astype(np.float32)
        outr, outi = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        for k1 in range(R):                                         # remaining DFT over n2 -> k2 ; k = k1 + R*k2
            a, b = rec(Zr[k1], Zi[k1], radices[1:])
            outr[k1::R], outi[k1::R] = a, b
        return outr, outi

    r, i = rec(xr, xi, list(radices))
    return r.astype(np.float64) + 1j * i.astype(np.float64)


def flops_per_symbol(radices, terms, pruned_last=0.25):
    """Real flops of the matrix products per 4096-point symbol: 4 real products per complex one, 2 flops per MAC,
    terms*(terms+1)/2 cross products; the last stage evaluates only `pruned_last` of its outputs (bins < 1024)."""
    N = int(np.prod(radices))
    cross = terms * (terms + 1) // 2
    tot = 0.0
    for s, R in enumerate(radices):
        f = N * R * 4 * 2 * cross
        tot += f * (pruned_last if s == len(radices) - 1 else 1.0)
    return tot


def main():
    rng = np.random.default_rng(0)
    N = 4096
    X = np.zeros(N, dtype=complex)                                   # an OFDM symbol: 1024 occupied carriers of unit-power 16QAM
    lv = np.array([-3, -1, 1, 3]) / np.sqrt(10)
    X[:1024] = rng.choice(lv, 1024) + 1j * rng.choice(lv, 1024)
    x = np.fft.ifft(X) * N / np.sqrt(1024) + 0.1 * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    ref = np.fft.fft(x)
    print(f"{'stages':12s} {'operand type':14s} {'terms':5s} {'rel L2 error':>13s} {'MFLOP/symbol':>13s} {'symbols/s at 2.25 PFLOP/s':>26s}")
    for radices in ((64, 64), (16, 16, 16)):
        for name, bits, terms in (("TF32", 10, 1), ("TF32", 10, 2), ("TF32", 10, 3), ("FP16*", 10, 2), ("BF16", 7, 2), ("BF16", 7, 3)):
            y = dft_by_gemm(x, radices, bits, terms)
            err = np.linalg.norm(y[:1024] - ref[:1024]) / np.linalg.norm(ref[:1024])
            f = flops_per_symbol(radices, terms)
            peak = 2.25e15 if name != "TF32" else 1.125e15
            print(f"{'x'.join(map(str, radices)):12s} {name:14s} {terms:<5d} {err:13.2e} {f / 1e6:13.2f} {peak / f / 1e6:22.0f} M")
    print("(* FP16 modelled by its mantissa width only: operands need a per-symbol scale to stay inside its exponent range.)")
    print("FP32 SIMT radix-16 x 3 kernel today: 110 M symbols/s, ~0.2 MFLOP per symbol.")


if __name__ == "__main__":
    main()
