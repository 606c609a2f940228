#!/usr/bin/env python
"""Tiny invocations of the fused kernels for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ofdm_b200 as G  # noqa: E402
import oracle as O  # noqa: E402
from oracle import chains as OC  # noqa: E402

ctx = G.Context(0, "f32")
rng = np.random.default_rng(0)
# Task-5 shape: TX fast path, fused channel, RX-4096 kernel
p = OC.params_task5(comb=4)
lp = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues)
B = 3
bits = ctx.bits(rng.integers(0, 2, B * p.stream_bits).astype(np.uint8))
h, _ = O.get_MP_channel_resp([[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]], p.Nfft)
tx, psum = ctx.tx_chain(lp, bits, B, want_power=True)
rx = ctx.channel_t5(tx, snr_db=20.0, h_dev=ctx.cplx(h), seed=1, power_sum=psum)
rx_own = ctx.channel_t5(tx, snr_db=20.0, h_dev=ctx.cplx(h), seed=1)
r5 = ctx.rx_chain_t5(lp, rx, B, tx_bits_dev=bits, near_eps=1e-3)
# Task-4 shape: autocorrelation + warp-per-symbol chain
p4 = OC.params_task4()
lp4 = ctx.link_params(p4.Nfft, p4.T_Guard, p4.N_carrier, p4.N_symb, p4.Amount_ODFM_SpF, p4.Constellation, p4.dataCarriers, p4.pilotCarriers, p4.pilotValues)
b4 = rng.integers(0, 2, (2, p4.stream_bits)).astype(np.uint8)
rx4 = []
for i, (sto, cfo) in enumerate([(611, 3.3), (900, 0.24)]):
    t, _, _ = OC.tx_chain(p4, b4[i], fast=True)
    rx4.append(OC.impair_task4(p4, t, SNR_dB=28, Time_Delay=sto, Freq_Shift=cfo, taps=[[0, 1], [4, .6], [10, .3]], rng=rng))
r4 = ctx.rx_chain_t4_fused(lp4, ctx.cplx(np.stack(rx4)), tx_bits_dev=ctx.bits(b4.ravel()), near_eps=1e-3, want_H=True)
# PAPR kernels
w = ctx.window_papr(tx.reshape(B, -1)[:, :3 * 4608].contiguous(), 4096)
xs, cc = ctx.ccdf(w.reshape(-1)[:5000].contiguous())
ctx.sync()
print("T5 counts", r5["counts"].cpu().numpy(), "T4 counts", r4["counts"].cpu().numpy(), "papr windows", tuple(w.shape), "ccdf points", xs.numel())
if os.environ.get("SAN_BIG"):
    # persistent loops of the three-CTA rx4096 kernel (more streams than CTAs), the split Task-4 chain (>= 512 streams),
    # the sweep entry point and the Batch-OMP kernel
    from ofdm_b200 import layouts
    B5 = 460                                             # > 3 x 148 CTAs: some CTAs walk on to a second stream
    lp5 = layouts.task5_link(ctx, comb=4)
    bits5 = torch.randint(-2**31, 2**31 - 1, (B5 * lp5.stream_bits // 32,), dtype=torch.int32, device=ctx.device)
    tx5, ps5 = ctx.tx_chain(lp5, bits5, B5, want_power=True)
    rx5 = ctx.channel_t5(tx5, snr_db=15.0, h_dev=ctx.cplx(h), seed=3, power_sum=ps5)
    big5 = ctx.rx_chain_t5(lp5, rx5, B5, tx_bits_dev=bits5, near_eps=1e-4, want_bits=False, want_H=False)
    Bs = 512
    rx4b = ctx.cplx(np.tile(np.stack(rx4), (Bs // 2, 1)))
    b4b = ctx.bits(np.tile(b4, (Bs // 2, 1)).ravel())
    big4 = ctx.rx_chain_t4_fused(lp4, rx4b, tx_bits_dev=b4b, near_eps=1e-3, want_H=True)
    pil = np.sort(rng.permutation(1024)[:64]) + 1
    yo = (torch.randn(8, 64, device=ctx.device) + 1j * torch.randn(8, 64, device=ctx.device)).to(torch.complex64)
    om = ctx.omp(yo, 1024, 5, Ldict=1024, pilot_loc=pil)
    ctx.sync()
    print("big T5", big5["counts"].cpu().numpy(), "big T4", big4["counts"].cpu().numpy(), "omp", type(om).__name__)
