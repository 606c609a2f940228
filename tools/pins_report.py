#!/usr/bin/env python
"""The reference-held numbers (tests/reference_pins.py) as measured on the GPU path, for DESIGN.md §2:  python tools/pins_report.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ofdm_b200 as G
from oracle import chains as OC
import reference_pins as RP

def lpar(ctx, p, scramble=True):
    return ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues, scramble=scramble)

ctx = G.default_context("f32")
print("P1 BER(SNR): constellation SNR figure gpu ratio")
for cname, pts in RP.BER_SNR_FIGURE.items():
    p = OC.params_task4(alternate=False, Constellation=cname)
    lp = lpar(ctx, p)
    B = 2048
    bits_d = ctx.bits(np.tile(RP.eagle_bits(p.stream_bits), B))
    tx = ctx.tx_chain(lp, bits_d, B)
    for i, (snr, fig) in enumerate(pts.items()):
        rx = ctx.channel_t5(tx, snr_db=float(snr), h_dev=None, seed=100 + i)
        res = ctx.rx_chain_t4_fused(lp, rx, tx_bits_dev=bits_d, time_desync=False, freq_desync=False, mp_desync=False, want_bits=False)
        ctx.sync()
        c = res["counts"].cpu().numpy()
        print(f"  {cname:6s} {snr:3d} dB  {fig:.4g}  {c[0] / c[1]:.4g}  {c[0] / c[1] / fig:.3f}")
p = OC.params_task4(percent=0.25, alternate=False)
lp = lpar(ctx, p)
sums = torch.zeros(2, dtype=torch.float64, device=ctx.device)
bits_d = ctx.bits(np.tile(RP.eagle_bits(p.stream_bits), 2048))
tx = ctx.tx_chain(lp, bits_d, 2048)
for c0 in range(0, 10240, 2048):
    rx = ctx.channel_t5(tx, snr_db=25.0, h_dev=None, seed=7, first_stream_id=c0)
    ctx.mer(ctx.get_payload(ctx.demodulate(rx, p.Nfft, p.T_Guard), p.dataCarriers).reshape(-1), "16QAM", sums)
s = sums.cpu().numpy()
print(f"P2 MER at 25 dB over 10,240 streams: {10 * np.log10(s[0] / s[1]):.4f} dB (reference {RP.MER_AWGN_25DB})")
for prec in ("f64", "f32"):
    c = G.default_context(prec)
    p = OC.params_task4(percent=50)
    lp = lpar(c, p)
    tx = c.tx_chain(lp, c.bits(RP.eagle_bits(p.stream_bits)), 1)
    h, _ = c.mp_channel_resp(RP.TAPS_T4, p.Nfft)
    grid = c.demodulate(c.apply_fir(tx.reshape(1, -1), c.cplx(h)).reshape(1, p.N_symb, -1), p.Nfft, p.T_Guard)
    Hs, Hp = c.estimate_channel(grid, np.arange(1, p.Nfft + 1), p.pilotCarriers, p.pilotValues)
    Hl = c.interpolate(Hp, p.pilotCarriers, p.N_carrier, "linear")
    last = int(p.pilotCarriers[-2]); inner = p.dataCarriers[p.dataCarriers < last]
    def mer(H):
        sm = c.mer(c.get_payload(c.equalize(grid, H, p.N_carrier), inner).reshape(-1), "16QAM").cpu().numpy()
        return 10 * np.log10(sm[0] / sm[1])
    Hc = np.ones(p.N_carrier, dtype=complex)
    Hc[:last] = RP.keys_cubic(p.pilotCarriers[:-1], Hp[0].cpu().numpy()[:-1], np.arange(1.0, last + 1))
    print(f"P3 {prec}: linear {mer(Hl):.2f}  cubic {mer(c.cplx(Hc)[None]):.2f}  spline {mer(Hs):.2f} dB (reference 60 / 108 / 130)")
c = G.default_context("f64")
p = OC.params_task4(percent=1, scale=2.0, alternate=True)
b = c.bits(RP.eagle_bits(p.stream_bits))
print(f"P4 PAPR plain {float(c.papr(c.tx_chain(lpar(c, p, False), b, 1).reshape(1, -1))[0]):.3f} dB, scrambled {float(c.papr(c.tx_chain(lpar(c, p, True), b, 1).reshape(1, -1))[0]):.3f} dB (reference 23 / 10)")
