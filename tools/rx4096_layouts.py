#!/usr/bin/env python
"""rx4096_kernel on other pilot layouts, three-CTA against two-CTA instantiation: python tools/rx4096_layouts.py [streams]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ofdm_b200 as G
from ofdm_b200 import layouts
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = G.Context(0, "f32")
for comb in (4, 7, 8):
    lp = layouts.task5_link(ctx, comb=comb)
    bits = torch.randint(-2**31, 2**31 - 1, (B * lp.stream_bits // 32,), dtype=torch.int32, device=ctx.device)
    tx, ps = ctx.tx_chain(lp, bits, B, want_power=True)
    h = ctx.cplx(ctx.mp_channel_resp(layouts.TAPS_TASK5, lp.Nfft)[0])
    rx = ctx.channel_t5(tx, snr_db=20.0, h_dev=h, seed=1, power_sum=ps)
    del tx
    for slim in (True, False):
        if slim: os.environ.pop("OFDM_B200_NO_SLIM", None)
        else: os.environ["OFDM_B200_NO_SLIM"] = "1"
        for _ in range(2): out = ctx.rx_chain_t5(lp, rx, B, tx_bits_dev=bits, want_bits=False, want_H=False, near_eps=1e-4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = ctx.rx_chain_t5(lp, rx, B, tx_bits_dev=bits, want_bits=False, want_H=False, near_eps=1e-4)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        c = out["counts"].cpu().numpy()
        print(f"comb {comb} {'three-CTA' if slim else 'two-CTA  '}: {B * lp.S / ms / 1e3:7.2f} M symbols/s  errors {int(c[0])} near {int(c[2])}")
    del rx
