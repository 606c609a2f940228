#!/usr/bin/env python
"""Stage times of one sweep tile (M5, Task-5 chain) with CUDA events: python tools/m5_stages.py [streams]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ofdm_b200 as G
from ofdm_b200 import layouts
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = G.Context(0, "f32")
lp = layouts.task5_link(ctx, comb=4)
lp_raw = layouts.task5_link(ctx, comb=4, scramble=False)
h_dev = ctx.cplx(ctx.mp_channel_resp(layouts.TAPS_TASK5, lp.Nfft)[0])
words = lp.stream_bits // 32
def T(label, fn, reps=5):
    r = fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:34s} {e0.elapsed_time(e1) / reps:8.3f} ms")
    return r
sbits = torch.zeros(n * words, dtype=torch.int32, device=ctx.device)
T("payload_bits (Philox)", lambda: ctx._chk(ctx.lib.ofdm_payload_bits(ctx.h, ctx.p(sbits), n, words, 7, 0)))
bits = T("descramble (payload = D(s))", lambda: ctx.scramble(sbits, n * 2, lp.frame_bits, descramble=True))
import ctypes as C
tx = ctx.empty_c(n, lp.S, lp.Nfft + lp.Tg)
rx = torch.empty_like(tx)
ps = torch.empty(n, dtype=torch.float64, device=ctx.device)
T("tx_chain_p with scrambler", lambda: ctx._chk(ctx.lib.ofdm_tx_chain_p(ctx.h, C.byref(lp), ctx.p(bits), n, ctx.p(tx), ctx.p(ps))))
T("tx_chain_p without (on s)", lambda: ctx._chk(ctx.lib.ofdm_tx_chain_p(ctx.h, C.byref(lp_raw), ctx.p(sbits), n, ctx.p(tx), ctx.p(ps))))
T("channel_t5_p", lambda: ctx.channel_t5(tx, snr_db=10.0, h_dev=h_dev, seed=1, first_stream_id=0, power_sum=ps, out=rx.reshape(n, -1)))
T("rx_chain_t5 (counters only)", lambda: ctx.rx_chain_t5(lp, rx, n, tx_bits_dev=bits, want_bits=False, want_H=False, near_eps=1e-4))
