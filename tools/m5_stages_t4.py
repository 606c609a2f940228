#!/usr/bin/env python
"""Stage times of one sweep tile (M5, Task-4 chain) with CUDA events: python tools/m5_stages_t4.py [streams]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ofdm_b200 as G
from ofdm_b200 import layouts
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = G.Context(0, "f32")
lp = layouts.task4_link(ctx)
lp_raw = layouts.task4_link(ctx, scramble=False)
h_dev = ctx.cplx(ctx.mp_channel_resp(layouts.TAPS_TASK4, lp.Nfft)[0])
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
bits = T("descramble (payload = D(s))", lambda: ctx.scramble(sbits, n * (lp.S // lp.SpF), lp.frame_bits, descramble=True))
tx = T("tx_chain without scrambler (on s)", lambda: ctx.tx_chain(lp_raw, sbits, n)).reshape(n, -1)
rng = np.random.default_rng(0)
sto = rng.integers(0, 1153, n); cfo = rng.integers(0, 31, n) + rng.random(n) - 0.5
a = T("add_noise", lambda: ctx.add_noise(tx, 25.0, seed=1)[0])
b = T("add_sto", lambda: ctx.add_sto(a, sto))
c = T("add_cfo", lambda: ctx.add_cfo(b, cfo, lp.Nfft))
d = T("apply_fir", lambda: ctx.apply_fir(c, h_dev))
T("rx_chain_t4 (counters only)", lambda: ctx.rx_chain_t4_fused(lp, d, tx_bits_dev=bits, want_bits=False, near_eps=1e-4))
