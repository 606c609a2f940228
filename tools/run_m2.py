#!/usr/bin/env python
"""One fused Task-4 chain call for profiling: python tools/run_m2.py [streams]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ofdm_b200 as G
from ofdm_b200 import layouts
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ctx = G.Context(0, "f32")
lp = layouts.task4_link(ctx)
torch.manual_seed(0)
bits = torch.randint(-2**31, 2**31 - 1, (B * lp.stream_bits // 32,), dtype=torch.int32, device=ctx.device)
tx = ctx.tx_chain(lp, bits, B).reshape(B, -1)
rng = np.random.default_rng(0)
rx, _ = ctx.add_noise(tx, 25.0, seed=1)
h = ctx.cplx(ctx.mp_channel_resp(layouts.TAPS_TASK4, lp.Nfft)[0])
rx = ctx.apply_fir(ctx.add_cfo(ctx.add_sto(rx, rng.integers(0, 1153, B)), rng.integers(0, 31, B) + rng.random(B) - 0.5, lp.Nfft), h)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ctx.rx_chain_t4_fused(lp, rx, tx_bits_dev=bits, want_bits=True)
    e1.record(); torch.cuda.synchronize()
    c = out["counts"].cpu().numpy()
    print(B, "streams", e0.elapsed_time(e1), "ms  BER", c[0] / c[1], "fails", int(out["fail"].sum()))
