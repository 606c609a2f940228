#!/usr/bin/env python
"""Per-stage timing of the composed Task-4 chain (M2) to see where a fused kernel would pay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ofdm_b200 as G, oracle as O
from oracle import chains as OC

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

ctx = G.Context(0, "f32")
p = OC.params_task4()
lp = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
bits = torch.randint(-2**31, 2**31 - 1, ((B * p.stream_bits + 31) // 32,), dtype=torch.int32, device=ctx.device)
tx = ctx.tx_chain(lp, bits, B).reshape(B, -1)
rng = np.random.default_rng(0)
rx, _ = ctx.add_noise(tx, 25.0, seed=1)
rx = ctx.apply_fir(ctx.add_cfo(ctx.add_sto(rx, rng.integers(0, 1153, B)), rng.integers(0, 31, B) + rng.random(B) - 0.5, p.Nfft), ctx.cplx(O.get_MP_channel_resp([[0, 1], [4, .6], [10, .3]], p.Nfft)[0]))
t, (_, tg, fo, fail) = timed(lambda: ctx.cp_autocorr(rx, 128, 1024)); print(f"cp_autocorr      {t:8.3f} ms")
t, x = timed(lambda: ctx.add_sto(rx, tg)); print(f"add_sto          {t:8.3f} ms")
t, x = timed(lambda: ctx.add_sto(x, -1152)); print(f"add_sto (2)      {t:8.3f} ms")
t, x = timed(lambda: ctx.add_cfo(x, -fo, 1024)); print(f"add_cfo          {t:8.3f} ms")
t, (x, ifo) = timed(lambda: ctx.remove_ifo(x, 1024)); print(f"remove_ifo       {t:8.3f} ms")
t, g = timed(lambda: ctx.demodulate(x.reshape(B, 50, 1152), 1024, 128)); print(f"demodulate       {t:8.3f} ms")
t, (g, tau, ph) = timed(lambda: ctx.fine_sync(g, p.pilotCarriers, p.pilotValues, 1, 1)); print(f"fine_sync        {t:8.3f} ms")
t, (H, _) = timed(lambda: ctx.estimate_channel(g, np.arange(1, 1025), p.pilotCarriers, p.pilotValues)); print(f"estimate_channel {t:8.3f} ms")
t, g = timed(lambda: ctx.equalize(g, H, 400)); print(f"equalize         {t:8.3f} ms")
t, iq = timed(lambda: ctx.get_payload(g, p.dataCarriers)); print(f"get_payload      {t:8.3f} ms")
t, raw = timed(lambda: ctx.demap(iq.reshape(-1), "16QAM")); print(f"demap            {t:8.3f} ms")
t, d = timed(lambda: ctx.scramble(raw, B * 10, 6640, descramble=True)); print(f"descramble       {t:8.3f} ms")
t, c = timed(lambda: ctx.ber_count(bits, d, B * p.stream_bits)); print(f"ber_count        {t:8.3f} ms")
print("stream bytes", rx.numel() * 8 / 1e6, "MB  -> one pass at 6.5 TB/s =", rx.numel() * 8 / 6.55e9, "ms")
