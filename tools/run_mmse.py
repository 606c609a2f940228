#!/usr/bin/env python
"""MMSE_CE timing: per-stream Levinson vs the shared-statistics tensor-core path.  python tools/run_mmse.py [streams] [comb]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ofdm_b200 as G
from ofdm_b200 import layouts
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
comb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = G.Context(0, "f32")
pil, _ = layouts.pilot_layout_comb(1024, comb)
pv = layouts.pilot_values(len(pil), 14, "16QAM", 4 / 3, False)
Y = (torch.randn(B, 1, 4096, device=ctx.device) + 1j * torch.randn(B, 1, 4096, device=ctx.device)).to(torch.complex64)
h = ctx.cplx(np.array([1, 0, 0, 0, .8, 0, 0, 0, 0, 0, .6, 0, 0, 0, 0, .4, 0, 0, 0, 0, 0, .2, 0, 0, 0, .1], dtype=complex))
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: ctx.mmse_ce_shared(Y, pv, pil, 1024, h, 20.0))
Np = len(pil)
print(f"shared  Np {Np} B {B}: {ms:.3f} ms  {B / ms * 1e3:.0f} streams/s  split-TF32 GEMM {2.0 * B * (2 * Np) * (6 * Np) / ms / 1e9:.1f} TFLOP/s (incl. W build and spline)")
Bl = min(B, 2048)
ms = timed(lambda: ctx.mmse_ce(Y[:Bl], pv, pil, 1024, h[None].expand(Bl, -1).contiguous(), 20.0), reps=2)
print(f"levinson Np {Np} B {Bl}: {ms:.3f} ms  {Bl / ms * 1e3:.0f} streams/s  FP64 {32.0 * Np * Np * Bl / ms / 1e6:.0f} GFLOP/s")
