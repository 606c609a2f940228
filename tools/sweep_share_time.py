#!/usr/bin/env python
"""Time one rank's share of the sweep on one GPU (any rank / world): python tools/sweep_share_time.py task4 8192 4"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ofdm_b200 as G
from ofdm_b200 import layouts, sweep
chain = sys.argv[1] if len(sys.argv) > 1 else "task4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
world = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = G.Context(0, "f32")
lp = layouts.task4_link(ctx) if chain == "task4" else layouts.task5_link(ctx, comb=4)
taps = layouts.TAPS_TASK4 if chain == "task4" else layouts.TAPS_TASK5
snrs = np.arange(0.0, 30.0 + 1e-9, 0.5)
sweep.ber_sweep(ctx, lp, snrs[:2], 8192 * world, taps, chain, seed=7, rank=0, world=world, tile=8192)
for rank in range(world):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        acc = sweep.sweep_local(ctx, lp, snrs, spp, taps, chain, seed=7, rank=rank, world=world, tile=8192, near_eps=1e-4)
        e1.record(); torch.cuda.synchronize()
        print(f"{chain} rank {rank}/{world} rep {rep}: {e0.elapsed_time(e1):8.2f} ms")
