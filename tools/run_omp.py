#!/usr/bin/env python
"""One M4 call (Batch-OMP on the partial-DFT descriptor, or the dense entry) for profiling: python tools/run_omp.py [frames] [dense|desc|tc]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ofdm_b200 as G  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = sys.argv[2] if len(sys.argv) > 2 else "desc"
if mode == "tc":
    os.environ["OFDM_B200_NO_DFT_PROBE"] = "1"
ctx = G.Context(0, "f32")
pil = np.sort(np.random.default_rng(1).permutation(1024)[:256]) + 1
gsel = torch.Generator(device=ctx.device); gsel.manual_seed(7)
hbig = torch.zeros(F, 4096, dtype=torch.complex64, device=ctx.device)
taps = torch.randint(0, 200, (F, 6), device=ctx.device, generator=gsel)
gains = (torch.randn(F, 6, device=ctx.device, generator=gsel) + 1j * torch.randn(F, 6, device=ctx.device, generator=gsel)).to(torch.complex64)
hbig.scatter_(1, taps, gains)
y = ctx.fft(hbig)[:, torch.as_tensor(pil - 1, device=ctx.device)].contiguous()
y = y + 0.05 * (torch.randn(F, 256, device=ctx.device) + 1j * torch.randn(F, 256, device=ctx.device)).to(torch.complex64)
del hbig
A = None
if mode != "desc":
    l = np.arange(4096)
    A = ctx.cplx(np.exp(-2j * np.pi * np.outer(l, pil - 1) / 4096).ravel())      # column-major Np x Ldict
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if A is None:
        out = ctx.omp(y, 4096, 9, Ldict=4096, pilot_loc=pil, tie_eps=1e-4)
    else:
        out = ctx.omp(y, 4096, 9, A_dev=A, tie_eps=1e-4)
    e1.record()
    torch.cuda.synchronize()
    print(mode, F, "frames", e0.elapsed_time(e1), "ms", "near-tie frames", int((out[4] > 0).sum()), "iters mean", float(out[3].float().mean()))
