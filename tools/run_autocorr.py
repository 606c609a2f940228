import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ofdm_b200 as G
ctx = G.Context(0, "f32")
B = 2000
rx = (torch.randn(B, 57600, device=ctx.device) + 1j * torch.randn(B, 57600, device=ctx.device)).to(torch.complex64)
for _ in range(3):
    ctx.cp_autocorr(rx, 128, 1024)
torch.cuda.synchronize()
