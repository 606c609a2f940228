import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import ofdm_b200 as G, oracle as O
from oracle import chains as OC
ctx = G.Context(0, "f32")
p = OC.params_task5(comb=4)
lp = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues)
h, _ = O.get_MP_channel_resp([[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]], p.Nfft)
h_dev = ctx.cplx(np.asarray(h))
n = 8192; words = lp.stream_bits // 32
def T(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f"{label:12s} {1e3*(time.perf_counter()-t):9.2f} ms"); return r
for rep in range(3):
    print("rep", rep)
    gen = torch.Generator(device=ctx.device); gen.manual_seed(5 + rep)
    bits = T("randint", lambda: torch.randint(-2**31, 2**31 - 1, (n * words,), dtype=torch.int32, device=ctx.device, generator=gen))
    tx, ps = T("tx_chain", lambda: ctx.tx_chain(lp, bits, n, want_power=True))
    rx = T("channel_t5", lambda: ctx.channel_t5(tx, snr_db=10.0, h_dev=h_dev, seed=1, first_stream_id=rep * n, power_sum=ps))
    res = T("rx_chain_t5", lambda: ctx.rx_chain_t5(lp, rx, n, tx_bits_dev=bits, want_bits=False, want_H=False))
    T("counts.cpu", lambda: res["counts"].cpu().numpy())
    del tx, rx
