#!/usr/bin/env python
"""Secondary workloads of SURVEY 8d (M2 sync chain, M3 channel estimation, M4 OMP/MP) timed with CUDA events on
one B200.  bench.py stays the headline (M1); these numbers go to DESIGN.md / profiles/.
usage: python tools/bench_configs.py [--m2-streams 2000] [--m3-streams 8192] [--m4-frames 4096]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ofdm_b200 as G  # noqa: E402
import oracle as O  # noqa: E402
from oracle import chains as OC  # noqa: E402

PEAK = 6551.4


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m2-streams", type=int, default=2000)
    ap.add_argument("--m3-streams", type=int, default=8192)
    ap.add_argument("--m4-frames", type=int, default=4096)
    a = ap.parse_args()
    ctx = G.Context(0, "f32")
    out = {}

    # ---------------- M2: Task-4 sync + CE chain (composed from the per-function kernels)
    p = OC.params_task4()
    lp = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues)
    B = a.m2_streams
    words = (B * p.stream_bits + 31) // 32
    bits = torch.randint(-2**31, 2**31 - 1, (words,), dtype=torch.int32, device=ctx.device)
    tx = ctx.tx_chain(lp, bits, B).reshape(B, -1)
    rng = np.random.default_rng(0)
    sto = rng.integers(0, 1153, B)
    cfo = rng.integers(0, 31, B) + rng.random(B) - 0.5
    h = ctx.cplx(O.get_MP_channel_resp([[0, 1], [4, .6], [10, .3]], p.Nfft)[0])
    rx, _ = ctx.add_noise(tx, 25.0, seed=1)
    rx = ctx.apply_fir(ctx.add_cfo(ctx.add_sto(rx, sto), cfo, p.Nfft), h)
    res = {}
    def m2():
        res["o"] = ctx.rx_chain_t4(lp, rx, tx_bits_dev=bits)
    ms = timed(m2)
    cnt = res["o"]["counts"].cpu().numpy()
    syms = B * p.N_symb
    out["M2_task4_sync_chain"] = {"streams": B, "symbols": syms, "ms": ms, "symbols_per_s": syms / ms * 1e3, "algorithmic_B_per_symbol": 9548,
                                  "GBps_algorithmic": 9548 * syms / ms / 1e6, "frac_of_measured_hbm": 9548 * syms / ms / 1e6 / PEAK,
                                  "ber": float(cnt[0]) / float(cnt[1]), "note": "13 kernels composed on the device, not yet fused"}

    def m2f():
        res["f"] = ctx.rx_chain_t4_fused(lp, rx, tx_bits_dev=bits)
    ms = timed(m2f)
    cntf = res["f"]["counts"].cpu().numpy()
    out["M2_task4_sync_chain_fused"] = {"streams": B, "symbols": syms, "ms": ms, "symbols_per_s": syms / ms * 1e3, "algorithmic_B_per_symbol": 9548,
                                        "GBps_algorithmic": 9548 * syms / ms / 1e6, "frac_of_measured_hbm": 9548 * syms / ms / 1e6 / PEAK,
                                        "ber": float(cntf[0]) / float(cntf[1]),
                                        "detector_failures": int(res["f"]["fail"].sum().item()),
                                        "ifo_not_found": int((res["f"]["IFO"] < 0).sum().item()),
                                        "note": "autocorr prefix scan (+ gated full-length re-scan) + one persistent fused kernel"}
    ms_ac = timed(lambda: ctx.cp_autocorr(rx, p.T_Guard, p.Nfft))
    out["M2_task4_sync_chain_fused"]["autocorr_ms"] = ms_ac

    # ---------------- M3: LS / MMSE + interpolate + equalise on post-FFT grids
    for comb in (4, 1):
        if comb == 1:
            p5 = OC.LinkParams()
            p5.pilotCarriers, p5.dataCarriers = O.pilot_layout_percent(1024, 100, 4096, last_gap=1)
            p5.pilotValues, _ = OC.make_pilot_values(1024, 14, "16QAM", 4 / 3, False)
        else:
            p5 = OC.params_task5(comb=comb)
        Bm = a.m3_streams
        Y = (torch.randn(Bm, 14, 4096, dtype=torch.float32, device=ctx.device) + 1j * torch.randn(Bm, 14, 4096, dtype=torch.float32, device=ctx.device)).to(torch.complex64)
        ms_ls = timed(lambda: ctx.equalize(Y, ctx.ls_ce(Y, p5.pilotValues, p5.pilotCarriers, 1024), 1024))
        Hls = ctx.ls_ce(Y, p5.pilotValues, p5.pilotCarriers, 1024)
        hch = ctx.fft(Hls, inverse=True)
        nmm = min(Bm, 1024 if comb == 1 else Bm)
        ms_mm = timed(lambda: ctx.mmse_ce(Y[:nmm], p5.pilotValues, p5.pilotCarriers, 1024, hch[:nmm], 20.0), reps=2, warm=1)
        syms = Bm * 14
        Np = len(p5.pilotCarriers)
        out[f"M3_ce_comb{comb}"] = {"streams": Bm, "ls_equalize_ms": ms_ls, "ls_symbols_per_s": syms / ms_ls * 1e3,
                                    "ls_GBps_full_grid_rw": (2 * 8 * 4096 * 14 + 8192) * Bm / ms_ls / 1e6,
                                    "mmse_streams": nmm, "mmse_ms": ms_mm, "mmse_streams_per_s": nmm / ms_mm * 1e3,
                                    "mmse_fp64_GFLOPs": 32.0 * Np * Np * nmm / ms_mm / 1e6}

    # ---------------- M4: OMP / MP, Np 256
    F = a.m4_frames
    pil = np.sort(np.random.default_rng(1).permutation(1024)[:256]) + 1
    y = (torch.randn(F, 256, device=ctx.device) + 1j * torch.randn(F, 256, device=ctx.device)).to(torch.complex64)
    A = ctx.cplx(np.asfortranarray(O.sensing_matrix_dft(pil, 4096, 4096)).ravel(order="F"))
    # tensor-core path (tcgen05 TF32 correlation + exact re-scoring) at the survey's batch, 65,536 frames
    Fbig = 65536
    # realistic measurements: 6-tap sparse CIRs seen through the pilot mask + noise (random y would have no dominant tap)
    gsel = torch.Generator(device=ctx.device); gsel.manual_seed(7)
    hbig = torch.zeros(Fbig, 4096, dtype=torch.complex64, device=ctx.device)
    taps = torch.randint(0, 200, (Fbig, 6), device=ctx.device, generator=gsel)
    gains = (torch.randn(Fbig, 6, device=ctx.device, generator=gsel) + 1j * torch.randn(Fbig, 6, device=ctx.device, generator=gsel)).to(torch.complex64)
    hbig.scatter_(1, taps, gains)
    ybig = ctx.fft(hbig)[:, torch.as_tensor(pil - 1, device=ctx.device)].contiguous()
    ybig = ybig + 0.05 * (torch.randn(Fbig, 256, device=ctx.device) + 1j * torch.randn(Fbig, 256, device=ctx.device)).to(torch.complex64)
    del hbig
    l0 = ctx.launches
    ms = timed(lambda: ctx.omp(ybig, 4096, 9, A_dev=A), reps=2, warm=1)
    out["M4_omp_dense_L4096_K9_65536_frames"] = {"frames": Fbig, "ms": ms, "frames_per_s": Fbig / ms * 1e3, "GBps_algorithmic": 67620.0 * Fbig / ms / 1e6,
                                            "dense_corr_TFLOPs_whole_call": 8.0 * 256 * 4096 * 9 * Fbig / ms / 1e9,
                                            "kernels_per_call": (ctx.launches - l0) // 3,
                                            "note": "dense dictionary recognised as a partial DFT by the probe kernel -> Batch-OMP (Gram vector + one fused kernel); an unstructured dictionary takes the tcgen05 path (OFDM_B200_NO_DFT_PROBE=1 forces it); outputs H,h (4.3 GB written)"}
    del ybig
    os.environ["OFDM_B200_NO_TC"] = "1"
    for name, fn in (("omp_dense_L4096_K9", lambda: ctx.omp(y, 4096, 9, A_dev=A)),
                     ("omp_dftdesc_L4096_K9", lambda: ctx.omp(y, 4096, 9, Ldict=4096, pilot_loc=pil)),
                     ("mp_dftdesc_L4096_K9", lambda: ctx.mp(y, 4096, 9, Ldict=4096, pilot_loc=pil))):
        ms = timed(fn, reps=2, warm=1)
        out["M4_" + name] = {"frames": F, "ms": ms, "frames_per_s": F / ms * 1e3, "GBps_algorithmic": 67620.0 * F / ms / 1e6,
                             "dense_corr_TFLOPs": (8.0 * 256 * 4096 * 9 * F / ms / 1e9) if "dense" in name else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
