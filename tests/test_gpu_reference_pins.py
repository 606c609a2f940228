"""The CUDA path (through the C ABI) against the numbers the reference itself holds -- see `tests/reference_pins.py`.
Statistical pins run over >= 10^4 streams per experiment (2,048 per BER point: 34 ... 136 Mbit), so their Monte-Carlo
error is far below the figure-reading error; deterministic pins run in the FP64 mode."""
import numpy as np
import pytest
import torch

from oracle import chains as OC

import reference_pins as RP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import ofdm_b200
    return ofdm_b200


def _lp(ctx, p, scramble=True):
    return ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers,
                           p.pilotCarriers, p.pilotValues, scramble=scramble)


def _replicated_payload(ctx, p, B):
    """The reference transmits the same eagle.tiff payload in every run; B streams = B noise realisations of it."""
    bits = RP.eagle_bits(p.stream_bits)
    return bits, ctx.bits(np.tile(bits, B))


@pytest.mark.parametrize("cname", list(RP.BER_SNR_FIGURE))
def test_P1_ber_snr_curve_on_gpu(G, cname):
    """`Task 3/Main_model_Task_3.m:192-268`: TX chain -> Noise -> OFDM_demodulator -> get_payload -> demapping ->
    DeScrambler -> BER, fused kernels, Philox noise, 2,048 streams per SNR point."""
    ctx = G.default_context("f32")
    p = OC.params_task4(alternate=False, Constellation=cname)
    lp = _lp(ctx, p)
    B = 2048
    _, bits_d = _replicated_payload(ctx, p, B)
    tx = ctx.tx_chain(lp, bits_d, B)
    for i, (snr, fig) in enumerate(RP.BER_SNR_FIGURE[cname].items()):
        rx = ctx.channel_t5(tx, snr_db=float(snr), h_dev=None, seed=100 + i)
        res = ctx.rx_chain_t4_fused(lp, rx, tx_bits_dev=bits_d, time_desync=False, freq_desync=False, mp_desync=False, want_bits=False)
        ctx.sync()
        c = res["counts"].cpu().numpy()
        assert c[1] == B * p.stream_bits
        ber = c[0] / c[1]
        assert abs(ber / fig - 1) < RP.ber_tolerance(fig, cname), (cname, snr, ber, fig)


def test_P2_mer_at_25dB_on_gpu(G):
    """`Task 3/README.md:53-55` (MER 29.0341 dB, BER 0 at 25 dB) over 10,240 streams."""
    ctx = G.default_context("f32")
    p = OC.params_task4(percent=0.25, alternate=False)
    lp = _lp(ctx, p)
    B, chunk = 10240, 2048
    sums = torch.zeros(2, dtype=torch.float64, device=ctx.device)
    errors = 0
    _, bits_d = _replicated_payload(ctx, p, chunk)
    tx = ctx.tx_chain(lp, bits_d, chunk)
    for c0 in range(0, B, chunk):
        rx = ctx.channel_t5(tx, snr_db=25.0, h_dev=None, seed=7, first_stream_id=c0)
        grid = ctx.demodulate(rx, p.Nfft, p.T_Guard)
        iq = ctx.get_payload(grid, p.dataCarriers)
        ctx.mer(iq.reshape(-1), "16QAM", sums)
        res = ctx.rx_chain_t4_fused(lp, rx, tx_bits_dev=bits_d, time_desync=False, freq_desync=False, mp_desync=False, want_bits=False)
        ctx.sync()
        errors += int(res["counts"][0].item())
    s = sums.cpu().numpy()
    mer = 10 * np.log10(s[0] / s[1])
    assert abs(mer - RP.MER_AWGN_25DB) < 0.06, mer
    assert errors == 0


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_P3_task4_mer_table_on_gpu(G, prec):
    """`Task 4/README.md:179-183`: noise-free 3-tap channel, estimate_channel / interpolate / equalize_signal / MER_func
    on the GPU; pilot step 2 (see reference_pins.P3).  FP32 cannot resolve 130 dB: there the spline figure is bounded by
    the arithmetic (> 100 dB) and only the linear one is compared."""
    ctx = G.default_context(prec)
    p = OC.params_task4(percent=50)
    lp = _lp(ctx, p)
    bits = RP.eagle_bits(p.stream_bits)
    tx = ctx.tx_chain(lp, ctx.bits(bits), 1)
    h, _ = ctx.mp_channel_resp(RP.TAPS_T4, p.Nfft)
    rx = ctx.apply_fir(tx.reshape(1, -1), ctx.cplx(h)).reshape(1, p.N_symb, -1)
    grid = ctx.demodulate(rx, p.Nfft, p.T_Guard)
    H_spline, Hp = ctx.estimate_channel(grid, np.arange(1, p.Nfft + 1), p.pilotCarriers, p.pilotValues)
    H_lin = ctx.interpolate(Hp, p.pilotCarriers, p.N_carrier, "linear")
    last_uniform = int(p.pilotCarriers[-2])
    inner = p.dataCarriers[p.dataCarriers < last_uniform]

    def mer(H):
        eq = ctx.equalize(grid, H, p.N_carrier)
        s = ctx.mer(ctx.get_payload(eq, inner).reshape(-1), "16QAM").cpu().numpy()
        return 10 * np.log10(s[0] / s[1])

    m_lin, m_spl = mer(H_lin), mer(H_spline)
    assert abs(m_lin - RP.MER_TABLE_T4["linear"]) < 0.5, m_lin
    if prec == "f64":
        assert abs(m_spl - RP.MER_TABLE_T4["spline"]) < 0.5, m_spl
        # cubic convolution is not a library function; apply it on the host to the GPU's pilot estimates
        Hc = np.zeros(p.N_carrier, dtype=complex)
        Hc[:last_uniform] = RP.keys_cubic(p.pilotCarriers[:-1], Hp[0].cpu().numpy()[:-1], np.arange(1.0, last_uniform + 1))
        Hc[last_uniform:] = 1.0
        assert abs(mer(ctx.cplx(Hc)[None]) - RP.MER_TABLE_T4["cubic"]) < 1.5
    else:
        assert m_spl > 100.0, m_spl


def test_P4_task2_papr_on_gpu(G):
    ctx = G.default_context("f64")
    p = OC.params_task4(percent=1, scale=2.0, alternate=True)
    bits_d = ctx.bits(RP.eagle_bits(p.stream_bits))
    plain = float(ctx.papr(ctx.tx_chain(_lp(ctx, p, scramble=False), bits_d, 1).reshape(1, -1))[0].item())
    scr = float(ctx.papr(ctx.tx_chain(_lp(ctx, p, scramble=True), bits_d, 1).reshape(1, -1))[0].item())
    assert abs(plain - RP.PAPR_T2["plain"]) < 1.5 and abs(scr - RP.PAPR_T2["scrambled"]) < 1.5
    assert abs(plain - 22.483) < 0.01 and abs(scr - 11.199) < 0.01
