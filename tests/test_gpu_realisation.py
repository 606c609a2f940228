"""The committed golden realisation (tests/golden/realisation_small.mat) through the CUDA path via the C ABI:
TX chain, channel with the IMPORTED noise realisation, M1-style RX chain -- against the stored reference outputs."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_golden_realisation_through_the_gpu_chain(prec):
    import ofdm_b200 as G
    from ofdm_b200 import realisations as R
    spec = importlib.util.spec_from_file_location("mkfix", os.path.join(GOLD, "make_realisation_fixture.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = R.import_realisation(os.path.join(GOLD, "realisation_small.mat"))
    p = mk.small_params()
    ctx = G.default_context(prec)
    lp = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues)
    bits_dev = ctx.bits(g["input_bits"])
    tx = ctx.tx_chain(lp, bits_dev, 1)
    tol = 1e-12 if prec == "f64" else 3e-6
    ref_tx, ref_rx = g["ref_tx"].ravel(), g["ref_rx"].ravel()
    assert np.linalg.norm(tx.cpu().numpy().ravel() - ref_tx) / np.linalg.norm(ref_tx) < tol
    h, _ = ctx.mp_channel_resp(g["channel_taps"], p.Nfft)
    normals = ctx.real(g["noise_normals"].T[None].copy(), ctx.rdtype)            # 1 x 2 x L: real block, imaginary block
    rx = ctx.channel_t5(tx.reshape(1, -1), snr_db=g["SNR_dB"], h_dev=ctx.cplx(h), normals_dev=normals)
    assert np.linalg.norm(rx.cpu().numpy().ravel() - ref_rx) / np.linalg.norm(ref_rx) < tol
    # RX on the REFERENCE stream (so that the comparison of decisions does not inherit the channel's rounding)
    res = ctx.rx_chain_t5(lp, ctx.cplx(ref_rx[None]), 1, tx_bits_dev=bits_dev, want_bits=True, want_H=True, near_eps=1e-3)
    H = res["H"].cpu().numpy().ravel()
    assert np.linalg.norm(H - g["ref_H_LS"].ravel()) / np.linalg.norm(g["ref_H_LS"]) < (1e-11 if prec == "f64" else 2e-5)
    got = ctx.host_bits(res["bits"], p.stream_bits)
    ref_bits = g["ref_bits"].ravel().astype(np.uint8)
    counts = res["counts"].cpu().numpy()
    assert counts[1] == p.stream_bits
    if prec == "f64":
        assert np.array_equal(got, ref_bits) and counts[0] == int(g["ref_errors"].ravel()[0])
    else:
        assert int(np.sum(got != ref_bits)) <= 3 * p.bps * int(counts[2])
