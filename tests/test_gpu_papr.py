"""PAPR / windowed PAPR / CCDF kernels against the oracle's direct restatement of
`calculatePAPR.m`, `calculate_window_PAPR.m` (the O(L*Nfft) definition) and `calculateCCDF.m` (MATLAB ecdf)."""
import numpy as np
import pytest

import oracle as O
from oracle import chains as OC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import ofdm_b200
    return ofdm_b200


def _streams(rng, B, L):
    x = rng.standard_normal((B, L)) + 1j * rng.standard_normal((B, L))
    x[0, 100] = 9.0 + 2.0j                       # a dominant peak: windows that contain it have a flat maximum
    return x


@pytest.mark.parametrize("prec,tol", [("f64", 1e-11), ("f32", 2e-5)])
def test_papr_and_window_papr(G, prec, tol):
    ctx = G.default_context(prec)
    rng = np.random.default_rng(2)
    for (B, L, W) in [(3, 5000, 1024), (2, 4096 + 37, 4096), (2, 700, 64), (1, 256, 256)]:
        x = _streams(rng, B, L)
        xd = ctx.cplx(x)
        xh = xd.cpu().numpy().astype(np.complex128)          # the oracle sees the values the device holds
        p = ctx.papr(xd).cpu().numpy()
        w = ctx.window_papr(xd, W).cpu().numpy()
        assert w.shape == (B, L - W + 1)
        for b in range(B):
            assert abs(p[b] - O.calculatePAPR(xh[b])) < tol * 10
            assert np.max(np.abs(w[b] - O.calculate_window_PAPR(xh[b], W))) < tol * 10


def test_window_papr_on_a_tx_stream_task2_shape(G):
    """Task 2's use (`Task 2/Main_model_Task_2.m:72-82`): PAPR of the unscrambled and of the scrambled TX stream."""
    ctx = G.default_context("f32")
    p = OC.params_task5(comb=4)
    rng = np.random.default_rng(3)
    bits = (rng.random(p.stream_bits) < 0.05).astype(np.uint8)          # strongly biased payload
    lp = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues, scramble=False)
    lps = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues, scramble=True)
    tx = ctx.tx_chain(lp, ctx.bits(bits), 1).reshape(1, -1)
    txs = ctx.tx_chain(lps, ctx.bits(bits), 1).reshape(1, -1)
    w = ctx.window_papr(tx, p.Nfft)
    ref = O.calculate_window_PAPR(tx.cpu().numpy()[0].astype(np.complex128), p.Nfft)
    assert np.max(np.abs(w.cpu().numpy()[0] - ref)) < 1e-3
    for t in (tx, txs):
        assert abs(float(ctx.papr(t)[0]) - O.calculatePAPR(t.cpu().numpy()[0].astype(np.complex128))) < 1e-4


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_ccdf_matches_ecdf(G, prec):
    ctx = G.default_context(prec)
    rng = np.random.default_rng(4)
    for n in (1, 2, 1000, 4097, 60417):
        v = np.round(rng.standard_normal(n) * 3, 1 if n > 100 else 3)    # many ties
        vd = ctx.real(v, ctx.rdtype)
        xs, cc = ctx.ccdf(vd)
        rx, rc = O.calculateCCDF(vd.cpu().numpy().astype(np.float64))
        assert xs.numel() == rx.size
        assert np.array_equal(xs.cpu().numpy().astype(np.float64), rx)
        assert np.max(np.abs(cc.cpu().numpy() - rc)) < (1e-12 if prec == "f64" else 1e-6)
