"""Task 5 part 2 on the GPU against the oracle's restatement of `Task5_part2.m:148-304`:
one sweep point = one pilot layout, the Monte-Carlo runs are the batch axis.  The fading realisations come from
the device generator and are handed to the oracle, the AWGN realisation is imported on both sides."""
import numpy as np
import pytest

from oracle import chains as OC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import ofdm_b200
    return ofdm_b200


@pytest.mark.parametrize("profile", ["EPA", "EVA", "ETU"])
def test_tdl_channel_matches_published_model(G, profile):
    ctx = G.default_context("f64")
    fs = 4e7
    n, hl, delays = ctx.tdl_info(profile, fs)
    d_ns, p_db = OC.TDL_PROFILES[profile]
    assert n == len(d_ns) and np.allclose(delays, np.asarray(d_ns) * 1e-9 * fs)
    B = 4096
    h, g = ctx.tdl_channel(profile, fs, B, seed=3, first_stream_id=10, want_gains=True)
    h, g = h.cpu().numpy(), g.cpu().numpy()
    assert h.shape == (B, hl)
    for b in (0, 1, B - 1):                                            # impulse response = windowed-sinc interpolation of the gains
        ref = OC.tdl_impulse_response(profile, fs, g[b])
        assert np.linalg.norm(h[b] - ref) / np.linalg.norm(ref) < 1e-5   # generator runs in FP32
    pw = np.mean(np.abs(g) ** 2, axis=0)                               # Rayleigh taps, normalised profile powers
    assert np.allclose(pw, OC.tdl_amplitudes(profile) ** 2, rtol=0.12, atol=2e-4) and abs(pw.sum() - 1) < 0.05
    # keyed by the global stream id: any split of the batch gives the same realisations
    h2 = ctx.tdl_channel(profile, fs, 16, seed=3, first_stream_id=10 + 100).cpu().numpy()
    assert np.array_equal(h2, h[100:116])


CASES = [("comb8", dict(comb=8), 8), ("comb37", dict(comb=37), 37), ("random64", None, None)]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name,kw,comb", CASES)
def test_part2_point_matches_oracle(G, name, kw, comb, prec):
    from ofdm_b200 import part2
    ctx = G.default_context(prec)
    rng = np.random.default_rng(11)
    if kw is None:
        mask = np.sort(rng.permutation(1024)[:64]) + 1                 # `sort(randperm(1024, Np))`, Task5_part2.m:63
        p = OC.params_part2(pilotCarriers=mask)
        Ldict = 4096
    else:
        p = OC.params_part2(**kw)
        Ldict = -(-4096 // comb)
    runs = 3
    bits = rng.integers(0, 2, p.stream_bits).astype(np.uint8)
    normals = rng.standard_normal((1, 2, p.stream_len))
    nmse, errs, nbits, ex = part2.run_point(ctx, p.pilotCarriers, p.dataCarriers, bits, runs, profile="EPA", snr_db=20.0, seed=5,
                                            point_id=2, Ldict=Ldict, noise_normals=ctx.real(normals, ctx.rdtype))
    assert nbits == p.stream_bits
    h = ex["h"].cpu().numpy().astype(np.complex128)
    tx, _, _ = OC.tx_chain(p, bits, scramble=False)
    txn, _ = OC.F.Noise(20.0, tx, normals=normals[0])
    n_paths = len(OC.TDL_PROFILES["EPA"][0])
    tol = 1e-7 if prec == "f64" else 2e-3
    ref_err = np.zeros(4, dtype=np.int64)
    for r in range(runs):
        rn, re_, rex = OC.part2_run(p, txn, h[r], bits, n_paths, 20.0, Ldict)
        ref_err += re_
        for j, k in enumerate(part2.ESTIMATORS):
            Hd = ex["H"][k][r].cpu().numpy()[:1024]
            Hr = np.asarray(rex["H"][j]).ravel()[:1024]
            assert np.linalg.norm(Hd - Hr) / np.linalg.norm(Hr) < tol, (k, r)
            assert abs(nmse[r, j] - rn[j]) <= tol * max(rn[j], 1e-3) * 10, (k, r, nmse[r, j], rn[j])
    if prec == "f64":
        assert np.array_equal(errs, ref_err)
    else:
        assert np.all(np.abs(errs - ref_err) <= 0.002 * runs * nbits + 8)


def test_part2_sweep_is_split_invariant(G):
    """Two points, few runs: the result does not depend on how the Monte-Carlo runs are cut into batches (fading
    realisations are keyed by (point, run), the AWGN realisation by the point)."""
    from ofdm_b200 import part2
    ctx = G.default_context("f32")
    pay = lambda n: np.random.default_rng(n).integers(0, 2, n).astype(np.uint8)   # noqa: E731  (same bits for a given layout)
    N4, B4 = part2.sweep(ctx, [8, 32], 8, pay, block=4, profile="EVA", seed=9)
    N8, B8 = part2.sweep(ctx, [8, 32], 8, pay, block=8, profile="EVA", seed=9)
    assert N4.shape == (4, 2) and B4.shape == (4, 2)
    assert np.all(np.isfinite(N4)) and np.all((B4 >= 0) & (B4 <= 1))
    assert np.allclose(N4, N8, rtol=1e-6, atol=0) and np.array_equal(B4, B8)
    assert B4[0, 0] < 0.05                   # LS with 128 pilots at 20 dB on EVA: a working link
