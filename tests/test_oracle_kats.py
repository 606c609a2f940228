"""Known-answer tests that pin the CPU oracle.  The reference has no tests of its own, so these
KATs are derived by hand from the reference source (KAT 1-9)."""
import numpy as np
import pytest

import oracle as O
from oracle import chains as C


def test_kat1_scrambler_first_32_bits():
    # Task 5/Scrambler.m:20-21 -> taps at register cells 13,14
    out, reg = O.Scrambler(O.DEFAULT_REGISTER, np.zeros(32, dtype=np.uint8))
    assert "".join(map(str, out)) == "00000111111011000010000011010001"
    # final register = last 15 outputs, newest first
    assert list(reg) == list(out[::-1][:15])


def test_kat1_scrambler_recurrence_and_roundtrip():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 2, 700).astype(np.uint8)
    s, _ = O.Scrambler(O.DEFAULT_REGISTER, x)
    hist = list(O.DEFAULT_REGISTER)  # s[1-m] = Register0(m)
    full = hist[::-1] + []
    seq = list(hist[::-1])
    for i in range(x.size):
        seq.append(int(x[i]) ^ seq[-13] ^ seq[-14])
    assert np.array_equal(s, np.array(seq[15:], dtype=np.uint8))
    d, _ = O.DeScrambler(O.DEFAULT_REGISTER, s)
    assert np.array_equal(d, x)


def test_kat2_log_depth_form():
    rng = np.random.default_rng(2)
    L = 6640
    x = rng.integers(0, 2, L).astype(np.uint8)
    s, _ = O.Scrambler(O.DEFAULT_REGISTER, x)
    reg = O.DEFAULT_REGISTER
    t = x.copy()
    for i in range(14):  # fold pre-history into the input
        a = reg[12 - i] if 12 - i >= 0 else 0
        b = reg[13 - i] if 13 - i >= 0 else 0
        t[i] ^= a ^ b
    j = 0
    while 13 * (1 << j) < L:
        s13, s14 = 13 << j, 14 << j
        u = t.copy()
        u[s13:] ^= t[:-s13]
        if s14 < L:
            u[s14:] ^= t[:-s14]
        t = u
        j += 1
    assert j == 9
    assert np.array_equal(t, s)


@pytest.mark.parametrize("name,bps", [("BPSK", 1), ("QPSK", 2), ("8PSK", 3), ("16QAM", 4)])
def test_constellations_unit_power_and_roundtrip(name, bps):
    d, b = O.constellation_func(name)
    assert b == bps and d.size == 2 ** bps
    assert abs(np.mean(np.abs(d) ** 2) - 1) < 1e-15
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2, 1201).astype(np.uint8)
    iq, pad = O.mapping(bits, name)
    assert pad == (-1 if 1201 % bps == 0 else bps - 1201 % bps)
    assert np.array_equal(O.demapping(pad, iq, name), bits)


def test_16qam_separable_table():
    d, _ = O.constellation_func("16QAM")
    I = {0: -3, 1: -1, 2: 3, 3: 1}
    Q = {0: 3, 1: 1, 2: -3, 3: -1}
    for idx in range(16):
        assert np.isclose(d[idx], (I[idx >> 2] + 1j * Q[idx & 3]) / np.sqrt(10))


def test_demapping_ties_and_nan():
    assert list(O.demapping(-1, np.array([0 + 0j]), "16QAM")) == [0, 1, 0, 1]  # first min: index 5 (-1+1i)
    assert list(O.demapping(-1, np.array([np.nan + 0j]), "16QAM")) == [0, 0, 0, 0]


def test_kat3_loopback_ber_zero_all_constellations():
    rng = np.random.default_rng(4)
    for name in ["BPSK", "QPSK", "8PSK", "16QAM"]:
        p = C.params_task4(Constellation=name)
        bits = rng.integers(0, 2, p.stream_bits).astype(np.uint8)
        tx, grid, _ = C.tx_chain(p, bits)
        assert tx.size == 57600
        X = tx.reshape((p.Nfft + p.T_Guard, p.N_symb), order="F")
        Y = O.OFDM_demodulator(X, p.T_Guard)
        assert np.max(np.abs(Y - grid)) < 1e-12
        out = O.demapping(-1, O.get_payload(Y, p.dataCarriers).ravel(order="F"), name)
        dsc = C.scramble_frames(p, out, descramble=True)
        assert O.BER_func(bits, dsc) == 0.0


def test_task4_layout():
    p = C.params_task4()
    assert len(p.pilotCarriers) == 68 and len(p.dataCarriers) == 332
    assert p.pilotCarriers[0] == 1 and p.pilotCarriers[-2] == 397 and p.pilotCarriers[-1] == 400
    assert p.frame_bits == 6640
    assert abs(abs(p.pilotValues[0, 0]) - 1.78885438) < 1e-7
    assert p.pilotValues[1, 0].real < 0


def test_task5_layouts():
    p = C.params_task5(comb=4)
    assert len(p.pilotCarriers) == 256 and len(p.dataCarriers) == 768 and p.frame_bits == 21504
    assert p.stream_len == 64512
    pil, data = O.pilot_layout_percent(1024, 100, 4096, last_gap=1)
    assert len(pil) == 1024 and len(data) == 0
    pil, data = O.pilot_layout_percent(400, 25, 1024, last_gap=2)
    assert len(pil) == 101 and len(data) == 299           # Task 1
    pil, data = O.pilot_layout_percent(400, 1, 1024, last_gap=2)
    assert list(pil) == [1, 101, 201, 301, 400]            # Task 2


TAPS5 = [[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]]


def _task5_comb1_rx():
    p = C.LinkParams()
    p.pilotCarriers, p.dataCarriers = O.pilot_layout_percent(1024, 100, 4096, last_gap=1)
    p.pilotValues, amp = C.make_pilot_values(1024, p.N_symb, "16QAM", 4 / 3, alternate=False)
    grid = np.zeros((p.Nfft, p.N_symb), dtype=complex)
    grid[p.pilotCarriers - 1, :] = p.pilotValues
    tx = O.OFDM_modulator(grid, p.T_Guard).ravel(order="F")
    rx = C.channel_task5(p, tx, None, TAPS5)
    Y = O.OFDM_demodulator(rx.reshape((p.Nfft + p.T_Guard, p.N_symb), order="F"), p.T_Guard)
    return p, amp, Y


def test_kat4_static_channel_equals_fft_of_taps():
    p, amp, Y = _task5_comb1_rx()
    h, H = O.get_MP_channel_resp(TAPS5, p.Nfft)
    assert h.size == 26
    for s in range(p.N_symb):
        assert np.max(np.abs(Y[:1024, s] / amp - H[:1024])) < 1e-13


def test_kat5_omp_mp_selected_taps():
    p, amp, Y = _task5_comb1_rx()
    A = O.sensing_matrix_dft(p.pilotCarriers, p.Nfft, 1024)
    y = Y[p.pilotCarriers - 1, 0] / amp
    H, h, index = O.OMP_estimate(y, A, p.Nfft, 6, 20)
    assert list(index) == [1, 5, 11, 16, 21, 27]
    Hm, hm = O.MP_estimate(y, A, p.Nfft, 6)
    assert list(np.nonzero(hm)[0] + 1) == sorted([1, 5, 11, 16, 21, 6])


def test_kat6_dictionary_correlation_is_ifft():
    rng = np.random.default_rng(6)
    for pil in [np.arange(1, 1025, 4), np.sort(rng.permutation(1024)[:256]) + 1]:
        A = O.sensing_matrix_dft(pil, 4096, 4096)
        r = rng.standard_normal(pil.size) + 1j * rng.standard_normal(pil.size)
        z = np.zeros(4096, dtype=complex)
        z[pil - 1] = r
        assert np.allclose(A.conj().T @ r, 4096 * np.fft.ifft(z), atol=1e-9)
        assert np.allclose(np.sum(np.abs(A) ** 2, axis=0), pil.size)


def test_kat7_spline_is_linear_operator_not_a_knot():
    rng = np.random.default_rng(7)
    loc = np.arange(1, 1025, 4)
    a = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    b = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    fa, fb, fab = (O.interpolate(v, loc, 1024, "spline") for v in (a, b, 2 * a - 3j * b))
    assert np.allclose(fab, 2 * fa - 3j * fb, atol=1e-10)
    assert np.allclose(fa[loc - 1], a)
    # a cubic is reproduced exactly by a not-a-knot spline, including the appended end point
    x = np.arange(1, 1025, dtype=float)
    cubic = lambda t: 1e-8 * t ** 3 - 2e-5 * t ** 2 + 3e-3 * t - 1
    knots = np.array([1., 6., 30., 31., 500., 1024.])
    assert np.allclose(O.interp1_spline(knots, cubic(knots), x), cubic(x), atol=1e-9)
    # linear method
    lin = O.interpolate(a, loc, 1024, "linear")
    assert np.allclose(lin[loc - 1], a) and np.allclose(lin[1], a[0] + (a[1] - a[0]) / 4)
    # comb 1 -> identity
    c = rng.standard_normal(1024) + 0j
    assert np.allclose(O.interpolate(c, np.arange(1, 1025), 1024, "spline"), c)


def test_kat8_mmse_rpp_structure_and_limit():
    p, amp, Y = _task5_comb1_rx()
    H_ls = O.LS_CE(Y, p.pilotValues, p.pilotCarriers, p.N_carrier)
    _, Htrue = O.get_MP_channel_resp(TAPS5, p.Nfft)
    assert np.max(np.abs(H_ls - Htrue[:1024])) < 1e-12
    H_mmse = O.MMSE_CE(Y, p.pilotValues, p.pilotCarriers, p.Nfft, p.N_carrier, np.fft.ifft(H_ls), 20)
    assert H_mmse.shape == (1024,)
    # H = Ht - Rpp^{-1} Ht / snr  (rows 1:Np of Rhp equal Rpp - I/snr)
    h = np.fft.ifft(H_ls)
    k = np.arange(1024)
    pw = np.abs(h) ** 2
    r = np.sum(pw * k) / pw.sum()
    r2 = np.sum(pw * k * k) / pw.sum()
    c = 2j * np.pi * np.sqrt(r2 - r * r) / 1024
    Rpp = 1 / (1 + c * (k[:, None] - k[None, :])) + np.eye(1024) / 100
    assert np.allclose(Rpp, Rpp.conj().T)
    assert np.allclose(H_mmse, H_ls - np.linalg.solve(Rpp, H_ls) / 100, atol=1e-9)


@pytest.mark.parametrize("sto,cfo", [(0, 0.0), (37, 7.24), (300, 25.24), (900, 0.24)])
def test_kat9_task4_coarse_sync(sto, cfo):
    rng = np.random.default_rng(9)
    p = C.params_task4()
    bits = rng.integers(0, 2, p.stream_bits).astype(np.uint8)
    tx, _, _ = C.tx_chain(p, bits)
    rx = C.impair_task4(p, tx, SNR_dB=None, Time_Delay=sto, Freq_Shift=cfo)
    ac, tg, fo = O.AutoCorrFunction(rx, p.T_Guard, p.Nfft)
    assert ac.size == 57600 - 128 - 1024
    frac = cfo - round(cfo)
    assert abs(fo - frac) < 2e-3
    r = O.add_STO(O.add_STO(rx, tg), -(p.Nfft + p.T_Guard))
    r = O.add_CFO(r, -fo, p.Nfft)
    r, ifo = O.remove_IFO(r, p.Nfft)
    assert ifo == round(cfo)
    out = C.rx_chain_task4(p, rx, bits, mp_desync=False)
    # the reference's own pass criterion (Main_model_Task_4.m:367).  (300, 25.24) lands the common
    # phase near +-pi, where fine_sync's mean of wrapped angles (fine_sync.m:52) breaks down -- a
    # property of the reference algorithm that the oracle reproduces rather than fixes.
    if (sto, cfo) != (300, 25.24):
        assert out["errors"] / out["n_bits"] < 0.2
        assert abs(out["tau"] - 8 / 2048) < 1e-5   # constant residual of -8 samples left for fine_sync


def test_autocorr_fallback_65():
    rng = np.random.default_rng(10)
    x = rng.standard_normal(4000) + 1j * rng.standard_normal(4000)
    ac, tg, fo = O.AutoCorrFunction(x, 128, 1024)
    assert tg == 65


def test_sto_cfo_basic():
    y = np.arange(1, 7) + 0j
    assert list(O.add_STO(y, 2).real) == [3, 4, 5, 6, 0, 0]
    assert list(O.add_STO(y, -2).real) == [0, 0, 1, 2, 3, 4]
    z = O.add_CFO(np.ones(8), 1.0, 8)
    assert np.allclose(z, np.exp(2j * np.pi * np.arange(8) / 8))


def test_mer_awgn_sanity():
    # Task 3/README.md:53-55: MER ~ SNR + 10log10(1024/400) at 25 dB
    rng = np.random.default_rng(11)
    p = C.params_task4()
    bits = rng.integers(0, 2, p.stream_bits).astype(np.uint8)
    tx, _, _ = C.tx_chain(p, bits)
    rx, nvar = O.Noise(25, tx, rng=rng)
    Y = O.OFDM_demodulator(rx.reshape((1152, 50), order="F"), 128)
    mer = O.MER_func(O.get_payload(Y, p.dataCarriers).ravel(order="F"), "16QAM")
    # pilots at +-4/3*max boost total power, so the offset is close to, not exactly, 4.08 dB
    assert 27.5 < mer < 30.5


def test_equalize_zero_rows_and_ber():
    X = np.ones((8, 2), dtype=complex)
    out = O.equalize_signal(X, 2 * np.ones(8), 3)
    assert np.allclose(out[:3], 0.5) and np.all(out[3:] == 0)
    assert O.BER_func([0, 1, 1, 0], [0, 1, 0, 1]) == 0.5


def test_payload_reader_matches_survey_stats():
    import os
    path = "/root/reference/Task 5/eagle.tiff"
    if not os.path.exists(path):
        pytest.skip("reference fixture only exists in the build container")
    bits = C.read_payload_bits(path, 129600)
    assert bits.size == 129600 and abs(bits.mean() - 0.337) < 0.005


@pytest.mark.parametrize("L", [1, 5, 13, 14, 15, 16, 100, 6640])
def test_fast_scrambler_forms_equal_the_loops(L):
    rng = np.random.default_rng(L)
    x = rng.integers(0, 2, L).astype(np.uint8)
    for reg in (O.DEFAULT_REGISTER, rng.integers(0, 2, 15).astype(np.uint8)):
        a, ra = O.Scrambler(reg, x)
        b, rb = O.Scrambler_fast(reg, x)
        assert np.array_equal(a, b) and np.array_equal(ra, rb)
        c, rc = O.DeScrambler(reg, a)
        d, rd = O.DeScrambler_fast(reg, a)
        assert np.array_equal(c, d) and np.array_equal(rc, rd) and np.array_equal(c, x)
