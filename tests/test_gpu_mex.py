"""The MEX gateway (mex/ofdm_mex.c) executed on the GPU through the mx* shim (mex/shim): what MATLAB / Octave
would do when a reference script calls a wrapper from matlab/.  Both complex-storage models are exercised
(R2018a interleaved and Octave / legacy split).  Results are compared with the float64 oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
from oracle import chains as OC

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Mex:
    def __init__(self, variant):
        self.lib = C.CDLL(os.path.join(ROOT, "mex", "lib", f"libofdm_mex_{variant}.so"))
        L = self.lib
        L.mxCreateDoubleMatrix.restype = C.c_void_p
        L.mxCreateDoubleMatrix.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        L.mxCreateString.restype = C.c_void_p
        L.mxCreateString.argtypes = [C.c_char_p]
        L.mxDestroyArray.argtypes = [C.c_void_p]
        L.mxGetM.restype = C.c_size_t; L.mxGetM.argtypes = [C.c_void_p]
        L.mxGetN.restype = C.c_size_t; L.mxGetN.argtypes = [C.c_void_p]
        L.mxIsComplex.argtypes = [C.c_void_p]
        L.ofdm_mex_shim_set.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double]
        L.ofdm_mex_shim_get.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ofdm_mex_shim_call.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
        L.ofdm_mex_shim_error.restype = C.c_char_p

    def to_mx(self, a):
        if isinstance(a, str):
            return self.lib.mxCreateString(a.encode())
        a = np.atleast_2d(np.asarray(a))
        if a.ndim == 2 and a.shape[0] == 1 and getattr(a, "_col", False):
            a = a.T
        cplx = np.iscomplexobj(a)
        m = self.lib.mxCreateDoubleMatrix(a.shape[0], a.shape[1], 1 if cplx else 0)
        flat = a.ravel(order="F")
        for i, v in enumerate(flat):
            self.lib.ofdm_mex_shim_set(m, i, float(np.real(v)), float(np.imag(v)))
        return m

    def from_mx(self, m):
        r, c = self.lib.mxGetM(m), self.lib.mxGetN(m)
        out = np.zeros(r * c, dtype=complex)
        re, im = C.c_double(), C.c_double()
        for i in range(r * c):
            self.lib.ofdm_mex_shim_get(m, i, C.byref(re), C.byref(im))
            out[i] = re.value + 1j * im.value
        out = out.reshape((r, c), order="F")
        return out if self.lib.mxIsComplex(m) else out.real

    def call(self, op, *args, nout=1):
        ins = [self.to_mx(op)] + [self.to_mx(a) for a in args]
        arr_in = (C.c_void_p * len(ins))(*ins)
        arr_out = (C.c_void_p * max(nout, 1))()
        rc = self.lib.ofdm_mex_shim_call(nout, arr_out, len(ins), arr_in)
        for m in ins:
            self.lib.mxDestroyArray(m)
        if rc:
            raise RuntimeError(self.lib.ofdm_mex_shim_error().decode())
        outs = [self.from_mx(arr_out[i]) for i in range(nout)]
        for i in range(nout):
            self.lib.mxDestroyArray(arr_out[i])
        return outs[0] if nout == 1 else outs


def col(a):
    return np.asarray(a).reshape(-1, 1)


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


@pytest.fixture(scope="module", params=["interleaved", "split"])
def mex(request):
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "mex")], check=True, capture_output=True)
    m = Mex(request.param)
    yield m
    m.lib.ofdm_mex_shim_exit()


def test_mex_bits_and_mapping(mex):
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, 701).astype(float)
    reg = O.DEFAULT_REGISTER.astype(float)
    s, r = mex.call("Scrambler", reg, bits, nout=2)
    s_ref, r_ref = O.Scrambler(reg, bits)
    assert s.shape == (1, 701) and np.array_equal(s.ravel(), s_ref) and np.array_equal(r.ravel(), r_ref)
    d, _ = mex.call("DeScrambler", reg, s, nout=2)
    assert np.array_equal(d.ravel(), bits)
    iq, pad = mex.call("mapping", col(bits), "16QAM", nout=2)
    iq_ref, pad_ref = O.mapping(bits, "16QAM")
    assert iq.shape == (1, 176) and pad[0, 0] == pad_ref and np.max(np.abs(iq.ravel() - iq_ref)) < 1e-7
    back = mex.call("demapping", float(pad_ref), iq, "16QAM")
    assert np.array_equal(back.ravel(), bits)
    dic, bps = mex.call("constellation_func", "8PSK", nout=2)
    assert bps[0, 0] == 3 and np.max(np.abs(dic.ravel() - O.constellation_func("8PSK")[0])) < 1e-15
    assert mex.call("BER_func", bits, 1 - bits)[0, 0] == 1.0
    assert abs(mex.call("MER_func", iq.ravel() + 0.01, "16QAM")[0, 0] - O.MER_func(iq_ref + 0.01, "16QAM")) < 1e-3


def test_mex_papr_and_ccdf(mex):
    """calculatePAPR / calculate_window_PAPR / calculateCCDF through the gateway (`Task 2/Main_model_Task_2.m:72-82`)."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3000) + 1j * rng.standard_normal(3000)
    assert abs(mex.call("calculatePAPR", x)[0, 0] - O.calculatePAPR(x)) < 1e-4
    w = mex.call("calculate_window_PAPR", x, 256.0)
    assert w.shape == (1, 3000 - 256 + 1) and np.max(np.abs(w.ravel() - O.calculate_window_PAPR(x, 256))) < 1e-4
    xs, cc = mex.call("calculateCCDF", np.round(w, 1), nout=2)
    rx, rc = O.calculateCCDF(np.round(w, 1).astype(np.float32).astype(np.float64))
    assert xs.shape == (rx.size, 1) and cc.shape == (rx.size, 1)
    assert np.max(np.abs(xs.ravel() - rx)) < 1e-6 and np.max(np.abs(cc.ravel() - rc)) < 1e-6


def test_mex_ofdm_and_channel_estimation(mex):
    rng = np.random.default_rng(2)
    p = OC.params_task5(comb=16)
    bits = rng.integers(0, 2, p.stream_bits).astype(np.uint8)
    iq, _ = O.mapping(OC.scramble_frames(p, bits, fast=True), "16QAM")
    grid = mex.call("OFDM_map_carriers", iq, p.N_symb, p.Nfft, p.dataCarriers.astype(float), p.pilotCarriers.astype(float), p.pilotValues)
    grid_ref = O.OFDM_map_carriers(iq, p.N_symb, p.Nfft, p.dataCarriers, p.pilotCarriers, p.pilotValues)
    assert grid.shape == (4096, 14) and np.max(np.abs(grid - grid_ref)) < 1e-6
    tx = mex.call("OFDM_modulator", grid_ref, p.T_Guard)
    assert tx.shape == (4608, 14) and rel(tx, O.OFDM_modulator(grid_ref, p.T_Guard)) < 2e-6
    rx = OC.channel_task5(p, O.OFDM_modulator(grid_ref, p.T_Guard).ravel(order="F"), 20.0, [[0, 1], [4, .8], [10, .6]], rng=rng)
    X = rx.reshape((4608, 14), order="F")
    Y = mex.call("OFDM_demodulator", X, p.T_Guard)
    Y_ref = O.OFDM_demodulator(X, p.T_Guard)
    assert rel(Y, Y_ref) < 2e-6
    H = mex.call("LS_CE", Y_ref, p.pilotValues, p.pilotCarriers.astype(float), p.N_carrier)
    H_ref = O.LS_CE(Y_ref, p.pilotValues, p.pilotCarriers, p.N_carrier)
    assert H.shape == (1, 1024) and rel(H, H_ref) < 2e-5
    Hm = mex.call("MMSE_CE", Y_ref, p.pilotValues, p.pilotCarriers.astype(float), p.Nfft, p.N_carrier, np.fft.ifft(H_ref), 20.0)
    assert rel(Hm, O.MMSE_CE(Y_ref, p.pilotValues, p.pilotCarriers, p.Nfft, p.N_carrier, np.fft.ifft(H_ref), 20.0)) < 5e-5
    eq = mex.call("equalize_signal", Y_ref, H_ref, p.N_carrier)
    assert eq.shape == (4096, 14) and rel(eq, O.equalize_signal(Y_ref, H_ref, p.N_carrier)) < 1e-6 and np.all(eq[1024:] == 0)
    pay = mex.call("get_payload", eq, p.dataCarriers.astype(float))
    assert pay.shape == (len(p.dataCarriers), 14)
    Hi = mex.call("interpolate", H_ref[::16], p.pilotCarriers.astype(float), 1024, "linear")
    assert rel(Hi, O.interpolate(H_ref[::16], p.pilotCarriers, 1024, "linear")) < 1e-6
    A = O.sensing_matrix_dft(p.pilotCarriers, p.Nfft, 256)
    y = Y_ref[p.pilotCarriers - 1, 0] / p.pilotValues[:, 0]
    Ho, ho, idx = mex.call("OMP_estimate", col(y), A, p.Nfft, 5, 20.0, nout=3)
    Ho_ref, ho_ref, idx_ref = O.OMP_estimate(y, A, p.Nfft, 5, 20.0)
    assert list(idx.ravel().astype(int)) == list(idx_ref) and Ho.shape == (1, 4096) and ho.shape == (1, 4096) and rel(Ho, Ho_ref) < 2e-4
    Hp, hp = mex.call("MP_estimate", col(y), A, p.Nfft, 5, nout=2)
    assert Hp.shape == (1, 4096) and hp.shape == (4096, 1) and rel(Hp, O.MP_estimate(y, A, p.Nfft, 5)[0]) < 2e-4


def test_mex_sync_chain_and_errors(mex):
    rng = np.random.default_rng(3)
    p = OC.params_task4()
    bits = rng.integers(0, 2, p.stream_bits).astype(np.uint8)
    tx, _, _ = OC.tx_chain(p, bits, fast=True)
    rx = OC.impair_task4(p, tx, SNR_dB=30, Time_Delay=37, Freq_Shift=7.24, rng=rng)
    ac, tg, fo = mex.call("AutoCorrFunction", col(rx), p.T_Guard, p.Nfft, nout=3)
    ac_ref, tg_ref, fo_ref = O.AutoCorrFunction(rx, p.T_Guard, p.Nfft)
    assert ac.shape == (1, 56448) and tg[0, 0] == tg_ref and abs(fo[0, 0] - fo_ref) < 2e-5
    r = mex.call("add_STO", col(rx), tg_ref)
    r = mex.call("add_STO", r, -(p.Nfft + p.T_Guard))
    r = mex.call("add_CFO", r, -fo_ref, p.Nfft)
    r_ref = O.add_CFO(O.add_STO(O.add_STO(rx, tg_ref), -(p.Nfft + p.T_Guard)), -fo_ref, p.Nfft)
    assert r.shape == (57600, 1) and rel(r, r_ref) < 1e-6
    fixed, ifo = mex.call("remove_IFO", col(r_ref), p.Nfft, nout=2)
    assert ifo[0, 0] == 7 and rel(fixed, O.remove_IFO(r_ref, p.Nfft)[0]) < 1e-6
    Y = O.OFDM_demodulator(O.remove_IFO(r_ref, p.Nfft)[0].reshape((1152, 50), order="F"), 128)
    fs = mex.call("fine_sync", Y, p.pilotCarriers.astype(float), p.pilotValues, 1, 1)
    assert rel(fs, O.fine_sync(Y, p.pilotCarriers, p.pilotValues, 1, 1)) < 2e-4
    He, Hp = mex.call("estimate_channel", Y, np.arange(1, 1025, dtype=float), p.pilotCarriers.astype(float), p.pilotValues, nout=2)
    He_ref, Hp_ref = O.estimate_channel(Y, np.arange(1, 1025), p.pilotCarriers, p.pilotValues)
    assert He.shape == (1, 1024) and Hp.shape == (68, 1) and rel(Hp, Hp_ref) < 1e-6 and rel(He[0, :400], He_ref[:400]) < 2e-5
    h, Hf = mex.call("get_MP_channel_resp", np.array([[0, 1], [4, .6], [10, .3]]), 1024, nout=2)
    assert h.shape == (1, 11) and rel(Hf, O.get_MP_channel_resp([[0, 1], [4, .6], [10, .3]], 1024)[1]) < 1e-6
    normals = rng.standard_normal((2, 1000))
    n_out, nvar = mex.call("Noise", 20.0, col(tx[:1000]), normals.T, nout=2)
    n_ref, nvar_ref = O.Noise(20.0, tx[:1000], normals=normals)
    assert rel(n_out, n_ref) < 1e-6 and abs(nvar[0, 0] - nvar_ref) < 1e-9
    with pytest.raises(RuntimeError, match="remove_IFO"):
        mex.call("remove_IFO", col(1e-3 * rx), p.Nfft, nout=2)
    with pytest.raises(RuntimeError, match="unknown operation"):
        mex.call("no_such_function", 1.0)
    with pytest.raises(RuntimeError, match="too few"):
        mex.call("LS_CE", Y)


def _link_args(p):
    """The ten positional LINK arguments of the batched gateway ops (what matlab/ofdm_link.m builds from the script's struct)."""
    return [float(p.Nfft), float(p.T_Guard), float(p.N_carrier), float(p.N_symb), float(p.Amount_ODFM_SpF), p.Constellation,
            p.dataCarriers.astype(float), p.pilotCarriers.astype(float), p.pilotValues, O.DEFAULT_REGISTER.astype(float)]


def test_mex_batched_chain_ops_against_oracle(mex):
    """tx_chain / channel_t5 / rx_chain_t5 through the gateway with B = 3 streams as columns (the batched form of the loop
    `Task 5/Main_model_Task_5.m:303-346`), each stage compared with the oracle."""
    TAPS5 = [[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]]
    p = OC.params_task5(comb=4)
    rng = np.random.default_rng(17)
    B = 3
    bits = rng.integers(0, 2, (p.stream_bits, B)).astype(float)            # one stream per column
    link = _link_args(p)
    tx = mex.call("tx_chain", *link, bits)
    assert tx.shape == (p.stream_len, B)
    refs = [OC.tx_chain(p, bits[:, b], fast=True)[0] for b in range(B)]
    for b in range(B):
        assert rel(tx[:, b], refs[b]) < 2e-6
    h, _ = O.get_MP_channel_resp(TAPS5, p.Nfft)
    clean = mex.call("channel_t5", tx, np.zeros((0, 0)), h, 0.0)            # no noise: pure multipath, comparable with the oracle
    for b in range(B):
        assert rel(clean[:, b], O.apply_channel(tx[:, b], h)) < 2e-6
    rx = mex.call("channel_t5", tx, 18.0, h, 5.0)                          # Philox noise, seed 5
    snr_meas = 10 * np.log10(np.mean(np.abs(tx) ** 2) / np.mean(np.abs(rx - clean) ** 2 / np.sum(np.abs(h) ** 2)))
    assert abs(snr_meas - 18.0) < 0.2
    out_bits, H, counts = mex.call("rx_chain_t5", *link, rx, bits, 1e-3, nout=3)
    assert out_bits.shape == (p.stream_bits, B) and H.shape == (p.N_carrier, B) and counts.shape == (1, 3)
    mism, ref_err = 0, 0
    for b in range(B):
        ref = OC.rx_chain_task5(p, rx[:, b], bits[:, b], fast=True)         # the oracle on the very samples the gateway returned
        assert rel(H[:, b], ref["H"]) < 2e-5
        mism += int(np.sum(out_bits[:, b] != ref["bits"]))
        ref_err += ref["errors"]
    assert counts[0, 1] == B * p.stream_bits and counts[0, 0] == np.sum(out_bits != bits)
    assert mism <= 3 * p.bps * counts[0, 2] and abs(counts[0, 0] - ref_err) <= mism


def test_mex_channel_t4_against_oracle(mex):
    """The fused Task-4 channel through the gateway (B = 3 columns) against the oracle's add_STO / add_CFO / conv at an SNR
    where the noise is far below the comparison tolerance, and its noise level at 15 dB."""
    rng = np.random.default_rng(23)
    L, B, Nfft = 6000, 3, 1024
    tx = rng.standard_normal((L, B)) + 1j * rng.standard_normal((L, B))
    h, _ = O.get_MP_channel_resp([[0, 1], [4, .6], [10, .3]], Nfft)
    sto, cfo = np.array([[37.0, 0.0, 900.0]]), np.array([[7.24, -0.3, 12.5]])
    clean = mex.call("channel_t4", tx, 200.0, sto, cfo, float(Nfft), h, 3.0)
    for b in range(B):
        want = O.apply_channel(O.add_CFO(O.add_STO(tx[:, b], sto[0, b]), cfo[0, b], Nfft), h)
        assert rel(clean[:, b], want) < 2e-6
    noisy = mex.call("channel_t4", tx, 15.0, sto, cfo, float(Nfft), h, 3.0)
    for b in range(B):
        keep = slice(0, L - int(sto[0, b]) - len(h))                       # where the shifted stream has samples
        snr = 10 * np.log10(np.mean(np.abs(tx[:, b]) ** 2) / np.mean(np.abs((noisy - clean)[keep, b]) ** 2 / np.sum(np.abs(h) ** 2)))
        assert abs(snr - 15.0) < 0.5


def test_mex_sweep_ber(mex):
    """sweep_ber through the gateway == ofdm_sweep_ber through the ctypes binding (same seeds), both chains."""
    import ofdm_b200 as G
    from ofdm_b200 import sweep
    ctx = G.default_context("f32")
    for chain, p, taps, snrs, spp in (("task5", OC.params_task5(comb=4), [[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]], [5.0, 15.0], 6),
                                      ("task4", OC.params_task4(), [[0, 1], [4, .6], [10, .3]], [10.0, 30.0], 5)):
        got = mex.call("sweep_ber", *_link_args(p), np.asarray(snrs), float(spp), np.asarray(taps, dtype=float), chain, 3.0, 1e-3)
        lp = ctx.link_params(p.Nfft, p.T_Guard, p.N_carrier, p.N_symb, p.Amount_ODFM_SpF, p.Constellation, p.dataCarriers, p.pilotCarriers, p.pilotValues)
        want = sweep.ber_sweep(ctx, lp, snrs, spp, taps, chain, seed=3, near_eps=1e-3)
        assert got.shape == (len(snrs), 4) and np.array_equal(got.astype(np.int64), want)
        assert np.all(got[:, 1] == spp * p.stream_bits) and got[0, 0] > got[1, 0]
