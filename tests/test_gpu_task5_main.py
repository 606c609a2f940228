"""`Task 5/Main_model_Task_5.m` (estimator part) through the reference-named drop-ins against the oracle on a shared
noise realisation: the four channel estimates, their MSE and the taps OMP selects (the six delays of
`Main_model_Task_5.m:112-119`)."""
import importlib.util
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _example():
    spec = importlib.util.spec_from_file_location("main_model_task5", os.path.join(ROOT, "examples", "main_model_task5.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("prec,tol", [("f64", 1e-8), ("f32", 3e-4)])
def test_task5_main_model_estimators(prec, tol):
    m = _example()
    rng = np.random.default_rng(12)
    L = 14 * (4096 + 512)
    normals = rng.standard_normal((2, L))
    r = m.run(SNR_dB=20.0, normals=normals, precision=prec)
    # oracle, same script
    Nfft, Nc, Tg, S = 4096, 1024, 512, 14
    pc = np.arange(1, Nc + 1)
    amp = 4.0 / 3.0 * np.max(np.abs(O.constellation_func("16QAM")[0]))
    pv = np.tile(np.conj(np.full(Nc, amp * np.exp(1j * 0)))[:, None], (1, S))
    grid = np.zeros((Nfft, S), dtype=np.complex128); grid[pc - 1, :] = pv
    tx = O.OFDM_modulator(grid, Tg).ravel(order="F")
    assert np.linalg.norm(r["Tx"] - tx) / np.linalg.norm(tx) < (1e-12 if prec == "f64" else 3e-6)
    rx, _ = O.Noise(20.0, tx, normals=normals)
    taps = [[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]]
    h, Hf = O.get_MP_channel_resp(taps, Nfft)
    rx = O.apply_channel(rx, h)
    Y = O.OFDM_demodulator(rx.reshape((Nfft + Tg, S), order="F"), Tg)
    H_ls = O.LS_CE(Y, pv, pc, Nc)
    H_mmse = O.MMSE_CE(Y, pv, pc, Nfft, Nc, np.fft.ifft(H_ls), 20.0)
    A = O.sensing_matrix_dft(pc, Nfft, Nc)
    y = Y[pc - 1, 0] / amp
    H_mp, _ = O.MP_estimate(y, A, Nfft, 6)
    H_omp, _, idx = O.OMP_estimate(y, A, Nfft, 6, 20.0)
    ref = {"LS": H_ls, "MMSE": H_mmse, "MP": np.asarray(H_mp).ravel(), "OMP": np.asarray(H_omp).ravel()}
    for k, Hr in ref.items():
        Hd = np.asarray(r["H"][k]).ravel()[:Nc]
        assert np.linalg.norm(Hd - Hr[:Nc]) / np.linalg.norm(Hr[:Nc]) < tol, k
    assert np.array_equal(np.asarray(r["omp_index"]).ravel(), np.asarray(idx).ravel())            # tap indices exact
    assert sorted(np.asarray(idx).ravel().tolist()) == [1, 5, 11, 16, 22, 26]                        # delays 0 4 10 15 21 25, 1-based
    assert r["MSE_omp"] < r["MSE_l"]                    # the script's conclusion: the sparse estimate beats LS at 20 dB
