#!/usr/bin/env python
"""`Task 5/Main_model_Task_5.m` (channel-estimation part, lines 6-215) on the GPU drop-ins: pilot-only symbols
(comb 1) through AWGN + the six-tap channel, then LS / MMSE / MP / OMP channel estimates and their MSE against the
true frequency response.  Function names, argument order and 1-based indices are the reference's; plotting is left
out.  ``precision="f64"`` selects the closer-comparison mode.

    python examples/main_model_task5.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(SNR_dB=20.0, normals=None, seed=1, precision="f32"):
    import ofdm_b200 as G
    P = dict(precision=precision)
    # Main_model_Task_5.m:6-46
    Nfft = 4096; N_carrier = 1024; T_Guard = Nfft // 8
    N_symb = 2 * 7
    comb = 1
    allCarriers = np.arange(1, Nfft + 1)
    pilotCarriers = np.concatenate([allCarriers[0:N_carrier - 1:1], [allCarriers[N_carrier - 1]]])      # comb 1: the 100 % branch (:24-33)
    amount_pilots = len(pilotCarriers)
    dict_, bps = G.constellation_func("16QAM")
    amp_pilots = 4.0 / 3.0 * np.max(np.abs(dict_))
    pilotValues = np.tile(np.conj(np.full(amount_pilots, amp_pilots * np.exp(1j * 0)))[:, None], (1, N_symb))
    # :79-85 (comb == 1: pilots only)
    OFDM_mapped_carriers = np.zeros((Nfft, N_symb), dtype=np.complex128)
    OFDM_mapped_carriers[pilotCarriers - 1, :] = pilotValues
    Tx_OFDM_Signal = G.OFDM_modulator(OFDM_mapped_carriers, T_Guard, **P).ravel(order="F")
    # :104-127 channel: noise first, then multipath
    Rx_OFDM_Signal, _ = G.Noise(SNR_dB, Tx_OFDM_Signal, normals=normals, seed=seed, **P)
    channel_taps = [[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]]
    H_tau, H_freq = G.get_MP_channel_resp(channel_taps, Nfft, **P)
    Rx_OFDM_Signal = G.apply_channel(Rx_OFDM_Signal, H_tau, **P)                   # conv(...,'full') cut back to the stream length
    # :150-154
    Rx = np.asarray(Rx_OFDM_Signal).reshape((Nfft + T_Guard, N_symb), order="F")
    RX = G.OFDM_demodulator(Rx, T_Guard, **P)
    # :172-193 estimators
    H_est_LS_l = G.LS_CE(RX, pilotValues, pilotCarriers, N_carrier, **P)
    h_t_mmse = np.fft.ifft(H_est_LS_l)
    H_est_MMSE = G.MMSE_CE(RX, pilotValues, pilotCarriers, Nfft, N_carrier, h_t_mmse, SNR_dB, **P)
    Ldict = -(-N_carrier // comb)
    l = np.arange(Ldict)
    sensing_matrix = np.exp(-2j * np.pi * np.outer(pilotCarriers - 1, l) / Nfft)      # P*F with F = dftmtx(Nfft)(:,1:Ldict)
    Y = RX[pilotCarriers - 1, 0] / amp_pilots
    K = len(channel_taps)
    H_est_MP, h_t_MP = G.MP_estimate(Y, sensing_matrix, Nfft, K, **P)
    H_est_OMP, h_t_OMP, kk1 = G.OMP_estimate(Y, sensing_matrix, Nfft, K, SNR_dB, **P)
    # :195-205
    Hu = np.asarray(H_freq).ravel()[:N_carrier]
    mse = lambda H: float(np.real(np.vdot(Hu - np.asarray(H).ravel()[:N_carrier], Hu - np.asarray(H).ravel()[:N_carrier])) / N_carrier)   # noqa: E731
    return {"MSE_l": mse(H_est_LS_l), "MSE_mmse": mse(H_est_MMSE), "MSE_mp": mse(H_est_MP), "MSE_omp": mse(H_est_OMP), "omp_index": np.asarray(kk1),
            "H": {"LS": H_est_LS_l, "MMSE": H_est_MMSE, "MP": H_est_MP, "OMP": H_est_OMP}, "Tx": Tx_OFDM_Signal}


if __name__ == "__main__":
    r = run()
    print("MSE  LS %.3e  MMSE %.3e  MP %.3e  OMP %.3e   OMP taps (1-based) %s" % (r["MSE_l"], r["MSE_mmse"], r["MSE_mp"], r["MSE_omp"], r["omp_index"].tolist()))
