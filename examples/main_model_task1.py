#!/usr/bin/env python
"""`Task 1/Main_model.m` line by line on the GPU drop-ins (BASELINE config 1: the reference's own CPU-runnable case).

The script body is the reference's, with MATLAB calls replaced by the functions of the same name from ``ofdm_b200``
(each forwards to the C ABI / CUDA kernels; there is no CPU fallback).  Plotting is left out.

    python examples/main_model_task1.py [path/to/eagle.tiff]      # without a file: a synthetic payload of the same size
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(input_bits=None, File=None, SNR_dB=None, seed=1):
    import ofdm_b200 as G
    from ofdm_b200 import realisations as R
    # Main_model.m:6-24
    Nfft = 1024; N_carrier = 400; T_Guard = Nfft // 8
    Amount_OFDM_Frames = 10; Amount_ODFM_SpF = 5
    N_symb = Amount_OFDM_Frames * Amount_ODFM_SpF
    Percent_pilot = 25
    allCarriers = np.arange(1, Nfft + 1)
    amount_pilots = int(np.floor(Percent_pilot / 100 * N_carrier + 0.5))
    pilot_step = N_carrier // amount_pilots
    pilotCarriers = np.concatenate([allCarriers[0:N_carrier - 2:pilot_step], [allCarriers[N_carrier - 1]]])
    dataCarriers = allCarriers[:N_carrier][~np.isin(allCarriers[:N_carrier], pilotCarriers)]
    # :26-33
    Constellation = "16QAM"
    dict_, bps = G.constellation_func(Constellation)
    Size_Buffer = Amount_ODFM_SpF * Amount_OFDM_Frames * len(dataCarriers) * bps
    if input_bits is None:
        input_bits = R.file_reader(File, Size_Buffer) if File else (np.random.default_rng(seed).random(Size_Buffer) < 0.337).astype(np.uint8)
    input_bits = np.asarray(input_bits).ravel()[:Size_Buffer]
    # :35-45
    TX_IQ, pad = G.mapping(input_bits, Constellation)
    amp_pilots = 2 * np.max(np.abs(dict_))
    OFDM_mapped_carriers = G.OFDM_map_carriers_v1(TX_IQ, N_symb, Nfft, dataCarriers, pilotCarriers, amp_pilots)
    Tx_OFDM_Signal_matrix = G.OFDM_modulator(OFDM_mapped_carriers, T_Guard)
    Tx_OFDM_Signal = Tx_OFDM_Signal_matrix.ravel(order="F")
    PAPR = G.calculatePAPR(Tx_OFDM_Signal)
    # channel: none in Task 1 (:63); an AWGN option mirrors the later tasks
    Rx_OFDM_Signal = Tx_OFDM_Signal if SNR_dB is None else G.Noise(SNR_dB, Tx_OFDM_Signal, seed=seed)[0]
    # :66-88
    Rx = np.asarray(Rx_OFDM_Signal).reshape((Nfft + T_Guard, N_symb), order="F")
    RX_OFDM_mapped_carriers = G.OFDM_demodulator(Rx, T_Guard)
    RX_IQ = G.get_payload(RX_OFDM_mapped_carriers, dataCarriers).ravel(order="F")
    output_bits = G.demapping(pad, RX_IQ, Constellation)
    # :90-104
    BER = G.BER_func(input_bits, output_bits)
    MER = G.MER_func(RX_IQ, Constellation) if SNR_dB is not None else np.inf
    ok = bool(np.array_equal(np.asarray(output_bits).ravel(), input_bits))
    return {"ok": ok, "BER": BER, "MER_dB": MER, "PAPR_dB": PAPR, "n_bits": int(input_bits.size), "pilots": len(pilotCarriers), "picture": R.display_pic(output_bits)}


if __name__ == "__main__":
    r = run(File=sys.argv[1] if len(sys.argv) > 1 else None)
    print(("Проверка пройдена!" if r["ok"] else "Проверка НЕ пройдена!"), f"BER={r['BER']}  PAPR={r['PAPR_dB']:.2f} dB  bits={r['n_bits']}  pilots={r['pilots']}")
