#!/usr/bin/env python
"""Generate matlab/<RefName>.m: wrappers with the reference's exact signatures (cited) that forward to
the MEX gateway.  Run from the repo root; the output is committed."""
import os

W = [
    ("Scrambler", "[sc_sequence, Register]", "Register, sequence", "Task 5/Scrambler.m:1"),
    ("DeScrambler", "[dsc_sequence, Register]", "Register, sequence", "Task 5/DeScrambler.m:1"),
    ("constellation_func", "[Dictionary, Bit_depth_Dict]", "Constellation", "Task 5/constellation_func.m:4"),
    ("mapping", "[IQ,pad]", "bits, constellation", "Task 5/mapping.m:1"),
    ("demapping", "[de_bits]", "pad, IQ, Constellation", "Task 5/demapping.m:1"),
    ("OFDM_map_carriers", "mapped_carriers", "QAM_payload, N_symb, Nfft, dataCarriers, pilotCarriers, pilotValues", "Task 5/OFDM_map_carriers.m:2"),
    ("OFDM_modulator", "OFDM_time_guarded", "OFDM_symbols, T_guard", "Task 5/OFDM_modulator.m:2"),
    ("OFDM_demodulator", "TX_IQ", "OFDM_time_guarded, T_guard", "Task 5/OFDM_demodulator.m:2"),
    ("get_payload", "RX_IQ", "RX_OFDM_symbols, dataCarriers", "Task 5/get_payload.m:2"),
    ("add_STO", "y_STO", "y, nSTO", "Task 5/add_STO.m:1"),
    ("add_CFO", "y_CFO", "y, CFO, Nfft", "Task 5/add_CFO.m:1"),
    ("Noise", "[IQ_RX, N_var]", "SNR, IQ_TX, varargin", "Task 5/Noise.m:1"),
    ("get_MP_channel_resp", "[impulse_response,frequency_response]", "channel_taps, Nfft", "Task 5/get_MP_channel_resp.m:2"),
    ("AutoCorrFunction", "[AutoCorr, TgPosition, FreqOffset]", "RxSignal, WidthWindow, Nfft", "Task 5/AutoCorrFunction.m:1"),
    ("remove_IFO", "[fixed_rx_signal, IFO]", "rx_signal, Nfft", "Task 5/remove_IFO.m:1"),
    ("fine_sync", "sync_signal", "rx_signal, pilotCarriers, pilotValues, time_desync, freq_desync", "Task 4/fine_sync.m:1"),
    ("estimate_channel", "[H_est,Hest_at_pilots]", "rx_signal, allCarriers, pilotCarriers, pilotValues", "Task 5/estimate_channel.m:1"),
    ("LS_CE", "[H_LS]", "Y, Xp, pilot_loc, N_carrier", "Task 5/LS_CE.m:1"),
    ("MMSE_CE", "[H_MMSE]", "Y, Xp, pilot_loc, Nfft, N_carrier, h, SNR", "Task 5/MMSE_CE.m:1"),
    ("interpolate", "[H_interpolated]", "H, pilot_loc, Nfft, method", "Task 5/interpolate.m:1"),
    ("equalize_signal", "equalized_Hest", "OFDM_demod, Hest, N_carrier", "Task 5/equalize_signal.m:1"),
    ("OMP_estimate", "[H_OMP,h_impulse_est,index]", "Y, sensing_matrix, Nfft, dominant_taps, SNR_dB", "Task 5/OMP_estimate.m:2"),
    ("MP_estimate", "[H_MP,h_impulse_est]", "Y, sensing_matrix, Nfft, dominant_taps", "Task 5/MP_estimate.m:2"),
    ("BER_func", "[BER]", "Bit_Tx, Bit_Rx", "Task 5/BER_func.m:1"),
    ("MER_func", "[MER]", "IQ_RX, Constellation", "Task 5/MER_func.m:1"),
    ("calculatePAPR", "PAPR", "OFDM_signal", "Task 5/calculatePAPR.m:2"),
    ("calculate_window_PAPR", "PAPRs", "Tx_OFDM_Signal, Nfft", "Task 5/calculate_window_PAPR.m:2"),
    ("calculateCCDF", "[PAPR_ccdf, CCDF]", "PAPR_values", "Task 5/calculateCCDF.m:2"),
]
NOTES = {
    "Noise": "%   Optional third argument: an L-by-2 matrix of unit normals (column 1 real part, column 2 imaginary\n"
             "%   part, the order of the reference's two normrnd calls) to share a realisation, or a scalar Philox seed.\n",
    "mapping": "%   bits is a column of 0/1 doubles; IQ is 1-by-N, pad = -1 when nothing was padded.\n",
    "OFDM_map_carriers": "%   pilotValues: Np-by-N_symb matrix, or a scalar (broadcast, as `Task 3/Main_model_Task_3.m:59` does).\n",
}
os.makedirs("matlab", exist_ok=True)
for name, outs, ins, cite in W:
    call_ins = ins.replace("varargin", "varargin{:}")
    if name in ("mapping", "demapping", "MER_func", "constellation_func"):
        # MATLAB string scalars ("16QAM") -> char for the gateway
        var = {"mapping": "constellation", "demapping": "Constellation", "MER_func": "Constellation", "constellation_func": "Constellation"}[name]
        call_ins = call_ins.replace(var, f"char({var})")
    if name == "interpolate":
        call_ins = call_ins.replace("method", "char(method)")
    n_out = outs.count(",") + 1
    lhs = outs if outs.startswith("[") else f"[{outs}]" if n_out > 1 else outs
    body = (f"function {outs} = {name}({ins})\n"
            f"%{name.upper()}  GPU (libofdm_b200, sm_100a) drop-in for `{cite}` of ladnlav/OFDM-course.\n"
            f"%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.\n"
            f"{NOTES.get(name, '')}"
            f"    {lhs} = ofdm_mex('{name}', {call_ins});\n"
            f"end\n")
    open(os.path.join("matlab", name + ".m"), "w").write(body)
open("matlab/OFDM_map_carriers_v1.m", "w").write(
    "function mapped_carriers = OFDM_map_carriers_v1(QAM_payload, N_symb, Nfft, dataCarriers, pilotCarriers, amp_pilots)\n"
    "%OFDM_MAP_CARRIERS_V1  Task-1/2 variant (`Task 1/OFDM_map_carriers.m:2`): alternating +a / a*exp(1i*pi) pilots,\n"
    "%   repmat(...,1,50).  Rename to OFDM_map_carriers.m when running the Task 1-2 scripts.\n"
    "    pv = zeros(1, length(pilotCarriers)); pv(1:2:end) = amp_pilots*exp(1i*0); pv(2:2:end) = amp_pilots*exp(1i*pi);\n"
    "    mapped_carriers = ofdm_mex('OFDM_map_carriers', QAM_payload, N_symb, Nfft, dataCarriers, pilotCarriers, repmat(pv', 1, 50));\n"
    "end\n")
open("matlab/apply_channel.m", "w").write(
    "function y = apply_channel(x, h)\n"
    "%APPLY_CHANNEL  conv(x, h.', 'full') truncated to length(x) (`Task 5/Main_model_Task_5.m:126-127`).\n"
    "    y = ofdm_mex('apply_channel', x, h);\n"
    "end\n")
print(len(os.listdir("matlab")), "wrappers")
