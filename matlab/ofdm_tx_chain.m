function Tx = ofdm_tx_chain(P, bits)
%OFDM_TX_CHAIN  Scrambler (per-frame reset) -> mapping -> OFDM_map_carriers -> OFDM_modulator for B streams at once
%   (`Task 5/Main_model_Task_5.m:53-85` in one fused kernel).  bits: stream_bits x B of 0/1; Tx: (N_symb*(Nfft+T_Guard)) x B,
%   column b = the serial stream Tx_OFDM_Signal of stream b.
    L = ofdm_link(P);
    Tx = ofdm_mex('tx_chain', L{:}, bits);
end
