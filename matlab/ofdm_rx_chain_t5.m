function [bits, H, counts] = ofdm_rx_chain_t5(P, Rx, tx_bits, near_eps)
%OFDM_RX_CHAIN_T5  OFDM_demodulator -> LS_CE -> equalize_signal -> get_payload -> demapping -> DeScrambler -> BER count
%   for B streams in one pass (`Task 5/Task5_part2.m:169-174,269-303`).  Rx: L x B host matrix (the library chunks and
%   overlaps the transfers).  bits: stream_bits x B decided bits; H: N_carrier x B channel estimates;
%   counts = [bit errors, bits, symbols within near_eps of a decision boundary] (errors need tx_bits, else pass []).
    if nargin < 3, tx_bits = []; end
    if nargin < 4, near_eps = 0; end
    L = ofdm_link(P);
    [bits, H, counts] = ofdm_mex('rx_chain_t5', L{:}, Rx, tx_bits, near_eps);
end
