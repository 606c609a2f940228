% Task-5 main loop, batched: what `Task 5/Main_model_Task_5.m:303-346` (BER over SNR for the LS estimator) does one
% stream and one SNR point at a time, here B streams per call and then the whole sweep in one call.
% Needs ofdm_mex on the path (see INTEGRATION.md) and an sm_100 GPU; not executable in the authoring image.
P.Nfft = 4096; P.N_carrier = 1024; P.T_Guard = P.Nfft / 8;
P.Amount_OFDM_Frames = 2; P.Amount_ODFM_SpF = 7; P.N_symb = P.Amount_OFDM_Frames * P.Amount_ODFM_SpF;
comb = 4;
P.pilotCarriers = 1:comb:P.N_carrier;                                   % `Main_model_Task_5.m:18-22`
P.dataCarriers = setdiff(1:P.N_carrier, P.pilotCarriers);
P.Constellation = "16QAM";
[dict, bps] = constellation_func(P.Constellation);
amp_pilots = 2 * max(abs(dict));                                        % `Task5_part2.m:86-91`
pv = zeros(1, numel(P.pilotCarriers)); pv(1:2:end) = amp_pilots * exp(1i * 0); pv(2:2:end) = amp_pilots * exp(1i * pi);
P.pilotValues = repmat(pv', 1, P.N_symb);
P.Register = [1 0 0 1 0 1 0 1 0 0 0 0 0 0 0];
channel_taps = [0 1; 4 .8; 10 .6; 15 .4; 21 .2; 25 .1];                 % `Main_model_Task_5.m:112-119`
[h, ~] = get_MP_channel_resp(channel_taps, P.Nfft);

% (1) explicit batch: B streams through TX -> channel -> RX, three calls
B = 64; stream_bits = P.N_symb * numel(P.dataCarriers) * bps;
bits = double(rand(stream_bits, B) > 0.5);
Tx = ofdm_tx_chain(P, bits);
Rx = ofdm_channel_t5(Tx, 20, h, 1);
[rx_bits, H, counts] = ofdm_rx_chain_t5(P, Rx, bits, 1e-4);
fprintf('B = %d streams at 20 dB: BER = %g (%d near-boundary symbols)
', B, counts(1) / counts(2), counts(3));

% (2) the whole Monte-Carlo sweep in one call (payload, noise and channel generated on the GPU)
SNRs = 0:0.5:30;                                                        % `Task 3/Main_model_Task_3.m:192`
c = ofdm_sweep_ber(P, SNRs, 1024, channel_taps, 'task5', 1, 1e-4);
semilogy(SNRs, c(:, 1) ./ c(:, 2), 'LineWidth', 2); grid on; xlabel('SNR (dB)'); ylabel('BER');
