function counts = ofdm_sweep_ber(P, SNRs, streams_per_point, channel_taps, chain, seed, near_eps)
%OFDM_SWEEP_BER  The whole BER-vs-SNR Monte-Carlo loop on the GPU (`Task 3/Main_model_Task_3.m:192-268`,
%   `Task 5/Main_model_Task_5.m:303-346`; chain 'task4' adds the STO / CFO draws and the synchroniser of
%   `Task 4/Main_model_Task_4.m:95-110,277-366`).  counts: numel(SNRs) x 4 =
%   [bit errors, bits, near-boundary symbols, guard-interval detector failures]; BER = counts(:,1)./counts(:,2).
    if nargin < 5, chain = 'task5'; end
    if nargin < 6, seed = 1; end
    if nargin < 7, near_eps = 0; end
    L = ofdm_link(P);
    counts = ofdm_mex('sweep_ber', L{:}, SNRs, streams_per_point, channel_taps, char(chain), seed, near_eps);
end
