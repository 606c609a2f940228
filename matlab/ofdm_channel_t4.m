function Rx = ofdm_channel_t4(Tx, SNR_dB, nSTO, CFO, Nfft, h, seed)
%OFDM_CHANNEL_T4  Noise -> add_STO -> add_CFO -> multipath for B streams in one pass
%   (`Task 4/Main_model_Task_4.m:95,103,110,263-264`; the same samples as Noise, add_STO, add_CFO and conv called in turn).
%   Tx: L x B; SNR_dB, nSTO, CFO: scalars or 1 x B; h: impulse response from get_MP_channel_resp (1 for no multipath);
%   seed: Philox seed of the noise (stream b uses the counter stream (seed, b-1)).
    if nargin < 7, seed = 0; end
    Rx = ofdm_mex('channel_t4', Tx, SNR_dB, nSTO, CFO, Nfft, h, seed);
end
