function Rx = ofdm_channel_t5(Tx, SNR_dB, h, seed)
%OFDM_CHANNEL_T5  Noise then multipath for B streams (`Task 5/Main_model_Task_5.m:108,123-127`, one fused kernel).
%   Tx: L x B; SNR_dB: scalar, 1 x B or [] (no noise); h: impulse response from get_MP_channel_resp or [] (no multipath);
%   seed: Philox seed of the noise (stream b uses the counter stream (seed, b-1)).
    if nargin < 4, seed = 0; end
    Rx = ofdm_mex('channel_t5', Tx, SNR_dB, h, seed);
end
