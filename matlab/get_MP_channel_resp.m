function [impulse_response,frequency_response] = get_MP_channel_resp(channel_taps, Nfft)
%GET_MP_CHANNEL_RESP  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/get_MP_channel_resp.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [impulse_response,frequency_response] = ofdm_mex('get_MP_channel_resp', channel_taps, Nfft);
end
