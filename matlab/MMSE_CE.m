function [H_MMSE] = MMSE_CE(Y, Xp, pilot_loc, Nfft, N_carrier, h, SNR)
%MMSE_CE  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/MMSE_CE.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [H_MMSE] = ofdm_mex('MMSE_CE', Y, Xp, pilot_loc, Nfft, N_carrier, h, SNR);
end
