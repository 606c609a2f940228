function [H_OMP,h_impulse_est,index] = OMP_estimate(Y, sensing_matrix, Nfft, dominant_taps, SNR_dB)
%OMP_ESTIMATE  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/OMP_estimate.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [H_OMP,h_impulse_est,index] = ofdm_mex('OMP_estimate', Y, sensing_matrix, Nfft, dominant_taps, SNR_dB);
end
