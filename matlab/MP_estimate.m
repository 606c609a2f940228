function [H_MP,h_impulse_est] = MP_estimate(Y, sensing_matrix, Nfft, dominant_taps)
%MP_ESTIMATE  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/MP_estimate.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [H_MP,h_impulse_est] = ofdm_mex('MP_estimate', Y, sensing_matrix, Nfft, dominant_taps);
end
