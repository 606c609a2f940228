function equalized_Hest = equalize_signal(OFDM_demod, Hest, N_carrier)
%EQUALIZE_SIGNAL  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/equalize_signal.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    equalized_Hest = ofdm_mex('equalize_signal', OFDM_demod, Hest, N_carrier);
end
