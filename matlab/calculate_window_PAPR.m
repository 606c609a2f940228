function PAPRs = calculate_window_PAPR(Tx_OFDM_Signal, Nfft)
%CALCULATE_WINDOW_PAPR  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/calculate_window_PAPR.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    PAPRs = ofdm_mex('calculate_window_PAPR', Tx_OFDM_Signal, Nfft);
end
