function TX_IQ = OFDM_demodulator(OFDM_time_guarded, T_guard)
%OFDM_DEMODULATOR  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/OFDM_demodulator.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    TX_IQ = ofdm_mex('OFDM_demodulator', OFDM_time_guarded, T_guard);
end
