function OFDM_time_guarded = OFDM_modulator(OFDM_symbols, T_guard)
%OFDM_MODULATOR  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/OFDM_modulator.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    OFDM_time_guarded = ofdm_mex('OFDM_modulator', OFDM_symbols, T_guard);
end
