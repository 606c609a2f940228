function PAPR = calculatePAPR(OFDM_signal)
%CALCULATEPAPR  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/calculatePAPR.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    PAPR = ofdm_mex('calculatePAPR', OFDM_signal);
end
