function [PAPR_ccdf, CCDF] = calculateCCDF(PAPR_values)
%CALCULATECCDF  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/calculateCCDF.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [PAPR_ccdf, CCDF] = ofdm_mex('calculateCCDF', PAPR_values);
end
