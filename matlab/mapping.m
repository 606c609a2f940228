function [IQ,pad] = mapping(bits, constellation)
%MAPPING  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/mapping.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
%   bits is a column of 0/1 doubles; IQ is 1-by-N, pad = -1 when nothing was padded.
    [IQ,pad] = ofdm_mex('mapping', bits, char(constellation));
end
