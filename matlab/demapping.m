function [de_bits] = demapping(pad, IQ, Constellation)
%DEMAPPING  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/demapping.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [de_bits] = ofdm_mex('demapping', pad, IQ, char(Constellation));
end
