function [Dictionary, Bit_depth_Dict] = constellation_func(Constellation)
%CONSTELLATION_FUNC  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/constellation_func.m:4` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [Dictionary, Bit_depth_Dict] = ofdm_mex('constellation_func', char(Constellation));
end
