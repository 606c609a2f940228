function [MER] = MER_func(IQ_RX, Constellation)
%MER_FUNC  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/MER_func.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [MER] = ofdm_mex('MER_func', IQ_RX, char(Constellation));
end
