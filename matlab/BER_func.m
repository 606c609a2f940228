function [BER] = BER_func(Bit_Tx, Bit_Rx)
%BER_FUNC  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/BER_func.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [BER] = ofdm_mex('BER_func', Bit_Tx, Bit_Rx);
end
