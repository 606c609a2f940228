function [dsc_sequence, Register] = DeScrambler(Register, sequence)
%DESCRAMBLER  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/DeScrambler.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [dsc_sequence, Register] = ofdm_mex('DeScrambler', Register, sequence);
end
