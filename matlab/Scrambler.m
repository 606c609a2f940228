function [sc_sequence, Register] = Scrambler(Register, sequence)
%SCRAMBLER  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/Scrambler.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [sc_sequence, Register] = ofdm_mex('Scrambler', Register, sequence);
end
