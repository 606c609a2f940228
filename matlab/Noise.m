function [IQ_RX, N_var] = Noise(SNR, IQ_TX, varargin)
%NOISE  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/Noise.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
%   Optional third argument: an L-by-2 matrix of unit normals (column 1 real part, column 2 imaginary
%   part, the order of the reference's two normrnd calls) to share a realisation, or a scalar Philox seed.
    [IQ_RX, N_var] = ofdm_mex('Noise', SNR, IQ_TX, varargin{:});
end
