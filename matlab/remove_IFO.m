function [fixed_rx_signal, IFO] = remove_IFO(rx_signal, Nfft)
%REMOVE_IFO  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/remove_IFO.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [fixed_rx_signal, IFO] = ofdm_mex('remove_IFO', rx_signal, Nfft);
end
