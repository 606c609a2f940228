function [H_est,Hest_at_pilots] = estimate_channel(rx_signal, allCarriers, pilotCarriers, pilotValues)
%ESTIMATE_CHANNEL  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/estimate_channel.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [H_est,Hest_at_pilots] = ofdm_mex('estimate_channel', rx_signal, allCarriers, pilotCarriers, pilotValues);
end
