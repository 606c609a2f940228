function sync_signal = fine_sync(rx_signal, pilotCarriers, pilotValues, time_desync, freq_desync)
%FINE_SYNC  GPU (libofdm_b200, sm_100a) drop-in for `Task 4/fine_sync.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    sync_signal = ofdm_mex('fine_sync', rx_signal, pilotCarriers, pilotValues, time_desync, freq_desync);
end
