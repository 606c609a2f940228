function mapped_carriers = OFDM_map_carriers(QAM_payload, N_symb, Nfft, dataCarriers, pilotCarriers, pilotValues)
%OFDM_MAP_CARRIERS  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/OFDM_map_carriers.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
%   pilotValues: Np-by-N_symb matrix, or a scalar (broadcast, as `Task 3/Main_model_Task_3.m:59` does).
    mapped_carriers = ofdm_mex('OFDM_map_carriers', QAM_payload, N_symb, Nfft, dataCarriers, pilotCarriers, pilotValues);
end
