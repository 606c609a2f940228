function mapped_carriers = OFDM_map_carriers_v1(QAM_payload, N_symb, Nfft, dataCarriers, pilotCarriers, amp_pilots)
%OFDM_MAP_CARRIERS_V1  Task-1/2 variant (`Task 1/OFDM_map_carriers.m:2`): alternating +a / a*exp(1i*pi) pilots,
%   repmat(...,1,50).  Rename to OFDM_map_carriers.m when running the Task 1-2 scripts.
    pv = zeros(1, length(pilotCarriers)); pv(1:2:end) = amp_pilots*exp(1i*0); pv(2:2:end) = amp_pilots*exp(1i*pi);
    mapped_carriers = ofdm_mex('OFDM_map_carriers', QAM_payload, N_symb, Nfft, dataCarriers, pilotCarriers, repmat(pv', 1, 50));
end
