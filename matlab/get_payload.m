function RX_IQ = get_payload(RX_OFDM_symbols, dataCarriers)
%GET_PAYLOAD  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/get_payload.m:2` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    RX_IQ = ofdm_mex('get_payload', RX_OFDM_symbols, dataCarriers);
end
