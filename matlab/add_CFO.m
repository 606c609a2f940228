function y_CFO = add_CFO(y, CFO, Nfft)
%ADD_CFO  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/add_CFO.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    y_CFO = ofdm_mex('add_CFO', y, CFO, Nfft);
end
