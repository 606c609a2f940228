function y_STO = add_STO(y, nSTO)
%ADD_STO  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/add_STO.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    y_STO = ofdm_mex('add_STO', y, nSTO);
end
