function [AutoCorr, TgPosition, FreqOffset] = AutoCorrFunction(RxSignal, WidthWindow, Nfft)
%AUTOCORRFUNCTION  GPU (libofdm_b200, sm_100a) drop-in for `Task 5/AutoCorrFunction.m:1` of ladnlav/OFDM-course.
%   Same signature, shapes and orientation as the reference; forwards to the MEX gateway.
    [AutoCorr, TgPosition, FreqOffset] = ofdm_mex('AutoCorrFunction', RxSignal, WidthWindow, Nfft);
end
