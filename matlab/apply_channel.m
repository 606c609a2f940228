function y = apply_channel(x, h)
%APPLY_CHANNEL  conv(x, h.', 'full') truncated to length(x) (`Task 5/Main_model_Task_5.m:126-127`).
    y = ofdm_mex('apply_channel', x, h);
end
