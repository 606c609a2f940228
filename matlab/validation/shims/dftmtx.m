function F = dftmtx(n)
%DFTMTX  Shim (Signal Processing Toolbox / Octave signal package absent): fft(eye(n)) (`Task 5/Main_model_Task_5.m:182`).
    F = fft(eye(n));
end
