function L = ofdm_link(P)
%OFDM_LINK  Pack the link description of a reference script into the positional list the batched MEX ops take.
%   P has the script's own variable names (`Task 5/Main_model_Task_5.m:6-46`): Nfft, T_Guard, N_carrier, N_symb,
%   Amount_ODFM_SpF, Constellation, dataCarriers, pilotCarriers, pilotValues (Np x N_symb or Np x 1), Register.
    L = {P.Nfft, P.T_Guard, P.N_carrier, P.N_symb, P.Amount_ODFM_SpF, char(P.Constellation), ...
         P.dataCarriers, P.pilotCarriers, P.pilotValues, P.Register};
end
