function r = normrnd(mu, sigma, varargin)
%NORMRND  Shim (Statistics Toolbox / Octave statistics package absent): mu + sigma*randn(size), the only form the
%   reference uses (`Task 5/Noise.m:7-8`).  MATLAB's own normrnd draws from the same randn stream.
    r = mu + sigma .* randn(varargin{:});
end
