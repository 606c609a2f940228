function bits = int2bit(x, n)
%INT2BIT  Shim for hosts without it (Octave; MATLAB < R2021b): n-bit MSB-first binary expansion down the columns,
%   as used by the reference's `Task 5/demapping.m:15` (x a row vector -> n-by-numel(x) matrix).
    x = double(x(:)).';
    bits = zeros(n, numel(x));
    for k = 1:n
        bits(k, :) = mod(floor(x ./ 2^(n - k)), 2);
    end
end
