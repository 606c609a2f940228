function d = bi2de(b, flag)
%BI2DE  Shim (Communications Toolbox / Octave communications package absent): rows of bits -> decimal, with the
%   'left-msb' orientation the reference uses (`Task 5/mapping.m:18`).
    if nargin < 2, flag = 'right-msb'; end
    n = size(b, 2);
    if strcmp(flag, 'left-msb'), w = 2 .^ (n - 1:-1:0); else, w = 2 .^ (0:n - 1); end
    d = double(b) * w.';
end
