function bw = imbinarize(I)
%IMBINARIZE  Shim (Image Processing Toolbox absent): global Otsu threshold on the 256-bin histogram, pixels strictly
%   above it are true -- the default the reference relies on (`Task 5/file_reader.m:7`).  Same rule as
%   ofdm_b200.realisations.imbinarize (graythresh: mean of the maximising bins, (idx-1)/255).
    I = double(I);
    counts = accumarray(I(:) + 1, 1, [256 1]);
    p = counts / sum(counts);
    omega = cumsum(p);
    mu = cumsum(p .* (1:256).');
    sigma_b2 = (mu(end) * omega - mu) .^ 2 ./ (omega .* (1 - omega));
    sigma_b2(~isfinite(sigma_b2)) = -Inf;
    idx = find(sigma_b2 == max(sigma_b2));
    level = (mean(idx) - 1) / 255;
    bw = (I / 255) > level;
end
