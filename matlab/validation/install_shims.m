function added = install_shims()
%INSTALL_SHIMS  Put matlab/validation/shims on the path for every toolbox function the reference needs and this host
%   lacks (int2bit, bi2de, normrnd, dftmtx, imbinarize).  Functions the host already has are left alone: each shim
%   lives in its own check, and the shim directory is appended to the END of the path.
    here = fileparts(mfilename('fullpath'));
    names = {'int2bit', 'bi2de', 'normrnd', 'dftmtx', 'imbinarize'};
    added = {};
    for k = 1:numel(names)
        if ~(exist(names{k}, 'file') || exist(names{k}, 'builtin')), added{end + 1} = names{k}; end %#ok<AGROW>
    end
    if ~isempty(added), addpath(fullfile(here, 'shims'), '-end'); end
    fprintf('install_shims: %d shim(s) in use: %s\n', numel(added), strjoin(added, ', '));
end
