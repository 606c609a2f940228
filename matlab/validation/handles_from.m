function h = handles_from(dirpath, names)
%HANDLES_FROM  Function handles bound to the implementations found in DIRPATH (the reference's "Task 5" folder or this
%   repository's matlab/ folder): both sides use the same function names, so the handles are created while DIRPATH
%   is the current folder -- a handle keeps the file it resolved to at creation.
    old = cd(dirpath);
    c = onCleanup(@() cd(old));
    h = struct();
    for k = 1:numel(names)
        h.(names{k}) = str2func(names{k});
    end
end
