function U = cmtf_nvecs(Z,n,r)
% Drop-in replacement of functions/cmtf_nvecs.m (called by init_coupled_AOADMM_CMTF.m:52 when init_options.nvecs = 1):
% the r leading eigenvectors of X_(n)*X_(n)' for the object that contains mode n, computed on the GPU through
% aoadmm_nvecs_mex (include/aoadmm.h: aoadmm_nvecs) without forming the unfolding on the host.
% Columns come in descending eigenvalue order; eigs leaves their sign open, here the entry of largest magnitude of each
% column is positive.  Sparse tensors are not supported (aoadmm:unsupported).
    P = length(Z.object);
    for p = 1:P
        i = find(Z.modes{p} == n, 1);
        if ~isempty(i)
            X = Z.object{p};
            if isa(X,'tensor'), X = X.data; end
            if ~isa(X,'double') || issparse(X)
                error('aoadmm:unsupported','cmtf_nvecs on the GPU needs a dense double object');
            end
            U = aoadmm_nvecs_mex(X, i, r);
            return
        end
    end
    error('Mode %d is not part of any object.', n);
end
