function make_reference_golden(ref_root, golden_dir, varargin)
% MAKE_REFERENCE_GOLDEN  Pin the parity fixtures of the B200 engine to the REAL reference.
%
%   make_reference_golden(ref_root, golden_dir)
%   make_reference_golden(ref_root, golden_dir, 'b200', true)     % additionally run the B200 drop-in from MATLAB
%
% Runs the UNMODIFIED reference solver (functions/cmtf_AOADMM.m of AOADMM-DataFusionFramework/Matlab-Code, checked
% out at ref_root) on the inputs that tests/golden/make_reference_inputs.py wrote to <golden_dir>/ref_inputs/in_*.mat
% and saves the reference's own outputs as <golden_dir>/ref_<name>.mat:
%     Fac  the final state struct (fac, constraint_fac, constraint_dual_fac, coupling_fac, coupling_dual_fac,
%          P, DeltaB, mu_DeltaB)             - first output of cmtf_fun_AOADMM.m:1 / second of cmtf_AOADMM.m:1
%     out  func_val_conv, func_coupl_conv, func_constr_conv, func_PAR2_coupl, innerIters, OuterIterations,
%          exit_flag, f_* (cmtf_fun_AOADMM.m:480-494)
% `python -m pytest tests/test_reference_mat_goldens.py` then compares the CPU oracle (1e-9) and the CUDA engine
% (factors 1e-8, objective 1e-10) with those files.  Nothing else of this repository runs inside MATLAB here.
%
% Needs on the MATLAB path (README.md:7-10 of the reference; none of them is vendored there):
%   * MATLAB Tensor Toolbox v3.1 (tensor, ktensor, sptensor, mttkrp, tenmat)          https://www.tensortoolbox.org
%   * The Proximity Operator Repository (MATLAB folder: project_*, prox_* used by functions/constraints_to_prox.m)
%                                                                                    http://proximity-operator.net
%   * TV_Condat_v2.m (Laurent Condat's direct 1-D TV denoising, used by functions/prox_TV.m)
%   * L-BFGS-B-C is NOT needed: every case uses the Frobenius loss.
% MATLAB R2018a or newer; GNU Octave is not sufficient (inputParser / Tensor Toolbox classes).
%
% The call made for every case is the one of example_script6_matrix_matrix_CP_nonneg.m:137:
%     [Zhat,Fac,FacInit,out] = cmtf_AOADMM(Z,'alg_options',options,'init',G);
    p = inputParser;
    p.addParameter('b200', false, @(x) islogical(x) || isnumeric(x));
    p.parse(varargin{:});
    addpath(genpath(fullfile(ref_root, 'functions')));
    files = dir(fullfile(golden_dir, 'ref_inputs', 'in_*.mat'));
    if isempty(files), error('no inputs in %s: run tests/golden/make_reference_inputs.py first', fullfile(golden_dir, 'ref_inputs')); end
    for f = files'
        S = load(fullfile(f.folder, f.name));
        name = regexprep(f.name, '^in_(.*)\.mat$', '$1');
        [Z, G, options] = to_reference_structs(S);
        fprintf('%-24s ', name);
        tic
        [~, Fac, ~, out] = cmtf_AOADMM(Z, 'alg_options', options, 'init', G); %#ok<ASGLU>
        t = toc;
        fprintf('reference: %4d outer iterations, f_tensors %.12g, %.2f s\n', out.OuterIterations, out.f_tensors, t);
        ref_version = version; %#ok<NASGU>
        save(fullfile(golden_dir, ['ref_' name '.mat']), 'Fac', 'out', 'ref_version', '-v7');
        if p.Results.b200
            % the drop-in: matlab-code_b200/matlab first on the path shadows functions/cmtf_fun_AOADMM.m
            here = fileparts(mfilename('fullpath'));
            addpath(here, '-begin');
            [~, Fac_b200, ~, out_b200] = cmtf_AOADMM(Z, 'alg_options', options, 'init', G);
            rmpath(here);
            err = 0;
            for m = 1:numel(Fac.fac)
                if iscell(Fac.fac{m})
                    for k = 1:numel(Fac.fac{m})
                        err = max(err, norm(Fac_b200.fac{m}{k} - Fac.fac{m}{k}, 'fro') / norm(Fac.fac{m}{k}, 'fro'));
                    end
                else
                    err = max(err, norm(Fac_b200.fac{m} - Fac.fac{m}, 'fro') / norm(Fac.fac{m}, 'fro'));
                end
            end
            fprintf('%-24s B200 engine vs reference: max factor error %.2e, |df_tensors| %.2e\n', '', err, abs(out_b200.f_tensors - out.f_tensors));
        end
    end
end

function [Z, G, options] = to_reference_structs(S)
% scipy.io.savemat -> the classes the reference expects
    Z = S.Z;
    G = S.G;
    options = S.options;
    P = numel(Z.object);
    nb_modes = numel(Z.size);
    Z.model = cellfun(@char, Z.model(:)', 'UniformOutput', false);
    Z.loss_function = cellfun(@char, Z.loss_function(:)', 'UniformOutput', false);
    Z.modes = cellfun(@(x) double(x(:)'), Z.modes(:)', 'UniformOutput', false);
    Z.size = cellfun(@(x) double(x(:)'), Z.size(:)', 'UniformOutput', false);
    Z.constrained_modes = double(Z.constrained_modes(:)');
    Z.weights = double(Z.weights(:)');
    Z.coupling.lin_coupled_modes = double(Z.coupling.lin_coupled_modes(:)');
    Z.coupling.coupling_type = double(Z.coupling.coupling_type(:)');
    Z.coupling.coupl_trafo_matrices = Z.coupling.coupl_trafo_matrices(:)';
    if isfield(Z.coupling, 'coupl_trafo_matrices2')
        if all(cellfun(@isempty, Z.coupling.coupl_trafo_matrices2))
            Z.coupling = rmfield(Z.coupling, 'coupl_trafo_matrices2');
        else
            Z.coupling.coupl_trafo_matrices2 = Z.coupling.coupl_trafo_matrices2(:)';
        end
    end
    cons = cell(nb_modes, 1);
    for m = 1:nb_modes
        c = Z.constraints{m};
        if isempty(c), cons{m} = {}; continue; end
        c = c(:)';
        c{1} = char(c{1});
        if strcmp(c{1}, 'unimodality') && numel(c) > 1, c{2} = logical(c{2}); end
        cons{m} = c;
    end
    Z.constraints = cons;
    for q = 1:P
        if strcmp(Z.model{q}, 'CP')
            Z.object{q} = tensor(double(Z.object{q}));           % create_coupled_data.m:158 hands over `tensor` objects
        else
            Z.object{q} = cellfun(@double, Z.object{q}(:), 'UniformOutput', false);
        end
    end
    if isfield(Z, 'miss')
        for q = 1:P
            if isempty(Z.miss{q}), Z.miss{q} = []; continue; end
            if strcmp(Z.model{q}, 'CP')
                Z.miss{q} = tensor(logical(Z.miss{q}));          % true = observed (example_script12_CP_PAR2_EM.m:118-124)
            else
                Z.miss{q} = cellfun(@logical, Z.miss{q}(:), 'UniformOutput', false);
            end
        end
    end
    % the init struct: cell columns, per-slice fields as cell columns, [] where a field does not exist
    for fn = {'fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac', 'coupling_fac', 'P', 'DeltaB', 'mu_DeltaB'}
        if ~isfield(G, fn{1}), continue; end
        c = G.(fn{1});
        if ~iscell(c), c = {}; end
        c = c(:);
        for i = 1:numel(c)
            if iscell(c{i}), c{i} = cellfun(@double, c{i}(:), 'UniformOutput', false); else, c{i} = double(c{i}); end
        end
        G.(fn{1}) = c;
    end
    names = fieldnames(options);
    for i = 1:numel(names)
        v = options.(names{i});
        if ischar(v) || isstring(v), options.(names{i}) = char(v); else, options.(names{i}) = double(v); end
    end
end
