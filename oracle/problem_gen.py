"""Oracle-side problem fabrication (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the callers on the input side of the boundary so tests can build (Z, G, options):
  functions/create_coupled_data.m:55-183 (Frobenius / Gaussian noise; couplings 0,1,3,4 and 2),
  functions/init_coupled_AOADMM_CMTF.m:37-169 (random init, nvecs=0),
  functions_for_example_scripts/create_CP_data_example10piecewiseconstant.m:77-92,
and the configuration sections of example_script6 (:25-133), example_script1 (:21-123),
example_script10 (:23-122).  MATLAB's randn/randi streams cannot be reproduced from NumPy
(SURVEY.md 8c RNG note) so seeds here are our own.
"""
import numpy as np

from .prox import constraints_to_prox
from .tensor_ops import full_ktensor


def _size1(s):
    return int(s[0]) if isinstance(s, (list, tuple, np.ndarray)) else int(s)


def create_coupled_data(model, sz, modes, lambdas, noise, coupling, normalize_columns, distr_data, rng,
                        piecewise_constant_mode1=False):
    """create_coupled_data.m:48-183.  distr_data[n](rows, cols, rng) -> ndarray."""
    P = len(modes)
    nb_modes = len(sz)
    lin = coupling['lin_coupled_modes']
    ctype_l = coupling.get('coupling_type', [])
    trafo = coupling.get('coupl_trafo_matrices') or [None] * nb_modes
    A = [None] * nb_modes
    Delta = [None] * nb_modes

    def normcols(M):
        return M / np.sqrt(np.sum(M * M, axis=0))[None, :]

    for p in range(P):
        for n in modes[p]:
            if lin[n - 1] == 0:
                A[n - 1] = distr_data[n - 1](_size1(sz[n - 1]), len(lambdas[p]), rng)
                if normalize_columns:
                    A[n - 1] = normcols(A[n - 1])
                if model[p] == 'PAR2' and modes[p].index(n) == 1:
                    AA = A[n - 1]
                    A[n - 1] = [np.roll(AA, k, axis=0) for k in range(len(sz[n - 1]))]   # circshift (:64-72)
    if piecewise_constant_mode1:                                                    # example10 generator :77-92
        I = _size1(sz[0])
        for r in range(len(lambdas[0])):
            jumps = np.concatenate(([1], np.sort(rng.randint(1, I + 1, size=4)), [I]))
            values = -1 + 2 * rng.rand(5)
            for i in range(5):
                A[0][jumps[i] - 1:jumps[i + 1], r] = values[i]
        if normalize_columns:
            A[0] = normcols(A[0])
    nb_c = max(lin) if len(lin) else 0
    for i in range(1, nb_c + 1):
        ct = ctype_l[i - 1]
        cp_modes = [m for m in range(1, nb_modes + 1) if lin[m - 1] == i]
        mode1 = cp_modes[0]
        p1 = [p for p in range(P) if mode1 in modes[p]][0]
        R1 = len(lambdas[p1])
        if ct == 0:
            A[mode1 - 1] = distr_data[mode1 - 1](_size1(sz[mode1 - 1]), R1, rng)
            if normalize_columns:
                A[mode1 - 1] = normcols(A[mode1 - 1])
            for j in cp_modes[1:]:
                A[j - 1] = A[mode1 - 1].copy()
        elif ct == 1:
            mode1 = cp_modes[int(np.argmax([_size1(sz[m - 1]) for m in cp_modes]))]
            A[mode1 - 1] = distr_data[mode1 - 1](_size1(sz[mode1 - 1]), R1, rng)
            if normalize_columns:
                A[mode1 - 1] = normcols(A[mode1 - 1])
            Delta[i - 1] = trafo[mode1 - 1] @ A[mode1 - 1]
            for j in cp_modes:
                if j != mode1:
                    A[j - 1] = np.linalg.pinv(trafo[j - 1]) @ Delta[i - 1]
        elif ct == 2:
            Delta[i - 1] = distr_data[mode1 - 1](_size1(sz[mode1 - 1]), trafo[mode1 - 1].shape[1], rng)
            if normalize_columns:
                Delta[i - 1] = normcols(Delta[i - 1])
            for j in cp_modes:
                A[j - 1] = np.linalg.lstsq(trafo[j - 1].T, Delta[i - 1].T, rcond=None)[0].T
        elif ct == 3:
            Delta[i - 1] = distr_data[mode1 - 1](trafo[mode1 - 1].shape[1], R1, rng)
            if normalize_columns:
                Delta[i - 1] = normcols(Delta[i - 1])
            for j in cp_modes:
                A[j - 1] = trafo[j - 1] @ Delta[i - 1]
        elif ct == 4:
            Delta[i - 1] = distr_data[mode1 - 1](_size1(sz[mode1 - 1]), trafo[mode1 - 1].shape[0], rng)
            if normalize_columns:
                Delta[i - 1] = normcols(Delta[i - 1])
            for j in cp_modes:
                A[j - 1] = Delta[i - 1] @ trafo[j - 1]
    X = [None] * P
    for p in range(P):
        if model[p] == 'CP':
            Xp = full_ktensor([A[m - 1] for m in modes[p]], lambdas[p])
            N = rng.randn(*Xp.shape)
            sigma = noise[p] * np.sqrt(np.sum(Xp ** 2)) / np.sqrt(np.sum(N ** 2))
            X[p] = np.asfortranarray(Xp + sigma * N)
        else:
            m1, m2, m3 = modes[p]
            Cf = A[m3 - 1]
            Xp = []
            for k in range(Cf.shape[0]):
                Xk = A[m1 - 1] @ np.diag(np.asarray(lambdas[p], dtype=float)) @ np.diag(Cf[k, :]) @ A[m2 - 1][k].T
                Nk = rng.randn(*Xk.shape)
                sigma = noise[p] * np.sqrt(np.sum(Xk ** 2)) / np.sqrt(np.sum(Nk ** 2))
                Xp.append(np.asfortranarray(Xk + sigma * Nk))
            X[p] = Xp
    return X, A, Delta


def normalize_objects(X, model):
    """example_script6...m:97-113: divide every data set by its Frobenius norm."""
    out = []
    norms = []
    for p, Xp in enumerate(X):
        if model[p] == 'CP':
            nz = np.sqrt(np.sum(Xp ** 2))
            out.append(np.asfortranarray(Xp / nz))
        else:
            nz = np.sqrt(sum(np.sum(Xk ** 2) for Xk in Xp))
            out.append([np.asfortranarray(Xk / nz) for Xk in Xp])
        norms.append(nz)
    return out, norms


def znorm_const(Z):
    """cmtf_AOADMM.m:124-156 (Frobenius; with Z.miss the norm of the observed entries, :134-151)."""
    out = []
    miss = Z.get('miss') or [None] * len(Z['object'])
    for p, Xp in enumerate(Z['object']):
        Mp = miss[p]
        if Z['model'][p] == 'CP':
            Xe = Xp if Mp is None else np.asarray(Mp, dtype=np.float64) * Xp
            out.append(float(np.sqrt(np.sum(Xe ** 2)) ** 2))
        else:
            out.append(float(sum(np.sqrt(np.sum((Xk if Mp is None else np.asarray(Mp[k], dtype=np.float64) * Xk) ** 2)) ** 2
                                 for k, Xk in enumerate(Xp))))
    return out


def add_missing(Z, frac=0.2, seed=0, objects=None):
    """example_script12_CP_PAR2_EM.m:115-146: ~frac of the entries of the chosen objects missing at random
    (mask true = observed), missing entries initialised with 0."""
    rng = np.random.RandomState(seed)
    Z = dict(Z)
    Z['object'] = list(Z['object'])
    miss = [None] * len(Z['object'])
    for p, Xp in enumerate(Z['object']):
        if objects is not None and p not in objects:
            continue
        if Z['model'][p] == 'CP':
            mask = np.ones(Xp.size, dtype=bool)
            mask[rng.permutation(Xp.size)[:int(round(frac * Xp.size))]] = False
            mask = np.asfortranarray(mask.reshape(Xp.shape, order='F'))
            miss[p] = mask
            Z['object'][p] = np.asfortranarray(np.where(mask, Xp, 0.0))
        else:
            mk, xs = [], []
            for Xk in Xp:
                m = np.ones(Xk.size, dtype=bool)
                m[rng.permutation(Xk.size)[:int(round(frac * Xk.size))]] = False
                m = np.asfortranarray(m.reshape(Xk.shape, order='F'))
                mk.append(m)
                xs.append(np.asfortranarray(np.where(m, Xk, 0.0)))
            miss[p] = mk
            Z['object'][p] = xs
    Z['miss'] = miss
    return Z


def _leading_eigvecs(Y, r):
    """[U,~] = eigs(Y, r, 'LM') for a symmetric positive semi-definite Y: eigenvectors of the r largest eigenvalues in
    descending order.  eigs leaves the sign of a vector open; it is fixed here (entry of largest magnitude positive,
    first one on ties) so that results can be compared."""
    w, V = np.linalg.eigh((Y + Y.T) * 0.5)
    idx = np.argsort(-w, kind='stable')[:r]
    U = V[:, idx].copy()
    for c in range(U.shape[1]):
        i = int(np.argmax(np.abs(U[:, c])))
        if U[i, c] < 0:
            U[:, c] = -U[:, c]
    return U, w[idx]


def cmtf_nvecs(Z, n, r):
    """cmtf_nvecs.m:33-58: A = mode-n unfolding of the first object containing mode n (objects sharing the mode id
    are concatenated, :36-51 -- mode ids are unique per object in this framework); Y = A*A' (:55); eigs(Y,r,'LM') (:57)."""
    blocks = []
    for p, ms in enumerate(Z['modes']):
        if n in ms:
            if Z['model'][p] != 'CP':
                raise ValueError('cmtf_nvecs works on CP objects (double(Z.object{p}) at cmtf_nvecs.m:44-48)')
            i = list(ms).index(n)
            X = np.asarray(Z['object'][p], dtype=np.float64)
            blocks.append(np.moveaxis(X, i, 0).reshape(X.shape[i], -1, order='F'))
    A = np.concatenate(blocks, axis=1)
    return _leading_eigvecs(A @ A.T, r)[0]


def init_coupled_AOADMM_CMTF(Z, init_options, rng, Delta=None):
    """init_coupled_AOADMM_CMTF.m:37-169 (init_options['nvecs']: :50-69)."""
    sz = Z['size']
    lambdas = init_options['lambdas_init']
    distr = init_options['distr']
    normalize = init_options['normalize']
    modes = Z['modes']
    model = Z['model']
    lin = Z['coupling']['lin_coupled_modes']
    ctype_l = Z['coupling'].get('coupling_type', [])
    trafo = Z['coupling'].get('coupl_trafo_matrices') or [None] * len(sz)
    constrained = Z['constrained_modes']
    constraints = Z['constraints']
    P = len(modes)
    nb_modes = len(sz)
    nb_c = max(lin) if len(lin) else 0
    G = {'fac': [None] * nb_modes, 'coupling_fac': [None] * nb_c, 'constraint_fac': [None] * nb_modes,
         'coupling_dual_fac': [None] * nb_modes, 'constraint_dual_fac': [None] * nb_modes,
         'P': [None] * P, 'DeltaB': [None] * P, 'mu_DeltaB': [None] * P}

    def normcols(M):
        return M / np.sqrt(np.sum(M * M, axis=0))[None, :]

    nvecs = bool(init_options.get('nvecs', 0))
    for p in range(P):
        R = len(lambdas[p])
        for n in modes[p]:
            if nvecs:                                                   # :50-69
                if model[p] == 'CP':
                    G['fac'][n - 1] = cmtf_nvecs(Z, n, R)
                elif modes[p].index(n) == 0:
                    M = np.concatenate([np.asarray(Xk) for Xk in Z['object'][p]], axis=1)
                    G['fac'][n - 1] = _leading_eigvecs(M @ M.T, R)[0]
                elif modes[p].index(n) == 1:
                    G['DeltaB'][p] = rng.rand(R, R)
                    G['fac'][n - 1], G['P'][p], G['mu_DeltaB'][p] = [], [], []
                    for k in range(len(sz[n - 1])):
                        Xk = np.asarray(Z['object'][p][k])
                        G['fac'][n - 1].append(_leading_eigvecs(Xk.T @ Xk, R)[0])
                        G['P'][p].append(np.eye(int(sz[n - 1][k]), R))
                        G['mu_DeltaB'][p].append(rng.rand(int(sz[n - 1][k]), R))
                else:
                    G['fac'][n - 1] = np.ones((_size1(sz[n - 1]), R))
                continue
            if model[p] == 'PAR2' and modes[p].index(n) == 1:
                G['DeltaB'][p] = rng.rand(R, R)
                G['fac'][n - 1] = []
                G['P'][p] = []
                G['mu_DeltaB'][p] = []
                for k in range(len(sz[n - 1])):
                    Fk = distr[n - 1](int(sz[n - 1][k]), R, rng)
                    G['P'][p].append(np.eye(int(sz[n - 1][k]), R))
                    G['mu_DeltaB'][p].append(rng.rand(int(sz[n - 1][k]), R))
                    if normalize:
                        Fk = normcols(Fk)
                    G['fac'][n - 1].append(Fk)
            else:
                F = distr[n - 1](_size1(sz[n - 1]), R, rng)
                if normalize:
                    F = normcols(F)
                G['fac'][n - 1] = F
    if any(constrained):
        prox_ops, _ = constraints_to_prox(constrained, constraints, sz)
        for p in range(P):
            for n in modes[p]:
                if not constrained[n - 1]:
                    continue
                if model[p] == 'PAR2' and modes[p].index(n) == 1:
                    G['constraint_fac'][n - 1] = []
                    G['constraint_dual_fac'][n - 1] = []
                    for k in range(len(sz[n - 1])):
                        Zk = distr[n - 1](*G['fac'][n - 1][k].shape, rng)
                        if constraints[n - 1][0] != 'tPARAFAC2':
                            Zk = prox_ops[n - 1](Zk, 1.0)
                        G['constraint_fac'][n - 1].append(Zk)
                        G['constraint_dual_fac'][n - 1].append(rng.rand(*G['fac'][n - 1][k].shape))
                else:
                    Zc = distr[n - 1](*G['fac'][n - 1].shape, rng)
                    G['constraint_fac'][n - 1] = prox_ops[n - 1](Zc, 1.0)
                    G['constraint_dual_fac'][n - 1] = rng.rand(*G['fac'][n - 1].shape)
    for n in range(1, nb_c + 1):
        cmodes = [m for m in range(1, nb_modes + 1) if lin[m - 1] == n]
        mode1 = cmodes[0]
        ct = ctype_l[n - 1]
        F1 = G['fac'][mode1 - 1]
        H1 = trafo[mode1 - 1]
        if ct == 0:
            G['coupling_fac'][n - 1] = rng.rand(*F1.shape)
            for m in cmodes:
                G['coupling_dual_fac'][m - 1] = rng.rand(*G['coupling_fac'][n - 1].shape)
        elif ct == 1:
            G['coupling_fac'][n - 1] = rng.rand(H1.shape[0], F1.shape[1])
            for m in cmodes:
                G['coupling_dual_fac'][m - 1] = rng.rand(*G['coupling_fac'][n - 1].shape)
        elif ct == 2:
            G['coupling_fac'][n - 1] = rng.rand(F1.shape[0], H1.shape[1])
            for m in cmodes:
                G['coupling_dual_fac'][m - 1] = rng.rand(*G['coupling_fac'][n - 1].shape)
        elif ct == 3:
            G['coupling_fac'][n - 1] = rng.rand(H1.shape[1], F1.shape[1])
            for m in cmodes:
                G['coupling_dual_fac'][m - 1] = rng.rand(*G['fac'][m - 1].shape)
        elif ct == 4:
            G['coupling_fac'][n - 1] = rng.rand(F1.shape[0], H1.shape[0])
            for m in cmodes:
                G['coupling_dual_fac'][m - 1] = rng.rand(*G['fac'][m - 1].shape)
        elif ct == 5:
            G['coupling_fac'][n - 1] = rng.rand(*Delta[n - 1].shape)
            for m in cmodes:
                G['coupling_dual_fac'][m - 1] = rng.rand(G['coupling_fac'][n - 1].shape[0], G['fac'][m - 1].shape[1])
    return G


# ----------------------------------------------------------------------------------------------
# configuration builders
# ----------------------------------------------------------------------------------------------
def d_rand(r, c, rng):
    return rng.rand(r, c)


def d_randn(r, c, rng):
    return rng.randn(r, c)


def d_rand01(r, c, rng):
    return rng.rand(r, c) + 0.1


def default_options(**kw):
    """example_script6...m:120-132."""
    o = dict(Display='no', DisplayIters=10, MaxOuterIters=4000, MaxInnerIters=5, AbsFuncTol=1e-4,
             OuterRelTol=1e-8, innerRelPrTol_coupl=1e-3, innerRelPrTol_constr=1e-3,
             innerRelDualTol_coupl=1e-3, innerRelDualTol_constr=1e-3, bsum=0, eps_log=1e-10)
    o.update(kw)
    return o


def _finish(model, sz, modes, lambdas, noise, coupling, distr, constrained_modes, constraints, weights, rng,
            normalize_columns=0, init_normalize=1, piecewise=False, Delta_shapes=None):
    X, Atrue, Delta = create_coupled_data(model, sz, modes, lambdas, noise, coupling, normalize_columns, distr, rng,
                                          piecewise_constant_mode1=piecewise)
    obj, norms = normalize_objects(X, model)
    Z = {'loss_function': ['Frobenius'] * len(modes), 'model': model, 'modes': modes, 'size': sz,
         'coupling': coupling, 'constrained_modes': constrained_modes, 'constraints': constraints,
         'weights': weights, 'object': obj}
    init_options = {'lambdas_init': lambdas, 'nvecs': 0, 'distr': distr, 'normalize': init_normalize}
    G = init_coupled_AOADMM_CMTF(Z, init_options, rng, Delta=Delta_shapes)
    return Z, G, {'Atrue': Atrue, 'norms': norms, 'Delta': Delta}


def config_script6(seed=0, sz=(50, 60, 40, 50, 70, 60, 80), R=3, noise=0.2):
    """C1: example_script6_matrix_matrix_CP_nonneg.m:25-92 (3-way CP + two coupled matrices, nonneg)."""
    rng = np.random.RandomState(seed)
    sz = list(sz)
    modes = [[1, 2, 3], [4, 5], [6, 7]]
    lambdas = [[1.0] * R] * 3
    distr = [d_rand, d_rand, d_randn, d_rand, d_rand, d_rand, d_rand]
    coupling = {'lin_coupled_modes': [1, 2, 0, 1, 0, 2, 0], 'coupling_type': [0, 0],
                'coupl_trafo_matrices': [None] * 7}
    constrained = [1, 0, 0, 1, 1, 1, 1]          # mode 2 unconstrained although a constraint is listed (quirk 9)
    nn = ('non-negativity',)
    constraints = [nn, nn, None, nn, nn, nn, nn]
    return _finish(['CP', 'CP', 'CP'], sz, modes, lambdas, [noise] * 3, coupling, distr, constrained, constraints,
                   [1 / 3, 1 / 3, 1 / 3], rng)


def config_cp_matrix(I, J, K, M, R, seed=0, noise=0.2):
    """C2/C3 family: 3-way CP I x J x K coupled in mode 1 with an I x M matrix, all modes nonneg
    (modes={[1 2 3],[4 5]}, lin_coupled_modes=[1 0 0 1 0], weights [1/2 1/2])."""
    rng = np.random.RandomState(seed)
    sz = [I, J, K, I, M]
    modes = [[1, 2, 3], [4, 5]]
    lambdas = [[1.0] * R] * 2
    distr = [d_rand] * 5
    coupling = {'lin_coupled_modes': [1, 0, 0, 1, 0], 'coupling_type': [0], 'coupl_trafo_matrices': [None] * 5}
    nn = ('non-negativity',)
    return _finish(['CP', 'CP'], sz, modes, lambdas, [noise] * 2, coupling, distr, [1] * 5, [nn] * 5,
                   [0.5, 0.5], rng)


def config_cp_par2(I=20, J=30, K=40, Jk=30, Kp=20, R=3, seed=0, noise=0.0):
    """C4 family: example_script1_CP_PAR2_nonneg.m:21-63 (CP coupled in mode 1 with a regular PARAFAC2)."""
    rng = np.random.RandomState(seed)
    sz = [I, J, K, I, [Jk] * Kp, Kp]
    modes = [[1, 2, 3], [4, 5, 6]]
    lambdas = [[1.0] * R] * 2
    distr = [d_rand, d_randn, d_randn, d_rand, d_rand, d_rand01]
    coupling = {'lin_coupled_modes': [1, 0, 0, 1, 0, 0], 'coupling_type': [0], 'coupl_trafo_matrices': [None] * 6}
    nn = ('non-negativity',)
    constrained = [1, 0, 0, 1, 1, 1]
    constraints = [nn, None, None, nn, nn, nn]
    return _finish(['CP', 'PAR2'], sz, modes, lambdas, [noise] * 2, coupling, distr, constrained, constraints,
                   [0.5, 0.5], rng)


def config_cp_tv(I=60, J=50, K=70, R=3, seed=0, noise=0.8, eta=1e-3, mode1=('TV regularization',)):
    """C5 family: example_script10_CP_TVreg.m:23-57 (TV on a piecewise-constant mode 1, l2-ball on 2,3)."""
    rng = np.random.RandomState(seed)
    sz = [I, J, K]
    modes = [[1, 2, 3]]
    lambdas = [[1.0] * R]
    distr = [d_randn] * 3
    coupling = {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3}
    c1 = tuple(mode1) + ((eta,) if len(mode1) == 1 else ())
    constraints = [c1, ('l2-ball', 1.0), ('l2-ball', 1.0)]
    return _finish(['CP'], sz, modes, lambdas, [noise], coupling, distr, [1, 1, 1], constraints, [1.0], rng,
                   normalize_columns=1, piecewise=True)


def config_single_cp(sz=(12, 10, 8), R=3, seed=0, noise=0.1, constraints=None, constrained=None, distr=None):
    """Single uncoupled 3-way (or N-way) CP tensor with arbitrary per-mode constraints."""
    rng = np.random.RandomState(seed)
    n = len(sz)
    sz = list(sz)
    modes = [list(range(1, n + 1))]
    coupling = {'lin_coupled_modes': [0] * n, 'coupling_type': [], 'coupl_trafo_matrices': [None] * n}
    constraints = constraints if constraints is not None else [None] * n
    constrained = constrained if constrained is not None else [1 if c else 0 for c in constraints]
    return _finish(['CP'], sz, modes, [[1.0] * R], [noise], coupling, distr or [d_rand] * n, constrained,
                   constraints, [1.0], rng)


def config_single_par2(I=18, Jk=(12, 9, 15, 11), R=3, seed=0, noise=0.05, constrained=(1, 0, 1), constraints=None,
                       ridge=None):
    """One uncoupled PARAFAC2 object with (possibly irregular, example_script4 style) slice heights J_k.
    X_k = A diag(c_k) B_k' with B_k = Q_k H (Q_k orthonormal J_k x R) so that B_k'B_k is the same for every k."""
    rng = np.random.RandomState(seed)
    K = len(Jk)
    sz = [I, [int(j) for j in Jk], K]
    modes = [[1, 2, 3]]
    A = rng.rand(I, R)
    C = rng.rand(K, R) + 0.1
    H = rng.rand(R, R) + np.eye(R)
    X = []
    for k in range(K):
        Q, _ = np.linalg.qr(rng.randn(int(Jk[k]), R))
        Xk = A @ np.diag(C[k, :]) @ (Q @ H).T
        Nk = rng.randn(*Xk.shape)
        X.append(Xk + noise * np.linalg.norm(Xk) / np.linalg.norm(Nk) * Nk)
    obj, _ = normalize_objects([X], ['PAR2'])
    nn = ('non-negativity',)
    constraints = list(constraints) if constraints is not None else [nn if c else None for c in constrained]
    coupling = {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3}
    Z = {'loss_function': ['Frobenius'], 'model': ['PAR2'], 'modes': modes, 'size': sz, 'coupling': coupling,
         'constrained_modes': list(constrained), 'constraints': constraints, 'weights': [1.0], 'object': obj}
    if ridge is not None:
        Z['ridge'] = list(ridge)
    init_options = {'lambdas_init': [[1.0] * R], 'nvecs': 0, 'distr': [d_rand, d_rand, d_rand01], 'normalize': 1}
    G = init_coupled_AOADMM_CMTF(Z, init_options, rng)
    return Z, G, {'Atrue': [A, None, C]}


def config_linear_coupling(ctype, seed=0, noise=0.05, constrained=True, second='matrix'):
    """A 3-way CP object linearly coupled (coupling types 1..5, cmtf_fun_AOADMM.m:278-389) with a matrix or a second
    CP tensor, sizes in the style of example_script3 (type 4), example_script13 (type 5), example_script14 (type 1)."""
    rng = np.random.RandomState(seed)
    nn = ('non-negativity',)
    sub = np.zeros((20, 40))
    sub[np.arange(20), 2 * np.arange(20)] = 1.0                       # "take every second entry" (script 13/14)
    part = np.vstack([np.eye(3), np.zeros((1, 3))])                   # partially shared components (script 3/13)
    if ctype == 1:      # H F = Delta
        I1, I4, R1, R2 = 20, 40, 3, 3
        H = {1: np.eye(20), 4: sub}
        H2 = {}
    elif ctype == 2:    # F H = Delta
        I1, I4, R1, R2 = 30, 30, 4, 3
        H = {1: part.copy(), 4: np.eye(3)}
        H2 = {}
    elif ctype == 3:    # F = H Delta
        I1, I4, R1, R2 = 30, 24, 3, 3
        H = {1: rng.rand(30, 10), 4: rng.rand(24, 10)}
        H2 = {}
    elif ctype == 4:    # F = Delta H
        I1, I4, R1, R2 = 30, 30, 4, 3
        H = {1: np.eye(4), 4: part.copy()}
        H2 = {}
    elif ctype == 5:    # H F = Delta H2
        I1, I4, R1, R2 = 20, 40, 4, 3
        H = {1: np.eye(20), 4: sub}
        H2 = {1: np.eye(4), 4: part.copy()}
    else:
        raise ValueError(ctype)
    if second == 'matrix':
        sz = [I1, 14, 12, I4, 18]
        modes = [[1, 2, 3], [4, 5]]
    elif second == 'par2':       # the first (A) mode of a PARAFAC2 object takes the place of the second tensor's mode
        sz = [I1, 14, 12, I4, [11] * 7, 7]
        modes = [[1, 2, 3], [4, 5, 6]]
    elif second == 'par2c':      # the third (C) mode of a PARAFAC2 object (K = I4 slices) is the coupled one
        sz = [I1, 14, 12, 16, [11] * I4, I4]
        modes = [[1, 2, 3], [4, 5, 6]]
    else:
        sz = [I1, 14, 12, I4, 11, 9]
        modes = [[1, 2, 3], [4, 5, 6]]
    nm = len(sz)
    trafo = [None] * nm
    trafo2 = [None] * nm
    second_mode = 6 if second == 'par2c' else 4
    for m, h in H.items():
        trafo[(m if m == 1 else second_mode) - 1] = h
    for m, h in H2.items():
        trafo2[(m if m == 1 else second_mode) - 1] = h
    lin = [0] * nm
    lin[0] = lin[second_mode - 1] = 1
    coupling = {'lin_coupled_modes': lin, 'coupling_type': [ctype], 'coupl_trafo_matrices': trafo}
    if ctype == 5:
        coupling['coupl_trafo_matrices2'] = trafo2
    lambdas = [[1.0] * R1, [1.0] * R2]
    distr = [d_rand] * nm
    model = ['CP', 'PAR2' if second in ('par2', 'par2c') else 'CP']
    Delta_shapes = None
    if second == 'par2c':
        free = {'lin_coupled_modes': [0] * nm, 'coupling_type': [], 'coupl_trafo_matrices': [None] * nm}
        X, _, _ = create_coupled_data(model, sz, modes, lambdas, [noise] * 2, free, 0, distr, rng)
        if ctype == 5:
            Delta_shapes = [rng.rand(20, 4)]
    elif ctype == 5 and second == 'par2':
        Delta_shapes = [rng.rand(20, 4)]
        free = {'lin_coupled_modes': [0] * nm, 'coupling_type': [], 'coupl_trafo_matrices': [None] * nm}
        X, _, _ = create_coupled_data(model, sz, modes, lambdas, [noise] * 2, free, 0, distr, rng)
    elif ctype == 5:
        # ground truth: Delta (q1 x q2); F1 = Delta H2_1 (H1_1 = I); rows of F4 picked by H1_4 equal Delta H2_4
        Delta = rng.rand(20, 4)
        A = [None] * nm
        for p in range(2):
            for n in modes[p]:
                A[n - 1] = rng.rand(sz[n - 1], len(lambdas[p]))
        A[0] = Delta @ H2[1]
        A[3][::2, :] = Delta @ H2[4]
        X = []
        for p in range(2):
            Xp = full_ktensor([A[m - 1] for m in modes[p]], lambdas[p])
            N = rng.randn(*Xp.shape)
            X.append(np.asfortranarray(Xp + noise * np.linalg.norm(Xp) / np.linalg.norm(N) * N))
        Delta_shapes = [Delta]
    else:
        X, _, _ = create_coupled_data(model, sz, modes, lambdas, [noise] * 2, coupling, 0, distr, rng)
    obj, _ = normalize_objects(X, model)
    cm = [0] * nm
    cons = [None] * nm
    if constrained:
        for m in ((1, 4, 6) if second in ('par2', 'par2c') else (1, 4, 5)):
            cm[m - 1] = 1
            cons[m - 1] = nn
    Z = {'loss_function': ['Frobenius'] * 2, 'model': model, 'modes': modes, 'size': sz, 'coupling': coupling,
         'constrained_modes': cm, 'constraints': cons, 'weights': [0.5, 0.5], 'object': obj}
    init_options = {'lambdas_init': lambdas, 'nvecs': 0, 'distr': distr, 'normalize': 1}
    G = init_coupled_AOADMM_CMTF(Z, init_options, rng, Delta=Delta_shapes)
    return Z, G, {}


def config_tparafac2(I=16, J=14, K=9, R=3, seed=0, noise=0.05, eta=0.05, drift=0.05):
    """example_script11 style: PARAFAC2 whose B_k evolve smoothly over k (B_k = B_0 + k*drift*D), nonneg A and C and the
    temporal-smoothness ('tPARAFAC2') regulariser on the B_k mode."""
    rng = np.random.RandomState(seed)
    A = rng.rand(I, R)
    C = rng.rand(K, R) + 0.5
    B0 = rng.rand(J, R)
    D = rng.randn(J, R)
    X = []
    for k in range(K):
        Xk = A @ np.diag(C[k, :]) @ (B0 + k * drift * D).T
        Nk = rng.randn(*Xk.shape)
        X.append(Xk + noise * np.linalg.norm(Xk) / np.linalg.norm(Nk) * Nk)
    obj, _ = normalize_objects([X], ['PAR2'])
    nn = ('non-negativity',)
    coupling = {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3}
    Z = {'loss_function': ['Frobenius'], 'model': ['PAR2'], 'modes': [[1, 2, 3]], 'size': [I, [J] * K, K],
         'coupling': coupling, 'constrained_modes': [1, 1, 1], 'constraints': [nn, ('tPARAFAC2', eta), nn],
         'weights': [1.0], 'object': obj}
    init_options = {'lambdas_init': [[1.0] * R], 'nvecs': 0, 'distr': [d_rand, d_rand, d_rand01], 'normalize': 1}
    G = init_coupled_AOADMM_CMTF(Z, init_options, rng)
    return Z, G, {}


def config_script14(I=20, J=14, K=12, I2=16, Jk=10, Kp=40, R=3, seed=0, noise=0.05, constrained_c=True):
    """example_script14_CP_PAR2_couplC_doublesamplingrate.m:24-38: a CP tensor whose first mode is coupled (type 1,
    H C = Delta) with the THIRD mode of a PARAFAC2 model sampled at twice the rate (H picks every second row)."""
    rng = np.random.RandomState(seed)
    sz = [I, J, K, I2, [Jk] * Kp, Kp]
    modes = [[1, 2, 3], [4, 5, 6]]
    H6 = np.zeros((I, Kp))
    H6[np.arange(I), 2 * np.arange(I)] = 1.0
    trafo = [np.eye(I), None, None, None, None, H6]
    coupling = {'lin_coupled_modes': [1, 0, 0, 0, 0, 1], 'coupling_type': [1], 'coupl_trafo_matrices': trafo}
    lambdas = [[1.0] * R] * 2
    distr = [d_rand, d_randn, d_randn, d_rand, d_rand, d_rand01]
    nn = ('non-negativity',)
    cm = [1, 0, 0, 1, 0, 1 if constrained_c else 0]
    cons = [nn, None, None, nn, None, nn if constrained_c else None]
    return _finish(['CP', 'PAR2'], sz, modes, lambdas, [noise] * 2, coupling, distr, cm, cons, [0.5, 0.5], rng)


def config_script2(I=20, J=30, I2=20, Jk=12, Kp=8, R=3, seed=0, noise=0.1):
    """example_script2_matrix_PAR2_nonneg.m:20-64: a matrix exactly coupled (mode 1) with the first mode of a regular
    PARAFAC2 object, every mode non-negative, weights [1/2 1/2]."""
    rng = np.random.RandomState(seed)
    sz = [I, J, I2, [Jk] * Kp, Kp]
    modes = [[1, 2], [3, 4, 5]]
    coupling = {'lin_coupled_modes': [1, 0, 1, 0, 0], 'coupling_type': [0], 'coupl_trafo_matrices': [None] * 5}
    nn = ('non-negativity',)
    return _finish(['CP', 'PAR2'], sz, modes, [[1.0] * R] * 2, [noise] * 2, coupling, [d_rand] * 5, [1] * 5, [nn] * 5,
                   [0.5, 0.5], rng)


def config_script15(I=28, dims1=(13, 11), dims2=(9, 12), M=17, seed=0, noise=0.1):
    """example_script15_realdata.m:35-77 with synthetic data of the same structure: two 3-way CP tensors (3 and 5
    components) and a matrix (5 components) that share the sample mode through ONE type-4 coupling C_m = Delta*H_m with
    three members (Delta: I x 6; H_1 = [I_3; 0], H_4 = [I_5; 0], H_7 = T, :47-51), all modes non-negative, weights 1/3."""
    rng = np.random.RandomState(seed)
    sz = [I, dims1[0], dims1[1], I, dims2[0], dims2[1], I, M]
    modes = [[1, 2, 3], [4, 5, 6], [7, 8]]
    lambdas = [[1.0] * 3, [1.0] * 5, [1.0] * 5]
    T = np.vstack([np.hstack([np.eye(4), np.zeros((4, 1))]), np.zeros((1, 5)), np.array([[0, 0, 0, 0, 1.0]])])
    trafo = [None] * 8
    trafo[0] = np.vstack([np.eye(3), np.zeros((3, 3))])
    trafo[3] = np.vstack([np.eye(5), np.zeros((1, 5))])
    trafo[6] = T
    coupling = {'lin_coupled_modes': [1, 0, 0, 1, 0, 0, 1, 0], 'coupling_type': [4], 'coupl_trafo_matrices': trafo}
    # ground truth with the coupling structure: C_m = Delta * H_m
    Delta = rng.rand(I, 6)
    A = [None] * 8
    for p in range(3):
        for n in modes[p]:
            A[n - 1] = rng.rand(sz[n - 1], len(lambdas[p]))
    for m in (1, 4, 7):
        A[m - 1] = Delta @ trafo[m - 1]
    X = []
    for p in range(3):
        Xp = full_ktensor([A[m - 1] for m in modes[p]], lambdas[p])
        N = rng.randn(*Xp.shape)
        X.append(np.asfortranarray(Xp + noise * np.linalg.norm(Xp) / np.linalg.norm(N) * N))
    model = ['CP'] * 3
    obj, _ = normalize_objects(X, model)
    nn = ('non-negativity',)
    Z = {'loss_function': ['Frobenius'] * 3, 'model': model, 'modes': modes, 'size': sz, 'coupling': coupling,
         'constrained_modes': [1] * 8, 'constraints': [nn] * 8, 'weights': [1 / 3] * 3, 'object': obj}
    init_options = {'lambdas_init': lambdas, 'nvecs': 0, 'distr': [d_rand] * 8, 'normalize': 0}
    G = init_coupled_AOADMM_CMTF(Z, init_options, rng)
    return Z, G, {'Delta': Delta}
