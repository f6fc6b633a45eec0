"""Oracle: the AO-ADMM solver (TEST INFRASTRUCTURE, see oracle/__init__.py).

Line-by-line NumPy float64 restatement of the Frobenius / dense paths of
functions/cmtf_fun_AOADMM.m (whole file), functions/evaluate_stopping_conditions.m,
functions/make_exit_flag.m, functions/cp_func.m and functions/pca_func.m.
Non-Frobenius (L-BFGS-B) branches (:128-130, :136, :612-613, :1365-1418) are out of scope
(component 1) and raise.  EM imputation of missing entries (Z.miss,
:408-441, masked objective :1224-1226 / :1249-1252) is restated.

Conventions (mirroring the MATLAB structs, 1-based labels kept where they are DATA):
  Z : dict with
      'object'  list (len P): ndarray (CP / matrix) or list of K ndarrays (PAR2 slices X_k, I x J_k)
      'model'   list of 'CP' | 'PAR2'
      'modes'   list of lists of 1-based global mode ids
      'size'    list (len nb_modes): int, or list of J_k for the PAR2 B_k mode
      'coupling' dict: 'lin_coupled_modes' (len nb_modes, 0 = uncoupled, else coupling id 1..),
                 'coupling_type' (per coupling id), 'coupl_trafo_matrices' (per mode or None),
                 optional 'coupl_trafo_matrices2'
      'constrained_modes', 'constraints', 'weights', 'loss_function', optional 'ridge'
      'prox_operators', 'reg_func' (lists of callables; filled by constraints_to_prox when absent)
  G : dict 'fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac' (lists per mode),
      'coupling_fac' (list per coupling id), 'P', 'DeltaB', 'mu_DeltaB' (lists per object, PAR2 only)
  options : dict with the fields of example_script6...m:120-132.
"""
import copy
import time

import numpy as np
import scipy.linalg

from .prox import constraints_to_prox
from .tensor_ops import mttkrp


def evaluate_stopping_conditions(f_tensors, f_couplings, f_constraints, f_PAR2_couplings,
                                 f_tensors_old, f_couplings_old, f_constraints_old,
                                 f_PAR2_couplings_old, options):
    """evaluate_stopping_conditions.m:3-44."""
    def one(f, f_old):
        if f_old > 0:
            rel = abs(f_old - f) / f_old
        else:
            rel = abs(f_old - f)
        return (f < options['AbsFuncTol']) or (rel < options['OuterRelTol'])
    return (one(f_tensors, f_tensors_old) and one(f_couplings, f_couplings_old)
            and one(f_constraints, f_constraints_old) and one(f_PAR2_couplings, f_PAR2_couplings_old))


def make_exit_flag(it, f_tensors, f_couplings, f_constraints, f_PAR2_couplings, options, illconditioned):
    """make_exit_flag.m:4-29."""
    if it > options['MaxOuterIters']:
        return 'maxIterations'
    if illconditioned:
        return 'illconditioned lin system'
    tol = options['AbsFuncTol']
    return {
        'f_tensors': 'AbsFuncTol' if f_tensors < tol else 'RelFuncTol',
        'f_couplings': 'AbsFuncTol' if f_couplings < tol else 'RelFuncTol',
        'f_constraints': 'AbsFuncTol' if f_constraints < tol else 'RelFuncTol',
        'f_PAR2_couplings': 'AbsFuncTol' if f_PAR2_couplings < tol else 'RelFuncTol',
    }


def cp_func(X, A, Znormsqr, weight):
    """cp_func.m:19-56."""
    R = A[0].shape[1]
    W = np.ones((R, R))
    for a in A:
        W = W * (a.T @ a)
    U = mttkrp(X, A, 0)
    f_2 = np.sum(A[0] * U)
    f_3 = np.sum(W)
    return weight * (Znormsqr - 2.0 * f_2 + f_3)


def pca_func(X, A, Znormsqr, weight):
    """pca_func.m:18-40."""
    U, V = A[0], A[1]
    f2 = 0.0
    for r in range(U.shape[1]):
        f2 += U[:, r] @ X @ V[:, r]
    W = (U.T @ U) * (V.T @ V)
    return weight * (Znormsqr - 2.0 * f2 + np.sum(W))


def _fro(x):
    return np.sqrt(np.sum(np.asarray(x) ** 2))


def _rsolve_chol(A_inner, L):
    """(A_inner / L') / L  (cmtf_fun_AOADMM.m:609): right-solve with the lower Cholesky factor."""
    # X L' = A  ->  L X' = A'
    Y = scipy.linalg.solve_triangular(L, A_inner.T, lower=True)
    # Z L = Y'  ->  L' Z' = Y
    Zt = scipy.linalg.solve_triangular(L.T, Y, lower=False)
    return Zt.T


def cmtf_fun_AOADMM(Z, Znorm_const, G, fh=None, gh=None, lscalar=None, uscalar=None, options=None,
                    mttkrp_fn=None, trace=None):
    """[G,out] = cmtf_fun_AOADMM(Z,Znorm_const,G,fh,gh,lscalar,uscalar,options)  (cmtf_fun_AOADMM.m:1).

    `mttkrp_fn(p, X, U, n)` optionally replaces the dense MTTKRP (used by the multi-rank CPU tests to
    inject a sharded + all-reduced product); default is the Tensor-Toolbox-style one.
    `trace`, if a list, receives (iter, mode, name, array) tuples for debugging/parity drilling.
    """
    G = copy.deepcopy(G)
    options = dict(options)
    if 'iter_start_PAR2Bkconstraint' not in options:
        options['iter_start_PAR2Bkconstraint'] = 0                                    # :7-9
    options.setdefault('bsum', 0)
    lin = list(Z['coupling']['lin_coupled_modes'])
    couplings = sorted(set(lin))                                                       # :10 unique() sorts
    nb_modes = len(Z['size'])
    modes = [list(mm) for mm in Z['modes']]
    P = len(Z['object'])
    which_p = [None] * (nb_modes + 1)                                                  # 1-based
    for i in range(1, nb_modes + 1):
        which_p[i] = [p for p in range(P) if i in modes[p]][0]
    for p in range(P):
        if Z['loss_function'][p] != 'Frobenius':
            raise NotImplementedError('oracle covers the Frobenius loss only (SURVEY.md section 2)')
    miss = Z.get('miss')
    has_missing = miss is not None and any(x is not None for x in miss)                # :29
    if has_missing:
        # Z.object is modified in place by the imputation (:420, :432): work on a copy
        Z = dict(Z)
        Z['object'] = [([np.array(xk, copy=True) for xk in X] if isinstance(X, list) else
                        (np.array(X, copy=True) if X is not None else None)) for X in Z['object']]
        miss = [([np.asarray(mk, dtype=bool) for mk in M] if isinstance(M, list) else
                 (np.asarray(M, dtype=bool) if M is not None else None)) for M in miss]
    state_missing = {'f_rel_missing': float('nan')}
    if 'prox_operators' not in Z or Z['prox_operators'] is None:
        prox_ops, reg_func = constraints_to_prox(Z['constrained_modes'], Z['constraints'], Z['size'])
    else:
        prox_ops, reg_func = Z['prox_operators'], Z.get('reg_func', [None] * nb_modes)
    constrained = [bool(c) for c in Z['constrained_modes']]
    weights = list(Z['weights'])
    ridge = Z.get('ridge')
    trafo = Z['coupling'].get('coupl_trafo_matrices') or [None] * nb_modes
    trafo2 = Z['coupling'].get('coupl_trafo_matrices2') or [None] * nb_modes
    ctypes_ = list(Z['coupling'].get('coupling_type', []))
    if mttkrp_fn is None:
        mttkrp_fn = lambda p, X, U, n: mttkrp(X, U, n)

    def fac(m):
        return G['fac'][m - 1]

    def K_of(p):
        return len(Z['size'][modes[p][1] - 1])

    def is_par2_mode(m, which):
        p = which_p[m]
        return Z['model'][p] == 'PAR2' and modes[p].index(m) == which - 1

    GtG = [None] * (nb_modes + 1)
    A = [None] * (nb_modes + 1)
    C = [None] * (nb_modes + 1)
    B = [None] * (nb_modes + 1)
    B2 = [None] * (nb_modes + 1)
    L = [None] * (nb_modes + 1)
    rho = [None] * (nb_modes + 1)
    HcI = {}
    last_m = [0] * P
    last_mttkrp = [None] * P
    last_had = [None] * P
    innerIters = []
    state = {'iter': 1}

    # ------------------------------------------------------------------ nested helpers
    def update_constraint(m, rho_m):
        """:1420-1429"""
        oldZ = G['constraint_fac'][m - 1]
        r = np.max(rho_m) if np.ndim(rho_m) > 0 and np.size(rho_m) > 1 else float(np.reshape(rho_m, -1)[0])
        G['constraint_fac'][m - 1] = prox_ops[m - 1](fac(m) + G['constraint_dual_fac'][m - 1], r)
        G['constraint_dual_fac'][m - 1] = G['constraint_dual_fac'][m - 1] + fac(m) - G['constraint_fac'][m - 1]
        return oldZ

    def eval_res_ADMM_constr(mds, oldZ):
        """:1079-1096"""
        pr = 0.0
        du = 0.0
        for mm in mds:
            pr += _fro(fac(mm) - G['constraint_fac'][mm - 1]) / _fro(fac(mm))
            scaling = _fro(G['constraint_dual_fac'][mm - 1])
            if scaling > 0:
                du += _fro(G['constraint_fac'][mm - 1] - oldZ[mm]) / scaling
            else:
                du += _fro(G['constraint_fac'][mm - 1] - oldZ[mm])
        return pr / len(mds), du / len(mds)

    def eval_res_ADMM_coupl(ctype, mds, coupl_id, oldDelta):
        """:1099-1210"""
        pr = 0.0
        du = 0.0
        Delta = G['coupling_fac'][coupl_id - 1]
        for mm in mds:
            H = trafo[mm - 1]
            H2 = trafo2[mm - 1]
            F = fac(mm)
            if ctype == 0:
                pr += _fro(F - Delta) / _fro(F)
                dterm = Delta - oldDelta
            elif ctype == 1:
                pr += _fro(H @ F - Delta) / _fro(H @ F)
                dterm = Delta - oldDelta
            elif ctype == 2:
                pr += _fro(F @ H - Delta) / _fro(F @ H)
                dterm = Delta - oldDelta
            elif ctype == 3:
                pr += _fro(F - H @ Delta) / _fro(F)
                dterm = H @ (Delta - oldDelta)
            elif ctype == 4:
                pr += _fro(F - Delta @ H) / _fro(F)
                dterm = (Delta - oldDelta) @ H
            elif ctype == 5:
                pr += _fro(H @ F - Delta @ H2) / _fro(F)
                dterm = (Delta - oldDelta) @ H2
            scaling = _fro(G['coupling_dual_fac'][mm - 1])
            if scaling > 0:
                du += _fro(dterm) / scaling
            else:
                du += _fro(dterm)
        return pr / len(mds), du / len(mds)

    def ADMM_constrained_only(A_m, L_m, m, p, options):
        """:591-623"""
        inner_iter = 1
        rp = np.inf
        rd = np.inf
        oldZ = [None] * (nb_modes + 1)
        while inner_iter <= options['MaxInnerIters'] and (rp > options['innerRelPrTol_constr']
                                                         or rd > options['innerRelDualTol_constr']):
            if is_par2_mode(m, 3):
                for kk in range(K_of(p)):
                    A_inner = A_m[kk] + rho[m][kk] / 2 * (G['constraint_fac'][m - 1][kk, :] - G['constraint_dual_fac'][m - 1][kk, :])
                    G['fac'][m - 1][kk, :] = scipy.linalg.cho_solve((L_m[kk], True), A_inner)
            else:
                A_inner = A_m + rho[m] / 2 * (G['constraint_fac'][m - 1] - G['constraint_dual_fac'][m - 1])
                G['fac'][m - 1] = _rsolve_chol(A_inner, L_m)
            oldZ[m] = update_constraint(m, rho[m])
            inner_iter += 1
            rp, rd = eval_res_ADMM_constr([m], oldZ)
        return inner_iter - 1

    def ADMM_B_Parafac2(A_m, L_m, m, p, rho_m, options):
        """:509-589"""
        K = K_of(p)
        inner_iter = 1
        rpc = rdc = rpk = rdk = np.inf
        oldP = [None] * K
        con_active = constrained[m - 1] and state['iter'] >= options['iter_start_PAR2Bkconstraint']
        while inner_iter <= options['MaxInnerIters'] and (rpk > options['innerRelPrTol_coupl'] or rpc > options['innerRelPrTol_constr']
                                                         or rdk > options['innerRelDualTol_coupl'] or rdc > options['innerRelDualTol_constr']):
            rpc = rdc = rpk = rdk = 0.0
            for kk in range(K):
                A_inner = A_m[kk] + rho_m[kk] / 2 * (G['P'][p][kk] @ G['DeltaB'][p] - G['mu_DeltaB'][p][kk])
                if con_active:
                    A_inner = A_inner + rho_m[kk] / 2 * (G['constraint_fac'][m - 1][kk] - G['constraint_dual_fac'][m - 1][kk])
                G['fac'][m - 1][kk] = _rsolve_chol(A_inner, L_m[kk])
                U, _, Vt = np.linalg.svd((G['fac'][m - 1][kk] + G['mu_DeltaB'][p][kk]) @ G['DeltaB'][p].T, full_matrices=False)
                oldP[kk] = G['P'][p][kk]
                G['P'][p][kk] = U @ Vt
            oldDeltaB = G['DeltaB'][p]
            newD = np.zeros_like(oldDeltaB)
            sum_rho_k = 0.0
            for kk in range(K):
                newD = newD + rho_m[kk] * G['P'][p][kk].T @ (G['fac'][m - 1][kk] + G['mu_DeltaB'][p][kk])
                sum_rho_k += rho_m[kk]
            G['DeltaB'][p] = newD / sum_rho_k
            for kk in range(K):
                G['mu_DeltaB'][p][kk] = G['mu_DeltaB'][p][kk] + G['fac'][m - 1][kk] - G['P'][p][kk] @ G['DeltaB'][p]
            if con_active:
                oldZ = G['constraint_fac'][m - 1]
                if Z['constraints'][m - 1][0] == 'tPARAFAC2':
                    G['constraint_fac'][m - 1] = prox_ops[m - 1]([G['fac'][m - 1][kk] + G['constraint_dual_fac'][m - 1][kk] for kk in range(K)], rho_m)
                else:
                    G['constraint_fac'][m - 1] = [prox_ops[m - 1](G['fac'][m - 1][kk] + G['constraint_dual_fac'][m - 1][kk], rho_m[kk]) for kk in range(K)]
                G['constraint_dual_fac'][m - 1] = list(G['constraint_dual_fac'][m - 1])
                for kk in range(K):
                    G['constraint_dual_fac'][m - 1][kk] = G['constraint_dual_fac'][m - 1][kk] + G['fac'][m - 1][kk] - G['constraint_fac'][m - 1][kk]
                    rpc += _fro(G['fac'][m - 1][kk] - G['constraint_fac'][m - 1][kk]) / _fro(G['fac'][m - 1][kk]) / K
                    scaling = _fro(G['constraint_dual_fac'][m - 1][kk])
                    if scaling > 0:
                        rdc += _fro(oldZ[kk] - G['constraint_fac'][m - 1][kk]) / scaling / K
                    else:
                        rdc += _fro(oldZ[kk] - G['constraint_fac'][m - 1][kk]) / K
            for kk in range(K):
                rpk += _fro(G['fac'][m - 1][kk] - G['P'][p][kk] @ G['DeltaB'][p]) / _fro(G['fac'][m - 1][kk]) / K
                with np.errstate(divide='ignore', invalid='ignore'):
                    rdk += _fro(oldP[kk] @ oldDeltaB - G['P'][p][kk] @ G['DeltaB'][p]) / _fro(G['mu_DeltaB'][p][kk]) / K   # unguarded (:584)
            inner_iter += 1
        return inner_iter - 1

    def ADMM_coupled(ctype, coupled_modes, coupl_id, options):
        """ADMM_coupled_case0..5 (:625-1075) folded into one function; branches follow the case."""
        inner_iter = 1
        rpk = rpc = rdk = rdc = np.inf
        oldZ = [None] * (nb_modes + 1)
        cid = coupl_id - 1
        while inner_iter <= options['MaxInnerIters'] and (rpk > options['innerRelPrTol_coupl'] or rpc > options['innerRelPrTol_constr']
                                                         or rdk > options['innerRelDualTol_coupl'] or rdc > options['innerRelDualTol_constr']):
            Delta = G['coupling_fac'][cid]
            for mm in coupled_modes:
                pp = which_p[mm]
                H = trafo[mm - 1]
                H2 = trafo2[mm - 1]
                mu = G['coupling_dual_fac'][mm - 1]
                if is_par2_mode(mm, 3):
                    K = K_of(pp)
                    if ctype in (1, 5):
                        rhoC = np.mean(rho[mm])
                        A_large = np.concatenate([A[mm][kk].reshape(-1) for kk in range(K)])
                        if ctype == 1:
                            inner = (Delta.T - mu.T).reshape(-1, order='F')
                        else:
                            inner = ((Delta @ H2).T - mu.T).reshape(-1, order='F')
                        A_inner = A_large + rhoC / 2 * (HcI[mm].T @ inner)
                        if constrained[mm - 1]:
                            A_inner = A_inner + rhoC / 2 * (G['constraint_fac'][mm - 1].T - G['constraint_dual_fac'][mm - 1].T).reshape(-1, order='F')
                        vec = scipy.linalg.cho_solve((L[mm], True), A_inner)
                        Rm = fac(mm).shape[1]
                        G['fac'][mm - 1] = vec.reshape((Rm, fac(mm).shape[0]), order='F').T.copy()
                    else:
                        for kk in range(K):
                            if ctype == 0:
                                cpl = Delta[kk, :] - mu[kk, :]
                            elif ctype == 2:
                                cpl = (Delta[kk, :] - mu[kk, :]) @ H.T
                            elif ctype == 3:
                                cpl = H[kk, :] @ Delta - mu[kk, :]
                            elif ctype == 4:
                                cpl = Delta[kk, :] @ H - mu[kk, :]
                            A_inner = A[mm][kk].reshape(-1) + rho[mm][kk] / 2 * cpl
                            if constrained[mm - 1]:
                                A_inner = A_inner + rho[mm][kk] / 2 * (G['constraint_fac'][mm - 1][kk, :] - G['constraint_dual_fac'][mm - 1][kk, :])
                            G['fac'][mm - 1][kk, :] = scipy.linalg.cho_solve((L[mm][kk], True), A_inner)
                else:
                    if ctype == 0:
                        cpl = Delta - mu
                    elif ctype == 1:
                        cpl = H.T @ (Delta - mu)
                    elif ctype == 2:
                        cpl = (Delta - mu) @ H.T
                    elif ctype == 3:
                        cpl = H @ Delta - mu
                    elif ctype == 4:
                        cpl = Delta @ H - mu
                    elif ctype == 5:
                        cpl = H.T @ (Delta @ H2 - mu)
                    A_inner = A[mm] + rho[mm] / 2 * cpl
                    if constrained[mm - 1]:
                        A_inner = A_inner + rho[mm] / 2 * (G['constraint_fac'][mm - 1] - G['constraint_dual_fac'][mm - 1])
                    if ctype in (1, 5):
                        G['fac'][mm - 1] = scipy.linalg.solve_sylvester(B2[mm], B[mm], A_inner)      # :728, :1016
                    else:
                        G['fac'][mm - 1] = _rsolve_chol(A_inner, L[mm])

            # ---- Delta update
            oldDelta = G['coupling_fac'][cid]
            if ctype == 0:                                                                        # :661-675
                newD = np.zeros_like(oldDelta)
                sum_rho = 0.0
                for jj in coupled_modes:
                    if is_par2_mode(jj, 3):
                        newD = newD + np.asarray(rho[jj])[:, None] * (fac(jj) + G['coupling_dual_fac'][jj - 1])
                    else:
                        newD = newD + rho[jj] * (fac(jj) + G['coupling_dual_fac'][jj - 1])
                    sum_rho = sum_rho + np.asarray(rho[jj])
                if np.ndim(sum_rho) > 0 and np.size(sum_rho) > 1:
                    G['coupling_fac'][cid] = (1.0 / sum_rho)[:, None] * newD
                else:
                    G['coupling_fac'][cid] = (1.0 / float(sum_rho)) * newD
            elif ctype == 1:                                                                      # :738-749
                newD = np.zeros_like(oldDelta)
                sum_rho = 0.0
                for jj in coupled_modes:
                    newD = newD + np.sum(rho[jj]) * (trafo[jj - 1] @ fac(jj) + G['coupling_dual_fac'][jj - 1])
                    sum_rho += np.sum(rho[jj])
                G['coupling_fac'][cid] = 1.0 / sum_rho * newD
            elif ctype == 2:                                                                      # :808-815
                newD = np.zeros_like(oldDelta)
                sum_rho = 0.0
                for jj in coupled_modes:
                    rj = np.asarray(rho[jj])
                    term = fac(jj) @ trafo[jj - 1] + G['coupling_dual_fac'][jj - 1]
                    newD = newD + (rj[:, None] * term if rj.ndim > 0 and rj.size > 1 else float(rj) * term)
                    sum_rho = sum_rho + rj
                if np.ndim(sum_rho) > 0 and np.size(sum_rho) > 1:
                    G['coupling_fac'][cid] = (1.0 / sum_rho)[:, None] * newD
                else:
                    G['coupling_fac'][cid] = (1.0 / float(sum_rho)) * newD
            elif ctype == 3:                                                                      # :873-881
                H1 = trafo[coupled_modes[0] - 1]
                AA = np.zeros((H1.shape[1], H1.shape[1]))
                BB = np.zeros((H1.shape[1], fac(coupled_modes[0]).shape[1]))
                for jj in coupled_modes:
                    Hj = trafo[jj - 1]
                    rj = np.asarray(rho[jj])
                    rcol = rj[:, None] if rj.ndim > 0 and rj.size > 1 else float(rj)
                    AA = AA + Hj.T @ (rcol * Hj)
                    BB = BB + Hj.T @ (rcol * (fac(jj) + G['coupling_dual_fac'][jj - 1]))
                G['coupling_fac'][cid] = np.linalg.solve(AA, BB)
            elif ctype in (4, 5):                                                                 # :938-963, :1026-1054
                if ctype == 4:
                    Hk = trafo
                    H1 = trafo[coupled_modes[0] - 1]
                    AA = np.zeros((H1.shape[0], H1.shape[0]))
                    BB = np.zeros((fac(coupled_modes[0]).shape[0], H1.shape[0]))
                else:
                    Hk = trafo2
                    H21 = trafo2[coupled_modes[0] - 1]
                    AA = np.zeros((H21.shape[0], H21.shape[0]))
                    BB = np.zeros((trafo[coupled_modes[0] - 1].shape[0], H21.shape[0]))
                PAR2_flag = False
                AA_PAR2 = None
                for jj in coupled_modes:
                    if ctype == 5:
                        rhoC = np.mean(rho[coupled_modes[-1]])     # `mm` left over from the solve loop (:1032)
                    Hj = Hk[jj - 1]
                    if is_par2_mode(jj, 3):
                        PAR2_flag = True
                        AAA = Hj @ Hj.T
                        AA_PAR2 = [rho[jj][kk] * AAA for kk in range(K_of(which_p[jj]))]
                    else:
                        AA = AA + (rho[jj] if ctype == 4 else rhoC) * Hj @ Hj.T
                    if ctype == 4:
                        rj = np.asarray(rho[jj])
                        rcol = rj[:, None] if rj.ndim > 0 and rj.size > 1 else float(rj)
                        BB = BB + (rcol * (fac(jj) + G['coupling_dual_fac'][jj - 1])) @ Hj.T
                    else:
                        BB = BB + rhoC * (trafo[jj - 1] @ fac(jj) + G['coupling_dual_fac'][jj - 1]) @ Hj.T
                if PAR2_flag:
                    newD = np.array(oldDelta, copy=True)
                    for kk in range(newD.shape[0]):
                        newD[kk, :] = np.linalg.solve((AA + AA_PAR2[kk]).T, BB[kk, :])
                    G['coupling_fac'][cid] = newD
                else:
                    G['coupling_fac'][cid] = np.linalg.solve(AA.T, BB.T).T
            Delta = G['coupling_fac'][cid]

            # ---- duals / constraints
            for mm in coupled_modes:
                H = trafo[mm - 1]
                H2 = trafo2[mm - 1]
                F = fac(mm)
                if ctype == 0:
                    upd = F - Delta
                elif ctype == 1:
                    upd = H @ F - Delta
                elif ctype == 2:
                    upd = F @ H - Delta
                elif ctype == 3:
                    upd = F - H @ Delta
                elif ctype == 4:
                    upd = F - Delta @ H
                elif ctype == 5:
                    upd = H @ F - Delta @ H2
                G['coupling_dual_fac'][mm - 1] = G['coupling_dual_fac'][mm - 1] + upd
                if constrained[mm - 1]:
                    oldZ[mm] = update_constraint(mm, rho[mm])
            inner_iter += 1
            rpk, rdk = eval_res_ADMM_coupl(ctype, coupled_modes, coupl_id, oldDelta)
            cm = [mm for mm in coupled_modes if constrained[mm - 1]]
            if cm:
                rpc, rdc = eval_res_ADMM_constr(cm, oldZ)
            else:
                rpc, rdc = 0.0, 0.0
        return inner_iter - 1

    def func_eval(first):
        """CMTF_AOADMM_func_eval (:1213-1363)."""
        fp = np.zeros(P)
        for pp in range(P):
            U = [G['fac'][mm - 1] for mm in modes[pp]]
            if has_missing and miss[pp] is not None and Z['model'][pp] == 'CP':                # :1224-1226
                from .tensor_ops import full_ktensor
                Mm = miss[pp] * full_ktensor(U)
                fp[pp] = weights[pp] * (Znorm_const[pp] - 2 * np.sum(Z['object'][pp] * Mm) + np.sum(Mm ** 2))
            elif has_missing and miss[pp] is not None:                                           # :1249-1252
                m1, m2, m3 = modes[pp]
                for kk in range(K_of(pp)):
                    Rk = Z['object'][pp][kk] - fac(m1) @ np.diag(fac(m3)[kk, :]) @ fac(m2)[kk].T
                    fp[pp] += _fro(miss[pp][kk] * Rk) ** 2
                fp[pp] = weights[pp] * fp[pp]
            elif Z['model'][pp] == 'CP':
                if first:
                    if Z['object'][pp].ndim >= 3:
                        fp[pp] = cp_func(Z['object'][pp], U, Znorm_const[pp], weights[pp])
                    else:
                        fp[pp] = pca_func(Z['object'][pp], U, Znorm_const[pp], weights[pp])
                else:
                    f_2 = np.sum(last_mttkrp[pp] * fac(last_m[pp]))
                    f_3 = np.sum(last_had[pp] * GtG[last_m[pp]])
                    fp[pp] = weights[pp] * (Znorm_const[pp] - 2 * f_2 + f_3)
            else:
                m1, m2, m3 = modes[pp]
                if (not first) and last_m[pp] == 1:
                    f_2 = np.sum(last_mttkrp[pp] * fac(m1))
                    f_3 = np.sum(last_had[pp] * GtG[m1])
                    fp[pp] = Znorm_const[pp] - 2 * f_2 + f_3
                else:
                    for kk in range(K_of(pp)):
                        fp[pp] += _fro(Z['object'][pp][kk] - fac(m1) @ np.diag(fac(m3)[kk, :]) @ fac(m2)[kk].T) ** 2
                fp[pp] = weights[pp] * fp[pp]
        f_tensors = float(np.sum(fp))
        for n in range(nb_modes):                                                        # :1272-1288
            if reg_func[n] is not None:
                if isinstance(G['constraint_fac'][n], list):
                    if Z['constraints'][n][0] == 'tPARAFAC2':
                        f_tensors += reg_func[n](G['fac'][n])
                    else:
                        for kk in range(len(G['constraint_fac'][n])):
                            f_tensors += reg_func[n](G['fac'][n][kk])
                else:
                    f_tensors += reg_func[n](G['fac'][n])
        if ridge is not None:                                                            # :1290-1300
            for n in range(nb_modes):
                if isinstance(G['fac'][n], list):
                    cf = G['constraint_fac'][n]                      # :1293 loops over length(G.constraint_fac{n})
                    for kk in range(len(cf) if isinstance(cf, list) else 0):
                        f_tensors += ridge[n] * _fro(G['fac'][n][kk]) ** 2
                else:
                    f_tensors += ridge[n] * _fro(G['fac'][n]) ** 2
        nb_couplings = max(lin) if lin else 0                                            # :1303-1329
        coupling_p = np.zeros(max(nb_couplings, 0))
        for n in range(1, nb_couplings + 1):
            ct = ctypes_[n - 1]
            cmodes = [mm for mm in range(1, nb_modes + 1) if lin[mm - 1] == n]
            Delta = G['coupling_fac'][n - 1]
            for mm in cmodes:
                F = fac(mm)
                H = trafo[mm - 1]
                H2 = trafo2[mm - 1]
                if ct == 0:
                    coupling_p[n - 1] += _fro(F - Delta) / _fro(F)
                elif ct == 1:
                    coupling_p[n - 1] += _fro(H @ F - Delta) / _fro(H @ F)
                elif ct == 2:
                    coupling_p[n - 1] += _fro(F @ H - Delta) / _fro(F @ H)
                elif ct == 3:
                    coupling_p[n - 1] += _fro(F - H @ Delta) / _fro(F)
                elif ct == 4:
                    coupling_p[n - 1] += _fro(F - Delta @ H) / _fro(F)
                elif ct == 5:
                    coupling_p[n - 1] += _fro(H @ F - Delta @ H2) / _fro(H @ F)
        f_couplings = float(np.sum(coupling_p))
        if f_couplings > 0:
            f_couplings = f_couplings / np.count_nonzero(coupling_p)
        f_constraint_p = np.zeros(nb_modes)                                              # :1332-1348
        for n in range(nb_modes):
            cf = G['constraint_fac'][n]
            if cf is not None and not (isinstance(cf, (list, np.ndarray)) and len(cf) == 0):
                if isinstance(cf, list):
                    for kk in range(len(cf)):
                        f_constraint_p[n] += _fro(G['fac'][n][kk] - cf[kk]) / _fro(G['fac'][n][kk])
                    f_constraint_p[n] /= len(cf)
                else:
                    f_constraint_p[n] = _fro(G['fac'][n] - cf) / _fro(G['fac'][n])
        f_constraints = float(np.sum(f_constraint_p))
        if f_constraints > 0:
            f_constraints = f_constraints / np.count_nonzero(f_constraint_p)
        f_par2_p = np.zeros(P)                                                           # :1351-1362
        for pp in range(P):
            if Z['model'][pp] == 'PAR2':
                m2 = modes[pp][1]
                for kk in range(K_of(pp)):
                    f_par2_p[pp] += _fro(fac(m2)[kk] - G['P'][pp][kk] @ G['DeltaB'][pp]) / _fro(fac(m2)[kk])
        f_PAR2 = float(np.sum(f_par2_p))
        if f_PAR2 > 0:
            szl = Z['size'][modes[P - 1][1] - 1]                 # loop variable pp left at P (:1361)
            f_PAR2 = f_PAR2 / (len(szl) if isinstance(szl, (list, tuple, np.ndarray)) else 1)
        return f_tensors, f_couplings, f_constraints, f_PAR2

    # ------------------------------------------------------------------ main (:32-506)
    f_tensors, f_couplings, f_constraints, f_PAR2 = func_eval(True)
    func_val = [f_tensors]
    func_coupl = [f_couplings]
    func_constr = [f_constraints]
    func_PAR2_coupl = [f_PAR2]
    func_rel_missing = [float('nan')]                                                  # :38-40
    tstart = time.perf_counter()
    time_at_it = [0.0]

    for m in range(1, nb_modes + 1):                                                     # :62-81
        p = which_p[m]
        if Z['model'][p] == 'CP':
            GtG[m] = fac(m).T @ fac(m)
        else:
            pos = modes[p].index(m) + 1
            if pos == 1:
                GtG[m] = fac(m).T @ fac(m)
            elif pos == 2:
                GtG[m] = [fk.T @ fk for fk in fac(m)]

    stop = False
    it = 1
    while it <= options['MaxOuterIters'] and not stop:
        state['iter'] = it
        innerIters.append(np.zeros(nb_modes))
        for coupl_id in couplings:
            coupled_modes = [mm for mm in range(1, nb_modes + 1) if lin[mm - 1] == coupl_id]
            for p in sorted(set(which_p[mm] for mm in coupled_modes)):
                w = weights[p]
                if Z['model'][p] == 'CP':
                    for m in [mm for mm in coupled_modes if which_p[mm] == p]:
                        X = Z['object'][p]
                        if X.ndim >= 3:                                                 # :96-103
                            A[m] = w * mttkrp_fn(p, X, [G['fac'][j - 1] for j in modes[p]], modes[p].index(m))
                            C[m] = np.ones_like(GtG[m])
                            for j in modes[p]:
                                if j != m:
                                    C[m] = C[m] * GtG[j]
                        else:                                                           # :105-113
                            if modes[p].index(m) == 0:
                                A[m] = w * X @ fac(modes[p][1])
                                C[m] = GtG[modes[p][1]]
                            else:
                                A[m] = w * X.T @ fac(modes[p][0])
                                C[m] = GtG[modes[p][0]]
                        Rm = C[m].shape[0]
                        rho[m] = np.trace(C[m]) / Rm                                    # :115
                        B[m] = w * C[m]
                        if ridge is not None:
                            B[m] = B[m] + ridge[m - 1] * np.eye(Rm)
                        last_mttkrp[p] = A[m] * 1 / w
                        last_had[p] = C[m]
                        last_m[p] = m
                        if options['bsum']:
                            A[m] = A[m] + options['bsum_weight'] / 2 * fac(m)
                            B[m] = B[m] + options['bsum_weight'] / 2 * np.eye(Rm)
                        if trace is not None:
                            trace.append((it, m, 'mttkrp', A[m].copy()))
                        if coupl_id == 0:
                            if not constrained[m - 1]:
                                G['fac'][m - 1] = np.linalg.solve(B[m].T, A[m].T).T      # A/B (:134)
                                inner_iters = 1
                            else:
                                B[m] = B[m] + rho[m] / 2 * np.eye(Rm)
                                L[m] = np.linalg.cholesky(B[m].T)
                                inner_iters = ADMM_constrained_only(A[m], L[m], m, p, options)
                            innerIters[-1][m - 1] = inner_iters
                            GtG[m] = fac(m).T @ fac(m)
                else:  # PAR2 (:157-250)
                    m1, m2, m3 = modes[p]
                    K = K_of(p)
                    for m in [mm for mm in coupled_modes if which_p[mm] == p]:
                        pos = modes[p].index(m) + 1
                        if pos == 1:
                            Rm = fac(m).shape[1]
                            A[m] = np.zeros_like(fac(m))
                            C[m] = np.zeros((Rm, Rm))
                            for k in range(K):
                                ck = fac(m3)[k, :]
                                A[m] = A[m] + Z['object'][p][k] @ fac(m2)[k] @ np.diag(ck)
                                C[m] = C[m] + np.diag(ck) @ GtG[m2][k] @ np.diag(ck)
                            last_had[p] = C[m]
                            last_mttkrp[p] = A[m]
                            last_m[p] = 1
                            A[m] = w * A[m]
                            rho[m] = np.trace(C[m]) / Rm
                            B[m] = w * C[m]
                            if ridge is not None:
                                B[m] = B[m] + ridge[m - 1] * np.eye(Rm)
                            if options['bsum']:
                                A[m] = A[m] + options['bsum_weight'] / 2 * fac(m)
                                B[m] = B[m] + options['bsum_weight'] / 2 * np.eye(Rm)
                            if coupl_id == 0:
                                if not constrained[m - 1]:
                                    G['fac'][m - 1] = np.linalg.solve(B[m].T, A[m].T).T
                                    inner_iters = 1
                                else:
                                    B[m] = B[m] + rho[m] / 2 * np.eye(Rm)
                                    L[m] = np.linalg.cholesky(B[m].T)
                                    inner_iters = ADMM_constrained_only(A[m], L[m], m, p, options)
                            # NOTE (:189-190): executed also for the coupled case with a stale inner_iters;
                            # the coupled branch overwrites innerIters(m,iter) at :392.
                            if coupl_id == 0:
                                innerIters[-1][m - 1] = inner_iters
                            GtG[m] = fac(m).T @ fac(m)
                        elif pos == 2:
                            Rm = fac(m1).shape[1]
                            A[m] = [None] * K
                            C[m] = [None] * K
                            B[m] = [None] * K
                            L[m] = [None] * K
                            rho[m] = np.zeros(K)
                            for k in range(K):
                                ck = fac(m3)[k, :]
                                A[m][k] = w * Z['object'][p][k].T @ fac(m1) @ np.diag(ck)
                                C[m][k] = np.diag(ck) @ GtG[m1] @ np.diag(ck)
                                rho[m][k] = np.trace(C[m][k]) / Rm
                                if 'increase_factor_rhoBk' in options:
                                    rho[m][k] = options['increase_factor_rhoBk'] * rho[m][k]
                                B[m][k] = w * C[m][k]
                                B[m][k] = B[m][k] + rho[m][k] / 2 * np.eye(Rm)
                                if ridge is not None:
                                    B[m][k] = B[m][k] + ridge[m - 1] * np.eye(Rm)
                                if options['bsum']:
                                    A[m][k] = A[m][k] + options['bsum_weight'] / 2 * fac(m)[k]
                                    B[m][k] = B[m][k] + options['bsum_weight'] / 2 * np.eye(Rm)
                                last_m[p] = 2
                                if constrained[m - 1] and it >= options['iter_start_PAR2Bkconstraint']:
                                    B[m][k] = B[m][k] + rho[m][k] / 2 * np.eye(Rm)
                                L[m][k] = np.linalg.cholesky(B[m][k])
                            inner_iters = ADMM_B_Parafac2(A[m], L[m], m, p, rho[m], options)
                            innerIters[-1][m - 1] = inner_iters
                            GtG[m] = [fk.T @ fk for fk in fac(m)]
                        else:
                            Rm = fac(m1).shape[1]
                            A[m] = [None] * K
                            C[m] = [None] * K
                            B[m] = [None] * K
                            L[m] = [None] * K
                            rho[m] = np.zeros(K)
                            inner_iters = 0
                            for k in range(K):
                                A[m][k] = w * np.diag(fac(m1).T @ Z['object'][p][k] @ fac(m2)[k]).copy()
                                C[m][k] = GtG[m1] * GtG[m2][k]
                                rho[m][k] = np.trace(C[m][k]) / Rm
                                B[m][k] = w * C[m][k]
                                if ridge is not None:
                                    B[m][k] = B[m][k] + ridge[m - 1] * np.eye(Rm)
                                last_m[p] = 3
                                if options['bsum']:
                                    A[m][k] = A[m][k] + options['bsum_weight'] / 2 * fac(m)[k, :]
                                    B[m][k] = B[m][k] + options['bsum_weight'] / 2 * np.eye(Rm)
                                if coupl_id == 0:
                                    if not constrained[m - 1]:
                                        G['fac'][m - 1][k, :] = np.linalg.solve(B[m][k], A[m][k])
                                        inner_iters = 1
                                    else:
                                        B[m][k] = B[m][k] + rho[m][k] / 2 * np.eye(Rm)
                                        L[m][k] = np.linalg.cholesky(B[m][k].T)
                            if constrained[m - 1] and coupl_id == 0:
                                inner_iters = ADMM_constrained_only(A[m], L[m], m, p, options)
                            if coupl_id == 0:
                                innerIters[-1][m - 1] = inner_iters

            if coupl_id != 0:                                                            # :253-404
                ctype = ctypes_[coupl_id - 1]
                for m in coupled_modes:
                    p = which_p[m]
                    par2c = is_par2_mode(m, 3)
                    H = trafo[m - 1]
                    if ctype in (0, 2, 3, 4):
                        if par2c:
                            for k in range(K_of(p)):
                                Rm = B[m][k].shape[0]
                                if ctype == 2:
                                    B[m][k] = B[m][k] + rho[m][k] / 2 * H @ H.T
                                else:
                                    B[m][k] = B[m][k] + rho[m][k] / 2 * np.eye(Rm)
                                if constrained[m - 1]:
                                    B[m][k] = B[m][k] + rho[m][k] / 2 * np.eye(Rm)
                                L[m][k] = np.linalg.cholesky(B[m][k].T)
                        else:
                            Rm = B[m].shape[0]
                            if ctype == 2:
                                B[m] = B[m] + rho[m] / 2 * H @ H.T
                            else:
                                B[m] = B[m] + rho[m] / 2 * np.eye(Rm)
                            if constrained[m - 1]:
                                B[m] = B[m] + rho[m] / 2 * np.eye(Rm)
                            L[m] = np.linalg.cholesky(B[m].T)
                    else:  # 1, 5
                        if par2c:
                            Rm = fac(m).shape[1]
                            HcI[m] = np.kron(H, np.eye(Rm))
                            B2[m] = np.mean(rho[m]) / 2 * (HcI[m].T @ HcI[m])
                            B2[m] = scipy.linalg.block_diag(*B[m]) + B2[m]
                        else:
                            B2[m] = rho[m] / 2 * H.T @ H
                        if constrained[m - 1]:
                            B2[m] = B2[m] + np.mean(rho[m]) / 2 * np.eye(B2[m].shape[0])
                        if par2c:
                            L[m] = np.linalg.cholesky(B2[m])
                inner_iters = ADMM_coupled(ctype, coupled_modes, coupl_id, options)
                for m in coupled_modes:
                    innerIters[-1][m - 1] = inner_iters
                    if not is_par2_mode(m, 3):
                        GtG[m] = fac(m).T @ fac(m)

        if has_missing:                                                                  # :408-441 EM imputation
            from .tensor_ops import full_ktensor
            num_sq = 0.0
            den_sq = 0.0
            for p in range(P):
                if miss[p] is None:
                    continue
                if Z['model'][p] == 'CP':
                    M_full = full_ktensor([G['fac'][j - 1] for j in modes[p]])
                    mm = ~miss[p]
                    old_vals = Z['object'][p][mm]
                    new_vals = M_full[mm]
                    Z['object'][p][mm] = new_vals
                    num_sq += float(np.sum((new_vals - old_vals) ** 2))
                    den_sq += float(np.sum(old_vals ** 2))
                else:
                    m1, m2, m3 = modes[p]
                    for k in range(K_of(p)):
                        M_k = fac(m1) @ np.diag(fac(m3)[k, :]) @ fac(m2)[k].T
                        mk = ~miss[p][k]
                        old_k = Z['object'][p][k][mk]
                        new_k = M_k[mk]
                        num_sq += float(np.sum((new_k - old_k) ** 2))
                        den_sq += float(np.sum(old_k ** 2))
                        Z['object'][p][k][mk] = new_k
            state_missing['f_rel_missing'] = np.sqrt(num_sq / den_sq) if den_sq > 0 else np.sqrt(num_sq)
        f_tensors_old, f_couplings_old, f_constraints_old, f_PAR2_old = f_tensors, f_couplings, f_constraints, f_PAR2
        f_tensors, f_couplings, f_constraints, f_PAR2 = func_eval(False)
        func_val.append(f_tensors)
        func_coupl.append(f_couplings)
        func_constr.append(f_constraints)
        func_PAR2_coupl.append(f_PAR2)
        time_at_it.append(time.perf_counter() - tstart)
        stop = evaluate_stopping_conditions(f_tensors, f_couplings, f_constraints, f_PAR2,
                                            f_tensors_old, f_couplings_old, f_constraints_old, f_PAR2_old, options)
        func_rel_missing.append(state_missing['f_rel_missing'])                          # :454
        if has_missing:
            stop = stop and (state_missing['f_rel_missing'] < options['OuterRelTol'])    # :457-459
        it += 1

    out = {
        'f_tensors': f_tensors, 'f_couplings': f_couplings, 'f_constraints': f_constraints,
        'f_PAR2_couplings': f_PAR2, 'f_rel_missing': float(state_missing['f_rel_missing']),
        'func_rel_missing': np.array(func_rel_missing),
        'exit_flag': make_exit_flag(it, f_tensors, f_couplings, f_constraints, f_PAR2, options, 0),
        'OuterIterations': it - 1,
        'func_val_conv': np.array(func_val), 'func_coupl_conv': np.array(func_coupl),
        'func_constr_conv': np.array(func_constr), 'func_PAR2_coupl': np.array(func_PAR2_coupl),
        'time_at_it': np.array(time_at_it),
        'innerIters': np.array(innerIters).T if innerIters else np.zeros((nb_modes, 0)),
    }
    return G, out
