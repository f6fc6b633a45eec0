"""CPU: the oracle is validated by known-answer checks; the reference ships no
golden vectors (parity unpinned), so these are what pins the oracle."""
import itertools
import os

import numpy as np
import pytest

from oracle import problem_gen as pg
from oracle import prox as P
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM, evaluate_stopping_conditions, make_exit_flag
from oracle.tensor_ops import full_ktensor, khatrirao, mttkrp
from _cases import golden_case, rel


def test_mttkrp_vs_einsum():
    rng = np.random.RandomState(0)
    X = np.asfortranarray(rng.randn(7, 6, 5))
    U = [rng.randn(s, 3) for s in X.shape]
    assert rel(mttkrp(X, U, 0), np.einsum('ijk,jr,kr->ir', X, U[1], U[2])) < 1e-14
    assert rel(mttkrp(X, U, 1), np.einsum('ijk,ir,kr->jr', X, U[0], U[2])) < 1e-14
    assert rel(mttkrp(X, U, 2), np.einsum('ijk,ir,jr->kr', X, U[0], U[1])) < 1e-14
    X4 = rng.randn(3, 4, 5, 6)
    U4 = [rng.randn(s, 2) for s in X4.shape]
    assert rel(mttkrp(X4, U4, 1), np.einsum('ijkl,ir,kr,lr->jr', X4, U4[0], U4[2], U4[3])) < 1e-14


def test_full_ktensor_and_khatrirao():
    rng = np.random.RandomState(1)
    U = [rng.randn(4, 3), rng.randn(5, 3), rng.randn(6, 3)]
    assert rel(full_ktensor(U, [2.0, 1.0, 0.5]), np.einsum('r,ir,jr,kr->ijk', [2.0, 1.0, 0.5], *U)) < 1e-14
    kr = khatrirao([U[0], U[1]])
    assert kr.shape == (20, 3) and abs(kr[1 + 4 * 2, 1] - U[0][1, 1] * U[1][2, 1]) < 1e-15


def test_shortcut_objective_equals_explicit_residual():
    Z, G, _ = pg.config_script6(seed=0, sz=(12, 14, 10, 12, 16, 14, 18))
    Gout, out = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=15))
    f = 0.0
    for p, X in enumerate(Z['object']):
        f += Z['weights'][p] * np.sum((X - full_ktensor([Gout['fac'][m - 1] for m in Z['modes'][p]])) ** 2)
    assert abs(f - out['f_tensors']) < 1e-13


def test_noise_free_recovery_and_attainable_fit():
    Z, G, info = pg.config_script6(seed=1, noise=0.0, sz=(15, 16, 14, 15, 18, 16, 20))
    Gout, out = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=2000, AbsFuncTol=1e-10))
    assert out['f_tensors'] < 1e-5
    Z, G, info = pg.config_script6(seed=0)   # noise 0.2 -> attainable fit ~ 100/(1+0.04) = 96.15 %
    Gout, out = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options())
    assert out['OuterIterations'] == 102     # the value the survey probe measured with RandomState(0)
    for p, X in enumerate(Z['object']):
        fit = 100 * (1 - np.sum((X - full_ktensor([Gout['fac'][m - 1] for m in Z['modes'][p]])) ** 2) / np.sum(X ** 2))
        assert 95.5 < fit < 97.0


def test_ordering_quirks():
    """uncoupled modes first; mixed constrained/unconstrained group (script 6: mode 2 unconstrained, mode 6 constrained)."""
    Z, G, _ = pg.config_script6(seed=0, sz=(10, 12, 8, 10, 14, 12, 16))
    tr = []
    cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=1), trace=tr)
    assert [t[1] for t in tr] == [3, 5, 7, 1, 4, 2, 6]
    Gout, out = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=3))
    assert Gout['constraint_fac'][1] is None and Gout['constraint_fac'][5] is not None
    assert out['innerIters'][2, 0] == 1      # unconstrained least-squares update counts one inner iteration


def _brute_force_projection(v, feasible_grid):
    best, bx = np.inf, None
    for x in feasible_grid:
        d = np.sum((np.asarray(x) - v) ** 2)
        if d < best:
            best, bx = d, np.asarray(x)
    return bx


def test_simplex_and_l1_projection_kkt():
    rng = np.random.RandomState(0)
    for _ in range(20):
        v = rng.randn(9) * 2
        w = P._project_simplex_vec(v, 1.5)
        assert abs(w.sum() - 1.5) < 1e-12 and w.min() >= 0
        theta = (v - w)[w > 0]
        assert np.ptp(theta) < 1e-12 and np.all((v - theta[0])[w == 0] <= 1e-12)   # KKT
    X = rng.randn(11, 4) * 3
    Y = P.project_L1(X, 2.0)
    assert np.all(np.abs(Y).sum(axis=0) <= 2.0 + 1e-12)
    small = rng.randn(5, 2) * 0.01
    assert np.array_equal(P.project_L1(small, 2.0), small)
    assert rel(P.project_simplex(X.T, 1.0, 2), P.project_simplex(X, 1.0, 1).T) < 1e-15


def test_tv_prox_optimality():
    """Condat's output must satisfy the TV-prox optimality conditions: x = y - D'u, |u|<=lam, u_i = lam*sign(dx_i)."""
    rng = np.random.RandomState(3)
    for n, lam in [(1, 0.5), (2, 0.3), (17, 0.4), (64, 1.5), (200, 0.05)]:
        y = np.cumsum(rng.randn(n)) * 0.3 + rng.randn(n) * 0.2
        x = P.tv_condat(y, lam)
        u = np.cumsum(y - x)[:-1]            # dual variable
        assert abs(np.sum(y - x)) < 1e-10
        if n > 1:
            assert np.all(np.abs(u) <= lam + 1e-10)
            dx = np.diff(x)
            assert np.all(np.abs(u[dx > 1e-12] + lam) < 1e-9) or np.all(np.abs(np.abs(u[np.abs(dx) > 1e-12]) - lam) < 1e-9)
    assert np.array_equal(P.tv_condat(np.array([1.0, 2.0, 0.5]), 0.0), np.array([1.0, 2.0, 0.5]))
    assert rel(P.tv_condat(np.array([0.0, 10.0]), 100.0), np.array([5.0, 5.0])) < 1e-15


def test_monotone_and_unimodal_bruteforce():
    rng = np.random.RandomState(5)
    grid = np.linspace(-1, 1, 5)
    for _ in range(5):
        v = rng.choice(grid, size=4)
        iso = P._pava_nondecreasing(v)
        assert np.all(np.diff(iso) >= -1e-14)
        # PAVA optimum is at least as good as any non-decreasing grid vector
        cands = [c for c in itertools.product(grid, repeat=4) if all(np.diff(c) >= 0)]
        bf = _brute_force_projection(v, cands)
        assert np.sum((iso - v) ** 2) <= np.sum((bf - v) ** 2) + 1e-12
        uni = P.project_unimodal_vector(v, False)
        k = int(np.argmax(uni))
        assert np.all(np.diff(uni[:k + 1]) >= -1e-14) and np.all(np.diff(uni[k:]) <= 1e-14)
        cands = [c for c in itertools.product(grid, repeat=4)
                 if any(all(np.diff(c[:m + 1]) >= 0) and all(np.diff(c[m:]) <= 0) for m in range(4))]
        bf = _brute_force_projection(v, cands)
        assert np.sum((uni - v) ** 2) <= np.sum((bf - v) ** 2) + 1e-12
    x = rng.randn(30)
    u = P.project_unimodal_vector(x, True)
    assert u.min() >= 0
    assert rel(P.project_unimodal_vector(np.array([0., 1., 3., 2., 1.]), False), np.array([0., 1., 3., 2., 1.])) < 1e-15


def test_elementwise_and_norm_prox_definitions():
    rng = np.random.RandomState(7)
    X = rng.randn(13, 5)
    assert np.array_equal(P.project_box(X, 0, np.inf), np.maximum(X, 0))
    assert np.array_equal(P.prox_abs(X, 0.3), np.sign(X) * np.maximum(np.abs(X) - 0.3, 0))
    Y = P.project_L2(X, 1.0)
    assert np.all(np.linalg.norm(Y, axis=0) <= 1 + 1e-14)
    Y = P.prox_normalized_nonneg(X)
    assert np.allclose(np.linalg.norm(Y, axis=0), 1) and Y.min() >= 0
    neg = -np.abs(X)
    Y = P.prox_normalized_nonneg(neg)
    assert np.all(Y.sum(axis=0) == 1) and np.all(Y.argmax(axis=0) == neg.argmax(axis=0))
    g = 0.8
    Y = P.prox_L2(X, g)
    for r in range(5):
        n = np.linalg.norm(X[:, r])
        assert rel(Y[:, r], X[:, r] * max(0, 1 - g / n)) < 1e-15
    ops, regs = P.constraints_to_prox([1, 1, 1], [('GL smoothness', 0.5), ('TV regularization', 0.1), ('l1 regularization', 0.2)],
                                      [13, 13, 13])
    L = P.gl_laplacian(13)
    assert rel((2 * 0.5 / 2.0 * L + np.eye(13)) @ ops[0](X, 2.0), X) < 1e-13
    assert abs(regs[1](X) - 0.1 * np.sum(X[-1] - X[0])) < 1e-12         # telescoping, no abs (constraints_to_prox.m:81)
    assert abs(regs[2](X) - 0.2 * np.abs(X).sum()) < 1e-12
    with pytest.raises(ValueError):
        P.constraints_to_prox([1], [None], [3])


def test_stopping_conditions_and_exit_flag():
    o = dict(AbsFuncTol=1e-4, OuterRelTol=1e-8, MaxOuterIters=10)
    assert evaluate_stopping_conditions(1e-5, 0, 0, 0, 1.0, 0, 0, 0, o)          # abs tol; zeros: abs change 0 < tol
    assert not evaluate_stopping_conditions(0.5, 0, 0, 0, 1.0, 0, 0, 0, o)
    assert evaluate_stopping_conditions(1.0, 0, 0, 0, 1.0 + 1e-9, 0, 0, 0, o)    # relative change
    assert not evaluate_stopping_conditions(1.0, 0.5, 0, 0, 1.0, 0.0, 0, 0, o)   # f_old<=0 -> absolute change
    assert make_exit_flag(11, 1, 1, 1, 1, o, 0) == 'maxIterations'
    fl = make_exit_flag(5, 1e-5, 1, 1e-9, 0, o, 0)
    assert fl == {'f_tensors': 'AbsFuncTol', 'f_couplings': 'RelFuncTol', 'f_constraints': 'AbsFuncTol',
                  'f_PAR2_couplings': 'AbsFuncTol'}


def test_par2_oracle_converges_noise_free():
    Z, G, _ = pg.config_cp_par2(I=10, J=9, K=8, Jk=12, Kp=6, R=2, seed=1)
    Gout, out = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(
        MaxOuterIters=400, AbsFuncTol=1e-9, innerRelPrTol_coupl=1e-5, innerRelPrTol_constr=1e-5,
        innerRelDualTol_coupl=1e-5, innerRelDualTol_constr=1e-5))
    assert out['func_val_conv'][-1] < 1e-3 * out['func_val_conv'][0]
    for k, Pk in enumerate(Gout['P'][1]):
        assert rel(Pk.T @ Pk, np.eye(2)) < 1e-10                                   # P_k'P_k = I (:532-534)


@pytest.mark.parametrize('name', ['script6_small', 'cp_matrix_small', 'cp_tv_small'])
def test_oracle_reproduces_golden(name):
    Z, G, opts, gold = golden_case(name)
    Gout, out = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=opts)
    assert out['OuterIterations'] == int(gold['OuterIterations'])
    assert np.max(np.abs(out['func_val_conv'] - gold['func_val_conv'])) < 1e-12
    for i, F in enumerate(Gout['fac']):
        assert rel(F, gold['out_fac_%d' % i]) < 1e-10
    assert np.array_equal(out['innerIters'], gold['innerIters'])


@pytest.mark.parametrize('ctype', [1, 2, 3, 4, 5])
def test_linear_coupling_oracle_satisfies_the_coupling_constraint(ctype):
    """Noise-free data generated WITH the coupling: the oracle's ADMM_coupled_case1..5 drive both the fit and the
    coupling residual ||G(F) - D(Delta)|| / ||.|| to (near) zero - a known answer that needs no MATLAB."""
    Z, G, _ = pg.config_linear_coupling(ctype, seed=3, noise=0.0, constrained=False)
    Go, oo = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=400, AbsFuncTol=1e-9))
    # type 5 with partially shared components stalls in a local minimum of the fit (f ~ 1.4e-3, coupling 3e-8)
    assert oo['f_tensors'] < (5e-3 if ctype == 5 else 1e-5) and oo['f_couplings'] < 1e-3
    H = Z['coupling']['coupl_trafo_matrices']
    H2 = Z['coupling'].get('coupl_trafo_matrices2', [None] * 6)
    D = Go['coupling_fac'][0]
    for m in (1, 4):
        F = Go['fac'][m - 1]
        lhs = {1: lambda: H[m - 1] @ F, 2: lambda: F @ H[m - 1], 3: lambda: F, 4: lambda: F, 5: lambda: H[m - 1] @ F}[ctype]()
        rhs = {1: lambda: D, 2: lambda: D, 3: lambda: H[m - 1] @ D, 4: lambda: D @ H[m - 1], 5: lambda: D @ H2[m - 1]}[ctype]()
        assert np.linalg.norm(lhs - rhs) / np.linalg.norm(lhs) < 5e-3


def test_em_imputation_oracle_recovers_missing_entries():
    """Known answer for the EM restatement (cmtf_fun_AOADMM.m:408-441): with 20 % of a low-noise CP tensor hidden, the
    model fitted on the observed entries predicts the hidden ones, and f_rel_missing decreases towards 0."""
    Z, G, _ = pg.config_cp_matrix(30, 26, 22, 40, 3, seed=4, noise=0.02)
    Zm = pg.add_missing(Z, 0.2, seed=1, objects=[0])
    Go, oo = cmtf_fun_AOADMM(Zm, pg.znorm_const(Zm), G, options=pg.default_options(MaxOuterIters=150))
    hidden = ~Zm['miss'][0]
    Xhat = full_ktensor([Go['fac'][m] for m in range(3)])
    assert np.corrcoef(Z['object'][0][hidden], Xhat[hidden])[0, 1] > 0.99
    assert np.isnan(oo['func_rel_missing'][0]) and oo['func_rel_missing'][1] > 10 * oo['f_rel_missing']
    assert oo['f_rel_missing'] < 1e-2
    # without a mask the same call reports NaN (cmtf_fun_AOADMM.m:30)
    _, o2 = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=3))
    assert np.isnan(o2['f_rel_missing'])


def _fms(U, V):
    U = U / np.linalg.norm(U, axis=0)
    V = V / np.linalg.norm(V, axis=0)
    M = np.abs(U.T @ V)
    return max(np.mean([M[i, p[i]] for i in range(U.shape[1])]) for p in itertools.permutations(range(U.shape[1])))


def test_oracle_recovers_the_references_own_ground_truth_script11():
    """The only fixtures with a known answer that the reference ships: noisy_dataset.mat + gnd_factors.mat of
    example_script11_tPARAFAC2.m (tests/golden/script11_tparafac2.npz, see make_script11_fixture.py).  With the script's
    own configuration the oracle must find the ground-truth factors (factor match scores as computed at :150-158)."""
    import make_script11_fixture as mk
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'script11_tparafac2.npz'))
    Z, G = mk.script11_problem(fx['dataset'])
    Go, oo = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=mk.script11_options(100))
    K = fx['C'].shape[0]
    assert _fms(Go['fac'][0], fx['A']) > 0.99
    assert _fms(Go['fac'][2], fx['C']) > 0.99
    assert _fms(np.vstack(Go['fac'][1]), np.vstack([fx['B'][k] for k in range(K)])) > 0.95
    assert oo['func_val_conv'][-1] < oo['func_val_conv'][0]
    assert np.all(Go['fac'][2] >= -1e-12 * np.abs(Go['fac'][2]).max() - 1e-3)     # C is (nearly) non-negative via Z_C
    # and the committed oracle state is reproducible
    Go30, oo30 = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=mk.script11_options(30))
    assert rel(Go30['fac'][0], fx['oracle_A']) < 1e-9 and rel(oo30['func_val_conv'], fx['oracle_func_val']) < 1e-10


def test_oracle_nvecs_known_answers():
    """cmtf_nvecs.m:33-58 restatement: (i) equals the left singular vectors of the explicit unfolding, (ii) spans the
    range of the true factor for noise-free low-rank data, (iii) the nvecs branch of the init (:50-69) gives
    orthonormal factors, ones for PARAFAC2 mode C."""
    rng = np.random.RandomState(0)
    A, B, C = rng.rand(12, 3), rng.rand(9, 3), rng.rand(7, 3)
    X = np.einsum('ir,jr,kr->ijk', A, B, C)
    Z = {'model': ['CP'], 'modes': [[1, 2, 3]], 'size': [12, 9, 7], 'object': [X]}
    for n, F in ((1, A), (2, B), (3, C)):
        U = pg.cmtf_nvecs(Z, n, 3)
        Xn = np.moveaxis(X, n - 1, 0).reshape(X.shape[n - 1], -1, order='F')
        Us = np.linalg.svd(Xn, full_matrices=False)[0][:, :3]
        assert np.allclose(np.abs(U.T @ Us), np.eye(3), atol=1e-8)
        assert np.linalg.norm(F - U @ (U.T @ F)) < 1e-10
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-12)
    Zp, _, _ = pg.config_cp_par2(seed=1, noise=0.05)
    io = {'lambdas_init': [[1.0] * 3] * 2, 'nvecs': 1, 'distr': [pg.d_rand] * len(Zp['size']), 'normalize': 1}
    G = pg.init_coupled_AOADMM_CMTF(Zp, io, np.random.RandomState(2))
    pm = Zp['modes'][1]
    assert np.allclose(G['fac'][pm[2] - 1], 1.0)
    for Fk in G['fac'][pm[1] - 1]:
        assert np.allclose(Fk.T @ Fk, np.eye(3), atol=1e-12)
    assert np.allclose(G['fac'][pm[0] - 1].T @ G['fac'][pm[0] - 1], np.eye(3), atol=1e-12)


def test_oracle_fixed_point_satisfies_kkt_of_the_coupled_problem():
    """Independent of the algorithm: at the fixed point of the restated AO-ADMM on the example_script6 structure (3-way CP
    + two matrices, two exact couplings, non-negativity on 6 of 7 modes) the factors are a KKT point of
        min  sum_p w_p || X_p - [[factors]] ||^2   s.t.  coupled factors equal, constrained factors >= 0:
    the gradients of the coupled objects are individually non-zero but cancel (or are complementary to the active
    bounds), the unconstrained mode has zero gradient."""
    from oracle.tensor_ops import mttkrp
    Z, G, _ = pg.config_script6(seed=3, noise=0.1)
    opts = pg.default_options(MaxOuterIters=2000, AbsFuncTol=0.0, OuterRelTol=1e-15, MaxInnerIters=20,
                              innerRelPrTol_coupl=1e-8, innerRelPrTol_constr=1e-8, innerRelDualTol_coupl=1e-8,
                              innerRelDualTol_constr=1e-8)
    Go, oo = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=opts)
    assert oo['f_couplings'] < 1e-12 and oo['f_constraints'] < 1e-12
    F = Go['fac']
    g = {}
    for p, ms in enumerate(Z['modes']):
        X, w = Z['object'][p], Z['weights'][p]
        for pos, m in enumerate(ms):
            had = np.ones((F[m - 1].shape[1],) * 2)
            for q in ms:
                if q != m:
                    had = had * (F[q - 1].T @ F[q - 1])
            if X.ndim > 2:
                M = mttkrp(X, [F[q - 1] for q in ms], pos)
            else:
                M = X @ F[ms[1] - 1] if pos == 0 else X.T @ F[ms[0] - 1]
            g[m] = 2 * w * (F[m - 1] @ had - M)
    lin = Z['coupling']['lin_coupled_modes']
    groups = {}
    for m in range(1, 8):
        groups.setdefault(('c', lin[m - 1]) if lin[m - 1] else ('m', m), []).append(m)
    for key, ms in groups.items():
        total = sum(g[m] for m in ms)
        Fm = F[ms[0] - 1]
        for m in ms:
            assert np.linalg.norm(F[m - 1] - Fm) < 1e-10                        # exact coupling holds
        if any(Z['constrained_modes'][m - 1] for m in ms):
            assert Fm.min() > -1e-12
            viol = np.linalg.norm(np.minimum(Fm, total))                         # F >= 0, grad >= 0, F .* grad = 0
        else:
            viol = np.linalg.norm(total)
        assert viol < 1e-9, (key, viol)
        if len(ms) > 1:                                                          # the single gradients do NOT vanish
            assert min(np.linalg.norm(g[m]) for m in ms) > 1e-5


def test_oracle_fixed_point_satisfies_kkt_with_l1_and_l2_ball():
    """example_script10 structure with the l1 variant: min ||X - [[A,B,C]]||^2 + eta*||A||_1 s.t. ||b_r||, ||c_r|| <= 1.
    At the fixed point: 0 in grad_A + eta*sign(A) (|grad| <= eta on the zeros, about half of the entries at this eta),
    and on the ball-constrained modes every column gradient is -lambda_r * column with lambda_r >= 0 and ||column|| = 1."""
    from oracle.tensor_ops import mttkrp
    eta = 6e-2
    Z, G, _ = pg.config_cp_tv(I=30, J=20, K=25, R=3, seed=2, noise=0.3, eta=eta, mode1=('l1 regularization',))
    opts = pg.default_options(MaxOuterIters=2500, AbsFuncTol=0.0, OuterRelTol=1e-15, MaxInnerIters=20,
                              innerRelPrTol_coupl=1e-8, innerRelPrTol_constr=1e-8, innerRelDualTol_coupl=1e-8,
                              innerRelDualTol_constr=1e-8)
    Go, oo = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=opts)
    F, Zc = Go['fac'], Go['constraint_fac']
    X = Z['object'][0]
    g = []
    for pos in range(3):
        had = np.ones((3, 3))
        for q in range(3):
            if q != pos:
                had = had * (F[q].T @ F[q])
        g.append(2 * (F[pos] @ had - mttkrp(X, F, pos)))
    A = Zc[0]
    assert np.linalg.norm(F[0] - A) < 1e-10
    nz = np.abs(A) > 1e-12
    assert 0.2 < nz.mean() < 0.9                                   # genuinely sparse, genuinely non-trivial
    assert np.abs(g[0][nz] + eta * np.sign(A[nz])).max() < 1e-9
    assert np.maximum(np.abs(g[0][~nz]) - eta, 0).max() < 1e-9
    for pos in (1, 2):
        for r in range(3):
            col, gc = F[pos][:, r], g[pos][:, r]
            lam = -(gc @ col) / (col @ col)
            assert abs(np.linalg.norm(col) - 1.0) < 1e-9 and lam > 0
            assert np.linalg.norm(gc + lam * col) < 1e-9


def test_oracle_parafac2_fixed_point_satisfies_kkt():
    """PARAFAC2 (irregular slices, non-negative A and C): at the fixed point of the restated ADMM_B_Parafac2 scheme the
    factors are a stationary point of  min sum_k ||X_k - A D_k B_k'||^2  s.t.  B_k = P_k DeltaB, P_k'P_k = I, A, C >= 0:
    the coupling holds, every P_k is the polar factor of X_k' A D_k DeltaB' (P_k'M_k symmetric positive definite), the
    gradient w.r.t. DeltaB (through all slices) and w.r.t. C vanishes, A satisfies complementarity."""
    Z, G, _ = pg.config_single_par2(I=14, Jk=(10, 8, 11, 9, 12), R=3, seed=3, noise=0.1)
    opts = pg.default_options(MaxOuterIters=1200, AbsFuncTol=0.0, OuterRelTol=1e-15, MaxInnerIters=10,
                              innerRelPrTol_coupl=1e-9, innerRelPrTol_constr=1e-9, innerRelDualTol_coupl=1e-9,
                              innerRelDualTol_constr=1e-9)
    Go, oo = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=opts)
    m1, m2, m3 = Z['modes'][0]
    A, Bk, C = Go['fac'][m1 - 1], Go['fac'][m2 - 1], Go['fac'][m3 - 1]
    P, dB = Go['P'][0], Go['DeltaB'][0]
    X = Z['object'][0]
    gA, gC, gdB = np.zeros_like(A), np.zeros_like(C), np.zeros_like(dB)
    for k in range(len(X)):
        Dk = np.diag(C[k])
        E = X[k] - A @ Dk @ Bk[k].T
        gA += -2 * E @ Bk[k] @ Dk
        gC[k] = -2 * np.diag(A.T @ E @ Bk[k])
        gdB += P[k].T @ (-2 * E.T @ A @ Dk)
        assert np.linalg.norm(Bk[k] - P[k] @ dB) < 1e-5 * np.linalg.norm(Bk[k])
        assert np.linalg.norm(P[k].T @ P[k] - np.eye(3)) < 1e-12
        S = P[k].T @ (X[k].T @ A @ Dk @ dB.T)
        assert np.linalg.norm(S - S.T) < 1e-5 * np.linalg.norm(S)
        assert np.linalg.eigvalsh((S + S.T) / 2).min() > 0
    assert A.min() > -1e-9
    assert np.linalg.norm(np.minimum(A, gA)) < 1e-4                   # A >= 0, grad >= 0, A .* grad = 0
    assert np.linalg.norm(gC) < 1e-6 and C.min() > 0
    assert np.linalg.norm(gdB) < 1e-6


def test_oracle_linear_coupling_fixed_point_satisfies_kkt():
    """Linear couplings H F = Delta (type 1, Sylvester solves) and F = Delta H (type 4, Cholesky solves), unconstrained:
    at the fixed point the coupling equations hold and the gradients of the two objects — each of order 1e-3 — cancel in
    the combination the constraint prescribes (H4' g1 + g4 = 0, resp. g1 H1' + g4 H4' = 0); the uncoupled modes have zero
    gradient."""
    from oracle.tensor_ops import mttkrp
    opts = pg.default_options(MaxOuterIters=1500, AbsFuncTol=0.0, OuterRelTol=1e-15, MaxInnerIters=20,
                              innerRelPrTol_coupl=1e-9, innerRelPrTol_constr=1e-9, innerRelDualTol_coupl=1e-9,
                              innerRelDualTol_constr=1e-9)
    for ct in (1, 4):
        Z, G, _ = pg.config_linear_coupling(ct, seed=ct, constrained=False)
        Go, _ = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=opts)
        F, D, H = Go['fac'], Go['coupling_fac'][0], Z['coupling']['coupl_trafo_matrices']
        g = {}
        for p, ms in enumerate(Z['modes']):
            X, w = Z['object'][p], Z['weights'][p]
            for pos, m in enumerate(ms):
                R = F[m - 1].shape[1]
                had = np.ones((R, R))
                for q in ms:
                    if q != m:
                        had = had * (F[q - 1].T @ F[q - 1])
                M = mttkrp(X, [F[q - 1] for q in ms], pos) if X.ndim > 2 else (X @ F[ms[1] - 1] if pos == 0 else X.T @ F[ms[0] - 1])
                g[m] = 2 * w * (F[m - 1] @ had - M)
        if ct == 1:
            feas = [H[0] @ F[0] - D, H[3] @ F[3] - D]
            stat = g[4] + H[3].T @ g[1]                    # H1 = I: F1 = H4 F4
        else:
            feas = [F[0] - D @ H[0], F[3] - D @ H[3]]
            stat = g[1] @ H[0].T + g[4] @ H[3].T
        assert max(np.linalg.norm(f) for f in feas) < 1e-8, ct
        assert np.linalg.norm(stat) < 1e-8 and min(np.linalg.norm(g[1]), np.linalg.norm(g[4])) > 1e-4, ct
        assert max(np.linalg.norm(g[m]) for m in (2, 3, 5)) < 1e-8, ct


def test_oracle_em_fixed_point_is_a_kkt_point_of_the_masked_problem():
    """Missing data (Z.miss, EM imputation :408-441): at the fixed point the imputed entries equal the model, so the
    factors must be a KKT point of the MASKED problem  min sum_observed (x - [[A,B,C]])^2, A,B,C >= 0  — checked with the
    explicitly masked residual, which the algorithm never forms."""
    from oracle.tensor_ops import full_ktensor, mttkrp
    nn = ('non-negativity',)
    Z, G, _ = pg.config_single_cp(sz=(14, 12, 10), R=3, seed=4, noise=0.1, constraints=[nn, nn, nn])
    Zm = pg.add_missing(Z, 0.3, seed=2)
    X0, M = Zm['object'][0].copy(), Zm['miss'][0]
    opts = pg.default_options(MaxOuterIters=2000, AbsFuncTol=0.0, OuterRelTol=1e-15, MaxInnerIters=20,
                              innerRelPrTol_coupl=1e-9, innerRelPrTol_constr=1e-9, innerRelDualTol_coupl=1e-9,
                              innerRelDualTol_constr=1e-9)
    Go, oo = cmtf_fun_AOADMM(Zm, pg.znorm_const(Zm), G, options=opts)
    assert oo['f_rel_missing'] < 1e-10
    F = Go['fac']
    E = np.where(M, X0 - full_ktensor(F), 0.0)
    active = 0
    for pos in range(3):
        g = -2 * mttkrp(E, F, pos)
        assert F[pos].min() > -1e-12 and np.linalg.norm(np.minimum(F[pos], g)) < 1e-9, pos
        active += np.linalg.norm(g) > 1e-4
    assert active >= 1            # at least one mode sits on its bounds with a non-zero gradient
