"""GPU (-m gpu): parity of the CUDA engine with the CPU oracle THROUGH THE C ABI, on identical seeded inputs.

Tolerances (north_star): factor matrices 1e-8 relative Frobenius error, objective/fit 1e-10; operator-level
checks are held to rounding level (1e-12)."""
import numpy as np
import pytest

from oracle import problem_gen as pg
from oracle import prox as P
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve
from oracle.tensor_ops import mttkrp as oracle_mttkrp
from _cases import FAC_TOL, FIT_TOL, ZERO_TOL, assert_state_close, golden_case, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _need_gpu(ab):
    assert ab.device_count() >= 1, 'GPU tests need a CUDA device (the engine has no CPU fallback)'


# ---------------------------------------------------------------------------------------------- MTTKRP
@pytest.mark.parametrize('shape,R', [
    ((50, 60, 40), 3),            # example_script6 sizes
    ((1, 1, 1), 1), ((2, 3, 4), 2), ((3, 129, 5), 7),   # degenerate / ragged
    ((130, 70, 33), 8), ((64, 64, 64), 16), ((200, 150, 90), 32), ((129, 257, 65), 64),
    ((96, 80, 72), 100), ((40, 36, 30), 256),           # multi-chunk ranks, maximum rank
    ((51, 61, 41), 5),            # odd leading dimension (padded on upload)
    ((40, 30), 4), ((300, 500), 32), ((7, 1), 3),       # matrices (cmtf_fun_AOADMM.m:106-113)
    ((20, 12, 10, 8), 6), ((9, 8, 7, 6, 5), 4),         # N-way tensors
])
def test_mttkrp_all_modes(ab, shape, R):
    rng = np.random.RandomState(sum(shape) + R)
    X = np.asfortranarray(rng.randn(*shape))
    U = [rng.randn(s, R) for s in shape]
    for n in range(1, len(shape) + 1):
        got = ab.mttkrp(X, U, n)
        assert rel(got, oracle_mttkrp(X, U, n - 1)) < 1e-12, (shape, R, n)


def test_mttkrp_linearity_and_scaling_large(ab):
    """size-independent properties at a size the oracle would not finish quickly: linearity in X and in a factor."""
    rng = np.random.RandomState(0)
    shape, R = (384, 320, 256), 32
    X1 = np.asfortranarray(rng.randn(*shape))
    X2 = np.asfortranarray(rng.randn(*shape))
    U = [rng.randn(s, R) for s in shape]
    for n in (1, 2, 3):
        a = ab.mttkrp(X1, U, n)
        b = ab.mttkrp(X2, U, n)
        c = ab.mttkrp(X1 + 2.0 * X2, U, n)
        assert rel(c, a + 2.0 * b) < 1e-12
    # rank-one tensor: closed form  M1 = a * (b'B .* c'C)
    a_, b_, c_ = rng.randn(shape[0]), rng.randn(shape[1]), rng.randn(shape[2])
    Xr = np.asfortranarray(np.einsum('i,j,k->ijk', a_, b_, c_))
    M1 = ab.mttkrp(Xr, U, 1)
    assert rel(M1, np.outer(a_, (b_ @ U[1]) * (c_ @ U[2]))) < 1e-12


def test_mttkrp_deterministic(ab):
    rng = np.random.RandomState(1)
    X = np.asfortranarray(rng.randn(150, 140, 130))
    U = [rng.randn(s, 16) for s in X.shape]
    for n in (1, 2, 3):
        assert np.array_equal(ab.mttkrp(X, U, n), ab.mttkrp(X, U, n))


# ---------------------------------------------------------------------------------------------- small ops
@pytest.mark.parametrize('rows,R', [(1, 1), (5, 3), (500, 7), (4097, 32), (3000, 64), (700, 100), (300, 256)])
def test_gram_and_chol_solve(ab, rows, R):
    rng = np.random.RandomState(rows + R)
    F = rng.randn(rows, R)
    G = ab.gram(F)
    assert rel(G, F.T @ F) < 1e-13
    B = F.T @ F / max(rows, 1) + (np.trace(F.T @ F) / rows / R + 0.1) * np.eye(R)
    A = rng.randn(max(rows // 3, 1), R)
    X = ab.chol_solve(B, A)
    assert rel(X @ B, A) < 1e-11
    assert rel(X, np.linalg.solve(B.T, A.T).T) < 1e-10


def test_chol_failure_is_reported(ab):
    B = np.array([[1.0, 2.0], [2.0, 1.0]])   # indefinite: chol() throws in the reference (cmtf_fun_AOADMM.m:142)
    with pytest.raises(ab.AoadmmError) as e:
        ab.chol_solve(B, np.ones((3, 2)))
    assert e.value.status_name == 'NOT_POSITIVE_DEFINITE'


CONSTRAINTS = [('non-negativity',), ('box', -0.3, 0.5), ('simplex column-wise', 1.0), ('simplex row-wise', 2.0),
               ('non-decreasing',), ('non-increasing',), ('unimodality', True), ('unimodality', False), ('l1-ball', 3.0),
               ('l2-ball', 1.0), ('non-negative l2-ball', 1.0), ('non-negative l2-sphere', 1.0), ('l1 regularization', 0.2),
               ('l0 regularization', 0.2), ('l2 regularization', 2.0), ('ridge', 0.3), ('GL smoothness', 0.5),
               ('TV regularization', 0.7)]


@pytest.mark.parametrize('con', CONSTRAINTS, ids=[c[0] + str(c[1:]) for c in CONSTRAINTS])
@pytest.mark.parametrize('rows,cols', [(1, 1), (2, 3), (200, 6), (2049, 8)])
def test_prox_matches_oracle(ab, con, rows, cols):
    rng = np.random.RandomState(rows * 7 + cols)
    V = rng.randn(rows, cols)
    if rows > 100:
        V[:, 0] = -np.abs(V[:, 0])         # an all-negative column (the non-negative sphere special case)
        V[rows // 2:, 1] = V[rows // 2, 1]  # ties / plateaus
    ops, _ = P.constraints_to_prox([1], [con], [rows])
    for rho in (1.0, 0.37):
        assert rel(ab.prox(con, V, rho=rho), ops[0](V, rho)) < 1e-12, (con, rows, cols, rho)


@pytest.mark.parametrize('rows', [2, 3, 17, 1000, 2327, 2328, 6000])
@pytest.mark.parametrize('eta', [1e-3, 0.3, 50.0])
def test_prox_tv_both_algorithms(ab, rows, eta):
    """TV prox: the dynamic-programming kernel (rows <= 2327, workspace in shared memory) and the direct kernel (longer
    columns) against the oracle's direct algorithm: noisy, piecewise-constant, constant and tied columns; tiny, moderate
    and huge eta (one flat segment)."""
    rng = np.random.RandomState(rows)
    V = np.stack([rng.randn(rows), np.repeat(rng.randn(rows // 20 + 1), 20)[:rows] + 0.05 * rng.randn(rows),
                  np.full(rows, 0.7), np.round(rng.randn(rows), 1), np.cumsum(rng.randn(rows)) * 0.1], axis=1)
    con = ('TV regularization', eta)
    ops, _ = P.constraints_to_prox([1], [con], [rows])
    for rho in (1.0, 0.05):
        got, ref = ab.prox(con, V, rho=rho), ops[0](V, rho)
        assert np.max(np.abs(got - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref))), (rows, eta, rho)


def test_prox_large_column_uses_global_scratch(ab):
    rng = np.random.RandomState(2)
    V = np.cumsum(rng.randn(20000, 2), axis=0) * 0.05
    for con in [('TV regularization', 0.5), ('non-decreasing',)]:
        ops, _ = P.constraints_to_prox([1], [con], [V.shape[0]])
        assert rel(ab.prox(con, V, rho=1.0), ops[0](V, 1.0)) < 1e-12
    Y = ab.prox(('simplex column-wise', 1.0), V)
    assert np.allclose(Y.sum(axis=0), 1.0) and Y.min() >= 0


def test_prox_idempotent_projections(ab):
    rng = np.random.RandomState(3)
    V = rng.randn(300, 5)
    for con in [('non-negativity',), ('simplex column-wise', 1.0), ('l2-ball', 1.0), ('unimodality', True),
                ('non-decreasing',), ('l1-ball', 2.0), ('non-negative l2-sphere', 1.0)]:
        once = ab.prox(con, V)
        assert rel(ab.prox(con, once), once) < 1e-12, con


@pytest.mark.parametrize('rows,cols', [(5, 5), (40, 3), (300, 16), (2049, 64), (500, 200)])
def test_prox_orthonormal_polar_factor(ab, rows, cols):
    """'orthonormal' (project_ortho.m:3-4: U*V' of the thin SVD) by one-sided Jacobi."""
    V = np.random.RandomState(rows + cols).randn(rows, cols)
    Y = ab.prox(('orthonormal',), V)
    assert rel(Y, P.project_ortho(V)) < 1e-11
    assert np.linalg.norm(Y.T @ Y - np.eye(cols)) < 1e-12 * cols


@pytest.mark.parametrize('rows,cols', [(1, 1), (7, 2), (120, 6), (700, 16)])
def test_prox_quadratic_regularization(ab, rows, cols):
    """'quadratic regularization' (constraints_to_prox.m:62-66) with a symmetric L: graph Laplacian and an indefinite one."""
    rng = np.random.RandomState(rows)
    V = rng.randn(rows, cols)
    S = rng.randn(rows, rows)
    for L, eta in ((P.gl_laplacian(rows), 0.5), (0.05 * (S + S.T) / max(rows, 1) ** 0.5, 0.3)):
        con = ('quadratic regularization', eta, L)
        ops, _ = P.constraints_to_prox([1], [con], [rows])
        for rho in (1.0, 0.37):
            assert rel(ab.prox(con, V, rho=rho), ops[0](V, rho)) < 1e-10, (rows, cols, rho)
    with pytest.raises(ab.AoadmmError) as e:
        ab.prox(('quadratic regularization', 0.1, np.triu(np.ones((rows, rows))) if rows > 1 else np.ones((2, 2))), V)
    assert e.value.status_name in ('UNSUPPORTED', 'INVALID_ARG')


def test_orthonormal_and_quadratic_in_the_loop(ab):
    I, J, K, R = 30, 26, 22, 3
    L = P.gl_laplacian(I)
    for cons in ([('quadratic regularization', 1e-2, L), ('orthonormal',), ('non-negativity',)],
                 [('orthonormal',), ('l2-ball', 1.0), ('quadratic regularization', 1e-3, P.gl_laplacian(K))]):
        Z, G, _ = pg.config_single_cp(sz=(I, J, K), R=R, seed=12, noise=0.1, constraints=cons, distr=[pg.d_randn] * 3)
        Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=20))
        _assert_out_close(od, oo)
        assert_state_close(Gd, Go)


def test_tparafac2_constraint(ab):
    """example_script11 style: temporal smoothness across the B_k of a regular PARAFAC2 (t_smoothness_prox.m)."""
    Z, G, _ = pg.config_tparafac2(seed=2, eta=0.05)   # B_k drifting smoothly over k, K=9 slices of 16 x 14
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=25))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)
    Zbad = dict(Z, size=[Z['size'][0], [14] * 8 + [13], 9])   # unequal slices: the penalty is undefined
    Zbad['object'] = [Z['object'][0][:8] + [Z['object'][0][8][:, :13]]]
    Gbad = dict(G)
    for key in ('fac', 'constraint_fac', 'constraint_dual_fac'):
        Gbad[key] = list(G[key])
        Gbad[key][1] = G[key][1][:8] + [G[key][1][8][:13]]
    Gbad['P'] = [G['P'][0][:8] + [G['P'][0][8][:13]]]
    Gbad['mu_DeltaB'] = [G['mu_DeltaB'][0][:8] + [G['mu_DeltaB'][0][8][:13]]]
    with pytest.raises(ab.AoadmmError):
        ab.cmtf_fun_AOADMM(Zbad, pg.znorm_const(Zbad), Gbad, None, None, None, None, pg.default_options(MaxOuterIters=2))


def test_custom_constraint_is_unsupported(ab):
    with pytest.raises(ab.AoadmmError) as e:
        ab.prox(('custom', None), np.zeros((3, 2)))
    assert e.value.status_name == 'UNSUPPORTED'


# ---------------------------------------------------------------------------------------------- solver
def _both(ab, Z, G, opts):
    zn = pg.znorm_const(Z)
    Go, oo = oracle_solve(Z, zn, G, options=opts)
    Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts)
    return Go, oo, Gd, od


def _assert_out_close(od, oo):
    assert od['OuterIterations'] == oo['OuterIterations']
    n = oo['OuterIterations'] + 1
    for key in ('func_val_conv', 'func_coupl_conv', 'func_constr_conv'):
        assert np.max(np.abs(od[key][:n] - oo[key][:n])) < FIT_TOL, key
    assert np.array_equal(od['innerIters'], oo['innerIters'])
    assert od['exit_flag'] == oo['exit_flag']
    for key in ('f_tensors', 'f_couplings', 'f_constraints'):
        assert abs(od[key] - oo[key]) < FIT_TOL


def test_config1_example_script6_full_run(ab):
    """C1: example_script6 at the script's own sizes and options (runs to convergence: 102 outer iterations)."""
    Z, G, _ = pg.config_script6(seed=0)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options())
    assert oo['OuterIterations'] == 102
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


def test_config1_fixed_iterations_zero_tolerances(ab):
    Z, G, _ = pg.config_script6(seed=2)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=40, **ZERO_TOL))
    assert np.all(oo['innerIters'][[0, 1, 3, 4, 5, 6], :] == 5) and np.all(oo['innerIters'][2, :] == 1)
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


@pytest.mark.parametrize('dims', [(120, 90, 70, 200, 8), (64, 48, 40, 80, 32), (33, 21, 17, 50, 3), (40, 36, 30, 64, 64)])
def test_config2_family_cp_coupled_matrix(ab, dims):
    Z, G, _ = pg.config_cp_matrix(*dims, seed=1)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=25))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


@pytest.mark.parametrize('mode1', [('TV regularization',), ('l1 regularization',), ('unimodality', False),
                                   ('non-negative l2-sphere', 1.0), ('simplex column-wise', 1.0), ('GL smoothness',),
                                   ('l2 regularization',), ('box', -0.2, 0.2), ('non-decreasing',), ('l1-ball', 2.0)])
def test_config5_family_prox_kernels_in_the_loop(ab, mode1):
    """C5 (example_script10 style): regulariser / constraint on mode 1, l2-ball on modes 2,3."""
    Z, G, _ = pg.config_cp_tv(I=40, J=30, K=26, R=3, seed=4, mode1=mode1, eta=1e-3 if len(mode1) == 1 else 1.0)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=30, AbsFuncTol=1e-7))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


def test_rank_sweep_small(ab):
    for R in (8, 16, 40, 128):
        Z, G, _ = pg.config_cp_tv(I=48, J=40, K=36, R=R, seed=R, mode1=('l1 regularization',), eta=1e-3)
        Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=6, **ZERO_TOL))
        _assert_out_close(od, oo)
        assert_state_close(Gd, Go)


def test_unconstrained_als_and_ridge_and_bsum(ab):
    Z, G, _ = pg.config_single_cp(sz=(24, 20, 18), R=3, seed=7, noise=0.1,
                                  constraints=[None, ('non-negativity',), None])
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=15))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)
    Zr = dict(Z, ridge=[1e-3, 2e-3, 1e-3])
    Go, oo, Gd, od = _both(ab, Zr, G, pg.default_options(MaxOuterIters=15, bsum=1, bsum_weight=1e-2))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


def test_four_way_tensor(ab):
    Z, G, _ = pg.config_single_cp(sz=(12, 10, 9, 8), R=3, seed=9, noise=0.1, constraints=[('non-negativity',)] * 4)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=20))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


@pytest.mark.parametrize('dims', [(120, 90, 70, 200, 8), (64, 48, 40, 80, 32), (40, 36, 30, 64, 64), (33, 21, 17, 50, 3)])
def test_dimension_tree_matches_three_pass_and_oracle(ab, dims):
    """engine knob options.dimtree=1: mode 3 is computed from the partial contraction emitted by the mode-2 pass."""
    Z, G, _ = pg.config_cp_matrix(*dims, seed=2)
    opts = pg.default_options(MaxOuterIters=25)
    Go, oo, Gd, od = _both(ab, Z, G, dict(opts, dimtree=1))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)
    Z, G, _ = pg.config_single_cp(sz=(50, 44, 38), R=5, seed=3, noise=0.1, constraints=[('non-negativity',)] * 3)
    Go, oo, Gd, od = _both(ab, Z, G, dict(pg.default_options(MaxOuterIters=20), dimtree=1))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


PAR2_KEYS = ('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac', 'coupling_fac', 'P', 'DeltaB',
             'mu_DeltaB')


def _assert_par2_out_close(od, oo):
    _assert_out_close(od, oo)
    n = oo['OuterIterations'] + 1
    assert np.max(np.abs(od['func_PAR2_coupl'][:n] - oo['func_PAR2_coupl'][:n])) < FIT_TOL
    assert abs(od['f_PAR2_couplings'] - oo['f_PAR2_couplings']) < FIT_TOL


@pytest.mark.parametrize('kw', [dict(I=20, J=30, K=40, Jk=30, Kp=20, R=3), dict(I=33, J=18, K=12, Jk=17, Kp=9, R=4),
                                dict(I=64, J=40, K=24, Jk=64, Kp=32, R=16), dict(I=50, J=20, K=16, Jk=70, Kp=5, R=7)])
def test_config4_family_cp_coupled_with_parafac2(ab, kw):
    """C4 (example_script1 style): CP coupled in mode 1 with a regular PARAFAC2, nonneg on A, B_k, C."""
    Z, G, _ = pg.config_cp_par2(seed=3, noise=0.1, **kw)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=20))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


@pytest.mark.parametrize('kw', [dict(I=90, J=20, K=10, Jk=80, Kp=6, R=70),      # R > 64: per-slice systems in global memory
                                dict(I=110, J=12, K=8, Jk=150, Kp=4, R=100),    # ... and slices that do not fit shared memory
                                dict(I=40, J=16, K=10, Jk=128, Kp=7, R=16),     # register Jacobi, four rows per lane
                                dict(I=30, J=16, K=10, Jk=200, Kp=5, R=12),     # long slices: CTA Jacobi
                                dict(I=30, J=16, K=10, Jk=40, Kp=5, R=20)])     # 16 < R <= 64: CTA Jacobi
def test_parafac2_rank_and_slice_size_variants(ab, kw):
    """Every code path of the per-slice PARAFAC2 work: the register-resident and the CTA polar-factor kernels, and ranks
    above 64 (reference: R <= J_k is the only limit, cmtf_AOADMM.m:55-65)."""
    Z, G, _ = pg.config_cp_par2(seed=8, noise=0.1, **kw)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=8, **ZERO_TOL))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def test_config4_zero_tolerances_and_noise_free(ab):
    Z, G, _ = pg.config_cp_par2(seed=5, noise=0.0)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=30, **ZERO_TOL))
    assert np.all(oo['innerIters'][[0, 3, 4, 5], :] == 5)
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


@pytest.mark.parametrize('case', [
    dict(constrained=(1, 0, 1)),                                   # explicit residual objective (mode C updated last)
    dict(constrained=(0, 0, 0)),                                   # ALS on A and row-wise ALS on C (:181, :236)
    dict(constrained=(1, 1, 1)),                                   # nonneg B_k: element-wise prox inside the slice kernel
    dict(constrained=(1, 1, 1), constraints=[('non-negativity',), ('l1 regularization', 1e-3), ('box', 0.0, 2.0)]),
    dict(constrained=(0, 1, 1), constraints=[None, ('unimodality', True), ('non-negativity',)]),   # deferred per-slice prox
    dict(constrained=(1, 1, 1), ridge=[1e-3, 2e-3, 1e-3]),
    dict(constrained=(1, 0, 1), Jk=(7, 7, 7), R=7),               # J_k == R
])
def test_single_parafac2_object_irregular_slices(ab, case):
    """example_script4 style irregular PARAFAC2 (varying J_k), every constraint placement.
    The synthetic B_k = Q_k H have mixed signs, so a sign / sparsity constraint on B_k makes the iteration erratic
    (objective not monotone) and rounding-level differences grow ~2x per outer iteration in BOTH implementations:
    a 1-ulp perturbation of the ORACLE's own input moves its P_k by 4e-8 after 10 iterations (ridge case) and by O(1)
    after 25.  Those ill-conditioned cases are compared over 6 outer iterations, the well-posed ones over 25."""
    opts = pg.default_options(MaxOuterIters=6 if case['constrained'][1] else 25)
    Z, G, _ = pg.config_single_par2(seed=8, **case)
    Go, oo, Gd, od = _both(ab, Z, G, opts)
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def test_parafac2_options_bsum_rho_factor_and_late_constraint(ab):
    Z, G, _ = pg.config_single_par2(seed=9, constrained=(1, 1, 1))
    opts = pg.default_options(MaxOuterIters=12, bsum=1, bsum_weight=1e-2, increase_factor_rhoBk=3.0,
                              iter_start_PAR2Bkconstraint=5)
    Go, oo, Gd, od = _both(ab, Z, G, opts)
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def test_parafac2_mode_c_coupled_with_matrix(ab):
    """third PARAFAC2 mode exactly coupled (row-wise rho_k in the Delta update, cmtf_fun_AOADMM.m:638-645, :666-669)."""
    Z, G, _ = pg.config_single_par2(seed=10, Jk=(10, 12, 9, 11, 13, 10), constrained=(1, 0, 1))
    rng = np.random.RandomState(3)
    K, R, M = 6, 3, 14
    Y = (rng.rand(K, R) + 0.1) @ rng.rand(M, R).T
    Y = np.asfortranarray(Y / np.linalg.norm(Y))
    nn = ('non-negativity',)
    Z2 = dict(Z, object=Z['object'] + [Y], model=['PAR2', 'CP'], modes=[[1, 2, 3], [4, 5]], size=Z['size'] + [K, M],
              loss_function=['Frobenius'] * 2, weights=[0.5, 0.5], constrained_modes=[1, 0, 1, 1, 1],
              constraints=[nn, None, nn, nn, nn],
              coupling={'lin_coupled_modes': [0, 0, 1, 1, 0], 'coupling_type': [0], 'coupl_trafo_matrices': [None] * 5})
    init_options = {'lambdas_init': [[1.0] * R] * 2, 'nvecs': 0, 'distr': [pg.d_rand, pg.d_rand, pg.d_rand01, pg.d_rand, pg.d_rand],
                    'normalize': 1}
    G2 = pg.init_coupled_AOADMM_CMTF(Z2, init_options, rng)
    Go, oo, Gd, od = _both(ab, Z2, G2, pg.default_options(MaxOuterIters=20))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


@pytest.mark.parametrize('ctype', [1, 2, 3, 4, 5])
@pytest.mark.parametrize('constrained', [True, False])
def test_linear_couplings_all_types(ab, ctype, constrained):
    """coupling types 1..5 (HC=D, CH=D, C=HD, C=DH, H1C=DH2; cmtf_fun_AOADMM.m:278-389, :698-1075), CP + matrix."""
    Z, G, _ = pg.config_linear_coupling(ctype, seed=ctype, constrained=constrained)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=30))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


@pytest.mark.parametrize('ctype', [1, 4, 5])
def test_linear_couplings_two_tensors_zero_tolerances(ab, ctype):
    Z, G, _ = pg.config_linear_coupling(ctype, seed=10 + ctype, second='tensor')
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=15, **ZERO_TOL))
    assert np.all(oo['innerIters'][[0, 3], :] == 5)
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


@pytest.mark.parametrize('ctype', [1, 2, 3, 4, 5])
@pytest.mark.parametrize('constrained', [True, False])
def test_linear_couplings_with_first_parafac2_mode(ab, ctype, constrained):
    """A CP mode linearly coupled (types 1..5) with the first (A) mode of a PARAFAC2 object: A{m}, B{m} come from
    :159-178, the coupled precompute and ADMM loops treat the mode like a CP mode (generic branches of :278-389)."""
    Z, G, _ = pg.config_linear_coupling(ctype, seed=20 + ctype, constrained=constrained, second='par2')
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=12))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


@pytest.mark.parametrize('ctype', [1, 2, 3, 4, 5])
@pytest.mark.parametrize('constrained', [True, False])
def test_linear_couplings_with_third_parafac2_mode(ab, ctype, constrained):
    """A CP mode linearly coupled with the third (C) mode of a PARAFAC2 object: row-wise rho_k and per-slice systems
    B{m}{k} (types 2-4: :305-311, :327-333, :349-355, :783-790, :848-855, :914-921; per-row Delta systems of type 4
    :940-960; type 1 is the (K*R)^2 system of example_script14)."""
    Z, G, _ = pg.config_linear_coupling(ctype, seed=30 + ctype, constrained=constrained, second='par2c')
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=12))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def _assert_missing_close(od, oo):
    n = oo['OuterIterations'] + 1
    a, b = od['func_rel_missing'][1:n], oo['func_rel_missing'][1:n]
    assert np.isnan(od['func_rel_missing'][0]) and np.max(np.abs(a - b)) < FIT_TOL
    assert abs(od['f_rel_missing'] - oo['f_rel_missing']) < FIT_TOL


@pytest.mark.parametrize('frac', [0.2, 0.6])
def test_em_imputation_cp_and_matrix(ab, frac):
    """Z.miss on a CP tensor and on its coupled matrix: EM imputation after every sweep (cmtf_fun_AOADMM.m:408-441),
    masked objective (:1224-1226), extra stopping condition (:457-459)."""
    Z, G, _ = pg.config_cp_matrix(45, 38, 33, 70, 4, seed=6, noise=0.1)
    Zm = pg.add_missing(Z, frac, seed=2)
    Go, oo, Gd, od = _both(ab, Zm, G, pg.default_options(MaxOuterIters=30))
    _assert_out_close(od, oo)
    _assert_missing_close(od, oo)
    assert_state_close(Gd, Go)
    Zm1 = pg.add_missing(Z, frac, seed=3, objects=[0])      # only the tensor has a mask
    Go, oo, Gd, od = _both(ab, Zm1, G, pg.default_options(MaxOuterIters=15, **ZERO_TOL))
    _assert_out_close(od, oo)
    _assert_missing_close(od, oo)
    assert_state_close(Gd, Go)


@pytest.mark.parametrize('sz', [(11, 9, 7, 6), (8, 7, 5, 4, 3)])
def test_em_imputation_higher_order_tensors(ab, sz):
    """Z.miss on 4- and 5-way CP tensors: the trailing modes are merged and their Khatri-Rao product is the third
    factor of the EM pass."""
    nn = ('non-negativity',)
    Z, G, _ = pg.config_single_cp(sz=sz, R=3, seed=9, noise=0.05, constraints=[nn] + [None] * (len(sz) - 1))
    Zm = pg.add_missing(Z, 0.3, seed=4)
    Go, oo, Gd, od = _both(ab, Zm, G, pg.default_options(MaxOuterIters=12, **ZERO_TOL))
    _assert_out_close(od, oo)
    _assert_missing_close(od, oo)
    assert_state_close(Gd, Go)


def test_em_imputation_cp_coupled_with_parafac2(ab):
    """example_script12_CP_PAR2_EM.m: ~20 % missing in the CP block and in every PARAFAC2 slice (:1249-1252)."""
    Z, G, _ = pg.config_cp_par2(I=24, J=20, K=18, Jk=16, Kp=10, R=3, seed=4, noise=0.05)
    Zm = pg.add_missing(Z, 0.2, seed=5)
    Go, oo, Gd, od = _both(ab, Zm, G, pg.default_options(MaxOuterIters=25))
    _assert_par2_out_close(od, oo)
    _assert_missing_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def test_em_imputation_large_ranks_and_front_end(ab):
    Z, G, _ = pg.config_cp_matrix(70, 66, 40, 90, 40, seed=7, noise=0.1)     # R > one rank chunk of the EM kernel
    Zm = pg.add_missing(Z, 0.3, seed=1)
    Go, oo, Gd, od = _both(ab, Zm, G, pg.default_options(MaxOuterIters=8, **ZERO_TOL))
    _assert_out_close(od, oo)
    _assert_missing_close(od, oo)
    assert_state_close(Gd, Go)
    Zhat, Fac, _, out = ab.cmtf_AOADMM(Zm, init=G, alg_options=pg.default_options(MaxOuterIters=8, **ZERO_TOL))
    assert abs(out['f_tensors'] - oo['f_tensors']) < FIT_TOL       # the front end computes the masked Znorm_const


@pytest.mark.parametrize('shape,R', [((50, 41, 6), 5), ((72, 70, 12), 70), ((130, 35, 9), 6)])
def test_em_imputation_kernel_variants(ab, shape, R):
    """Both EM kernels: fewer than 8 slabs (one CTA per tile and slab range, rank chunks of 32) and the pipelined kernel
    (TMA data ring + GEMM warps + comparison warps) with a rank that is neither a multiple of 4 nor below 64, partial
    tiles in both directions."""
    Z, G, _ = pg.config_cp_matrix(shape[0], shape[1], shape[2], 60, R, seed=12, noise=0.1)
    Zm = pg.add_missing(Z, 0.35, seed=6, objects=[0])
    Go, oo, Gd, od = _both(ab, Zm, G, pg.default_options(MaxOuterIters=6, **ZERO_TOL))
    _assert_out_close(od, oo)
    _assert_missing_close(od, oo)
    assert_state_close(Gd, Go)


def test_znorm_const_computed_on_device(ab):
    """Znorm_const = NaN asks the engine for ||X_p||^2 (observed entries with Z.miss), cmtf_AOADMM.m:124-156."""
    Z, G, _ = pg.config_cp_par2(I=21, J=17, K=13, Jk=11, Kp=6, R=3, seed=6, noise=0.1)
    Zm = pg.add_missing(Z, 0.3, seed=4)
    for Zx in (Z, Zm):
        opts = pg.default_options(MaxOuterIters=6)
        zn = pg.znorm_const(Zx)
        _, o1 = ab.cmtf_fun_AOADMM(Zx, zn, G, None, None, None, None, opts)
        _, o2 = ab.cmtf_fun_AOADMM(Zx, [float('nan')] * 2, G, None, None, None, None, opts)
        assert np.max(np.abs(o1['func_val_conv'] - o2['func_val_conv'])) < 1e-12


@pytest.mark.parametrize('constrained_c', [True, False])
def test_example_script14_type1_coupling_with_parafac2_mode_c(ab, constrained_c):
    """example_script14: CP mode 1 coupled (H C = Delta) with the third PARAFAC2 mode at twice the sampling rate: the
    (K*R) x (K*R) system of cmtf_fun_AOADMM.m:283-297 / :710-722 next to a Sylvester-type CP mode."""
    Z, G, _ = pg.config_script14(seed=1, constrained_c=constrained_c)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=25))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def test_example_script1a_smooth_bk_and_l2_balls(ab):
    """example_script1a_CP_PAR2_smooth_l2ball.m: l2-ball on the coupled modes, GL smoothness on every B_k (per-slice
    deferred prox), non-negative l2-ball on the PARAFAC2 C mode."""
    Z, G, _ = pg.config_cp_par2(I=24, J=20, K=18, Jk=16, Kp=10, R=3, seed=11, noise=0.05)
    cons = [('l2-ball', 1.0), None, None, ('l2-ball', 1.0), ('GL smoothness', 1.0), ('non-negative l2-ball', 1.0)]
    Z = dict(Z, constraints=cons, constrained_modes=[1, 0, 0, 1, 1, 1])
    rng = np.random.RandomState(5)
    init_options = {'lambdas_init': [[1.0] * 3] * 2, 'nvecs': 0, 'normalize': 1,
                    'distr': [pg.d_rand, pg.d_randn, pg.d_randn, pg.d_rand, pg.d_rand, pg.d_rand01]}
    G = pg.init_coupled_AOADMM_CMTF(Z, init_options, rng)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=20))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def test_references_own_dataset_example_script11(ab):
    """The reference's own data (noisy_dataset.mat, committed as tests/golden/script11_tparafac2.npz) with the
    configuration of example_script11_tPARAFAC2.m: engine == committed oracle state after 30 outer iterations."""
    import os
    import make_script11_fixture as mk
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'script11_tparafac2.npz'))
    Z, G = mk.script11_problem(fx['dataset'])
    Gd, od = ab.cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, None, None, None, None, mk.script11_options(30))
    assert rel(Gd['fac'][0], fx['oracle_A']) < FAC_TOL and rel(Gd['fac'][2], fx['oracle_C']) < FAC_TOL
    assert rel(np.stack(Gd['fac'][1]), fx['oracle_B']) < FAC_TOL
    # unnormalised data: the objective is ~4e6, so the 1e-10 fit tolerance is relative here
    assert rel(od['func_val_conv'], fx['oracle_func_val']) < FIT_TOL
    assert np.max(np.abs(od['func_PAR2_coupl'] - fx['oracle_func_par2'])) < 1e-9


def test_degenerate_shapes(ab):
    """Edge cases the reference never exercises: rank 1, single-row modes, a single PARAFAC2 slice, matrices only."""
    nn = ('non-negativity',)
    Z, G, _ = pg.config_single_cp(sz=(1, 5, 3), R=1, seed=1, noise=0.1, constraints=[nn, None, nn])
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=8))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)
    Z, G, _ = pg.config_single_par2(seed=3, Jk=(5,), R=2, constrained=(1, 0, 1))          # K = 1
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=8))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)
    Z, G, _ = pg.config_script6(seed=4, sz=(7, 9, 1, 7, 5, 9, 6), R=2)                    # a 7 x 9 x 1 "tensor"
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=10))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


def test_full_size_config2_properties(ab):
    """BASELINE configs[1] at full size (1000^3, R=32, 1000 x 5000 matrix; tensor generated on device): properties that
    need no oracle - the dimension-tree sweep equals the three-pass sweep, a fused-inner-loop run equals the plain one,
    the objective decreases, and the run is bit-reproducible."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    from perf_probe import build
    Z, G, facs = build(1000, 1000, 1000, 5000, 32)
    runs = {}
    for name, kw in (('tree', dict(dimtree=1)), ('three', dict(dimtree=0)), ('plain', dict(dimtree=1, fuse_inner=-1, graph=-1)),
                     ('tree2', dict(dimtree=1))):
        with ab.Solver(Z, [1.0, float(np.sum(Z['object'][1] ** 2))]) as s:
            s.generate_cp_data(1, list(facs), 0.2, 99)
            s.set_state(G)
            out = s.run(pg.default_options(MaxOuterIters=6, graph=kw.pop('graph', 0), **dict(ZERO_TOL, **kw)))
            runs[name] = (s.get_state(), out)
    f = runs['tree'][1]['func_val_conv']
    assert np.all(np.isfinite(f)) and f[-1] < 0.01 * f[0]     # (the random, unnormalised init makes the start erratic)
    for other in ('three', 'plain'):
        for m in range(5):
            assert rel(runs[other][0]['fac'][m], runs['tree'][0]['fac'][m]) < 1e-10, (other, m)
        assert np.max(np.abs(runs[other][1]['func_val_conv'] - f)) < 1e-12
    for m in range(5):
        assert np.array_equal(runs['tree2'][0]['fac'][m], runs['tree'][0]['fac'][m])      # deterministic reductions


REDUCED = {1: ('TF32 on tcgen05 / TMEM', 2.0 ** -11), 2: ('BF16 on tcgen05 / TMEM', 2.0 ** -8), 3: ('TF32 on mma.sync', 2.0 ** -11)}


@pytest.mark.parametrize('prec', [1, 2, 3])
@pytest.mark.parametrize('dims', [(130, 90, 70, 200, 8), (96, 80, 72, 120, 32), (129, 257, 65, 64, 64), (40, 36, 30, 64, 100),
                                  (300, 33, 5, 40, 16)])
def test_opt_in_reduced_precision_mttkrp_modes(ab, dims, prec):
    """options.mttkrp_precision (north_star: "TF32/BF16 opt-in"): 1 = TF32 and 2 = BF16 operands on the 5th-generation
    tensor cores (tcgen05.mma, FP32 accumulation in TMEM per slab, FP64 across slabs), 3 = TF32 on the mma.sync variant of
    the FP64 kernels.  Everything else stays FP64.  Not parity modes: the kernel must match the FP64 MTTKRP to the operand
    rounding (2^-11 / 2^-8 per operand), the run must track the FP64 run and still be deterministic."""
    name, eps = REDUCED[prec]
    Z, G, _ = pg.config_cp_matrix(*dims, seed=5)
    zn = pg.znorm_const(Z)
    with ab.Solver(ab._with_rank(Z, G), zn) as s:
        s.set_state(G)
        for pos in (1, 2, 3):
            M64, Mlo = s.object_mttkrp(1, pos, 0), s.object_mttkrp(1, pos, prec)
            others = [G['fac'][q] for q in range(3) if q != pos - 1]
            ref = np.einsum('ijk,jr,kr->ir', np.moveaxis(Z['object'][0], pos - 1, 0), *others)
            assert rel(M64, ref) < 1e-13
            # error scale: products of rounded operands, summed with random signs
            bound = np.einsum('ijk,jr,kr->ir', np.abs(np.moveaxis(Z['object'][0], pos - 1, 0)), *[np.abs(u) for u in others])
            err = np.max(np.abs(Mlo - ref) / bound)
            assert 1e-8 < err < 3 * eps, (name, pos, err)
            assert np.array_equal(Mlo, s.object_mttkrp(1, pos, prec)), (name, pos)     # deterministic
    opts = pg.default_options(MaxOuterIters=8, **ZERO_TOL)
    G64, o64 = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts)
    Glo, olo = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, dict(opts, mttkrp_precision=prec))
    Glob, _ = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, dict(opts, mttkrp_precision=prec, dimtree=1))
    # over-factored cases (R close to a dimension) amplify the perturbation through the ill-conditioned Grams
    ftol = (2e-2 if dims[4] <= 32 else 0.3) * (eps / 2.0 ** -11)
    for m in range(5):
        assert 1e-9 < rel(Glo['fac'][m], G64['fac'][m]) < ftol, (name, m, rel(Glo['fac'][m], G64['fac'][m]))
        assert rel(Glob['fac'][m], G64['fac'][m]) < ftol, (name, 'dimtree', m)
    otol = (5e-2 if dims[4] <= 32 else 0.25) * (eps / 2.0 ** -11)
    assert np.max(np.abs(olo['func_val_conv'] - o64['func_val_conv']) / o64['func_val_conv']) < otol
    Gloc, _ = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, dict(opts, mttkrp_precision=prec))
    assert all(np.array_equal(Gloc['fac'][m], Glo['fac'][m]) for m in range(5))
    with pytest.raises(ab.AoadmmError):
        ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, dict(opts, mttkrp_precision=7))


def test_reduced_precision_multi_wave_tensor(ab):
    """A tensor large enough for several waves of CTAs, slab-range splits and two rank chunks (640 x 520 x 96, R=80):
    TF32 / BF16 MTTKRP of every mode against the FP64 kernels."""
    rng = np.random.RandomState(3)
    I, J, K, R = 640, 520, 96, 80
    X = np.asfortranarray(rng.rand(I, J, K))
    U = [rng.rand(s, R) for s in (I, J, K)]
    Z = {'loss_function': ['Frobenius'], 'model': ['CP'], 'modes': [[1, 2, 3]], 'size': [I, J, K],
         'coupling': {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3},
         'constrained_modes': [0, 0, 0], 'constraints': [None] * 3, 'weights': [1.0], 'object': [X], 'rank': [R]}
    with ab.Solver(Z, [1.0]) as s:
        s.set_state({'fac': U})
        for pos in (1, 2, 3):
            M64 = s.object_mttkrp(1, pos, 0)
            for prec in (1, 2):
                assert rel(s.object_mttkrp(1, pos, prec), M64) < REDUCED[prec][1], (pos, prec)


def test_warm_restart_equals_continuous_run(ab):
    """checkpoint/resume of the reference = pass Fac back as 'init' (cmtf_AOADMM.m:15,:44-45)."""
    Z, G, _ = pg.config_cp_matrix(30, 24, 20, 40, 4, seed=11)
    zn = pg.znorm_const(Z)
    o10 = pg.default_options(MaxOuterIters=10, **ZERO_TOL)
    o20 = pg.default_options(MaxOuterIters=20, **ZERO_TOL)
    Ga, _ = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, o10)
    Gb, _ = ab.cmtf_fun_AOADMM(Z, zn, Ga, None, None, None, None, o10)
    Gc, _ = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, o20)
    for m in range(5):
        assert np.array_equal(Gb['fac'][m], Gc['fac'][m])


@pytest.mark.parametrize('name', ['script6_small', 'cp_matrix_small', 'cp_tv_small'])
def test_engine_reproduces_committed_golden_vectors(ab, name):
    Z, G, opts, gold = golden_case(name)
    Gd, od = ab.cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, None, None, None, None, opts)
    assert od['OuterIterations'] == int(gold['OuterIterations'])
    assert np.max(np.abs(od['func_val_conv'] - gold['func_val_conv'])) < FIT_TOL
    for i, F in enumerate(Gd['fac']):
        assert rel(F, gold['out_fac_%d' % i]) < FAC_TOL
    assert np.array_equal(od['innerIters'], gold['innerIters'])


def test_front_end_cmtf_AOADMM(ab):
    Z, G, _ = pg.config_script6(seed=0, sz=(20, 24, 16, 20, 28, 24, 32))
    Zhat, Fac, Ginit, out = ab.cmtf_AOADMM(Z, G, pg.default_options(MaxOuterIters=30))
    assert len(Zhat) == 3 and Zhat[0][0].shape == (20, 3) and out['OuterIterations'] <= 30
    fit = 100 * (1 - np.sum((Z['object'][1] - Zhat[1][0] @ Zhat[1][1].T) ** 2) / np.sum(Z['object'][1] ** 2))
    assert fit > 90


def test_solver_reports_launches_and_times(ab):
    Z, G, _ = pg.config_cp_matrix(64, 48, 40, 80, 8, seed=1)
    Z = dict(Z, rank=[8, 8])
    with ab.Solver(Z, pg.znorm_const(Z)) as s:
        s.set_state(G)
        out = s.run(pg.default_options(MaxOuterIters=3, **ZERO_TOL))
        assert out['OuterIterations'] == 3
        assert s.launch_count() > 50 and s.last_run_ms() > 0
        assert s.time_mttkrp(1, 2, 2) > 0


# ---- nvecs initialisation on the device (cmtf_nvecs.m, init_coupled_AOADMM_CMTF.m:50-69) ----------------
def _nvecs_close(Ud, Uo, Z=None, n=None, tol=1e-8):
    """Columns agree to `tol`, loosened by the conditioning of the eigenvector (rounding level / relative gap to the
    nearest other eigenvalue) when the unfolding is available."""
    assert Ud.shape == Uo.shape
    assert np.linalg.norm(Ud.T @ Ud - np.eye(Ud.shape[1])) < 1e-12          # orthonormal columns
    tols = np.full(Ud.shape[1], tol)
    if Z is not None:
        p = [q for q, ms in enumerate(Z['modes']) if n in ms][0]
        X = np.asarray(Z['object'][p])
        i = list(Z['modes'][p]).index(n)
        A = np.moveaxis(X, i, 0).reshape(X.shape[i], -1, order='F')
        w = np.sort(np.linalg.eigvalsh(A @ A.T))[::-1]
        for c in range(Ud.shape[1]):
            gap = min(abs(w[c] - w[d]) for d in range(len(w)) if d != c) if len(w) > 1 else w[0]
            tols[c] += 2e-13 * w[0] / max(gap, 1e-300)
    for c in range(Ud.shape[1]):
        assert np.linalg.norm(Ud[:, c] - Uo[:, c]) < tols[c], (c, np.linalg.norm(Ud[:, c] - Uo[:, c]), tols[c])


@pytest.mark.parametrize('shape,R', [((30, 20, 25), 3), ((33, 65, 17), 4), ((130, 90, 70), 5), ((300, 40, 9), 6),
                                     ((8, 9, 10), 8), ((7, 6, 5, 4), 2), ((20, 12, 6, 10, 5), 3), ((64, 50), 4),
                                     ((131, 259), 7),
                                     # long modes (n > 256) in every position: the TMA-fed Gram kernel, partial tiles,
                                     # a contiguous extent shorter than one TMA box (10 < 16), both map orientations
                                     ((36, 50, 300), 4), ((10, 301, 7, 9), 3), ((45, 290, 33), 5), ((257, 31, 29), 3)])
def test_nvecs_matches_oracle_every_mode(ab, shape, R):
    """Leading eigenvectors of X_(n) X_(n)' from the resident object, every mode position of 2..5-way objects (odd
    sizes: padded leading dimension, partial tiles), against numpy eigh of the explicit unfolding."""
    rng = np.random.RandomState(3)
    N = len(shape)
    facs = [rng.rand(s, R) * (1.0 + np.arange(R))[None, :] for s in shape]
    X = np.einsum(','.join(chr(97 + d) + 'r' for d in range(N)) + '->' + ''.join(chr(97 + d) for d in range(N)), *facs)
    X = X + 0.2 * np.linalg.norm(X) / np.sqrt(X.size) * rng.randn(*shape)
    Z = {'loss_function': ['Frobenius'], 'model': ['CP'], 'modes': [list(range(1, N + 1))], 'size': list(shape),
         'coupling': {'lin_coupled_modes': [0] * N, 'coupling_type': [], 'coupl_trafo_matrices': [None] * N},
         'constrained_modes': [0] * N, 'constraints': [None] * N, 'weights': [1.0], 'object': [X]}
    with ab.Solver(dict(Z, rank=[R]), [float('nan')]) as s:
        for n in range(1, N + 1):
            Ud, info = s.nvecs(n, min(R, shape[n - 1]), return_info=True)
            _nvecs_close(Ud, pg.cmtf_nvecs(Z, n, min(R, shape[n - 1])), Z, n)
            assert info['residual'] < 1e-12 and info['iterations'] < 1000, info
        with pytest.raises(ab.AoadmmError):
            s.nvecs(1, shape[0] + 1)
        with pytest.raises(ab.AoadmmError):
            s.nvecs(N + 1, 1)
    assert rel(ab.cmtf_nvecs(Z, 2, 2), pg.cmtf_nvecs(Z, 2, 2)) < 1e-8


def test_nvecs_rank_deficient_data(ab):
    """Noise-free rank-3 tensor: Y has 3 non-zero eigenvalues, the oversampled block of the subspace iteration is rank
    deficient.  The three vectors must span the range of the true factor."""
    rng = np.random.RandomState(11)
    A, B, Cc = rng.rand(90, 3), rng.rand(40, 3), rng.rand(50, 3)
    X = np.einsum('ir,jr,kr->ijk', A, B, Cc)
    Z = {'loss_function': ['Frobenius'], 'model': ['CP'], 'modes': [[1, 2, 3]], 'size': [90, 40, 50],
         'coupling': {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3},
         'constrained_modes': [0, 0, 0], 'constraints': [None] * 3, 'weights': [1.0], 'object': [X]}
    for n, F in ((1, A), (2, B), (3, Cc)):
        U = ab.cmtf_nvecs(Z, n, 3)
        _nvecs_close(U, pg.cmtf_nvecs(Z, n, 3))
        assert np.linalg.norm(F - U @ (U.T @ F)) < 1e-10 * np.linalg.norm(F)


@pytest.mark.parametrize('nvecs', [0, 1])
def test_init_front_end_matches_oracle(ab, nvecs):
    """init_coupled_AOADMM_CMTF mirror of the product package (device nvecs + device prox) against the oracle's, same
    random stream; then the solver started from the device-made state tracks the oracle run."""
    Z, _, _ = pg.config_cp_par2(seed=4, noise=0.1)
    R = 3
    r1, r2 = np.random.RandomState(21), np.random.RandomState(21)
    distr_o = [pg.d_rand] * len(Z['size'])
    distr_d = [lambda a, b: r1.rand(a, b)] * len(Z['size'])
    io = {'lambdas_init': [[1.0] * R] * 2, 'nvecs': nvecs, 'normalize': 1}
    Gd = ab.init_coupled_AOADMM_CMTF(Z, dict(io, distr=distr_d), rng=r1)
    Go = pg.init_coupled_AOADMM_CMTF(Z, dict(io, distr=distr_o), r2)
    assert_state_close(Gd, Go, keys=('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac', 'coupling_fac',
                                     'P', 'DeltaB', 'mu_DeltaB'))
    opts = pg.default_options(MaxOuterIters=5, **ZERO_TOL)
    _, Fd, _, od = ab.cmtf_AOADMM(Z, Gd, opts)
    Fo, oo = oracle_solve(Z, pg.znorm_const(Z), Go, options=opts)
    assert_state_close(Fd, Fo, tol=1e-7)
    assert abs(od['f_tensors'] - oo['f_tensors']) <= 1e-8 * abs(oo['f_tensors'])
    with pytest.raises(ValueError):
        ab.cmtf_AOADMM(Z, 'random', opts)
    _, _, _, o2 = ab.cmtf_AOADMM(Z, 'random', opts, init_options=dict(io, distr=distr_d))
    assert np.isfinite(o2['f_tensors'])


def test_nvecs_medium_size_uses_large_tiles(ab):
    """700 x 300 x 40: 128-wide Gram tiles with a partial last tile, split reduction; and the matrix modes."""
    Z, G, _ = pg.config_cp_matrix(700, 300, 40, 500, 6, seed=8)
    with ab.Solver(ab._with_rank(Z, G), pg.znorm_const(Z)) as s:
        for n in (1, 2, 3, 4, 5):
            _nvecs_close(s.nvecs(n, 6), pg.cmtf_nvecs(Z, n, 6), Z, n)


def test_handles_release_all_device_memory(ab):
    """create -> run -> destroy many times over every kind of problem (CP + matrix, PARAFAC2, linear coupling, missing
    data, nvecs, TF32 mode, graph replay): the free device memory afterwards is what it was before."""
    import ctypes
    cudart = ctypes.CDLL('libcudart.so')

    def free_bytes():
        f, t = ctypes.c_size_t(), ctypes.c_size_t()
        assert cudart.cudaMemGetInfo(ctypes.byref(f), ctypes.byref(t)) == 0
        return f.value

    cases = [pg.config_cp_matrix(40, 36, 30, 64, 5, seed=8)[:2], pg.config_cp_par2(seed=4, noise=0.1)[:2],
             pg.config_linear_coupling(1, seed=1, second='tensor')[:2], pg.config_linear_coupling(4, seed=4, second='par2')[:2],
             pg.config_script14()[:2]]
    Zc, Gc = cases[0]
    cases.append((pg.add_missing(Zc, 0.25, seed=3), Gc))

    def one_round():
        for Z, G in cases:
            for extra in (dict(), dict(mttkrp_precision=1, graph=1)):
                ab.cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, None, None, None, None,
                                   pg.default_options(MaxOuterIters=8, **dict(ZERO_TOL, **extra)))
        ab.cmtf_nvecs(cases[0][0], 1, 3)

    one_round()                      # module loading, library workspaces, caches
    before = free_bytes()
    for _ in range(3):
        one_round()
    after = free_bytes()
    assert before - after < (8 << 20), (before, after)


def test_engine_fixed_point_satisfies_kkt_without_the_oracle(ab):
    """Oracle-independent: run the ENGINE to its fixed point on the example_script6 structure and check the KKT
    conditions of the coupled, non-negativity constrained problem directly with NumPy (gradients of coupled objects
    cancel or are complementary to the bounds, coupled factors are equal)."""
    Z, G, _ = pg.config_script6(seed=3, noise=0.1)
    opts = pg.default_options(MaxOuterIters=2000, AbsFuncTol=0.0, OuterRelTol=1e-15, MaxInnerIters=20,
                              innerRelPrTol_coupl=1e-8, innerRelPrTol_constr=1e-8, innerRelDualTol_coupl=1e-8,
                              innerRelDualTol_constr=1e-8)
    Gd, od = ab.cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, None, None, None, None, opts)
    assert od['f_couplings'] < 1e-12 and od['f_constraints'] < 1e-12
    F = Gd['fac']
    g = {}
    for p, ms in enumerate(Z['modes']):
        X, w = Z['object'][p], Z['weights'][p]
        for pos, m in enumerate(ms):
            had = np.ones((F[m - 1].shape[1],) * 2)
            for q in ms:
                if q != m:
                    had = had * (F[q - 1].T @ F[q - 1])
            M = oracle_mttkrp(X, [F[q - 1] for q in ms], pos) if X.ndim > 2 else (X @ F[ms[1] - 1] if pos == 0 else X.T @ F[ms[0] - 1])
            g[m] = 2 * w * (F[m - 1] @ had - M)
    lin = Z['coupling']['lin_coupled_modes']
    groups = {}
    for m in range(1, 8):
        groups.setdefault(('c', lin[m - 1]) if lin[m - 1] else ('m', m), []).append(m)
    for key, ms in groups.items():
        total = sum(g[m] for m in ms)
        Fm = F[ms[0] - 1]
        for m in ms:
            assert np.linalg.norm(F[m - 1] - Fm) < 1e-10
        if any(Z['constrained_modes'][m - 1] for m in ms):
            assert Fm.min() > -1e-12 and np.linalg.norm(np.minimum(Fm, total)) < 1e-9, key
        else:
            assert np.linalg.norm(total) < 1e-9, key
        if len(ms) > 1:
            assert min(np.linalg.norm(g[m]) for m in ms) > 1e-5


def test_example_script2_matrix_coupled_with_first_parafac2_mode(ab):
    """example_script2_matrix_PAR2_nonneg.m: a matrix exactly coupled with mode A of a regular PARAFAC2 object."""
    Z, G, _ = pg.config_script2(seed=2)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=25))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)


def test_example_script15_structure_three_member_type4_coupling(ab):
    """example_script15_realdata.m structure: two CP tensors (3 and 5 components) and a matrix (5 components) share
    their sample mode through one type-4 coupling C_m = Delta*H_m with three members and a 6-column Delta."""
    Z, G, _ = pg.config_script15(seed=3)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=40))
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=10, **ZERO_TOL))
    assert np.all(oo['innerIters'][[0, 3, 6], :] == 5)
    _assert_out_close(od, oo)
    assert_state_close(Gd, Go)


def test_quadratic_regularization_on_parafac2_bk(ab):
    """'quadratic regularization' (eta*trace(B_k' L B_k), constraints_to_prox.m:60-67) on the B_k mode of a regular
    PARAFAC2 object: one L (second-difference Laplacian) for every slice, per-slice rho_k in the prox."""
    J = 9
    L = 2.0 * np.eye(J) - np.eye(J, k=1) - np.eye(J, k=-1)
    L[0, 0] = L[-1, -1] = 1.0
    nn = ('non-negativity',)
    Z, G, _ = pg.config_single_par2(I=14, Jk=(J,) * 5, R=3, seed=6, noise=0.05, constrained=(1, 1, 1),
                                    constraints=[nn, ('quadratic regularization', 0.05, L), nn])
    Go, oo, Gd, od = _both(ab, Z, G, pg.default_options(MaxOuterIters=10))
    _assert_par2_out_close(od, oo)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)
