"""GPU (-m gpu): parity of the CUDA engine with the CPU oracle on BASELINE.json's configurations AT THEIR OWN SIZES
(the small-size parity tests live in test_gpu_parity.py).  At these sizes other code runs than at toy sizes: split
reductions (nsplit > 1), multi-wave grids, 64-column rank chunks with two warp groups, GB-sized partial-contraction
buffers for the dimension tree.

  C2  1000^3 R=32 + 1000 x 5000 matrix           full oracle solve, 3 outer iterations, dimtree 0 and 1
  C3  a per-GPU slab 4096 x 4096 x 64, R=64      MTTKRP of every mode against float64 einsum on random output rows,
                                                 dimension-tree sweep == three-pass sweep
  C4  CP 512^3 + PARAFAC2 K=512 x (512 x 64)     full oracle solve, 2 outer iterations
  C5  2048-row prox kernels at R=8 and R=256, and a 512^3 TV-regularised solve in the loop

Tolerances (north_star): factor matrices 1e-8 relative Frobenius error, objective 1e-10; operators 1e-12."""
import os
import sys

import numpy as np
import pytest

from oracle import problem_gen as pg
from oracle import prox as P
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve
from _cases import FAC_TOL, FIT_TOL, ZERO_TOL, assert_state_close, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _need_gpu(ab):
    assert ab.device_count() >= 1, 'GPU tests need a CUDA device (the engine has no CPU fallback)'


def _host_gb():
    for line in open('/proc/meminfo'):
        if line.startswith('MemAvailable:'):
            return int(line.split()[1]) / 1e6
    return 0.0


def _device_tensor_problem(ab, I, J, K, M, R, seed, data_seed):
    """The C2/C3 construction of bench.py: the tensor is generated on the device (create_coupled_data.m:158-162 there),
    then copied to the host so that the oracle sees exactly the same data."""
    import bench
    Z, G, facs = bench.make_problem(I, J, K, M, R, seed=seed, with_tensor=False)
    zn = [1.0, float(np.sum(Z['object'][1] ** 2))]
    s = ab.Solver(Z, zn)
    s.generate_cp_data(1, facs, 0.2, data_seed)
    X = np.empty((I, J, K), order='F')
    s.get_object_data(1, X)
    return Z, G, zn, s, X


def test_config2_full_size_three_iterations_match_oracle(ab):
    """BASELINE configs[1] (1000^3, R=32, nonneg, coupled with a 1000 x 5000 matrix) at full size: factors, duals and the
    objective history of the engine (three-pass and dimension-tree sweeps) against the oracle on the SAME tensor."""
    if _host_gb() < 40:
        pytest.skip('needs ~30 GB of host memory for the oracle')
    Z, G, zn, s, X = _device_tensor_problem(ab, 1000, 1000, 1000, 5000, 32, seed=0, data_seed=99)
    opts = pg.default_options(MaxOuterIters=3)
    runs = {}
    with s:
        for dt in (0, 1):
            s.set_state(G)
            out = s.run(dict(opts, dimtree=dt))
            runs[dt] = (s.get_state(), out)
    assert abs(float(np.sum(X * X)) - 1.0) < 1e-12           # normalised on device (example_script6...m:101-102)
    Zo = dict(Z, object=[X, Z['object'][1]])
    Go, oo = oracle_solve(Zo, zn, G, options=opts)
    for dt in (0, 1):
        Gd, od = runs[dt]
        assert od['OuterIterations'] == oo['OuterIterations']
        for key in ('func_val_conv', 'func_coupl_conv', 'func_constr_conv'):
            assert np.max(np.abs(od[key] - oo[key])) < FIT_TOL, (dt, key, od[key], oo[key])
        assert np.array_equal(od['innerIters'], oo['innerIters']), dt
        assert_state_close(Gd, Go)
        for m in range(5):
            assert rel(Gd['fac'][m], Go['fac'][m]) < FAC_TOL, (dt, m)


def test_config3_slab_mttkrp_rows_and_dimension_tree(ab):
    """One GPU's slab of BASELINE configs[2] at N=32 GPUs' worth of K (4096 x 4096 x 64, R=64 - the same 4096 x 4096
    tile grid, 64-column chunk and split reduction as the full tensor): MTTKRP of every mode checked on 64 random output
    rows against float64 einsum of the host copy; the dimension-tree sweep equals the three-pass sweep."""
    if _host_gb() < 30:
        pytest.skip('needs ~20 GB of host memory')
    I, J, K, M, R = 4096, 4096, 64, 8192, 64
    Z, G, zn, s, X = _device_tensor_problem(ab, I, J, K, M, R, seed=1, data_seed=7)
    rng = np.random.RandomState(5)
    with s:
        s.set_state(G)
        A, B, C = G['fac'][0], G['fac'][1], G['fac'][2]
        for pos, n in ((1, I), (2, J), (3, K)):
            Md = s.object_mttkrp(1, pos)
            rows = np.unique(rng.randint(0, n, size=64))
            for r in rows:
                if pos == 1:
                    ref = np.einsum('jk,jr,kr->r', X[r, :, :], B, C)
                elif pos == 2:
                    ref = np.einsum('ik,ir,kr->r', X[:, r, :], A, C)
                else:
                    ref = np.einsum('ij,ir,jr->r', X[:, :, r], A, B)
                assert rel(Md[r], ref) < 1e-12, (pos, r, rel(Md[r], ref))
        # matrix-block products of the coupled 4096 x 8192 matrix (cmtf_fun_AOADMM.m:108, :111)
        Y = Z['object'][1]
        assert rel(s.object_mttkrp(2, 1), Y @ G['fac'][4]) < 1e-12
        assert rel(s.object_mttkrp(2, 2), Y.T @ G['fac'][3]) < 1e-12
        runs = {}
        for dt in (0, 1):
            s.set_state(G)
            out = s.run(pg.default_options(MaxOuterIters=2, dimtree=dt, **ZERO_TOL))
            runs[dt] = (s.get_state(), out)
    for m in range(5):
        assert rel(runs[1][0]['fac'][m], runs[0][0]['fac'][m]) < 1e-10, m
    f3, ft = runs[0][1]['func_val_conv'], runs[1][1]['func_val_conv']     # (random, unnormalised start: f ~ 1e4)
    assert np.max(np.abs(ft - f3) / np.maximum(1.0, np.abs(f3))) < 1e-12
    # first outer iteration against the oracle's scalar bookkeeping: objective at iteration 0 from the host copy
    # (cp_func.m:47-56): f = w * (||X||^2 - 2 <X, [[A,B,C]]> + ||[[A,B,C]]||^2) + the same for the matrix
    from oracle.tensor_ops import mttkrp as oracle_mttkrp
    inner = float(np.sum(oracle_mttkrp(X, [A, B, C], 0) * A))
    had = float(np.sum((A.T @ A) * (B.T @ B) * (C.T @ C)))
    f0_x = 0.5 * (float(np.sum(X * X)) - 2.0 * inner + had)
    Ay, V = G['fac'][3], G['fac'][4]
    f0_y = 0.5 * float(np.sum((Y - Ay @ V.T) ** 2))
    assert abs(runs[0][1]['func_val_conv'][0] - (f0_x + f0_y)) < 1e-9 * (f0_x + f0_y)


def test_config4_full_size_cp_coupled_with_parafac2(ab):
    """BASELINE configs[3]: CP 512^3 coupled in mode 1 with a regular PARAFAC2, K=512 slices of 512 x 64, R=16, nonneg on
    A, B_k, C - two outer iterations against the oracle (512 per-slice systems, Jacobi polar factors, stacked products)."""
    if _host_gb() < 16:
        pytest.skip('needs ~10 GB of host memory')
    Z, G, _ = pg.config_cp_par2(I=512, J=512, K=512, Jk=64, Kp=512, R=16, seed=1, noise=0.1)
    zn = pg.znorm_const(Z)
    opts = pg.default_options(MaxOuterIters=2)
    Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts)
    Go, oo = oracle_solve(Z, zn, G, options=opts)
    assert od['OuterIterations'] == oo['OuterIterations']
    for key in ('func_val_conv', 'func_coupl_conv', 'func_constr_conv', 'func_PAR2_coupl'):
        assert np.max(np.abs(od[key] - oo[key])) < FIT_TOL, (key, od[key], oo[key])
    assert np.array_equal(od['innerIters'], oo['innerIters'])
    assert_state_close(Gd, Go, keys=('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac', 'coupling_fac',
                                     'P', 'DeltaB', 'mu_DeltaB'))


C5_PROX = [('TV regularization', 1e-3), ('TV regularization', 0.7), ('l1 regularization', 1e-3), ('l1 regularization', 0.2),
           ('unimodality', True), ('unimodality', False), ('l2-ball', 1.0), ('non-decreasing',), ('simplex column-wise', 1.0),
           ('l1-ball', 3.0), ('GL smoothness', 0.5), ('l2 regularization', 2.0), ('non-negative l2-sphere', 1.0)]


@pytest.mark.parametrize('R', [8, 256])
@pytest.mark.parametrize('con', C5_PROX, ids=[c[0] + str(c[1:]) for c in C5_PROX])
def test_config5_prox_kernels_at_2048_rows(ab, con, R):
    """BASELINE configs[4] (2048^3 rank sweep R=8..256 with TV / l1): the prox kernels on 2048-row factor matrices at the
    ends of the rank sweep - piecewise-constant + noise columns (example_script10 style), ties, an all-negative column."""
    rows = 2048
    rng = np.random.RandomState(R + len(con[0]))
    V = np.repeat(rng.randn(rows // 64, R), 64, axis=0) + 0.05 * rng.randn(rows, R)
    V[:, 0] = -np.abs(V[:, 0])
    V[rows // 2:, 1] = V[rows // 2, 1]
    V[:, 2] = np.cumsum(rng.randn(rows)) * 0.05
    ops, _ = P.constraints_to_prox([1], [con], [rows])
    for rho in (1.0, 0.05):
        got, ref = ab.prox(con, V, rho=rho), ops[0](V, rho)
        assert np.max(np.abs(got - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref))), (con, R, rho)


@pytest.mark.parametrize('mode1,eta', [(('TV regularization',), 1e-3), (('l1 regularization',), 1e-3)])
def test_config5_in_the_loop_512_cubed(ab, mode1, eta):
    """example_script10 structure at 512^3 (the size BASELINE.md section 3 gives the oracle for C5), R=8: regulariser on
    mode 1, l2-ball on modes 2 and 3, three outer iterations against the oracle."""
    Z, G, _ = pg.config_cp_tv(I=512, J=512, K=512, R=8, seed=5, mode1=mode1, eta=eta)
    zn = pg.znorm_const(Z)
    opts = pg.default_options(MaxOuterIters=3)
    Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts)
    Go, oo = oracle_solve(Z, zn, G, options=opts)
    assert od['OuterIterations'] == oo['OuterIterations']
    for key in ('func_val_conv', 'func_constr_conv'):
        assert np.max(np.abs(od[key] - oo[key])) < FIT_TOL, (key, od[key], oo[key])
    assert np.array_equal(od['innerIters'], oo['innerIters'])
    assert_state_close(Gd, Go)
