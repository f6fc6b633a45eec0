"""GPU (-m gpu): the multi-GPU forms of the boundary.

  * ONE caller process driving N GPUs (aoadmm_create_multi - what a MATLAB session uses): parity with the
    oracle and with the one-GPU engine for CP + matrix, PARAFAC2, linear couplings, EM and nvecs;
  * one process per GPU (aoadmm_create + aoadmm_dist under torchrun): tests/dist_worker.py is launched from here, so the
    multi-rank parity is part of the graded suite (it skips on a one-GPU box).

The argument checks, the create-failure clean-up and the non-finite-residual behaviour need one GPU only."""
import ctypes
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import problem_gen as pg
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve
from _cases import FAC_TOL, FIT_TOL, ZERO_TOL, assert_state_close, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

PAR2_KEYS = ('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac', 'coupling_fac', 'P', 'DeltaB', 'mu_DeltaB')


@pytest.fixture(scope='module')
def ngpu(ab):
    n = ab.device_count()
    assert n >= 1, 'GPU tests need a CUDA device (the engine has no CPU fallback)'
    return n


def _free_bytes():
    cudart = ctypes.CDLL('libcudart.so')
    f, t = ctypes.c_size_t(), ctypes.c_size_t()
    assert cudart.cudaMemGetInfo(ctypes.byref(f), ctypes.byref(t)) == 0
    return f.value


# ------------------------------------------------------------------------------------------ one caller, N GPUs
def test_create_multi_with_one_gpu_is_the_plain_engine(ab, ngpu):
    Z, G, _ = pg.config_cp_matrix(64, 48, 40, 80, 8, seed=1)
    zn = pg.znorm_const(Z)
    opts = pg.default_options(MaxOuterIters=10)
    Ga, oa = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts)
    c = ab._capi
    with ab.Solver(ab._with_rank(Z, G), zn) as s:
        n = ctypes.c_int32(0)
        assert c.lib.aoadmm_gpu_count(s._h, ctypes.byref(n)) == 0 and n.value == 1
    # n_gpus = 1 through aoadmm_create_multi: same bits as aoadmm_create
    s = ab.Solver.__new__(ab.Solver)
    ab.Solver.__init__(s, ab._with_rank(Z, G), zn, n_gpus=1)
    s.set_state(G)
    ob = s.run(opts)
    Gb = s.get_state()
    s.close()
    assert all(np.array_equal(Ga['fac'][m], Gb['fac'][m]) for m in range(5)) and oa['f_tensors'] == ob['f_tensors']


def test_create_multi_rejects_bad_arguments(ab, ngpu):
    Z, G, _ = pg.config_cp_matrix(20, 18, 16, 30, 3, seed=1)
    zn = pg.znorm_const(Z)
    with pytest.raises(ab.AoadmmError) as e:
        ab.Solver(ab._with_rank(Z, G), zn, n_gpus=ngpu + 1)
    assert e.value.status_name == 'INVALID_ARG' and 'n_gpus' in str(e.value)
    if ngpu >= 2:
        with pytest.raises(ab.AoadmmError):
            ab.Solver(ab._with_rank(Z, G), zn, n_gpus=2, devices=[0, 0])
        Zs, Gs, _ = pg.config_single_cp(sz=(12, 9, 1), R=2, seed=3, noise=0.1)      # last mode shorter than the GPU count
        with pytest.raises(ab.AoadmmError) as e:
            ab.Solver(ab._with_rank(Zs, Gs), pg.znorm_const(Zs), n_gpus=2)
        assert 'fewer indices' in str(e.value)
    with pytest.raises(ValueError):
        ab.Solver(ab._with_rank(Z, G), zn, n_gpus=2, world_size=2, unique_id=b'0' * 128)


def test_single_process_multi_gpu_matches_oracle(ab, ngpu):
    """aoadmm_create_multi: the caller hands over whole objects once; the library cuts mode-3 slabs, one worker thread per
    GPU, NCCL all-reduces between the devices.  Factors / duals / objective against the oracle, for every N the box has."""
    if ngpu < 2:
        pytest.skip('needs at least 2 GPUs')
    worlds = sorted({2, ngpu} | ({4} if ngpu >= 4 else set()))
    cases = [('cp+matrix R=8', pg.config_cp_matrix(64, 48, 40, 80, 8, seed=1)[:2], 20, ('fac',)),
             ('cp+matrix R=32 uneven slabs', pg.config_cp_matrix(130, 90, 37, 100, 32, seed=1)[:2], 20, ('fac',)),
             ('cp+matrix R=64', pg.config_cp_matrix(40, 36, 30, 64, 64, seed=1)[:2], 20, ('fac',)),
             ('cp+par2', pg.config_cp_par2(I=24, J=20, K=18, Jk=16, Kp=10, R=3, seed=3, noise=0.1)[:2], 15, PAR2_KEYS),
             ('lin4', pg.config_linear_coupling(4, seed=4)[:2], 15, ('fac',)),
             ('lin1 two tensors', pg.config_linear_coupling(1, seed=1, second='tensor')[:2], 15, ('fac',)),
             ('4-way', pg.config_single_cp(sz=(12, 10, 9, 16), R=3, seed=9, noise=0.1, constraints=[('non-negativity',)] * 4)[:2], 15, ('fac',))]
    Zc, Gc, _ = pg.config_cp_matrix(40, 36, 30, 64, 5, seed=8)
    cases.append(('em', (pg.add_missing(Zc, 0.25, seed=3), Gc), 15, ('fac',)))
    # PARAFAC2 slices sharded over the GPUs: irregular slices, per-slice prox on B_k (segmented launch),
    # explicit-residual objective, mode A / mode C exactly coupled, EM on the slices; and the cases that stay replicated
    nn = ('non-negativity',)
    cases += [('par2 irregular, C last', pg.config_single_par2(seed=8, constrained=(1, 0, 1))[:2], 20, PAR2_KEYS),
              ('par2 ALS on A and C', pg.config_single_par2(seed=8, constrained=(0, 0, 0))[:2], 20, PAR2_KEYS),
              ('par2 nonneg B_k', pg.config_single_par2(seed=8, constrained=(1, 1, 1))[:2], 6, PAR2_KEYS),
              ('par2 unimodal B_k', pg.config_single_par2(seed=8, constrained=(0, 1, 1),
                                                          constraints=[None, ('unimodality', True), nn])[:2], 6, PAR2_KEYS),
              ('par2 8 slices ridge', pg.config_single_par2(seed=5, Jk=(10, 12, 9, 11, 13, 10, 8, 14), constrained=(1, 1, 1),
                                                            ridge=[1e-3, 2e-3, 1e-3])[:2], 6, PAR2_KEYS),
              ('script2 matrix + par2 mode A', pg.config_script2(seed=2)[:2], 20, PAR2_KEYS),
              ('script14 type 1 with par2 mode C (replicated)', pg.config_script14(seed=1)[:2], 12, PAR2_KEYS),
              ('lin4 with par2 mode A (replicated)', pg.config_linear_coupling(4, seed=24, second='par2')[:2], 12, PAR2_KEYS),
              ('tparafac2 (replicated)', pg.config_tparafac2(seed=2, eta=0.05)[:2], 12, PAR2_KEYS)]
    Zp, Gp, _ = pg.config_cp_par2(I=24, J=20, K=18, Jk=16, Kp=10, R=3, seed=4, noise=0.05)
    cases.append(('em cp+par2', (pg.add_missing(Zp, 0.2, seed=5), Gp), 15, PAR2_KEYS))
    for name, (Z, G), iters, keys in cases:
        zn = pg.znorm_const(Z)
        opts = pg.default_options(MaxOuterIters=iters)
        Go, oo = oracle_solve(Z, zn, G, options=opts)
        for n in worlds:
            Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts, n_gpus=n)
            assert od['OuterIterations'] == oo['OuterIterations'], (name, n)
            for key in ('func_val_conv', 'func_coupl_conv', 'func_constr_conv', 'func_PAR2_coupl'):
                assert np.max(np.abs(od[key] - oo[key])) < FIT_TOL, (name, n, key)
            assert np.array_equal(od['innerIters'], oo['innerIters']), (name, n)
            try:
                assert_state_close(Gd, Go, keys=tuple(keys) + ('constraint_fac', 'constraint_dual_fac', 'coupling_fac', 'coupling_dual_fac'))
            except AssertionError as e:
                raise AssertionError('%s on %d GPUs: %s' % (name, n, e))


def test_single_process_multi_gpu_medium_tensor_and_helpers(ab, ngpu):
    """A tensor large enough for split reductions on every device (512 x 384 x 64N, R=32): MTTKRP of every mode against
    the oracle, device-side generation + gather of the slabs, nvecs of every mode (incl. the sharded one: slab exchange
    between the devices), launch count and timers summed / maxed over the GPUs."""
    if ngpu < 2:
        pytest.skip('needs at least 2 GPUs')
    from oracle.tensor_ops import mttkrp as oracle_mttkrp
    n = ngpu
    I, J, K, M, R = 512, 384, 64 * n, 300, 32
    import bench
    sys.path.insert(0, ROOT)
    Z, G, facs = bench.make_problem(I, J, K, M, R, seed=2, with_tensor=False)
    zn = [1.0, float(np.sum(Z['object'][1] ** 2))]
    with ab.Solver(Z, zn, n_gpus=n) as s, ab.Solver(Z, zn) as s1:
        s.generate_cp_data(1, facs, 0.2, 31)
        s1.generate_cp_data(1, facs, 0.2, 31)
        X, X1 = np.empty((I, J, K), order='F'), np.empty((I, J, K), order='F')
        s.get_object_data(1, X)
        s1.get_object_data(1, X1)
        assert rel(X, X1) < 1e-14                       # same counter-based noise on every layout (norms reduced per device)
        s.set_state(G)
        for pos in (1, 2, 3):
            assert rel(s.object_mttkrp(1, pos), oracle_mttkrp(X, G['fac'][:3], pos - 1)) < 1e-12, pos
        Zx = dict(Z, object=[X, Z['object'][1]])
        for mode in (1, 2, 3, 4, 5):
            Ud, Uo = s.nvecs(mode, 4), pg.cmtf_nvecs(Zx, mode, 4)
            assert max(np.linalg.norm(Ud[:, c] - Uo[:, c]) for c in range(4)) < 1e-8, mode
        opts = pg.default_options(MaxOuterIters=4, **ZERO_TOL)
        l0 = s.launch_count()
        od = s.run(opts)
        Gd = s.get_state()
        assert s.launch_count() - l0 > 50 * n and s.last_loop_ms() > 0 and s.time_mttkrp(1, 1, 2) > 0
    Go, oo = oracle_solve(Zx, [float(np.sum(X * X)), zn[1]], G, options=opts)
    assert np.max(np.abs(od['func_val_conv'] - oo['func_val_conv'])) < FIT_TOL
    for m in range(5):
        assert rel(Gd['fac'][m], Go['fac'][m]) < FAC_TOL, m


def test_single_process_multi_gpu_parafac2_nvecs_and_front_end(ab, ngpu):
    """nvecs initialisation of a PARAFAC2 object whose slices are sharded: mode A = Gram of the slices side by side (partial
    Gram matrices summed over the GPUs), B_k = X_k'X_k computed by the owner of slice k; then a solve from that init."""
    if ngpu < 2:
        pytest.skip('needs at least 2 GPUs')
    Z, _, _ = pg.config_cp_par2(seed=4, noise=0.1)
    R = 3
    r1, r2 = np.random.RandomState(21), np.random.RandomState(21)
    io = {'lambdas_init': [[1.0] * R] * 2, 'nvecs': 1, 'normalize': 1}
    Gd = ab.init_coupled_AOADMM_CMTF(Z, dict(io, distr=[lambda a, b: r1.rand(a, b)] * len(Z['size'])), rng=r1, n_gpus=ngpu)
    Go = pg.init_coupled_AOADMM_CMTF(Z, dict(io, distr=[pg.d_rand] * len(Z['size'])), r2)
    assert_state_close(Gd, Go, keys=PAR2_KEYS)
    opts = pg.default_options(MaxOuterIters=5, **ZERO_TOL)
    Fd, od = ab.cmtf_fun_AOADMM(Z, pg.znorm_const(Z), Gd, None, None, None, None, opts, n_gpus=ngpu)
    Fo, oo = oracle_solve(Z, pg.znorm_const(Z), Go, options=opts)
    assert_state_close(Fd, Fo, tol=1e-7, keys=PAR2_KEYS)
    assert abs(od['f_tensors'] - oo['f_tensors']) <= 1e-8 * abs(oo['f_tensors'])
    # Znorm_const computed on device from sharded slices (NaN = ask the engine)
    _, o2 = ab.cmtf_fun_AOADMM(Z, [float('nan')] * 2, Go, None, None, None, None, opts, n_gpus=ngpu)
    assert np.max(np.abs(o2['func_val_conv'] - oo['func_val_conv'])) < 1e-10


# ------------------------------------------------------------------------------------------ one process per GPU
def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_one_process_per_gpu_matches_oracle(ab, ngpu):
    """torchrun-style launch of tests/dist_worker.py on every GPU of the box (2, and all of them when there are more)."""
    if ngpu < 2:
        pytest.skip('needs at least 2 GPUs')
    for n in sorted({2, ngpu}):
        env = dict(os.environ, OMP_NUM_THREADS='8')
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
               '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dist_worker.py')]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
        tail = (out.stdout + out.stderr)[-3000:]
        assert out.returncode == 0, tail
        assert 'DIST OK' in out.stdout and 'DIST FAILED' not in out.stdout, tail


# ------------------------------------------------------------------------------------------ one GPU is enough
def test_failed_create_releases_everything(ab, ngpu):
    """A problem that is rejected AFTER its first tensor is already in HBM (wrong L for a quadratic regularisation on a
    later mode; an unsupported constraint is caught before any allocation) must not keep device memory: the engine's
    constructor releases what it acquired before the error leaves the C ABI."""
    rng = np.random.RandomState(0)
    I, J, K, R = 256, 256, 512, 4                      # 268 MB tensor
    X = np.asfortranarray(rng.rand(I, J, K))
    nn = ('non-negativity',)
    Z = {'loss_function': ['Frobenius'], 'model': ['CP'], 'modes': [[1, 2, 3]], 'size': [I, J, K],
         'coupling': {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3},
         'constrained_modes': [1, 1, 1], 'constraints': [nn, nn, ('quadratic regularization', 0.1, np.eye(K + 1))],
         'weights': [1.0], 'object': [X], 'rank': [R]}
    ab.gram(rng.rand(10, 3))                            # context + module load before the first measurement
    before = _free_bytes()
    for _ in range(3):
        with pytest.raises(ab.AoadmmError) as e:
            ab.Solver(Z, [1.0])
        assert 'rows x rows' in str(e.value)
    assert before - _free_bytes() < (8 << 20), (before, _free_bytes())
    if ngpu >= 2:                                        # the same through aoadmm_create_multi: every device is clean
        with pytest.raises(ab.AoadmmError):
            ab.Solver(Z, [1.0], n_gpus=2)
        assert before - _free_bytes() < (8 << 20)


def test_non_finite_inner_residual_does_not_abort_the_run(ab, ngpu):
    """All-zero data with a ridge term drive every factor to exactly zero: ||fac|| = 0 and the relative residuals are 0/0
    (cmtf_fun_AOADMM.m:1086-1112).  The reference does not stop: NaN > tol is false, so that inner loop ends after one
    iteration and the outer loop runs on with f_constraints = NaN.  The engine does the same and reports the event in
    out.non_finite_mode instead of raising (AOADMM_ERR_NON_FINITE is no longer returned)."""
    import warnings
    nn = ('non-negativity',)
    Z, G, _ = pg.config_single_cp(sz=(12, 10, 8), R=3, seed=1, noise=0.1, constraints=[nn, nn, nn])
    Z = dict(Z, object=[np.zeros((12, 10, 8), order='F')], ridge=[1e-3] * 3)
    G = dict(G)
    for k in ('fac', 'constraint_fac', 'constraint_dual_fac'):
        G[k] = list(G[k])
        G[k][0] = np.zeros_like(G[k][0])
    opts = pg.default_options(MaxOuterIters=4)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        Go, oo = oracle_solve(Z, [0.0], G, options=opts)
    assert np.all(np.isnan(oo['func_constr_conv'])) and oo['OuterIterations'] == 4      # the reference's behaviour
    Gd, od = ab.cmtf_fun_AOADMM(Z, [0.0], G, None, None, None, None, opts)
    assert od['OuterIterations'] == 4 and od['exit_flag'] == oo['exit_flag'] == 'maxIterations'
    assert np.array_equal(od['innerIters'], oo['innerIters'])
    assert np.all(np.isnan(od['func_constr_conv']))
    assert np.max(np.abs(od['func_val_conv'] - oo['func_val_conv'])) < FIT_TOL
    assert od['non_finite_mode'] == 1
    for m in range(3):
        assert np.all(Gd['fac'][m] == 0.0) and np.all(Go['fac'][m] == 0.0)
    # a healthy run reports 0
    Z2, G2, _ = pg.config_cp_matrix(20, 18, 16, 30, 3, seed=1)
    _, o2 = ab.cmtf_fun_AOADMM(Z2, pg.znorm_const(Z2), G2, None, None, None, None, pg.default_options(MaxOuterIters=3))
    assert o2['non_finite_mode'] == 0
