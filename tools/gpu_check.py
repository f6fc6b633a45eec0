"""Quick on-GPU sanity run (developer tool): operator-level + solver-level comparison with the oracle."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab
from oracle import problem_gen as pg, prox as oprox
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve
from oracle.tensor_ops import mttkrp as omttkrp

def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)

print('devices', ab.device_count())
rng = np.random.RandomState(0)
ok = True
for shape, R in [((50, 60, 40), 3), ((130, 70, 33), 8), ((64, 64, 64), 16), ((200, 150, 90), 32), ((129, 257, 65), 64),
                 ((96, 80, 72), 100), ((51, 60, 40), 5), ((40, 30), 4), ((300, 500), 32), ((20, 12, 10, 8), 6)]:
    X = np.asfortranarray(rng.randn(*shape))
    U = [rng.randn(s, R) for s in shape]
    for n in range(1, len(shape) + 1):
        try:
            M = ab.mttkrp(X, U, n)
            e = rel(M, omttkrp(X, U, n - 1))
        except Exception as ex:
            e = float('nan'); print('EXC', ex)
        flag = 'OK ' if e < 1e-12 else 'BAD'
        ok &= e < 1e-12
        print(f'{flag} mttkrp shape={shape} R={R} mode={n} rel={e:.2e}')
X = rng.randn(500, 7)
G = ab.gram(X); print('gram', rel(G, X.T @ X))
B = X.T @ X + np.eye(7); A = rng.randn(33, 7)
print('chol_solve', rel(ab.chol_solve(B, A), A @ np.linalg.inv(B)))
V = rng.randn(200, 6)
cons = [('non-negativity',), ('box', -0.3, 0.5), ('simplex column-wise', 1.0), ('simplex row-wise', 2.0), ('non-decreasing',),
        ('non-increasing',), ('unimodality', True), ('unimodality', False), ('l1-ball', 3.0), ('l2-ball', 1.0),
        ('non-negative l2-ball', 1.0), ('non-negative l2-sphere', 1.0), ('l1 regularization', 0.2), ('l0 regularization', 0.2),
        ('l2 regularization', 2.0), ('ridge', 0.3), ('GL smoothness', 0.5), ('TV regularization', 0.7)]
ops, _ = oprox.constraints_to_prox([1] * len(cons), cons, [200] * len(cons))
for c, op in zip(cons, ops):
    try:
        e = rel(ab.prox(c, V, rho=1.7), op(V, 1.7))
    except Exception as ex:
        e = float('nan'); print('EXC', ex)
    print(('OK ' if e < 1e-12 else 'BAD'), 'prox', c, f'{e:.2e}')
    ok &= e < 1e-12

def compare(name, Z, G, opts):
    zn = pg.znorm_const(Z)
    t = time.time(); Go, oo = oracle_solve(Z, zn, G, options=opts); to = time.time() - t
    t = time.time(); Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts); td = time.time() - t
    errs = [rel(Gd['fac'][m], Go['fac'][m]) for m in range(len(Go['fac']))]
    print(f'{name}: iters oracle={oo["OuterIterations"]} dev={od["OuterIterations"]} t_oracle={to:.2f}s t_dev={td:.2f}s')
    print('   fac rel err', ['%.1e' % e for e in errs])
    print('   f_tensors', oo['f_tensors'], od['f_tensors'], 'f_coupl', oo['f_couplings'], od['f_couplings'], 'f_constr', oo['f_constraints'], od['f_constraints'])
    n = min(len(oo['func_val_conv']), len(od['func_val_conv']))
    print('   max |func_val diff|', np.max(np.abs(oo['func_val_conv'][:n] - od['func_val_conv'][:n])), 'inner equal', np.array_equal(oo['innerIters'][:, :n-1], od['innerIters'][:, :n-1]))
    return max(errs)

Z, G, _ = pg.config_script6(seed=0)
e = compare('script6 fixed 50 iters tol=0', Z, G, pg.default_options(MaxOuterIters=50, AbsFuncTol=0, OuterRelTol=0, innerRelPrTol_coupl=0, innerRelPrTol_constr=0, innerRelDualTol_coupl=0, innerRelDualTol_constr=0)); ok &= e < 1e-8
e = compare('script6 default options', Z, G, pg.default_options()); ok &= e < 1e-8
Z, G, _ = pg.config_cp_matrix(120, 90, 70, 200, 8, seed=1)
e = compare('cp+matrix 120x90x70 R=8', Z, G, pg.default_options(MaxOuterIters=30)); ok &= e < 1e-8
Z, G, _ = pg.config_cp_tv(seed=2)
e = compare('TV/l2ball', Z, G, pg.default_options(MaxOuterIters=40, AbsFuncTol=1e-7)); ok &= e < 1e-8
print('ALL OK' if ok else 'SOME FAILED')
