import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab
from oracle import problem_gen as pg
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
for (I,J,K,M,R) in [(120,90,70,200,8),(50,60,40,70,3),(50,60,40,70,8),(120,90,70,200,3),(120,90,70,100,8),(100,90,70,200,8)]:
    Z, G, _ = pg.config_cp_matrix(I,J,K,M,R, seed=1)
    for inner in (1,):
        opts = pg.default_options(MaxOuterIters=1, MaxInnerIters=inner)
        zn = pg.znorm_const(Z)
        Go, oo = oracle_solve(Z, zn, G, options=opts)
        Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts)
        print((I,J,K,M,R), 'inner', inner, 'f0', oo['func_val_conv'][0], od['func_val_conv'][0], 'f1', oo['func_val_conv'][1], od['func_val_conv'][1])
        for key in ('fac','constraint_fac','constraint_dual_fac','coupling_dual_fac'):
            print('   ', key, ['%.1e' % rel(Gd[key][m], Go[key][m]) if Go[key][m] is not None else '-' for m in range(5)])
        print('    delta', rel(Gd['coupling_fac'][0], Go['coupling_fac'][0]))
