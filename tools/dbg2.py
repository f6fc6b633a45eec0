import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab
from oracle import problem_gen as pg
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
Z, G, _ = pg.config_cp_matrix(120,90,70,200,8, seed=1)
zn = pg.znorm_const(Z)
opts = pg.default_options(MaxOuterIters=30)
Go, oo = oracle_solve(Z, zn, G, options=opts)
Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts)
for i in range(len(oo['func_val_conv'])):
    print(i, oo['func_val_conv'][i], od['func_val_conv'][i], abs(oo['func_val_conv'][i]-od['func_val_conv'][i])/abs(oo['func_val_conv'][i]))
# step-by-step restart: run oracle k iterations then compare one more iteration from identical state
state = G
for k in range(8):
    o1 = pg.default_options(MaxOuterIters=1)
    Go1, _ = oracle_solve(Z, zn, state, options=o1)
    Gd1, _ = ab.cmtf_fun_AOADMM(Z, zn, state, None, None, None, None, o1)
    print('step', k, ['%.1e' % rel(Gd1['fac'][m], Go1['fac'][m]) for m in range(5)])
    state = Go1
