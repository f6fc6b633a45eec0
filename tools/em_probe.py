"""Developer tool: cost of the EM imputation pass - outer iteration time of a 512^3 CP problem with and without Z.miss."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab
from oracle import problem_gen as pg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rng = np.random.RandomState(0)
A, B, C = rng.rand(N, R), rng.rand(N, R), rng.rand(N, R)
X = np.einsum('ir,jr,kr->ijk', A, B, C, optimize=True)
X += 0.05 * np.linalg.norm(X) / np.sqrt(X.size) * rng.standard_normal(X.shape)
X = np.asfortranarray(X / np.linalg.norm(X))
nn = ('non-negativity',)
Z = {'loss_function': ['Frobenius'], 'model': ['CP'], 'modes': [[1, 2, 3]], 'size': [N, N, N],
     'coupling': {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3},
     'constrained_modes': [1, 1, 1], 'constraints': [nn] * 3, 'weights': [1.0], 'object': [X]}
G = pg.init_coupled_AOADMM_CMTF(Z, {'lambdas_init': [[1.0] * R], 'nvecs': 0, 'distr': [pg.d_rand] * 3, 'normalize': 1}, rng)
mask = np.asfortranarray(rng.rand(N, N, N) > 0.2)
opts = pg.default_options(MaxOuterIters=10, AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0,
                          innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0, dimtree=1)
for name, Zx in (('complete', Z), ('20% missing', dict(Z, object=[np.asfortranarray(np.where(mask, X, 0.0))], miss=[mask]))):
    with ab.Solver(ab._with_rank(Zx, G), [float('nan')]) as s:
        s.set_state(G)
        s.run(dict(opts, MaxOuterIters=3))
        s.set_state(G)
        out = s.run(opts)
        print('%-12s %.3f ms per outer iteration, f=%.6g f_rel_missing=%s' % (name, s.last_loop_ms() / 10, out['f_tensors'], out['f_rel_missing']))
