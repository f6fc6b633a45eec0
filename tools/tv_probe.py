"""Developer tool: small CP problem with a long TV-regularised mode (2048 x R), to look at the prox kernels under ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab
from oracle import problem_gen as pg
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = sys.argv[2] if len(sys.argv) > 2 else 'TV regularization'
Z, G, _ = pg.config_cp_tv(I=2048, J=32, K=32, R=R, seed=5, mode1=(kind,), eta=1e-3)
zn = pg.znorm_const(Z)
opts = pg.default_options(MaxOuterIters=3, AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0,
                          innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0)
with ab.Solver(ab._with_rank(Z, G), zn) as s:
    s.set_state(G)
    out = s.run(opts)
    print('ms per iteration', s.last_loop_ms() / 3, out['f_tensors'])
