"""Developer tool: the C4 configuration (CP 512^3 coupled with PARAFAC2 K=512 x (512x64), R=16), 2 outer iterations, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab
from oracle import problem_gen as pg
Z, G, _ = pg.config_cp_par2(I=512, J=512, K=512, Jk=64, Kp=512, R=16, seed=1, noise=0.1)
zn = pg.znorm_const(Z)
opts = pg.default_options(MaxOuterIters=2, AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0,
                          innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0, dimtree=1)
with ab.Solver(ab._with_rank(Z, G), zn) as s:
    s.set_state(G)
    out = s.run(opts)
    print('ms per iteration', s.last_loop_ms() / 2, out['f_tensors'])
