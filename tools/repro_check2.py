"""Developer tool: isolate run-to-run differences (data generation, single MTTKRPs, one iteration)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200')); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import aoadmm_b200 as ab
from perf_probe import build
from oracle import problem_gen as pg
ZERO = dict(AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0, innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0)
I = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Z, G, facs = build(I, I, I, 5000, 32)
ref = None
for r in range(5):
    res = {}
    with ab.Solver(Z, [1.0, float(np.sum(Z['object'][1] ** 2))]) as s:
        s.generate_cp_data(1, list(facs), 0.2, 99)
        s.set_state(G)
        for pos in (1, 2, 3):
            res['mttkrp%d' % pos] = s.object_mttkrp(1, pos, 0)
            res['mttkrp%d_again' % pos] = s.object_mttkrp(1, pos, 0)
        for it, kw in ((1, dict(dimtree=0, fuse_inner=-1, graph=-1)), (1, dict(dimtree=1, fuse_inner=-1, graph=-1)), (1, dict(dimtree=1, graph=-1)), (3, dict(dimtree=1, graph=-1))):
            s.set_state(G)
            out = s.run(pg.default_options(MaxOuterIters=it, **dict(ZERO, **kw)))
            st = s.get_state()
            for m in range(5):
                res['it%d_%s_fac%d' % (it, sorted(kw.items()), m)] = st['fac'][m]
    if ref is None:
        ref = res
        for pos in (1, 2, 3):
            print('same-handle repeat mttkrp', pos, np.array_equal(res['mttkrp%d' % pos], res['mttkrp%d_again' % pos]))
        continue
    bad = [(k, float(np.max(np.abs(res[k] - ref[k])))) for k in res if not np.array_equal(res[k], ref[k])]
    print('run', r, 'differs in', bad[:12], flush=True)
