"""Developer tool: run the full-size C2 problem several times and report where runs differ bitwise."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200')); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import aoadmm_b200 as ab
from perf_probe import build
from oracle import problem_gen as pg
ZERO = dict(AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0, innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
extra = {}
for a in sys.argv[3:]:
    k, v = a.split('='); extra[k] = int(v)
Z, G, facs = build(1000, 1000, 1000, 5000, 32)
ref = None
for r in range(n):
    with ab.Solver(Z, [1.0, float(np.sum(Z['object'][1] ** 2))]) as s:
        s.generate_cp_data(1, list(facs), 0.2, 99)
        s.set_state(G)
        out = s.run(pg.default_options(MaxOuterIters=iters, **dict(ZERO, dimtree=1, **extra)))
        st = s.get_state()
    if ref is None:
        ref = (st, out)
        continue
    diffs = []
    for key in ('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_fac', 'coupling_dual_fac'):
        for m, a in enumerate(st[key]):
            if a is None: continue
            b = ref[0][key][m]
            if not np.array_equal(a, b):
                diffs.append((key, m, float(np.max(np.abs(a - b))), int(np.sum(a != b))))
    fd = np.max(np.abs(out['func_val_conv'] - ref[1]['func_val_conv']))
    print('run', r, 'diffs', diffs, 'func diff', fd, flush=True)
