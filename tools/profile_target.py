"""Developer tool: a short program that launches every MTTKRP kernel variant once or twice, meant to run under ncu.

  python tools/profile_target.py I J K M R
    - 1 outer iteration with dimtree=1 (lead, inner EPI0+EMIT, from_T kernels + the ADMM / Gram / reduction kernels)
    - then one launch of each stand-alone mode kernel (lead, inner EPI0, inner EPI1)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab  # noqa: E402
from perf_probe import build  # noqa: E402

if __name__ == '__main__':
    I, J, K, M, R = [int(x) for x in sys.argv[1:6]]
    Z, G, (A, B, C) = build(I, J, K, M, R)
    s = ab.Solver(Z, [1.0, 1.0])
    s.generate_cp_data(1, [A, B, C], 0.2, 1234)
    s.set_state(G)
    opts = dict(MaxOuterIters=1, MaxInnerIters=5, AbsFuncTol=0, OuterRelTol=0, innerRelPrTol_coupl=0, innerRelPrTol_constr=0,
                innerRelDualTol_coupl=0, innerRelDualTol_constr=0, dimtree=1,
                mttkrp_precision=int(os.environ.get('PROBE_PREC', '0')))
    s.run(opts)
    for pos in (1, 2, 3):
        ms = s.time_mttkrp(1, pos, 0)
        print('mode %d: %.3f ms' % (pos, ms))
    s.close()
