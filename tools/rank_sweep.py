"""Developer tool: per-mode MTTKRP time over a rank sweep on an N^3 tensor (SURVEY.md 8d C5 shape), one table line per R.

  python tools/rank_sweep.py N R1 R2 ...
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab  # noqa: E402
from perf_probe import build  # noqa: E402

if __name__ == '__main__':
    N = int(sys.argv[1])
    ranks = [int(x) for x in sys.argv[2:]]
    print('# N=%d   R  mode  ms  TFLOP/s  GB/s   | outer iteration (3-pass / dimtree) ms' % N)
    for R in ranks:
        Z, G, (A, B, C) = build(N, N, N, 1024, R)
        t = time.time()
        s = ab.Solver(Z, [1.0, 1.0])
        s.generate_cp_data(1, [A, B, C], 0.2, 1234)
        s.set_state(G)
        flops = 2.0 * N * N * N * R
        bytes_ = 8.0 * N * N * N
        for pos in (1, 2, 3):
            ms = s.time_mttkrp(1, pos, 3)
            print('R=%3d mode %d: %9.3f ms  %6.2f TFLOP/s  %7.1f GB/s' % (R, pos, ms, flops / ms * 1e-9, bytes_ / ms * 1e-6), flush=True)
        for dt in (0, 1):
            opts = dict(MaxOuterIters=2, MaxInnerIters=5, AbsFuncTol=0, OuterRelTol=0, innerRelPrTol_coupl=0,
                        innerRelPrTol_constr=0, innerRelDualTol_coupl=0, innerRelDualTol_constr=0, dimtree=dt)
            s.run(opts)
            opts['MaxOuterIters'] = 4
            s.run(opts)
            print('R=%3d dimtree=%d: %9.3f ms / outer iteration' % (R, dt, s.last_run_ms() / 4), flush=True)
        s.close()
