"""Developer tool: outer iterations/s of the parity-test configurations of BASELINE.json that are not bench.py's line
(SURVEY.md 8d C1, C4, C5), engine vs the oracle port on the host cores.  Writes one JSON line per configuration.

  python tools/bench_configs.py [c1] [c4] [c5] [--iters N]

C1: example_script6 at the script's sizes (full oracle run, both sides)
C4: CP 512x512x512 coupled in mode 1 with a regular PARAFAC2, K=512 slices of 512x64, R=16, nonneg on A, B_k, C
    (oracle timed on K=64 slices + a 512x128x128 CP, scaled by work)
C5: 2048^3 CP, R in {8,32,64}, TV (eta=1e-3) on mode 1 and l2-ball on modes 2,3; and the l1 variant
    (oracle timed on a 256^3 sample, scaled by tensor elements)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab  # noqa: E402
from oracle import problem_gen as pg  # noqa: E402
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve  # noqa: E402

ZERO = dict(AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0,
            innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0)


def engine_rate(Z, G, zn, iters, gen=None, dimtree=1):
    with ab.Solver(ab._with_rank(Z, G), zn) as s:
        if gen is not None:
            s.generate_cp_data(*gen)
        s.set_state(G)
        s.run(pg.default_options(MaxOuterIters=3, dimtree=dimtree, **ZERO))
        s.set_state(G)
        l0 = s.launch_count()
        out = s.run(pg.default_options(MaxOuterIters=iters, dimtree=dimtree, **ZERO))
        ms = s.last_run_ms()
        return iters / (ms * 1e-3), ms / iters, (s.launch_count() - l0) / iters, out


def oracle_rate(Z, G, zn, iters):
    oracle_solve(Z, zn, G, options=pg.default_options(MaxOuterIters=1, **ZERO))
    t = time.perf_counter()
    _, out = oracle_solve(Z, zn, G, options=pg.default_options(MaxOuterIters=iters, **ZERO))
    per = (out['time_at_it'][-1] - out['time_at_it'][0]) / iters
    return 1.0 / per, time.perf_counter() - t


def c1(iters):
    Z, G, _ = pg.config_script6(seed=0)
    zn = pg.znorm_const(Z)
    ev, ms, ln, _ = engine_rate(Z, G, zn, iters)
    ov, _ = oracle_rate(Z, G, zn, iters)
    return {'config': 'C1 example_script6 sizes (50x60x40 + 50x70 + 60x80, R=3)', 'engine_it_s': ev, 'engine_ms_per_it': ms,
            'launches_per_it': ln, 'oracle_it_s': ov, 'oracle_sample': 'full problem', 'iters': iters}


def c4(iters):
    R = 16
    Z, G, _ = pg.config_cp_par2(I=512, J=512, K=512, Jk=64, Kp=512, R=R, seed=1, noise=0.1)
    zn = pg.znorm_const(Z)
    ev, ms, ln, out = engine_rate(Z, G, zn, iters)
    Zs, Gs, _ = pg.config_cp_par2(I=512, J=128, K=128, Jk=64, Kp=64, R=R, seed=1, noise=0.1)
    ov, _ = oracle_rate(Zs, Gs, pg.znorm_const(Zs), 3)
    # work per iteration ~ CP tensor elements * 3 passes + PARAFAC2 elements * 3 products
    scale = (512.0 ** 3 * 3 + 512 * 64 * 512 * 3) / (512.0 * 128 * 128 * 3 + 512 * 64 * 64 * 3)
    return {'config': 'C4 CP 512^3 coupled (mode 1) with PARAFAC2 K=512 slices of 512x64, R=16, nonneg A,B_k,C',
            'engine_it_s': ev, 'engine_ms_per_it': ms, 'launches_per_it': ln, 'oracle_it_s': ov / scale,
            'oracle_sample': 'CP 512x128x128 + K=64 slices, it/s scaled by 1/%.1f ~ work' % scale, 'iters': iters,
            'f_tensors': out['f_tensors'], 'f_PAR2': out['f_PAR2_couplings']}


def c5(iters, R, mode1):
    N = 2048
    rng = np.random.RandomState(5)
    A, B, C = rng.randn(N, R), rng.randn(N, R), rng.randn(N, R)
    for r in range(R):   # piecewise constant mode 1 (create_CP_data_example10piecewiseconstant.m:77-92)
        jumps = np.concatenate(([0], np.sort(rng.randint(1, N, size=4)), [N]))
        for i in range(5):
            A[jumps[i]:jumps[i + 1], r] = -1 + 2 * rng.rand()
    A /= np.linalg.norm(A, axis=0)
    B /= np.linalg.norm(B, axis=0)
    C /= np.linalg.norm(C, axis=0)
    cons = [mode1, ('l2-ball', 1.0), ('l2-ball', 1.0)]
    Z = {'loss_function': ['Frobenius'], 'model': ['CP'], 'modes': [[1, 2, 3]], 'size': [N, N, N],
         'coupling': {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3},
         'constrained_modes': [1, 1, 1], 'constraints': cons, 'weights': [1.0], 'object': [None], 'rank': [R]}
    init_options = {'lambdas_init': [[1.0] * R], 'nvecs': 0, 'distr': [pg.d_randn] * 3, 'normalize': 1}
    G = pg.init_coupled_AOADMM_CMTF(Z, init_options, rng)
    ev, ms, ln, out = engine_rate(Z, G, [1.0], iters, gen=(1, [A, B, C], 0.8, 77))
    Zs, Gs, _ = pg.config_cp_tv(I=256, J=256, K=256, R=R, seed=5, mode1=mode1[:1], eta=mode1[1])
    ov, _ = oracle_rate(Zs, Gs, pg.znorm_const(Zs), 3)
    scale = (N / 256.0) ** 3
    return {'config': 'C5 2048^3 CP R=%d, %s(%g) on mode 1, l2-ball on modes 2,3, noise 0.8' % (R, mode1[0], mode1[1]),
            'engine_it_s': ev, 'engine_ms_per_it': ms, 'launches_per_it': ln, 'oracle_it_s': ov / scale,
            'oracle_sample': '256^3 sample, it/s scaled by 1/%g ~ tensor elements' % scale, 'iters': iters,
            'f_tensors': out['f_tensors']}


if __name__ == '__main__':
    which = [a for a in sys.argv[1:] if not a.startswith('--')] or ['c1', 'c4', 'c5']
    iters = int(sys.argv[sys.argv.index('--iters') + 1]) if '--iters' in sys.argv else 10
    if 'c1' in which:
        print(json.dumps(c1(100)), flush=True)
    if 'c4' in which:
        print(json.dumps(c4(iters)), flush=True)
    if 'c5' in which:
        for R in (8, 64):
            for m1 in (('TV regularization', 1e-3), ('l1 regularization', 1e-3)):
                print(json.dumps(c5(iters, R, m1)), flush=True)
