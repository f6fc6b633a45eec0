"""Generates the committed golden vectors (tests/golden/*.npz) from the CPU oracle.

The reference is MATLAB and cannot run offline (no MATLAB/Octave here) and it ships no expected
outputs, so these fixtures pin the ORACLE at fixed seeds: the CPU tests check that the oracle still reproduces
them, the GPU tests check the engine against them without needing the oracle at all.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import problem_gen as pg  # noqa: E402
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM  # noqa: E402

CASES = {
    # name: (builder, kwargs, options)
    'script6_small': (pg.config_script6, dict(seed=0, sz=(20, 24, 16, 20, 28, 24, 32)), dict(MaxOuterIters=25)),
    'cp_matrix_small': (pg.config_cp_matrix, dict(I=36, J=30, K=22, M=50, R=4, seed=3), dict(MaxOuterIters=20)),
    'cp_tv_small': (pg.config_cp_tv, dict(I=30, J=24, K=20, R=3, seed=2), dict(MaxOuterIters=20, AbsFuncTol=1e-7)),
}


def flatten_state(G):
    out = {}
    for key in ('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac', 'coupling_fac'):
        for i, v in enumerate(G.get(key, [])):
            if v is not None and not isinstance(v, list):
                out['%s_%d' % (key, i)] = np.asarray(v)
    return out


def main():
    for name, (builder, kw, okw) in CASES.items():
        Z, G, _ = builder(**kw)
        opts = pg.default_options(**okw)
        Gout, out = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=opts)
        blob = {'in_' + k: v for k, v in flatten_state(G).items()}
        blob.update({'out_' + k: v for k, v in flatten_state(Gout).items()})
        for p, X in enumerate(Z['object']):
            blob['object_%d' % p] = np.asarray(X)
        blob['func_val_conv'] = out['func_val_conv']
        blob['func_coupl_conv'] = out['func_coupl_conv']
        blob['func_constr_conv'] = out['func_constr_conv']
        blob['innerIters'] = out['innerIters']
        blob['OuterIterations'] = np.asarray(out['OuterIterations'])
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **blob)
        print(name, 'iters', out['OuterIterations'], 'f', out['f_tensors'])


if __name__ == '__main__':
    main()
