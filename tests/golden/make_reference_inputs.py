"""Writes the INPUTS of the reference-pinned golden cases as MATLAB files (tests/golden/ref_inputs/in_<name>.mat).

Nothing in this repository was produced by the MATLAB reference (neither MATLAB nor Octave exists offline),
so parity with the reference itself is "unpinned".  This script + matlab-code_b200/matlab/make_reference_golden.m close
the loop for anyone who has MATLAB:

    python tests/golden/make_reference_inputs.py                 # here: inputs (Z, G, options) -> in_<name>.mat
    matlab -batch "make_reference_golden('<reference>', '<repo>/tests/golden')"     # there: UNMODIFIED cmtf_AOADMM
    python -m pytest tests/test_reference_mat_goldens.py         # oracle (CPU) and engine (GPU) against ref_<name>.mat

Every in_<name>.mat holds the structs the reference's own call takes (functions/cmtf_AOADMM.m:1, example_script6_matrix_
matrix_CP_nonneg.m:84-138): Z (object, model, modes, size, coupling, constrained_modes, constraints, weights, loss_function
[, ridge, miss]), G (the init struct of init_coupled_AOADMM_CMTF.m) and options.  Cases cover the rows of SURVEY.md 8(a):
exact couplings (script 6), CP + matrix, TV / l2-ball prox, CP + PARAFAC2, irregular PARAFAC2, a linear coupling, EM.
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import problem_gen as pg  # noqa: E402

OUT = os.path.join(HERE, 'ref_inputs')


def _missing_case():
    Z, G, _ = pg.config_cp_matrix(18, 16, 14, 24, 3, seed=6, noise=0.1)
    return pg.add_missing(Z, 0.25, seed=2), G, None


CASES = {
    # name: (builder, kwargs, options)
    'script6_small': (pg.config_script6, dict(seed=0, sz=(20, 24, 16, 20, 28, 24, 32)), dict(MaxOuterIters=25)),
    'cp_matrix_small': (pg.config_cp_matrix, dict(I=36, J=30, K=22, M=50, R=4, seed=3), dict(MaxOuterIters=20)),
    'cp_tv_small': (pg.config_cp_tv, dict(I=30, J=24, K=20, R=3, seed=2), dict(MaxOuterIters=20, AbsFuncTol=1e-7)),
    'cp_par2_small': (pg.config_cp_par2, dict(I=20, J=18, K=16, Jk=14, Kp=8, R=3, seed=3, noise=0.1), dict(MaxOuterIters=15)),
    'par2_irregular': (pg.config_single_par2, dict(seed=8, constrained=(1, 0, 1)), dict(MaxOuterIters=15)),
    'lin_coupling_type4': (pg.config_linear_coupling, dict(ctype=4, seed=4), dict(MaxOuterIters=20)),
    'cp_matrix_missing': (_missing_case, dict(), dict(MaxOuterIters=15)),
}


def cell(items):
    """MATLAB cell column from a Python list (None -> [])."""
    c = np.empty((len(items), 1), dtype=object)
    for i, v in enumerate(items):
        c[i, 0] = to_mat(v)
    return c


def to_mat(v):
    if v is None:
        return np.zeros((0, 0))
    if isinstance(v, (list, tuple)):
        if len(v) > 0 and all(np.isscalar(x) and not isinstance(x, str) for x in v):
            return np.asarray(v, dtype=np.float64).reshape(1, -1)
        return cell(list(v))
    if isinstance(v, np.ndarray):
        return v.astype(np.float64) if v.dtype != np.bool_ else v
    return v


def constraint_cell(c):
    """('box', l, u) -> {'box', l, u}; None -> {}"""
    if c is None:
        return np.empty((0, 0), dtype=object)
    out = np.empty((1, len(c)), dtype=object)
    for i, x in enumerate(c):
        out[0, i] = x if isinstance(x, str) else (np.asarray(x, dtype=np.float64) if not isinstance(x, bool) else float(x))
    return out


def z_to_mat(Z):
    nb = len(Z['size'])
    size_c = np.empty((1, nb), dtype=object)
    for m, s in enumerate(Z['size']):
        size_c[0, m] = np.asarray(s, dtype=np.float64).reshape(1, -1) if isinstance(s, (list, tuple, np.ndarray)) else float(s)
    modes_c = np.empty((1, len(Z['modes'])), dtype=object)
    for p, ms in enumerate(Z['modes']):
        modes_c[0, p] = np.asarray(ms, dtype=np.float64).reshape(1, -1)
    cons = np.empty((nb, 1), dtype=object)
    for m in range(nb):
        cons[m, 0] = constraint_cell(Z['constraints'][m] if Z['constrained_modes'][m] else None)
    obj = np.empty((1, len(Z['object'])), dtype=object)
    for p, X in enumerate(Z['object']):
        obj[0, p] = cell(list(X)).T if isinstance(X, list) else np.asarray(X, dtype=np.float64)
    cp = Z['coupling']
    trafo = np.empty((1, nb), dtype=object)
    trafo2 = np.empty((1, nb), dtype=object)
    for m in range(nb):
        t1 = (cp.get('coupl_trafo_matrices') or [None] * nb)[m]
        t2 = (cp.get('coupl_trafo_matrices2') or [None] * nb)[m]
        trafo[0, m] = np.zeros((0, 0)) if t1 is None else np.asarray(t1, dtype=np.float64)
        trafo2[0, m] = np.zeros((0, 0)) if t2 is None else np.asarray(t2, dtype=np.float64)
    out = {'object': obj, 'model': np.array(Z['model'], dtype=object).reshape(1, -1), 'modes': modes_c, 'size': size_c,
           'loss_function': np.array(Z['loss_function'], dtype=object).reshape(1, -1),
           'coupling': {'lin_coupled_modes': np.asarray(cp['lin_coupled_modes'], dtype=np.float64).reshape(1, -1),
                        'coupling_type': np.asarray(cp.get('coupling_type', []), dtype=np.float64).reshape(1, -1),
                        'coupl_trafo_matrices': trafo, 'coupl_trafo_matrices2': trafo2},
           'constrained_modes': np.asarray(Z['constrained_modes'], dtype=np.float64).reshape(1, -1), 'constraints': cons,
           'weights': np.asarray(Z['weights'], dtype=np.float64).reshape(1, -1)}
    if Z.get('ridge') is not None:
        out['ridge'] = np.asarray(Z['ridge'], dtype=np.float64).reshape(1, -1)
    if Z.get('miss') is not None:
        miss = np.empty((1, len(Z['miss'])), dtype=object)
        for p, M in enumerate(Z['miss']):
            miss[0, p] = np.zeros((0, 0)) if M is None else (cell([np.asarray(m) != 0 for m in M]).T if isinstance(M, list)
                                                              else (np.asarray(M) != 0))
        out['miss'] = miss
    return out


def g_to_mat(G, nb_modes, P):
    out = {}
    for key in ('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac'):
        vals = list(G.get(key) or []) + [None] * nb_modes
        c = np.empty((nb_modes, 1), dtype=object)
        for m in range(nb_modes):
            v = vals[m]
            c[m, 0] = cell(v).T if isinstance(v, list) else to_mat(v)
        out[key] = c
    cf = list(G.get('coupling_fac') or [])
    out['coupling_fac'] = cell(cf) if cf else np.empty((0, 0), dtype=object)
    for key in ('P', 'DeltaB', 'mu_DeltaB'):
        vals = list(G.get(key) or []) + [None] * P
        c = np.empty((P, 1), dtype=object)
        for p in range(P):
            v = vals[p]
            c[p, 0] = cell(v).T if isinstance(v, list) else to_mat(v)
        out[key] = c
    return out


def options_to_mat(o):
    return {k: (v if isinstance(v, str) else float(v)) for k, v in o.items()}


def build_case(name):
    builder, kw, okw = CASES[name]
    Z, G, _ = builder(**kw)
    return Z, G, pg.default_options(**okw)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        Z, G, opts = build_case(name)
        blob = {'Z': z_to_mat(Z), 'G': g_to_mat(G, len(Z['size']), len(Z['object'])), 'options': options_to_mat(opts),
                'case_name': name}
        path = os.path.join(OUT, 'in_%s.mat' % name)
        sio.savemat(path, blob, do_compression=True, oned_as='row')
        print('%-22s -> %s (%d bytes)' % (name, os.path.relpath(path, ROOT), os.path.getsize(path)))


if __name__ == '__main__':
    main()
