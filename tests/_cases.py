"""Shared helpers for the test-suite."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden  # noqa: E402

ZERO_TOL = dict(AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0,
                innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0)

# north_star tolerances: factors 1e-8 relative Frobenius, fit/objective 1e-10
FAC_TOL = 1e-8
FIT_TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def golden_case(name):
    builder, kw, okw = make_golden.CASES[name]
    from oracle import problem_gen as pg
    Z, G, _ = builder(**kw)
    return Z, G, pg.default_options(**okw), np.load(os.path.join(HERE, 'golden', name + '.npz'))


def assert_state_close(Gd, Go, tol=FAC_TOL, keys=('fac', 'constraint_fac', 'constraint_dual_fac', 'coupling_dual_fac',
                                                  'coupling_fac')):
    for key in keys:
        for i, ref in enumerate(Go.get(key, [])):
            if ref is None:
                continue
            if isinstance(ref, list):
                for k, rk in enumerate(ref):
                    # per-slice duals can be pure rounding noise (K = 1: mu_DeltaB is exactly 0 in exact arithmetic)
                    err = np.linalg.norm(np.asarray(Gd[key][i][k]) - rk)
                    scale = max(np.linalg.norm(rk), 1e-300)
                    assert err < tol * max(scale, 1.0) or err / scale < tol, (key, i, k, err, scale)
            else:
                # duals can be exactly zero (inactive prox): compare absolutely then
                scale = max(np.linalg.norm(ref), 1e-300)
                err = np.linalg.norm(np.asarray(Gd[key][i]) - ref)
                assert err < tol * max(scale, 1.0) or err / scale < tol, (key, i, err, scale)
