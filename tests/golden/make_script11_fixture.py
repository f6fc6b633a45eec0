"""Builds tests/golden/script11_tparafac2.npz from the reference's own data fixtures (run in the build container, where
/root/reference exists; the .npz travels to the GPU box, the reference does not).

  functions_for_example_scripts/noisy_dataset.mat  'dataset' 100 x 80 x 25   (example_script11_tPARAFAC2.m:32)
  functions_for_example_scripts/gnd_factors.mat    'A' 100x3, 'B' 25x80x3, 'C' 25x3   (:20-22)

plus the state of the ORACLE after 30 outer iterations of the script's own configuration (:51-139: PARAFAC2, tPARAFAC2
penalty 1000 on B_k, non-negative C, ridge [100 0 100], unnormalised data and init) from a seeded init.
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import problem_gen as pg  # noqa: E402
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM  # noqa: E402

REF = '/root/reference/functions_for_example_scripts'


def script11_problem(dataset, seed=0):
    K = dataset.shape[2]
    X = [np.asfortranarray(dataset[:, :, k]) for k in range(K)]
    Z = {'loss_function': ['Frobenius'], 'model': ['PAR2'], 'modes': [[1, 2, 3]],
         'size': [dataset.shape[0], [dataset.shape[1]] * K, K],
         'coupling': {'lin_coupled_modes': [0, 0, 0], 'coupling_type': [], 'coupl_trafo_matrices': [None] * 3},
         'constrained_modes': [0, 1, 1], 'constraints': [None, ('tPARAFAC2', 1000), ('non-negativity',)],
         'weights': [1.0], 'object': [X], 'ridge': [100, 0, 100]}
    rng = np.random.RandomState(seed)
    G = pg.init_coupled_AOADMM_CMTF(Z, {'lambdas_init': [[1, 1, 1]], 'nvecs': 0, 'distr': [pg.d_rand] * 3, 'normalize': 0}, rng)
    return Z, G


def script11_options(iters):
    return pg.default_options(MaxOuterIters=iters, AbsFuncTol=1e-14, OuterRelTol=1e-8, innerRelPrTol_coupl=1e-4,
                              innerRelPrTol_constr=1e-4, innerRelDualTol_coupl=1e-4, innerRelDualTol_constr=1e-4)


if __name__ == '__main__':
    d = sio.loadmat(os.path.join(REF, 'noisy_dataset.mat'))['dataset']
    g = sio.loadmat(os.path.join(REF, 'gnd_factors.mat'))
    Z, G = script11_problem(d)
    Go, oo = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=script11_options(30))
    np.savez_compressed(os.path.join(HERE, 'script11_tparafac2.npz'), dataset=d, A=g['A'], B=g['B'], C=g['C'],
                        oracle_A=Go['fac'][0], oracle_B=np.stack(Go['fac'][1]), oracle_C=Go['fac'][2],
                        oracle_func_val=oo['func_val_conv'], oracle_func_par2=oo['func_PAR2_coupl'])
    print('written', os.path.getsize(os.path.join(HERE, 'script11_tparafac2.npz')), 'bytes')
