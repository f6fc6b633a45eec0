"""Developer tool: MTTKRP per-mode device time + outer-iteration time for a CP(+matrix) workload generated on device."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import aoadmm_b200 as ab

def build(I, J, K, M, R, seed=0):
    rng = np.random.RandomState(seed)
    sz = [I, J, K, I, M]
    A, B, C, V = rng.rand(I, R), rng.rand(J, R), rng.rand(K, R), rng.rand(M, R)
    Y = A @ V.T
    N = rng.randn(I, M)
    Y = Y + 0.2 * np.linalg.norm(Y) / np.linalg.norm(N) * N
    Y /= np.linalg.norm(Y)
    nn = ('non-negativity',)
    Z = {'loss_function': ['Frobenius'] * 2, 'model': ['CP', 'CP'], 'modes': [[1, 2, 3], [4, 5]], 'size': sz,
         'coupling': {'lin_coupled_modes': [1, 0, 0, 1, 0], 'coupling_type': [0], 'coupl_trafo_matrices': [None] * 5},
         'constrained_modes': [1] * 5, 'constraints': [nn] * 5, 'weights': [0.5, 0.5], 'object': [None, np.asfortranarray(Y)],
         'rank': [R, R]}
    def normc(F): return F / np.linalg.norm(F, axis=0)
    G = {'fac': [normc(rng.rand(s, R)) for s in sz], 'constraint_fac': [rng.rand(s, R) for s in sz],
         'constraint_dual_fac': [rng.rand(s, R) for s in sz], 'coupling_dual_fac': [rng.rand(I, R), None, None, rng.rand(I, R), None],
         'coupling_fac': [rng.rand(I, R)]}
    return Z, G, (A, B, C)

if __name__ == '__main__':
    I, J, K, M, R = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (1000, 1000, 1000, 5000, 32))]
    iters = int(sys.argv[6]) if len(sys.argv) > 6 else 10
    Z, G, (A, B, C) = build(I, J, K, M, R)
    t = time.time()
    s = ab.Solver(Z, [1.0, 1.0])
    s.generate_cp_data(1, [A, B, C], 0.2, 1234)
    s.set_state(G)
    print('setup %.2fs' % (time.time() - t))
    flops = 2.0 * I * J * K * R; bytes_ = 8.0 * I * J * K
    prec = int(os.environ.get('PROBE_PREC', '0'))   # 1: opt-in TF32 MTTKRP
    if prec:
        s.run(dict(MaxOuterIters=1, MaxInnerIters=1, AbsFuncTol=0, OuterRelTol=0, innerRelPrTol_coupl=0, innerRelPrTol_constr=0,
                   innerRelDualTol_coupl=0, innerRelDualTol_constr=0, mttkrp_precision=prec))
        s.set_state(G)
    for pos in (1, 2, 3):
        ms = s.time_mttkrp(1, pos, 5)
        print('mttkrp mode %d: %.3f ms  %.2f TFLOP/s  %.1f GB/s' % (pos, ms, flops / ms * 1e-9, bytes_ / ms * 1e-6))
    for pos in (1, 2):
        ms = s.time_mttkrp(2, pos, 5)
        print('matrix mode %d: %.3f ms' % (pos, ms))
    opts = dict(MaxOuterIters=3, MaxInnerIters=5, AbsFuncTol=0, OuterRelTol=0, innerRelPrTol_coupl=0, innerRelPrTol_constr=0,
                innerRelDualTol_coupl=0, innerRelDualTol_constr=0, mttkrp_precision=prec, dimtree=int(os.environ.get('PROBE_DIMTREE', '0')))
    s.run(opts)
    opts['MaxOuterIters'] = iters
    l0 = s.launch_count(); p0 = s.phase_ms().copy()
    t = time.time(); out = s.run(opts); wall = time.time() - t
    ms = s.last_run_ms()
    ph = s.phase_ms() - p0
    print('run %d iters: device %.2f ms (%.3f ms/iter, %.1f it/s) wall %.3fs launches/iter %.1f' % (iters, ms, ms / iters, iters / ms * 1e3, wall, (s.launch_count() - l0) / iters))
    print('phase ms per iter: mttkrp %.3f matrix %.3f' % (ph[0] / iters, ph[1] / iters))
    print('f', out['func_val_conv'][:4], out['func_val_conv'][-1])
