"""Developer tool: time aoadmm_create (host->device upload of the tensor) from pageable and from pinned host memory and
check that the resident tensor equals the source bit for bit."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200')); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch
import aoadmm_b200 as ab
from perf_probe import build
I, J, K = [int(a) for a in sys.argv[1:4]]
Z, G, facs = build(I, J, K, 64, 8)
rng = np.random.RandomState(0)
n = I * J * K
for how in ('pinned', 'pageable', 'pageable'):
    if how == 'pinned':
        t = torch.empty(n, dtype=torch.float64, pin_memory=True)
        X = t.numpy().reshape((I, J, K), order='F')
    else:
        X = np.empty((I, J, K), order='F')
    flat = X.reshape(-1, order='F')
    step = 1 << 24
    for s in range(0, n, step):
        flat[s:s + step] = np.arange(s, min(s + step, n), dtype=np.float64) * 1e-9
    Zh = dict(Z, object=[X, Z['object'][1]])
    t0 = time.perf_counter()
    s = ab.Solver(Zh, [1.0, 1.0])
    dt = time.perf_counter() - t0
    back = np.empty((I, J, K), order='F')
    s.get_object_data(1, back)
    s.close()
    print('%-9s create %.3f s  (%.1f GB/s for the %.1f GB tensor)  resident == source: %s' % (how, dt, 8 * n / dt / 1e9, 8 * n / 1e9, np.array_equal(back, X)), flush=True)
    del X, back
