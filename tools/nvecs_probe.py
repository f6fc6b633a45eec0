"""Time the device-side nvecs initialisation (cmtf_nvecs.m) on a device-generated tensor: python tools/nvecs_probe.py I J K R"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import aoadmm_b200 as ab  # noqa: E402
from perf_probe import build  # noqa: E402


def main():
    I, J, K, R = [int(a) for a in sys.argv[1:5]]
    Z, G, facs = build(I, J, K, 64, R)
    with ab.Solver(Z, [1.0, float(np.sum(Z['object'][1] ** 2))]) as s:
        s.generate_cp_data(1, list(facs), 0.2, 99)
        for n in (1, 2, 3):
            t0 = time.perf_counter()
            U, info = s.nvecs(n, R, return_info=True)
            dt = time.perf_counter() - t0
            rows = (I, J, K)[n - 1]
            gram_flops = 1.0 * rows * rows * I * J * K / rows * 2 / 2      # upper triangle only
            ortho = np.linalg.norm(U.T @ U - np.eye(R))
            # alignment with the generating factor (noise 0.2): cosines of the principal angles
            Q = np.linalg.qr(facs[n - 1])[0]
            cosines = np.linalg.svd(Q.T @ U, compute_uv=False)
            print('nvecs mode %d: %.3f s wall (%d iterations, residual %.1e, Gram %.2e flop => >= %.1f TFLOP/s if the Gram were all), '
                  '||U\'U-I|| %.1e, min cos(angle to true factor) %.4f' % (n, dt, info['iterations'], info['residual'], gram_flops,
                                                                     gram_flops / dt / 1e12, ortho, cosines.min()), flush=True)


if __name__ == '__main__':
    main()
