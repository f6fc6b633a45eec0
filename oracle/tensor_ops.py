"""Oracle: dense tensor helpers (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the Tensor Toolbox v3.1 calls made on the hot path (the toolbox itself is
not vendored, README.md:8): `mttkrp` (cmtf_fun_AOADMM.m:97, cp_func.m:47),
`full(ktensor(...))` (create_coupled_data.m:158), `norm` (cmtf_AOADMM.m:136).
Tensors are NumPy arrays in Fortran (column-major) order to mirror MATLAB.
"""
import numpy as np


def khatrirao(mats):
    """Column-wise Khatri-Rao product; the FIRST matrix varies fastest in the rows
    (Tensor Toolbox `khatrirao(U{end:-1:1})` convention expressed for a list given in
    natural order)."""
    R = mats[0].shape[1]
    out = mats[0]
    for M in mats[1:]:
        # rows index (i_prev + n_prev * i_new)
        out = (M[:, None, :] * out[None, :, :]).reshape(-1, R)
    return out


def mttkrp(X, U, n):
    """Matricised tensor times Khatri-Rao product, mode n (0-based): the dense
    Tensor Toolbox formulation (unfold + explicit Khatri-Rao + GEMM)."""
    N = X.ndim
    R = U[0].shape[1]
    others = [U[k] for k in range(N) if k != n]
    Xn = np.reshape(np.moveaxis(X, n, 0), (X.shape[n], -1), order='F')
    return Xn @ khatrirao(others) if others else Xn @ np.ones((1, R))


def full_ktensor(U, lambdas=None):
    """full(ktensor(lambda, U))."""
    R = U[0].shape[1]
    lam = np.ones(R) if lambdas is None else np.asarray(lambdas, dtype=np.float64).reshape(-1)
    shape = tuple(u.shape[0] for u in U)
    kr = khatrirao(U[1:]) if len(U) > 1 else np.ones((1, R))
    X1 = (U[0] * lam[None, :]) @ kr.T
    return np.reshape(X1, shape, order='F')
