"""Oracle: proximal / projection operators (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates functions/constraints_to_prox.m:13-91 and the in-repo operators
(prox_normalized_nonneg.m:3-11, prox_TV.m:6-8, project_unimodal.m:10-14,
project_unimodal_vector.m:10-88, project_ortho.m:3-4, t_smoothness_prox.m:23-56,
t_smoothness_penalty.m:5-9).  The Proximity Operator Repository functions and
TV_Condat_v2 are NOT vendored in the reference (README.md:9, List...txt:39); they
are restated here from their published mathematical definitions (each is the
unique minimiser of a strictly convex problem, so any exact algorithm agrees to
rounding).
"""
import numpy as np


# ----------------------------------------------------------------------------
# Proximity Operator Repository restatements (definitions)
# ----------------------------------------------------------------------------
def project_box(x, l, u):
    """Clip to [l,u] (call site constraints_to_prox.m:14,18)."""
    return np.minimum(np.maximum(x, l), u)


def _project_simplex_vec(v, eta):
    n = v.shape[0]
    u = np.sort(v)[::-1]
    css = np.cumsum(u) - eta
    ind = np.arange(1, n + 1)
    cond = u - css / ind > 0
    rho = ind[cond][-1]
    theta = css[rho - 1] / rho
    return np.maximum(v - theta, 0.0)


def project_simplex(x, eta, direction):
    """Euclidean projection onto {x>=0, sum(x)=eta} along `direction` (1: each column,
    2: each row).  Call sites constraints_to_prox.m:21,24."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    if direction == 1:
        for r in range(x.shape[1]):
            out[:, r] = _project_simplex_vec(x[:, r], eta)
    else:
        for i in range(x.shape[0]):
            out[i, :] = _project_simplex_vec(x[i, :], eta)
    return out


def _pava_nondecreasing(y):
    n = y.shape[0]
    level = np.empty(n)
    weight = np.empty(n)
    start = np.empty(n, dtype=np.int64)
    nb = 0
    for i in range(n):
        level[nb] = y[i]
        weight[nb] = 1.0
        start[nb] = i
        nb += 1
        while nb > 1 and level[nb - 2] > level[nb - 1]:
            w = weight[nb - 2] + weight[nb - 1]
            level[nb - 2] = (weight[nb - 2] * level[nb - 2] + weight[nb - 1] * level[nb - 1]) / w
            weight[nb - 2] = w
            nb -= 1
    out = np.empty(n)
    for b in range(nb):
        end = start[b + 1] if b + 1 < nb else n
        out[start[b]:end] = level[b]
    return out


def project_monotone(x, direction=1):
    """Column-wise isotonic (non-decreasing) regression (constraints_to_prox.m:26,28)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    for r in range(x.shape[1]):
        out[:, r] = _pava_nondecreasing(x[:, r])
    return out


def project_L1(x, eta, direction=1):
    """Column-wise projection onto the l1 ball of radius eta (constraints_to_prox.m:34)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    for r in range(x.shape[1]):
        v = x[:, r]
        if np.sum(np.abs(v)) <= eta:
            out[:, r] = v
        else:
            out[:, r] = np.sign(v) * _project_simplex_vec(np.abs(v), eta)
    return out


def project_L2(x, eta, direction=1):
    """Column-wise projection onto the l2 ball of radius eta (constraints_to_prox.m:37,40)."""
    x = np.asarray(x, dtype=np.float64)
    nrm = np.sqrt(np.sum(x * x, axis=0))
    scale = np.ones_like(nrm)
    big = nrm > eta
    scale[big] = eta / nrm[big]
    return x * scale[None, :]


def prox_abs(x, gamma):
    """Soft threshold (constraints_to_prox.m:48)."""
    return np.sign(x) * np.maximum(np.abs(x) - gamma, 0.0)


def prox_zero(x, gamma):
    """Hard threshold, prox of gamma*||x||_0: keep x where x^2 > 2*gamma (constraints_to_prox.m:52)."""
    return np.where(np.abs(x) > np.sqrt(2.0 * gamma), x, 0.0)


def prox_L2(x, gamma, direction=1):
    """Column-wise block soft threshold, prox of gamma*||x_col||_2 (constraints_to_prox.m:56)."""
    x = np.asarray(x, dtype=np.float64)
    nrm = np.sqrt(np.sum(x * x, axis=0))
    scale = np.zeros_like(nrm)
    big = nrm > gamma
    scale[big] = 1.0 - gamma / nrm[big]
    return x * scale[None, :]


def tv_condat(y, lam):
    """Exact 1-D total-variation prox argmin 0.5||x-y||^2 + lam*sum|x[i+1]-x[i]|
    (L. Condat, "A direct algorithm for 1-D total variation denoising", IEEE SPL 2013).
    Stands in for the un-vendored TV_Condat_v2 called at prox_TV.m:7."""
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0]
    x = np.empty(n)
    if n == 0:
        return x
    if lam <= 0:
        return y.copy()
    k = k0 = kplus = kminus = 0
    umin = lam
    umax = -lam
    vmin = y[0] - lam
    vmax = y[0] + lam
    twolam = 2.0 * lam
    minlam = -lam
    while True:
        while k == n - 1:
            if umin < 0.0:
                while True:
                    x[k0] = vmin
                    k0 += 1
                    if not k0 <= kminus:
                        break
                kminus = k = k0
                vmin = y[k]
                umin = lam
                umax = vmin + umin - vmax
            elif umax > 0.0:
                while True:
                    x[k0] = vmax
                    k0 += 1
                    if not k0 <= kplus:
                        break
                kplus = k = k0
                vmax = y[k]
                umax = minlam
                umin = vmax + umax - vmin
            else:
                vmin += umin / (k - k0 + 1)
                while True:
                    x[k0] = vmin
                    k0 += 1
                    if not k0 <= k:
                        break
                return x
        umin += y[k + 1] - vmin
        if umin < minlam:
            while True:
                x[k0] = vmin
                k0 += 1
                if not k0 <= kminus:
                    break
            kplus = kminus = k = k0
            vmin = y[k]
            vmax = vmin + twolam
            umin = lam
            umax = minlam
        else:
            umax += y[k + 1] - vmax
            if umax > lam:
                while True:
                    x[k0] = vmax
                    k0 += 1
                    if not k0 <= kplus:
                        break
                kplus = kminus = k = k0
                vmax = y[k]
                vmin = vmax - twolam
                umin = lam
                umax = minlam
            else:
                k += 1
                if umin >= lam:
                    kminus = k
                    vmin += (umin - lam) / (kminus - k0 + 1)
                    umin = lam
                if umax <= minlam:
                    kplus = k
                    vmax += (umax + lam) / (kplus - k0 + 1)
                    umax = minlam


# ----------------------------------------------------------------------------
# In-repo operators
# ----------------------------------------------------------------------------
def prox_normalized_nonneg(X):
    """prox_normalized_nonneg.m:3-11."""
    X = np.asarray(X, dtype=np.float64)
    Y = project_box(X, 0.0, np.inf)
    for r in range(Y.shape[1]):
        nrm = np.sqrt(np.sum(Y[:, r] ** 2))
        if nrm == 0:
            maxcoord = int(np.argmax(X[:, r]))  # first maximum, as MATLAB max
            Y[maxcoord, r] = 1.0
        else:
            Y[:, r] = Y[:, r] / nrm
    return Y


def prox_TV(X, lam):
    """prox_TV.m:5-8."""
    X = np.asarray(X, dtype=np.float64)
    out = np.zeros_like(X)
    for r in range(X.shape[1]):
        out[:, r] = tv_condat(X[:, r], lam)
    return out


def _prefix_isotonic_regression(y, non_negativity):
    """project_unimodal_vector.m:43-88 (1-based arrays kept: index 0 unused for clarity)."""
    n = y.shape[0]
    sumwy = np.concatenate(([0.0], y))
    sumwy2 = np.concatenate(([0.0], y * y))
    sumw = np.concatenate(([0.0], np.ones(n)))
    # MATLAB arrays are 1..n+1; use 1-based python arrays of length n+2
    s_wy = np.zeros(n + 2)
    s_wy2 = np.zeros(n + 2)
    s_w = np.zeros(n + 2)
    s_wy[1:] = sumwy
    s_wy2[1:] = sumwy2
    s_w[1:] = sumw
    level_set = np.zeros(n + 2)
    index_range = np.zeros(n + 2, dtype=np.int64)
    error = np.zeros(n + 2)
    level_set[1] = -np.inf
    if non_negativity:
        cumsumwy2 = np.zeros(n + 2)
        cumsumwy2[1:] = np.cumsum(sumwy2)
        threshold = np.zeros(n + 2, dtype=bool)
    for i in range(2, n + 2):
        level_set[i] = y[i - 2]
        index_range[i] = i
        while level_set[i] <= level_set[index_range[i] - 1]:
            merger = index_range[i] - 1
            s_wy[i] += s_wy[merger]
            s_wy2[i] += s_wy2[merger]
            s_w[i] += s_w[merger]
            level_set[i] = s_wy[i] / s_w[i]
            index_range[i] = index_range[index_range[i] - 1]
        levelerror = s_wy2[i] - (s_wy[i] ** 2 / s_w[i])
        if non_negativity and level_set[i] < 0:
            threshold[i] = True
            error[i] = cumsumwy2[i - 1]
        else:
            error[i] = levelerror + error[index_range[i] - 1]
    if non_negativity:
        level_set[threshold] = 0.0
    iso_level = level_set[2:].copy()           # iso(:,1)
    iso_range = index_range[2:].copy() - 1     # iso(:,2), 1-based positions in y
    return iso_level, iso_range, error[2:].copy()


def project_unimodal_vector(x, non_negativity):
    """project_unimodal_vector.m:10-41 (Stout 2008, prefix isotonic regression)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    lvl_l, rng_l, err_l = _prefix_isotonic_regression(x, non_negativity)
    lvl_r, rng_r, err_r = _prefix_isotonic_regression(x[::-1].copy(), non_negativity)

    # get_best_unimodality_index (:21-32), 1-based
    best_error = err_r[n - 1]
    best_idx = 1
    for i in range(2, n + 1):
        e = err_l[i - 1] + err_r[n - (i - 1) - 1]
        if e < best_error:
            best_error = e
            best_idx = i

    def compute_isotonic_from_index(mode_idx, level_set, index_range):
        y_iso = np.full(mode_idx, np.nan)
        idx = mode_idx
        while idx >= 1:
            lo = index_range[idx - 1]
            y_iso[lo - 1:idx] = level_set[idx - 1]
            idx = lo - 1
        return y_iso

    left = compute_isotonic_from_index(best_idx, lvl_l, rng_l)
    right = compute_isotonic_from_index(n - best_idx, lvl_r, rng_r)
    return np.concatenate((left, right[::-1]))


def project_unimodal(X, non_negativity):
    """project_unimodal.m:10-14."""
    X = np.asarray(X, dtype=np.float64)
    out = np.zeros_like(X)
    for r in range(X.shape[1]):
        out[:, r] = project_unimodal_vector(X[:, r], non_negativity)
    return out


def project_ortho(X):
    """project_ortho.m:3-4."""
    U, _, Vt = np.linalg.svd(X, full_matrices=False)
    return U @ Vt


def t_smoothness_prox(factor_matrices, rho, smoothness_l):
    """t_smoothness_prox.m:3-56 (Thomas algorithm over the slice index)."""
    K = len(factor_matrices)
    rho = np.asarray(rho, dtype=np.float64).reshape(-1)
    rhs = [rho[i] * factor_matrices[i] for i in range(K)]
    A = np.zeros((K, K))
    for i in range(K):
        for j in range(K):
            if i == j:
                A[i, j] = 4 * smoothness_l + rho[i]
            elif i == j - 1 or i == j + 1:
                A[i, j] = -2 * smoothness_l
    A[0, 0] -= 2 * smoothness_l
    A[K - 1, K - 1] -= 2 * smoothness_l
    for i in range(1, K):
        m = A[i, i - 1] / A[i - 1, i - 1]
        A[i, i] = A[i, i] - m * A[i - 1, i]
        rhs[i] = rhs[i] - m * rhs[i - 1]
    new = [None] * K
    new[K - 1] = rhs[K - 1] / A[K - 1, K - 1]
    q = new[K - 1]
    for k in range(K - 2, -1, -1):
        q = (rhs[k] - A[k, k + 1] * q) / A[k, k]
        new[k] = q
    return new


def t_smoothness_penalty(x, smoothness_l):
    """t_smoothness_penalty.m:5-9."""
    loss = 0.0
    for i in range(1, len(x)):
        loss += np.linalg.norm(x[i] - x[i - 1], 'fro') ** 2
    return smoothness_l * loss


def gl_laplacian(n):
    """constraints_to_prox.m:71-73."""
    L = 2.0 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    L[0, 0] = 1.0
    L[-1, -1] = 1.0
    return L


# ----------------------------------------------------------------------------
# constraints_to_prox.m:1-94
# ----------------------------------------------------------------------------
def constraints_to_prox(constrained_modes, constraints, sz):
    """Returns (prox_operators, reg_func): lists (per mode) of callables prox(x, rho) and
    reg(x) or None.  `constraints[m]` is a tuple/list (name, params...) as in
    Z.constraints{m}."""
    nb = len(constrained_modes)
    prox_operators = [None] * nb
    reg_func = [None] * nb
    for m in range(nb):
        if not constrained_modes[m]:
            continue
        c = constraints[m]
        if c is None or len(c) == 0:
            raise ValueError('No constraint provided for mode %d.' % (m + 1))
        name = c[0]
        if name == 'non-negativity':
            prox_operators[m] = lambda x, rho: project_box(x, 0.0, np.inf)
        elif name == 'box':
            l, u = c[1], c[2]
            prox_operators[m] = lambda x, rho, l=l, u=u: project_box(x, l, u)
        elif name == 'simplex column-wise':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: project_simplex(x, eta, 1)
        elif name == 'simplex row-wise':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: project_simplex(x, eta, 2)
        elif name == 'non-decreasing':
            prox_operators[m] = lambda x, rho: project_monotone(x, 1)
        elif name == 'non-increasing':
            prox_operators[m] = lambda x, rho: -project_monotone(-x, 1)
        elif name == 'unimodality':
            nn = bool(c[1])
            prox_operators[m] = lambda x, rho, nn=nn: project_unimodal(x, nn)
        elif name == 'l1-ball':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: project_L1(x, eta, 1)
        elif name == 'l2-ball':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: project_L2(x, eta, 1)
        elif name == 'non-negative l2-ball':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: project_L2(project_box(x, 0.0, np.inf), eta, 1)
        elif name == 'non-negative l2-sphere':
            prox_operators[m] = lambda x, rho: prox_normalized_nonneg(x)
        elif name == 'orthonormal':
            prox_operators[m] = lambda x, rho: project_ortho(x)
        elif name == 'l1 regularization':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: prox_abs(x, eta / rho)
            reg_func[m] = lambda x, eta=eta: eta * np.sum(np.abs(x))
        elif name == 'l0 regularization':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: prox_zero(x, eta / rho)
            reg_func[m] = lambda x, eta=eta: eta * float(np.count_nonzero(x))
        elif name == 'l2 regularization':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: prox_L2(x, eta / rho, 1)
            reg_func[m] = lambda x, eta=eta: eta * np.sum(np.sqrt(np.sum(x * x, axis=0)))
        elif name == 'ridge':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: 1.0 / (2.0 * (eta / rho) + 1.0) * x
            reg_func[m] = lambda x, eta=eta: eta * np.linalg.norm(x, 'fro') ** 2
        elif name == 'quadratic regularization':
            eta, L = c[1], np.asarray(c[2], dtype=np.float64)
            prox_operators[m] = lambda x, rho, eta=eta, L=L: np.linalg.solve(2.0 * eta / rho * L + np.eye(L.shape[0]), x)
            reg_func[m] = lambda x, eta=eta, L=L: eta * np.trace(x.T @ L @ x)
        elif name == 'GL smoothness':
            eta = c[1]
            szm = sz[m][0] if isinstance(sz[m], (list, tuple, np.ndarray)) else sz[m]
            L = gl_laplacian(int(szm))
            prox_operators[m] = lambda x, rho, eta=eta, L=L: np.linalg.solve(2.0 * eta / rho * L + np.eye(L.shape[0]), x)
            reg_func[m] = lambda x, eta=eta, L=L: eta * np.trace(x.T @ L @ x)
        elif name == 'TV regularization':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: prox_TV(x, eta / rho)
            # as written in the reference (no abs): constraints_to_prox.m:81
            reg_func[m] = lambda x, eta=eta: eta * np.sum(x[1:, :] - x[:-1, :])
        elif name == 'tPARAFAC2':
            eta = c[1]
            prox_operators[m] = lambda x, rho, eta=eta: t_smoothness_prox(x, rho, eta)
            reg_func[m] = lambda x, eta=eta: t_smoothness_penalty(x, eta)
        elif name == 'custom':
            prox_operators[m] = c[1]
            if len(c) > 2:
                reg_func[m] = c[2]
        else:
            # the reference silently leaves the handle empty for unknown names
            prox_operators[m] = None
    return prox_operators, reg_func
