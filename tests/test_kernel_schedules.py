"""Index arithmetic of the warp-level PARAFAC2 polar-factor kernel (csrc/par2.cu, par2_B_step1_reg_kernel) restated in
Python: the static register permutation of the round-robin ring (ring_step / ring_newpos) and the lane that ends up with
each total after the transposing butterfly (reduce_transpose / transpose_index).  The CUDA templates are the same few
lines; these checks pin the invariants the kernel relies on (no GPU needed)."""
import itertools

import numpy as np
import pytest


def ring_step(c):
    """one step of the tournament ring: t_0 stays, b_0 -> t_1, t_q -> t_{q+1}, t_{n-1} -> b_{n-1}, b_q -> b_{q-1}"""
    RE = len(c)
    n = RE // 2
    nt, nb = [None] * n, [None] * n
    nt[0] = c[0]
    nt[1] = c[1]
    for q in range(2, n):
        nt[q] = c[2 * (q - 1)]
    for q in range(n - 1):
        nb[q] = c[2 * (q + 1) + 1]
    nb[n - 1] = c[2 * (n - 1)]
    out = [None] * RE
    for q in range(n):
        out[2 * q], out[2 * q + 1] = nt[q], nb[q]
    return out


def ring_newpos(RE, pos):
    n, q = RE // 2, pos >> 1
    if pos & 1 == 0:
        return 0 if q == 0 else (2 * (n - 1) + 1 if q == n - 1 else 2 * (q + 1))
    return 2 if q == 0 else 2 * (q - 1) + 1


@pytest.mark.parametrize('RE', [4, 8, 16])
def test_ring_visits_every_pair_once_per_sweep_and_returns_home(RE):
    c = list(range(RE))
    seen = []
    for _ in range(RE - 1):
        seen += [frozenset((c[2 * q], c[2 * q + 1])) for q in range(RE // 2)]
        c2 = ring_step(c)
        assert all(c2[ring_newpos(RE, pos)] == c[pos] for pos in range(RE))    # the tracked norms follow their columns
        c = c2
    assert len(seen) == len(set(seen)) == RE * (RE - 1) // 2                      # every pair exactly once
    assert set(seen) == {frozenset(p) for p in itertools.combinations(range(RE), 2)}
    assert c == list(range(RE))                                                   # whole sweeps end where they began


def reduce_transpose(vals):
    """vals[lane][i] -> value held by every lane afterwards (sum over the lanes of one index)"""
    N = len(vals[0])
    v = [list(x) for x in vals]
    off, half = 16, N // 2
    while half >= 1:
        nv = [list(x) for x in v]
        for l in range(32):
            up, o = (l & off) != 0, l ^ off
            for i in range(half):
                sent = v[o][i] if (o & off) else v[o][i + half]
                keep = v[l][i + half] if up else v[l][i]
                nv[l][i] = keep + sent
        v, half, off = nv, half // 2, off // 2
    while off >= 1:
        v = [[v[l][0] + v[l ^ off][0]] + v[l][1:] for l in range(32)]
        off //= 2
    return [x[0] for x in v]


def transpose_index(N, lane):
    idx, off, half = 0, 16, N // 2
    while half >= 1:
        if lane & off:
            idx |= half
        half //= 2
        off //= 2
    return idx


@pytest.mark.parametrize('N', [2, 4, 8, 16])
def test_transposing_butterfly_leaves_total_i_in_its_lane_group(N):
    rng = np.random.RandomState(N)
    vals = rng.rand(32, N)
    out = reduce_transpose(vals.tolist())
    tot = vals.sum(0)
    for lane in range(32):
        assert abs(out[lane] - tot[transpose_index(N, lane)]) < 1e-12
    writers = [lane for lane in range(32) if (lane & (32 // N - 1)) == 0]         # one writer lane per index
    assert sorted(transpose_index(N, lane) for lane in writers) == list(range(N))
