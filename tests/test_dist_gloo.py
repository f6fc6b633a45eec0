"""CPU, world_size 2 (gloo): the multi-GPU decomposition of SURVEY.md 8e - mode-3 slabs, partial MTTKRPs of the
non-local modes all-reduced, local rows of the sharded mode gathered (all-reduce of zero-padded rows) - reproduces
the unsharded solve.  The engine does exactly this with NCCL; here the oracle's MTTKRP is replaced by the sharded one."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import problem_gen as pg
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM
from oracle.tensor_ops import mttkrp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import aoadmm_b200 as ab
    Z, G, _ = pg.config_cp_matrix(14, 12, 11, 20, 3, seed=5)
    K = Z['size'][2]
    lo, hi = ab.shard_range(K, rank, world)

    def sharded_mttkrp(p, X, U, n):
        if X.ndim < 3:
            return mttkrp(X, U, n)                      # matrices are replicated
        Xl = X[..., lo:hi]
        Ul = list(U)
        Ul[-1] = U[-1][lo:hi]
        if n == X.ndim - 1:
            out = np.zeros((X.shape[-1], U[0].shape[1]))
            out[lo:hi] = mttkrp(Xl, Ul, n)              # local rows, zero elsewhere
        else:
            out = mttkrp(Xl, Ul, n)                     # partial sum over the local slab
        t = torch.from_numpy(out)
        dist.all_reduce(t)
        return t.numpy()

    Gs, outs = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=12),
                               mttkrp_fn=sharded_mttkrp)
    if rank == 0:
        Gr, outr = cmtf_fun_AOADMM(Z, pg.znorm_const(Z), G, options=pg.default_options(MaxOuterIters=12))
        errs = [float(np.linalg.norm(Gs['fac'][m] - Gr['fac'][m]) / np.linalg.norm(Gr['fac'][m])) for m in range(5)]
        ret['errs'] = errs
        ret['df'] = float(abs(outs['f_tensors'] - outr['f_tensors']))
    # replicated state stays in lock-step: every rank holds bit-identical factors
    t = torch.from_numpy(np.ascontiguousarray(Gs['fac'][0]))
    t0 = t.clone()
    dist.broadcast(t0, 0)
    ret['same_%d' % rank] = bool(torch.equal(t, t0))
    dist.destroy_process_group()


def test_sharded_solve_matches_unsharded_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert max(ret['errs']) < 1e-10, ret['errs']
    assert ret['df'] < 1e-12
    assert ret['same_0'] and ret['same_1']


# ---------------------------------------------------------------------------------------------------------------------
# PARAFAC2 slices sharded over the ranks (SURVEY.md 8e last line; Engine Par2State::sharded): every rank owns a contiguous
# range of slices, per-slice work is local, the sums over k are all-reduced, the rows of the third mode are gathered by
# "zero the foreign rows + all-reduce".  The identities below are exactly what the engine relies on; they are checked
# against the unsharded formulas of cmtf_fun_AOADMM.m:159-178 (mode A), :509-547 (B_k / P_k / DeltaB) and :220-243 (C).
def _par2_worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.RandomState(11)                    # same numbers on every rank
    I, R, Jk = 9, 3, [7, 5, 8, 6, 9]
    K = len(Jk)
    X = [rng.randn(I, j) for j in Jk]
    A, C = rng.rand(I, R), rng.rand(K, R) + 0.1
    B = [rng.randn(j, R) for j in Jk]
    mu = [0.1 * rng.randn(j, R) for j in Jk]
    DeltaB = rng.randn(R, R)
    k0, k1 = (K * rank) // world, (K * (rank + 1)) // world

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    # mode A (:159-164): A_mttkrp = sum_k X_k B_k diag(c_k), C_had = sum_k diag(c_k) B_k'B_k diag(c_k)
    full_A = sum(X[k] @ B[k] @ np.diag(C[k]) for k in range(K))
    full_H = sum(np.diag(C[k]) @ B[k].T @ B[k] @ np.diag(C[k]) for k in range(K))
    part_A = sum((X[k] @ B[k] @ np.diag(C[k]) for k in range(k0, k1)), np.zeros((I, R)))
    part_H = sum((np.diag(C[k]) @ B[k].T @ B[k] @ np.diag(C[k]) for k in range(k0, k1)), np.zeros((R, R)))
    errA = float(np.linalg.norm(allreduce(part_A) - full_A) / np.linalg.norm(full_A))
    errH = float(np.linalg.norm(allreduce(part_H) - full_H) / np.linalg.norm(full_H))

    # B_k mode (:525-544): rho_k, P_k = polar((B_k + mu_k) DeltaB') are per-slice; DeltaB = sum_k rho_k P_k'(B_k+mu_k) / sum_k rho_k
    GA = A.T @ A
    rho = np.array([np.trace(np.diag(C[k]) @ GA @ np.diag(C[k])) / R for k in range(K)])

    def polar(M):
        U, _, Vt = np.linalg.svd(M, full_matrices=False)
        return U @ Vt
    P = [polar((B[k] + mu[k]) @ DeltaB.T) for k in range(K)]
    full_D = sum(rho[k] * P[k].T @ (B[k] + mu[k]) for k in range(K)) / rho.sum()
    part = np.zeros(R * R + 1)
    for k in range(k0, k1):                                 # the engine's redbuf: R*R numerator entries + sum of rho_k
        part[:R * R] += (rho[k] * P[k].T @ (B[k] + mu[k])).ravel()
        part[R * R] += rho[k]
    tot = allreduce(part)
    errD = float(np.linalg.norm(tot[:R * R].reshape(R, R) / tot[R * R] - full_D) / np.linalg.norm(full_D))
    # residual ratios averaged over ALL slices (:577-585): local sums, all-reduce, divide by K
    ratios = np.array([np.linalg.norm(B[k] - P[k] @ full_D) / np.linalg.norm(B[k]) for k in range(K)])
    err_res = float(abs(allreduce(np.array([ratios[k0:k1].sum()]))[0] / K - ratios.mean()))

    # mode C (:220-223): row k of the right-hand side = diag(A' X_k B_k); rows prepared by their owners, gathered by
    # zeroing the foreign rows and all-reducing
    full_rhs = np.stack([np.diag(A.T @ X[k] @ B[k]) for k in range(K)])
    mine = np.zeros((K, R))
    for k in range(k0, k1):
        mine[k] = np.diag(A.T @ X[k] @ B[k])
    err_rhs = float(np.linalg.norm(allreduce(mine) - full_rhs))
    gathered = allreduce(mine)
    t0 = torch.from_numpy(gathered.copy())
    dist.broadcast(t0, 0)
    ret['par2_%d' % rank] = dict(errA=errA, errH=errH, errD=errD, err_res=err_res, err_rhs=err_rhs,
                                 same=bool(np.array_equal(t0.numpy(), gathered)), owned=(k0, k1))
    dist.destroy_process_group()


def test_sharded_parafac2_slices_decomposition_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_par2_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    owned = []
    for r in (0, 1):
        d = ret['par2_%d' % r]
        assert d['errA'] < 1e-14 and d['errH'] < 1e-14 and d['errD'] < 1e-14, d
        assert d['err_res'] < 1e-15 and d['err_rhs'] < 1e-13 and d['same'], d
        owned.append(d['owned'])
    assert owned[0][0] == 0 and owned[0][1] == owned[1][0] and owned[1][1] == 5     # contiguous cover of the slices
