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
