"""torchrun --nproc-per-node N tools/dist_check.py : multi-GPU parity of the sharded engine with the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
import torch, torch.distributed as dist
import aoadmm_b200 as ab
from oracle import problem_gen as pg
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
def uid():
    t = torch.zeros(128, dtype=torch.uint8, device='cuda')
    if rank == 0: t.copy_(torch.tensor(list(ab.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0); return bytes(t.cpu().tolist())
ok = True
for dims in [(64, 48, 40, 80, 8), (130, 90, 37, 100, 32), (40, 36, 30, 64, 64)]:
    Z, G, _ = pg.config_cp_matrix(*dims, seed=1)
    opts = pg.default_options(MaxOuterIters=20)
    zn = pg.znorm_const(Z)
    Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts, rank=rank, world_size=world, device=lr, unique_id=uid())
    if rank == 0:
        Go, oo = oracle_solve(Z, zn, G, options=opts)
        errs = [np.linalg.norm(Gd['fac'][m] - Go['fac'][m]) / np.linalg.norm(Go['fac'][m]) for m in range(5)]
        print(dims, 'world', world, 'max fac err %.2e' % max(errs), 'df %.2e' % abs(od['f_tensors'] - oo['f_tensors']), 'iters', od['OuterIterations'], oo['OuterIterations'])
        ok &= max(errs) < 1e-8
    # all ranks must hold identical state
    t = torch.from_numpy(np.ascontiguousarray(Gd['fac'][2])).cuda(); t0 = t.clone(); dist.broadcast(t0, 0)
    same = bool(torch.equal(t, t0))
    if not same: print('rank', rank, 'state differs from rank 0'); ok = False
# PARAFAC2 objects (replicated) next to a sharded CP tensor, and a linear coupling (type 4, type 1)
extra = [('cp+par2', pg.config_cp_par2(I=24, J=20, K=18, Jk=16, Kp=10, R=3, seed=3, noise=0.1)[:2]),
         ('lin4', pg.config_linear_coupling(4, seed=4)[:2]), ('lin1', pg.config_linear_coupling(1, seed=1, second='tensor')[:2])]
Zc, Gc, _ = pg.config_cp_matrix(40, 36, 30, 64, 5, seed=8)
extra.append(('em', (pg.add_missing(Zc, 0.25, seed=3), Gc)))       # masks are sharded with the tensor
for name, (Z, G) in extra:
    opts = pg.default_options(MaxOuterIters=15)
    zn = pg.znorm_const(Z)
    Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts, rank=rank, world_size=world, device=lr, unique_id=uid())
    if rank == 0:
        Go, oo = oracle_solve(Z, zn, G, options=opts)
        errs = []
        for a, b in zip(Gd['fac'], Go['fac']):
            if isinstance(b, list):
                errs += [np.linalg.norm(x - y) / np.linalg.norm(y) for x, y in zip(a, b)]
            else:
                errs.append(np.linalg.norm(a - b) / np.linalg.norm(b))
        print(name, 'world', world, 'max fac err %.2e' % max(errs), 'df %.2e' % abs(od['f_tensors'] - oo['f_tensors']))
        ok &= max(errs) < 1e-8
# nvecs initialisation on the sharded tensor: partial Gram matrices of the slabs are all-reduced; for the sharded (last)
# mode the slabs are exchanged chunk by chunk (NCCL send/recv) so that slice pairs of different ranks meet
for dims in [(130, 90, 37, 100, 4), (33, 7, 41, 20, 3)]:
    Z, G, _ = pg.config_cp_matrix(*dims, seed=2)
    with ab.Solver(ab._with_rank(Z, G), pg.znorm_const(Z), rank=rank, world_size=world, device=lr, unique_id=uid()) as s:
        errs = [np.linalg.norm(s.nvecs(n, dims[4]) - pg.cmtf_nvecs(Z, n, dims[4])) for n in (1, 2, 3, 4, 5)]
        if rank == 0:
            print('nvecs', dims, 'world', world, 'err per mode', ['%.1e' % e for e in errs])
        ok &= max(errs) < 1e-8
# 4-way tensor: the sharded mode is the fourth one, the two middle modes are merged
Z4, G4, _ = pg.config_single_cp(sz=(12, 9, 7, 19), R=3, seed=3, noise=0.1)
with ab.Solver(ab._with_rank(Z4, G4), pg.znorm_const(Z4), rank=rank, world_size=world, device=lr, unique_id=uid()) as s:
    errs = [np.linalg.norm(s.nvecs(n, 3) - pg.cmtf_nvecs(Z4, n, 3)) for n in (1, 2, 3, 4)]
    if rank == 0:
        print('nvecs 4-way world', world, 'err per mode', ['%.1e' % e for e in errs])
    ok &= max(errs) < 1e-8
if rank == 0: print('DIST OK' if ok else 'DIST FAILED')
dist.destroy_process_group()
