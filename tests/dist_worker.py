"""Worker of tests/test_gpu_multi.py::test_one_process_per_gpu_matches_oracle (also runnable by hand):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_worker.py

One process per GPU (aoadmm_create + aoadmm_dist): parity of the sharded engine with the oracle (CP + matrix, PARAFAC2,
linear couplings, EM, nvecs of every mode incl. the sharded one), identical state on all ranks, and a per-rank
4096 x 4096 x 32 slab of BASELINE configs[2] whose MTTKRPs are checked row-wise against float64 einsum.
Prints DIST OK on rank 0 when everything holds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'matlab-code_b200'))
os.environ.setdefault('OMP_NUM_THREADS', '8')   # torchrun sets 1: the oracle would crawl
import torch, torch.distributed as dist
import aoadmm_b200 as ab
from oracle import problem_gen as pg
from oracle.cmtf_fun_aoadmm import cmtf_fun_AOADMM as oracle_solve

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
_UID = []
def uid():
    """one unique id for the whole run: the engine caches the communicator per id, so every handle after the first
    reuses it (a fresh id per handle also works and costs one NCCL bootstrap each)"""
    if not _UID:
        t = torch.zeros(128, dtype=torch.uint8, device='cuda')
        if rank == 0: t.copy_(torch.tensor(list(ab.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0); _UID.append(bytes(t.cpu().tolist()))
    return _UID[0]
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
nn = ('non-negativity',)
extra = [('cp+par2', pg.config_cp_par2(I=24, J=20, K=18, Jk=16, Kp=10, R=3, seed=3, noise=0.1)[:2]),
         # PARAFAC2 slices sharded over the ranks (8 irregular slices, per-slice prox on B_k, ridge)
         ('par2 sharded slices', pg.config_single_par2(seed=5, Jk=(10, 12, 9, 11, 13, 10, 8, 14), constrained=(1, 1, 1),
                                                       ridge=[1e-3, 2e-3, 1e-3])[:2]),
         ('par2 unimodal B_k', pg.config_single_par2(seed=8, Jk=(12, 9, 15, 11, 10, 13, 9, 12), constrained=(0, 1, 1),
                                                     constraints=[None, ('unimodality', True), nn])[:2]),
         ('lin4', pg.config_linear_coupling(4, seed=4)[:2]), ('lin1', pg.config_linear_coupling(1, seed=1, second='tensor')[:2])]
Zc, Gc, _ = pg.config_cp_matrix(40, 36, 30, 64, 5, seed=8)
extra.append(('em', (pg.add_missing(Zc, 0.25, seed=3), Gc)))       # masks are sharded with the tensor
for name, (Z, G) in extra:
    opts = pg.default_options(MaxOuterIters=6 if 'B_k' in name or 'sharded slices' in name else 15)
    zn = pg.znorm_const(Z)
    Gd, od = ab.cmtf_fun_AOADMM(Z, zn, G, None, None, None, None, opts, rank=rank, world_size=world, device=lr, unique_id=uid())
    if rank == 0:
        Go, oo = oracle_solve(Z, zn, G, options=opts)
        errs = []
        for a, b in list(zip(Gd['fac'], Go['fac'])) + [(x, y) for x, y in zip(Gd.get('P') or [], Go.get('P') or []) if y is not None]:
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
# ---- BASELINE configs[2] slab: 4096 x 4096 x 32 per rank, R=64, generated on device; MTTKRP rows against einsum --------
if os.environ.get('DIST_BIG', '1') != '0':
    import bench
    I, J, Kr, M, R = 4096, 4096, 32, 8192, 64
    K = Kr * world
    Z, G, facs = bench.make_problem(I, J, K, M, R, seed=3, with_tensor=False)
    lo, hi = ab.shard_range(K, rank, world)
    zn = [1.0, float(np.sum(Z['object'][1] ** 2))]
    with ab.Solver(Z, zn, rank=rank, world_size=world, device=lr, unique_id=uid(), shard=[(lo, hi), None]) as s:
        s.generate_cp_data(1, facs, 0.2, 4242)
        s.set_state(G)
        X = np.empty((I, J, hi - lo), order='F')
        s.get_object_data(1, X)
        A, B, C = G['fac'][0], G['fac'][1], G['fac'][2]
        rng = np.random.RandomState(9)
        worst = 0.0
        for pos, n in ((1, I), (2, J)):
            Md = s.object_mttkrp(1, pos)                       # summed over ranks by the engine (NCCL all-reduce)
            rows = np.unique(rng.randint(0, n, size=24))
            ref = np.stack([np.einsum('jk,jr,kr->r', X[r, :, :], B, C[lo:hi]) if pos == 1 else
                            np.einsum('ik,ir,kr->r', X[:, r, :], A, C[lo:hi]) for r in rows])
            t = torch.from_numpy(ref).cuda(); dist.all_reduce(t); ref = t.cpu().numpy()
            worst = max(worst, float(np.linalg.norm(Md[rows] - ref) / np.linalg.norm(ref)))
        Md = s.object_mttkrp(1, 3)                             # rows of the sharded mode: every rank checks its own
        for k in range(lo, hi, 7):
            ref = np.einsum('ij,ir,jr->r', X[:, :, k - lo], A, B)
            worst = max(worst, float(np.linalg.norm(Md[k] - ref) / np.linalg.norm(ref)))
        o1 = s.run(pg.default_options(MaxOuterIters=2, AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0, innerRelPrTol_constr=0.0,
                                      innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0))
        Gs = s.get_state()
        s.set_state(G)
        o2 = s.run(pg.default_options(MaxOuterIters=2, dimtree=1, AbsFuncTol=0.0, OuterRelTol=0.0, innerRelPrTol_coupl=0.0,
                                      innerRelPrTol_constr=0.0, innerRelDualTol_coupl=0.0, innerRelDualTol_constr=0.0))
        Gt = s.get_state()
    w = torch.tensor([worst], dtype=torch.float64, device='cuda'); dist.all_reduce(w, op=dist.ReduceOp.MAX)
    tree = max(float(np.linalg.norm(Gt['fac'][m] - Gs['fac'][m]) / np.linalg.norm(Gs['fac'][m])) for m in range(5))
    t = torch.from_numpy(np.ascontiguousarray(Gs['fac'][1])).cuda(); t0 = t.clone(); dist.broadcast(t0, 0)
    same = bool(torch.equal(t, t0))
    if rank == 0:
        print('c3 slab %dx%dx%d per rank, world %d: worst MTTKRP row err %.2e, tree vs three-pass %.2e, f %.6e -> %.6e, replicas identical %s'
              % (I, J, Kr, world, float(w.item()), tree, o1['func_val_conv'][0], o1['func_val_conv'][-1], same))
    ok &= float(w.item()) < 1e-12 and tree < 1e-10 and same and bool(np.isfinite(o1['f_tensors'])) and o1['f_tensors'] < o1['func_val_conv'][0]
flag = torch.tensor([1.0 if ok else 0.0], device='cuda'); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print('DIST OK' if flag.item() > 0 else 'DIST FAILED')
dist.destroy_process_group()
