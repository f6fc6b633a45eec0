"""Achievable pinned host->device bandwidth on this box (1-D copy), for the e2e leg's interpretation."""
import time
import torch
n = 1 << 30  # 8 GiB of doubles
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device='cuda')
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('pinned H2D 1-D: %.1f GB/s' % (8 * n / dt / 1e9), flush=True)
# 2-D copy with equal pitches (what aoadmm_create issues when the leading dimension needs no padding)
from cuda import cudart
for width in (8000, 32768):
    rows = (8 * n) // width
    torch.cuda.synchronize(); t0 = time.perf_counter()
    err, = cudart.cudaMemcpy2D(d.data_ptr(), width, h.data_ptr(), width, width, rows, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('pinned H2D 2-D width %d: %.1f GB/s (%s)' % (width, width * rows / dt / 1e9, err), flush=True)
p = torch.empty(n // 4, dtype=torch.float64)
p.fill_(1.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
d[: n // 4].copy_(p); torch.cuda.synchronize()
print('pageable H2D: %.1f GB/s' % (8 * (n // 4) / (time.perf_counter() - t0) / 1e9))
