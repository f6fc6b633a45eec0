// em.cu - EM imputation of missing entries and the masked objective terms.
//
// Reference: functions/cmtf_fun_AOADMM.m
//   :408-441   after every sweep the missing entries (Z.miss == 0) of each object are overwritten by the current
//              low-rank model; f_rel_missing = sqrt(sum (new-old)^2 / sum old^2)
//   :1224-1226 objective of a CP object with a mask:   w (Znorm - 2 <X, M.*model> + ||M.*model||^2)
//   :1249-1252 objective of a PARAFAC2 object with a mask: w sum_k || M_k .* (X_k - A D_k B_k') ||^2
// One pass over the object does all of it: the model value of every element is a rank-R product
//     m(i,j,k) = sum_r Fi(i,r) Fj(j,r) Fk(k,r)
// evaluated as a GEMM per slab k on the FP64 tensor cores (DMMA.8x8x4; 64 x 64 output tiles, factor tiles staged in
// shared memory once per CTA), compared with the stored element and the mask, and five sums are reduced deterministically (per-CTA partials, fixed-order final
// sum):  [0] sum_missing (m-x)^2   [1] sum_missing x^2   [2] sum_observed x*m   [3] sum_observed m^2
//        [4] sum_observed (x-m)^2
// A PARAFAC2 object is the K = 1 case on the stacked I x Jtot matrix with Fj(j,:) = B(j,:) .* C(seg(j),:).
#include "em.cuh"
#include "mttkrp.cuh"

#include <algorithm>

namespace aoadmm {

namespace {

constexpr int kTile = 64;          // tile rows (i)
constexpr int kTJ = 32;            // tile columns (j): 64 x 32 tiles with 4 warps give four resident CTAs per SM whose
                                   // load / tensor-core / epilogue phases interleave (64 x 64 with 8 warps: two)
constexpr int kEmThreads = 128;
constexpr int kRC = 32;          // rank chunk staged per pass
constexpr int kPitch = kTile + 4;  // smem row pitch: the 4 x 8 fragment footprint of a DMMA operand hits 32 distinct banks

constexpr int kMP = kTile + 2;     // pitch of the model tile: C-fragment stores and 16-byte row reads are conflict-free

// running sums of one element: observed -> [2] += x m, [3] += m^2 (, [4] += (x-m)^2);  missing -> [0] += (m-x)^2,
// [1] += x^2.  Written as two predicated groups (no selects: the FP64 pipe is shared with the DMMAs, and every select
// of a double is two more instructions on an issue-bound path).
__device__ __forceinline__ void em_accumulate4(double x, double m, bool observed, double& s0, double& s1, double& s2,
                                               double& s3) {
  const double d = m - x;
  if (observed) {
    s2 = fma(x, m, s2);
    s3 = fma(m, m, s3);
  } else {
    s0 = fma(d, d, s0);
    s1 = fma(x, x, s1);
  }
}

// The model of slab k is a rank-R GEMM, M_k = (Fi diag(Fk(k,:))) * Fj': it runs on the FP64 tensor cores like the
// MTTKRP (mma.m8n8k4, SASS DMMA.8x8x4).  CTA tile 64 (i) x 32 (j), 4 warps along i, warp tile 16 x 32 =
// 2 x 4 accumulator tiles.  The unscaled factor tiles are staged in shared memory once per CTA (once per rank chunk
// when R > 32) and reused for every k of the CTA's range; the k-dependent scale Fk(k,r) is applied to the A fragment
// in registers (2 DMUL per 8 DMMA).
// Epilogue per k: the accumulators go through shared memory so that the comparison with the stored data runs in the
// data's own layout - warp w owns 8 columns j, lane l the rows 2l, 2l+1: one 16-byte load of X and one 2-byte load of
// the mask per column and lane (a warp reads one whole 512-byte column segment), requested BEFORE the tensor-core loop
// so that the DRAM latency hides behind it.  (The first DMMA version compared in the accumulator layout: 32 scalar
// loads with their own 64-bit addresses per lane and k made it issue bound - 1080 instructions per warp and k, FP64
// pipe 7 % busy, profiles/r02_ncu_em_summary.md.)
__global__ void __launch_bounds__(kEmThreads, 4) em_kernel(EmArgs a, int kper) {
  extern __shared__ __align__(16) double em_smem[];   // 68.4 KB: beyond the static limit, two CTAs per SM
  double* Ms = em_smem;                                                     // kTJ x kMP model tile (column-major)
  double (*As)[kPitch] = reinterpret_cast<double (*)[kPitch]>(em_smem + kTJ * kMP);
  double (*Bs)[kPitch] = reinterpret_cast<double (*)[kPitch]>(em_smem + kTJ * kMP + kRC * kPitch);
  double* cs = em_smem + kTJ * kMP + 2 * kRC * kPitch;
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wi = warp, wj = 0;   // 4 warps along i, each 16 (i) x 32 (j)
  const int q = lane & 3, p = lane >> 2;
  const long long i0 = (long long)blockIdx.x * kTile, j0 = (long long)blockIdx.y * kTJ;
  const int k0 = blockIdx.z * kper, k1 = min(a.K, k0 + kper);
  const int nchunk = (a.R + kRC - 1) / kRC;
  // epilogue ownership: rows ie, ie+1 of the columns j0 + 8*warp + c, c = 0..7
  const long long ie = i0 + 2 * lane;
  const int nrow = (ie + 1 < a.I) ? 2 : ((ie < a.I) ? 1 : 0);
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = k0; k < k1; ++k) {
    double acc[2][4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    double x0[8], x1[8];
    unsigned mk[8];   // byte 0 / byte 1: rows 2l / 2l+1 observed (non-zero); outside the object: "observed", data 0
    {
      const long long colbase = ie + a.ldI * (j0 + 8 * warp + (long long)a.J * k);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const long long idx = colbase + a.ldI * c;
        const bool col_ok = (j0 + 8 * warp + c < a.J);
        x0[c] = x1[c] = 0.0;
        mk[c] = 0x0101u;
        if (col_ok && nrow == 2) {   // idx is even (ldI and ie are): 16-byte / 2-byte aligned vector loads
          const double2 xv = *reinterpret_cast<const double2*>(a.X + idx);
          const unsigned short mv = *reinterpret_cast<const unsigned short*>(a.mask + idx);
          x0[c] = xv.x;
          x1[c] = xv.y;
          mk[c] = mv;
        } else if (col_ok && nrow == 1) {
          x0[c] = a.X[idx];
          mk[c] = 0x0100u | a.mask[idx];   // second row outside
        }
      }
    }
    for (int c = 0; c < nchunk; ++c) {
      const int r0 = c * kRC;
      __syncthreads();   // every warp has finished reading cs / As / Bs / Ms of the previous (k, chunk)
      if (nchunk > 1 || k == k0) {
        for (int e = tid; e < kRC * kTile; e += kEmThreads) {
          const int ii = e % kTile, rr = e / kTile;
          const int r = r0 + rr;
          As[rr][ii] = (r < a.R && i0 + ii < a.I) ? a.Fi[i0 + ii + (long long)r * a.ldFi] : 0.0;
          if (ii < kTJ) Bs[rr][ii] = (r < a.R && j0 + ii < a.J) ? a.Fj[j0 + ii + (long long)r * a.ldFj] : 0.0;
        }
      }
      if (tid < kRC) {
        const int r = r0 + tid;
        cs[tid] = (r < a.R) ? ((a.Fk != nullptr) ? a.Fk[k + (long long)r * a.ldFk] : 1.0) : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < kRC / 4; ++ks) {
        const int rr = 4 * ks + q;
        const double ck = cs[rr];
        double af[2], bf[4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) af[mt] = As[rr][16 * wi + 8 * mt + p] * ck;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bf[nt] = Bs[rr][32 * wj + 8 * nt + p];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
      }
    }
    // accumulator (mt, nt, e) of this lane is model element (i = 16 wi + 8 mt + p, j = 32 wj + 8 nt + 2 q + e) of the tile
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        Ms[(32 * wj + 8 * nt + 2 * q) * kMP + 16 * wi + 8 * mt + p] = acc[mt][nt][0];
        Ms[(32 * wj + 8 * nt + 2 * q + 1) * kMP + 16 * wi + 8 * mt + p] = acc[mt][nt][1];
      }
    __syncthreads();
    {
      const long long colbase = ie + a.ldI * (j0 + 8 * warp + (long long)a.J * k);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double2 mv = *reinterpret_cast<const double2*>(&Ms[(8 * warp + c) * kMP + 2 * lane]);
        // outside the object data and model are both 0 (zero-padded factor tiles): no special case in the sums
        const bool ob0 = (mk[c] & 0xFFu) != 0, ob1 = (mk[c] & 0xFF00u) != 0;
        em_accumulate4(x0[c], mv.x, ob0, s[0], s[1], s[2], s[3]);
        em_accumulate4(x1[c], mv.y, ob1, s[0], s[1], s[2], s[3]);
        if (a.Fk == nullptr) {   // PARAFAC2 / matrix objects: the direct residual of :1249-1252
          const double d0 = x0[c] - mv.x, d1 = x1[c] - mv.y;
          if (ob0) s[4] = fma(d0, d0, s[4]);
          if (ob1) s[4] = fma(d1, d1, s[4]);
        }
        if (a.impute) {
          const long long idx = colbase + a.ldI * c;
          if (!ob0) a.X[idx] = mv.x;
          if (!ob1) a.X[idx + 1] = mv.y;
        }
      }
    }
  }
  const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    const double v = block_sum(s[t], red);
    if (tid == 0) a.partials[cta * 5 + t] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pipelined version for objects with many slabs (K >= 8, R <= 128): one persistent-style CTA per SM walks the slabs
// k of its (i-tile, j-tile) column, with the three phases of a slab on different warps so that they overlap ACROSS
// slabs instead of alternating inside a CTA (the kernel above keeps the FP64 pipe - DMMA and the epilogue's DFMAs
// share it - busy only 52 % of the time: profiles/r02_ncu_em_v4_summary.md):
//   warp 16     producer: TMA box load of the 64 x 32 data tile of slab k (dense, zero fill outside) and a bulk copy
//               of row k of the transposed third factor into a 4-stage ring;
//   warps 0-7   model GEMM of slab k on DMMA (16 x 16 each; two such warps per scheduler, one warp alone issues a
//               DMMA only every ~32 cycles: profiles/r02_ncu_em_v4_summary.md) from the factor tiles staged once per
//               CTA - for R <= 32 held as register fragments for the whole slab range - scaled by the ring's Fk row,
//               written to one of two model tiles in shared memory;
//   warps 8-15  comparison of model tile and data tile in the data layout (4 columns each, lane l rows 2l, 2l+1),
//               mask bytes prefetched one slab ahead from global memory (the mask's row stride is not TMA-aligned),
//               imputed values stored straight to global memory, four running sums.
// mbarriers: xfull / xempty per ring stage (expect_tx; 16 consumer-warp arrivals), mfull / mempty per model tile.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPStages = 4;
constexpr int kPXBytes = kTile * kTJ * 8;      // 16 KB data tile
constexpr int kPBPitch = kTJ + 4;              // Fj tile pitch
constexpr int kPMmaWarps = 8, kPEpiWarps = 8;
constexpr int kPEC = kTJ / kPEpiWarps;         // columns of a comparison warp
constexpr int kPThreads = (kPMmaWarps + kPEpiWarps + 1) * 32;
constexpr int kPRegKs = 8;                     // R <= 32: the GEMM warps keep their operand fragments in registers
constexpr int kPMaxR = 128;

struct EmPipeLayout {
  int Rp, offA, offB, offX, offC, offM, offBar, bytes;
};
__host__ __device__ inline EmPipeLayout em_pipe_layout(int R) {
  EmPipeLayout l;
  l.Rp = (R + 3) / 4 * 4;
  l.offX = 0;                                         // 1024-aligned: TMA destinations
  l.offC = l.offX + kPStages * kPXBytes;
  l.offA = l.offC + kPStages * l.Rp * 8;
  l.offB = l.offA + l.Rp * kPitch * 8;
  l.offM = l.offB + l.Rp * kPBPitch * 8;
  l.offBar = l.offM + 2 * kTJ * kMP * 8;
  l.bytes = l.offBar + (2 * kPStages + 4) * 8 + 1024;
  return l;
}

template <bool REGS>
__global__ void __launch_bounds__(kPThreads, 1) em_pipe_kernel(const __grid_constant__ CUtensorMap xmap, EmArgs a, int kper) {
  extern __shared__ uint8_t em_raw[];
  const EmPipeLayout L = em_pipe_layout(a.R);
  const uint32_t sbase = (smem_u32(em_raw) + 1023u) & ~1023u;
  uint8_t* gbase = em_raw + (sbase - smem_u32(em_raw));
  const uint32_t sX = sbase + L.offX, sC = sbase + L.offC, sBar = sbase + L.offBar;
  double* As = reinterpret_cast<double*>(gbase + L.offA);   // [Rp][kPitch]
  double* Bs = reinterpret_cast<double*>(gbase + L.offB);   // [Rp][kPBPitch]
  double* Ms = reinterpret_cast<double*>(gbase + L.offM);   // 2 x [kTJ][kMP]
  const double* Xs = reinterpret_cast<const double*>(gbase + L.offX);
  const double* Cs = reinterpret_cast<const double*>(gbase + L.offC);
  __shared__ double red[kPEpiWarps][4];
  const uint32_t xfull = sBar, xempty = sBar + kPStages * 8, mfull = sBar + 2 * kPStages * 8, mempty = mfull + 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long i0 = (long long)blockIdx.x * kTile, j0 = (long long)blockIdx.y * kTJ;
  const int k0 = blockIdx.z * kper, k1 = min(a.K, k0 + kper);
  const int Rp = L.Rp;

  if (tid == 0) {
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(xfull + s * 8, 1);
      mbar_init(xempty + s * 8, kPMmaWarps + kPEpiWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(mfull + b * 8, kPMmaWarps);
      mbar_init(mempty + b * 8, kPEpiWarps);
    }
    mbar_fence_init();
  }
  for (int e = tid; e < Rp * kTile; e += kPThreads) {
    const int ii = e % kTile, r = e / kTile;
    As[r * kPitch + ii] = (r < a.R && i0 + ii < a.I) ? a.Fi[i0 + ii + (long long)r * a.ldFi] : 0.0;
    if (ii < kTJ) Bs[r * kPBPitch + ii] = (r < a.R && j0 + ii < a.J) ? a.Fj[j0 + ii + (long long)r * a.ldFj] : 0.0;
  }
  __syncthreads();

  if (warp == kPMmaWarps + kPEpiWarps) {
    if (lane == 0) {
      prefetch_tensormap(&xmap);
      for (int k = k0; k < k1; ++k) {
        const int kl = k - k0, s = kl % kPStages;
        mbar_wait(xempty + s * 8, (uint32_t)(((kl / kPStages) & 1) ^ 1));
        mbar_expect_tx(xfull + s * 8, kPXBytes + Rp * 8);
        tma_load_3d(sX + s * kPXBytes, &xmap, (int)i0, (int)j0, k, xfull + s * 8);
        bulk_load_1d(sC + s * Rp * 8, a.fkT + (long long)k * Rp, Rp * 8, xfull + s * 8);
      }
    }
    return;
  }

  if (warp < kPMmaWarps) {
    // ===== model GEMM: 16 x 16 block (rows 16*wr, columns 16*wc) of the tile per warp.  One warp issues a DMMA only
    // every ~32 cycles, the pipe takes one every 16 per scheduler: two GEMM warps per scheduler.  (Measured variants,
    // 512^3 R = 32: 4 warps of 16 x 32: 0.51 ms; 8 of 16 x 16: 0.45; 16 of 8 x 16: 0.48; 8 of 8 x 32 with the producer
    // folded into a comparison warp: 0.50.) =====
    const int q = lane & 3, p = lane >> 2, wr = warp & 3, wc = warp >> 2;
    const double* Ap = As + q * kPitch + 16 * wr + p;
    const double* Bp = Bs + q * kPBPitch + 16 * wc + p;
    const int nks = Rp / 4;
    double afr[2][kPRegKs], bfr[2][kPRegKs];
    if (REGS) {
      // both operand tiles are the same for every slab: the fragments stay in registers, a slab costs one
      // shared-memory load (its Fk entry) and two multiplies per rank step
#pragma unroll
      for (int ks = 0; ks < kPRegKs; ++ks) {
        const bool in = ks < nks;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          afr[t][ks] = in ? Ap[4 * ks * kPitch + 8 * t] : 0.0;
          bfr[t][ks] = in ? Bp[4 * ks * kPBPitch + 8 * t] : 0.0;
        }
      }
    }
    for (int k = k0; k < k1; ++k) {
      const int kl = k - k0, s = kl % kPStages, b = kl & 1;
      double acc[2][2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
      mbar_wait(xfull + s * 8, (uint32_t)((kl / kPStages) & 1));
      const double* cs = Cs + s * Rp + q;
      if (REGS) {
#pragma unroll
        for (int ks = 0; ks < kPRegKs; ++ks) {
          const double ck = (ks < nks) ? cs[4 * ks] : 0.0;
          const double a0 = afr[0][ks] * ck, a1 = afr[1][ks] * ck;
          dmma884(acc[0][0][0], acc[0][0][1], a0, bfr[0][ks]);
          dmma884(acc[0][1][0], acc[0][1][1], a0, bfr[1][ks]);
          dmma884(acc[1][0][0], acc[1][0][1], a1, bfr[0][ks]);
          dmma884(acc[1][1][0], acc[1][1][1], a1, bfr[1][ks]);
        }
      } else {
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks) {
          const double ck = cs[4 * ks];
          const double a0 = Ap[4 * ks * kPitch] * ck, a1 = Ap[4 * ks * kPitch + 8] * ck;
          const double b0 = Bp[4 * ks * kPBPitch], b1 = Bp[4 * ks * kPBPitch + 8];
          dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
          dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
          dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
          dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
        }
      }
      mbar_wait(mempty + b * 8, (uint32_t)(((kl >> 1) & 1) ^ 1));
      // accumulator (mt, nt, e) of this lane is model element (i = 16 wr + 8 mt + p, j = 16 wc + 8 nt + 2 q + e)
      double* M = Ms + b * kTJ * kMP + (16 * wc + 2 * q) * kMP + 16 * wr + p;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          M[(8 * nt) * kMP + 8 * mt] = acc[mt][nt][0];
          M[(8 * nt + 1) * kMP + 8 * mt] = acc[mt][nt][1];
        }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(mfull + b * 8);
        mbar_arrive(xempty + s * 8);
      }
    }
    return;
  }

  // ===== comparison: columns kPEC*ew .. +kPEC-1, lane l rows 2l, 2l+1 =====
  // The FP64 pipe is shared with the DMMAs of the two GEMM warps on the same scheduler and serves the ready warps in
  // turn, so a comparison warp gets roughly one FP64 instruction per ~35 cycles: the sums are spread over eight warps
  // (two per scheduler) to keep the per-slab time of a comparison warp below that of the GEMM.
  const int ew = warp - kPMmaWarps;
  const long long ie = i0 + 2 * lane;
  const int nrow = (ie + 1 < a.I) ? 2 : ((ie < a.I) ? 1 : 0);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  // Elements outside the object need no special case in the sums: TMA zero-fills the data tile and the zero-padded
  // factor tiles make the model 0 there, so every product is 0; their mask word is "observed", which suppresses the
  // store.  Mask word of a column: byte 0 / byte 1 = rows 2l / 2l+1 (non-zero = observed).
  unsigned mk[kPEC], mkn[kPEC];
  const int ncol = (int)max(0LL, min((long long)kPEC, a.J - (j0 + kPEC * ew)));   // columns of this warp inside the object
  auto load_mask = [&](int k, unsigned (&m)[kPEC]) {
    const uint8_t* mp = a.mask + (ie + a.ldI * (j0 + kPEC * ew + (long long)a.J * k));
#pragma unroll
    for (int c = 0; c < kPEC; ++c) {
      unsigned v = 0x0101u;
      if (c < ncol) {
        if (nrow == 2) v = *reinterpret_cast<const unsigned short*>(mp);   // ie and ldI are even: 2-byte aligned
        else if (nrow == 1) v = 0x0100u | *mp;
      }
      m[c] = v;
      mp += a.ldI;
    }
  };
  if (k0 < k1) load_mask(k0, mkn);
  for (int k = k0; k < k1; ++k) {
    const int kl = k - k0, s = kl % kPStages, b = kl & 1;
#pragma unroll
    for (int c = 0; c < kPEC; ++c) mk[c] = mkn[c];
    if (k + 1 < k1) load_mask(k + 1, mkn);
    mbar_wait(xfull + s * 8, (uint32_t)((kl / kPStages) & 1));
    mbar_wait(mfull + b * 8, (uint32_t)((kl >> 1) & 1));
    const double* M = Ms + b * kTJ * kMP + kPEC * ew * kMP + 2 * lane;
    const double* X = Xs + s * (kPXBytes / 8) + kPEC * ew * kTile + 2 * lane;
    double* xp = a.X + (ie + a.ldI * (j0 + kPEC * ew + (long long)a.J * k));
#pragma unroll
    for (int c = 0; c < kPEC; ++c) {
      const double2 mv = *reinterpret_cast<const double2*>(M + c * kMP);
      const double2 xv = *reinterpret_cast<const double2*>(X + c * kTile);
      const bool ob0 = (mk[c] & 0xFFu) != 0, ob1 = (mk[c] & 0xFF00u) != 0;
      em_accumulate4(xv.x, mv.x, ob0, s0, s1, s2, s3);
      em_accumulate4(xv.y, mv.y, ob1, s0, s1, s2, s3);
      if (a.impute) {
        if (!ob0) xp[0] = mv.x;
        if (!ob1) xp[1] = mv.y;
      }
      xp += a.ldI;
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(mempty + b * 8);
      mbar_arrive(xempty + s * 8);
    }
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  s3 = warp_sum(s3);
  if (lane == 0) {
    red[ew][0] = s0;
    red[ew][1] = s1;
    red[ew][2] = s2;
    red[ew][3] = s3;
  }
  named_bar_sync(1, kPEpiWarps * 32);   // the comparison warps
  if (ew == 0 && lane < 5) {
    const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
    double v = 0.0;
    if (lane < 4) {
#pragma unroll
      for (int w = 0; w < kPEpiWarps; ++w) v += red[w][lane];   // fixed order
    }
    a.partials[cta * 5 + lane] = v;
  }
}

// fkT[r + Rp*k] = Fk(k, r), zero for R <= r < Rp: row k of the third factor as one aligned bulk copy
__global__ void em_transpose_fk_kernel(const double* __restrict__ Fk, long long ldFk, int K, int R, int Rp,
                                       double* __restrict__ out) {
  const long long n = (long long)K * Rp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % Rp);
    const long long k = idx / Rp;
    out[idx] = (r < R) ? Fk[k + (long long)r * ldFk] : 0.0;
  }
}

__global__ void em_reduce_kernel(const double* __restrict__ partials, long long nctas, double* __restrict__ out) {
  __shared__ double red[32];
  for (int t = 0; t < 5; ++t) {
    double v = 0.0;
    for (long long c = threadIdx.x; c < nctas; c += blockDim.x) v += partials[c * 5 + t];
    v = block_sum(v, red);
    if (threadIdx.x == 0) out[t] = v;
  }
}

// sum of squares of the (observed) elements of an object stored with leading dimension ld: Znorm_const of
// cmtf_AOADMM.m:124-156 (with Z.miss: norm(miss.*X)^2)
__global__ void __launch_bounds__(256) norm2_masked_kernel(const double* __restrict__ X, const uint8_t* __restrict__ mask,
                                                            long long I, long long ld, long long slab,
                                                            double* __restrict__ partials) {
  __shared__ double red[32];
  const long long n = I * slab;
  double a0 = 0.0, a1 = 0.0;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; idx + stride < n; idx += 2 * stride) {
    const long long i0 = idx % I + ld * (idx / I), i1 = (idx + stride) % I + ld * ((idx + stride) / I);
    const double v0 = (mask == nullptr || mask[i0] != 0) ? X[i0] : 0.0;
    const double v1 = (mask == nullptr || mask[i1] != 0) ? X[i1] : 0.0;
    a0 = fma(v0, v0, a0);
    a1 = fma(v1, v1, a1);
  }
  if (idx < n) {
    const long long i0 = idx % I + ld * (idx / I);
    const double v0 = (mask == nullptr || mask[i0] != 0) ? X[i0] : 0.0;
    a0 = fma(v0, v0, a0);
  }
  const double v = block_sum(a0 + a1, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double v = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) v += partials[c];
  v = block_sum(v, red);
  if (threadIdx.x == 0) *out = v;
}

__global__ void em_khatri_rao_kernel(KrArgs a, double* __restrict__ out, long long K, int R) {
  const long long n = K * R;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    long long k = idx % K;
    const long long r = idx / K;
    double v = 1.0;
    for (int q = 0; q < a.n; ++q) {
      v *= a.F[q][k % a.d[q] + r * a.ld[q]];
      k /= a.d[q];
    }
    out[idx] = v;
  }
}

void em_grid(const EmArgs& a, dim3& grid, int& kper) {
  const long long ti = ceil_div(a.I, kTile), tj = ceil_div(a.J, kTJ);
  // enough CTAs for a few waves of 148 SMs, but several k per CTA so that the Fj tile is reused
  long long kz = std::min<long long>(a.K, std::max<long long>(1, (148LL * 8) / std::max<long long>(ti * tj, 1)));
  kz = std::min<long long>(kz, 65535);
  kper = (int)ceil_div(a.K, kz);
  kz = ceil_div(a.K, kper);
  grid = dim3((unsigned)ti, (unsigned)tj, (unsigned)kz);
}

// pipelined kernel: one CTA per SM, a CTA's work is kper slabs (+ about two slab-times of prologue: factor tiles, ring
// fill) - pick the slab split whose waves of 148 CTAs finish first
void em_pipe_grid(const EmArgs& a, dim3& grid, int& kper) {
  const long long ti = ceil_div(a.I, kTile), tj = ceil_div(a.J, kTJ), tiles = ti * tj;
  long long best_kz = 1;
  double best_cost = 1e300;
  for (long long kz = 1; kz <= std::min<long long>(a.K, 65535); ++kz) {
    const long long kp = ceil_div(a.K, kz), kzz = ceil_div(a.K, kp);
    if (kzz != kz) continue;
    const double cost = (double)ceil_div(tiles * kz, 148) * ((double)kp + 2.0);
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best_kz = kz;
    }
  }
  kper = (int)ceil_div(a.K, best_kz);
  grid = dim3((unsigned)ti, (unsigned)tj, (unsigned)ceil_div(a.K, kper));
}

bool em_use_pipe(const EmArgs& a) {
  return a.Fk != nullptr && a.fkT != nullptr && a.K >= 8 && a.R <= kPMaxR && (a.ldI & 1) == 0 &&
         (reinterpret_cast<uintptr_t>(a.X) & 15) == 0 && ceil_div(a.J, kTJ) <= 65535;
}

}  // namespace

size_t em_partials_doubles(const EmArgs& a) {
  dim3 g;
  int kper;
  em_grid(a, g, kper);
  size_t n = (size_t)g.x * g.y * g.z * 5;
  if (a.K >= 8) {
    em_pipe_grid(a, g, kper);
    n = std::max(n, (size_t)g.x * g.y * g.z * 5);
  }
  return n;
}

int em_pass(const EmArgs& a, double* sums_out, cudaStream_t st) {
  if (a.I <= 0 || a.J <= 0 || a.K <= 0) return 0;
  dim3 g;
  int kper;
  if (em_use_pipe(a)) {
    em_pipe_grid(a, g, kper);
    const EmPipeLayout L = em_pipe_layout(a.R);
    CUtensorMap xmap;
    const uint64_t dims[3] = {(uint64_t)a.I, (uint64_t)a.J, (uint64_t)a.K};
    const uint64_t str[2] = {(uint64_t)a.ldI * 8, (uint64_t)a.ldI * (uint64_t)a.J * 8};
    const uint32_t box[3] = {(uint32_t)kTile, (uint32_t)kTJ, 1};
    encode_map3(&xmap, a.X, dims, str, box, false);
    const unsigned tb = (unsigned)std::min<long long>(ceil_div((long long)a.K * L.Rp, 256), 148 * 8);
    em_transpose_fk_kernel<<<tb, 256, 0, st>>>(a.Fk, a.ldFk, a.K, a.R, L.Rp, a.fkT);
    AO_CHECK_LAUNCH();
    if (L.Rp <= 4 * kPRegKs) {
      ensure_dynamic_smem(reinterpret_cast<const void*>(em_pipe_kernel<true>), (size_t)L.bytes, 0);
      em_pipe_kernel<true><<<g, kPThreads, (size_t)L.bytes, st>>>(xmap, a, kper);
    } else {
      ensure_dynamic_smem(reinterpret_cast<const void*>(em_pipe_kernel<false>), (size_t)L.bytes, 0);
      em_pipe_kernel<false><<<g, kPThreads, (size_t)L.bytes, st>>>(xmap, a, kper);
    }
    AO_CHECK_LAUNCH();
    em_reduce_kernel<<<1, 256, 0, st>>>(a.partials, (long long)g.x * g.y * g.z, sums_out);
    AO_CHECK_LAUNCH();
    return 3;
  }
  em_grid(a, g, kper);
  if (g.y > 65535) throw CudaError(2, "EM imputation: object too wide");
  constexpr size_t kEmSmem = (size_t)(kTJ * kMP + 2 * kRC * kPitch + kRC) * sizeof(double);
  ensure_dynamic_smem(reinterpret_cast<const void*>(em_kernel), kEmSmem);
  em_kernel<<<g, kEmThreads, kEmSmem, st>>>(a, kper);
  AO_CHECK_LAUNCH();
  em_reduce_kernel<<<1, 256, 0, st>>>(a.partials, (long long)g.x * g.y * g.z, sums_out);
  AO_CHECK_LAUNCH();
  return 2;
}

int em_khatri_rao(const KrArgs& a, double* out, long long K, int R, cudaStream_t st) {
  if (K <= 0) return 0;
  const unsigned blocks = (unsigned)std::min<long long>(ceil_div(K * R, 256), 148 * 8);
  em_khatri_rao_kernel<<<blocks, 256, 0, st>>>(a, out, K, R);
  AO_CHECK_LAUNCH();
  return 1;
}

int object_norm2(const double* X, const uint8_t* mask, long long I, long long ld, long long slab, double* partials,
                 double* out, cudaStream_t st) {
  const int ctas = 148 * 8;
  norm2_masked_kernel<<<ctas, 256, 0, st>>>(X, mask, I, ld, slab, partials);
  AO_CHECK_LAUNCH();
  sum_partials_kernel<<<1, 256, 0, st>>>(partials, ctas, out);
  AO_CHECK_LAUNCH();
  return 2;
}

}  // namespace aoadmm
