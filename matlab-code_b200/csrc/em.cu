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
    unsigned mk[8];   // bit 0 / bit 8: element 0 / 1 observed; 0xFFFF0000: column outside the object
    {
      const long long colbase = ie + a.ldI * (j0 + 8 * warp + (long long)a.J * k);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const long long idx = colbase + a.ldI * c;
        const bool col_ok = (j0 + 8 * warp + c < a.J);
        x0[c] = x1[c] = 0.0;
        mk[c] = 0xFFFF0000u;
        if (col_ok && nrow == 2) {   // idx is even (ldI and ie are): 16-byte / 2-byte aligned vector loads
          const double2 xv = *reinterpret_cast<const double2*>(a.X + idx);
          const unsigned short mv = *reinterpret_cast<const unsigned short*>(a.mask + idx);
          x0[c] = xv.x;
          x1[c] = xv.y;
          mk[c] = mv;
        } else if (col_ok && nrow == 1) {
          x0[c] = a.X[idx];
          mk[c] = 0x0000FF00u | a.mask[idx];   // second row outside: marked neither observed nor missing below
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
        const unsigned m16 = mk[c];
        const bool in0 = (m16 >> 16) == 0, in1 = in0 && ((m16 & 0xFF00u) != 0xFF00u || nrow == 2);
        // branch-free sums; only the store of an imputed value is predicated
        const bool ob0 = in0 && (m16 & 0xFFu) != 0, ob1 = in1 && (m16 & 0xFF00u) != 0;
        const bool mi0 = in0 && !ob0, mi1 = in1 && !ob1;
        const double xo0 = ob0 ? x0[c] : 0.0, mo0 = ob0 ? mv.x : 0.0, xo1 = ob1 ? x1[c] : 0.0, mo1 = ob1 ? mv.y : 0.0;
        const double xm0 = mi0 ? x0[c] : 0.0, mm0 = mi0 ? mv.x : 0.0, xm1 = mi1 ? x1[c] : 0.0, mm1 = mi1 ? mv.y : 0.0;
        s[2] = fma(xo0, mo0, s[2]);
        s[2] = fma(xo1, mo1, s[2]);
        s[3] = fma(mo0, mo0, s[3]);
        s[3] = fma(mo1, mo1, s[3]);
        const double do0 = xo0 - mo0, do1 = xo1 - mo1;
        s[4] = fma(do0, do0, s[4]);
        s[4] = fma(do1, do1, s[4]);
        const double dm0 = mm0 - xm0, dm1 = mm1 - xm1;
        s[0] = fma(dm0, dm0, s[0]);
        s[0] = fma(dm1, dm1, s[0]);
        s[1] = fma(xm0, xm0, s[1]);
        s[1] = fma(xm1, xm1, s[1]);
        if (a.impute) {
          const long long idx = colbase + a.ldI * c;
          if (mi0) a.X[idx] = mv.x;
          if (mi1) a.X[idx + 1] = mv.y;
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

}  // namespace

size_t em_partials_doubles(const EmArgs& a) {
  dim3 g;
  int kper;
  em_grid(a, g, kper);
  return (size_t)g.x * g.y * g.z * 5;
}

int em_pass(const EmArgs& a, double* sums_out, cudaStream_t st) {
  if (a.I <= 0 || a.J <= 0 || a.K <= 0) return 0;
  dim3 g;
  int kper;
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
