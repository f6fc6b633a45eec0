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

constexpr int kTile = 64;
constexpr int kRC = 32;          // rank chunk staged per pass
constexpr int kPitch = kTile + 4;  // smem row pitch: the 4 x 8 fragment footprint of a DMMA operand hits 32 distinct banks

// The model of slab k is a rank-R GEMM, M_k = (Fi diag(Fk(k,:))) * Fj': it runs on the FP64 tensor cores like the
// MTTKRP (mma.m8n8k4, SASS DMMA.8x8x4).  CTA tile 64 (i) x 64 (j), 8 warps as 4 (i) x 2 (j), warp tile 16 x 32 =
// 2 x 4 accumulator tiles.  The unscaled factor tiles are staged in shared memory once per CTA (once per rank chunk
// when R > 32) and reused for every k of the CTA's range; the k-dependent scale Fk(k,r) is applied to the A fragment
// in registers (2 DMUL per 8 DMMA).  Epilogue per k: the 16 model values of a lane are compared with the stored
// element and the mask byte, missing entries are overwritten, five sums are accumulated.
__global__ void __launch_bounds__(256, 2) em_kernel(EmArgs a, int kper) {
  __shared__ double As[kRC][kPitch];
  __shared__ double Bs[kRC][kPitch];
  __shared__ double cs[kRC];
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wi = warp & 3, wj = warp >> 2;
  const int q = lane & 3, p = lane >> 2;
  const long long i0 = (long long)blockIdx.x * kTile, j0 = (long long)blockIdx.y * kTile;
  const int k0 = blockIdx.z * kper, k1 = min(a.K, k0 + kper);
  const int nchunk = (a.R + kRC - 1) / kRC;
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = k0; k < k1; ++k) {
    double acc[2][4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    // The stored elements and mask bytes of this lane's 16 outputs are requested BEFORE the tensor-core loop, so their
    // DRAM latency hides behind it: accumulator (mt, nt, e) is element i = i0 + 16 wi + 8 mt + p, j = j0 + 32 wj + 8 nt + 2 q + e.
    double xv[2][4][2];
    uint8_t mk[2][4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const long long j = j0 + 32 * wj + 8 * nt + 2 * q + e;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const long long i = i0 + 16 * wi + 8 * mt + p;
          const bool in = (i < a.I) && (j < a.J);
          const long long idx = in ? i + a.ldI * (j + (long long)a.J * k) : 0;
          xv[mt][nt][e] = in ? a.X[idx] : 0.0;
          mk[mt][nt][e] = in ? (uint8_t)(a.mask[idx] != 0) : (uint8_t)2;   // any non-zero byte = observed; 2 = outside
        }
      }
    for (int c = 0; c < nchunk; ++c) {
      const int r0 = c * kRC;
      __syncthreads();   // every warp has finished reading cs / As / Bs of the previous (k, chunk)
      if (nchunk > 1 || k == k0) {
        for (int e = tid; e < kRC * kTile; e += 256) {
          const int ii = e % kTile, rr = e / kTile;
          const int r = r0 + rr;
          As[rr][ii] = (r < a.R && i0 + ii < a.I) ? a.Fi[i0 + ii + (long long)r * a.ldFi] : 0.0;
          Bs[rr][ii] = (r < a.R && j0 + ii < a.J) ? a.Fj[j0 + ii + (long long)r * a.ldFj] : 0.0;
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
    // epilogue: compare the model with the stored elements loaded above / impute
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const long long j = j0 + 32 * wj + 8 * nt + 2 * q + e;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const long long i = i0 + 16 * wi + 8 * mt + p;
          const double x = xv[mt][nt][e], m = acc[mt][nt][e];
          if (mk[mt][nt][e] == 1) {
            s[2] = fma(x, m, s[2]);
            s[3] = fma(m, m, s[3]);
            const double d = x - m;
            s[4] = fma(d, d, s[4]);
          } else if (mk[mt][nt][e] == 0) {
            const double d = m - x;
            s[0] = fma(d, d, s[0]);
            s[1] = fma(x, x, s[1]);
            if (a.impute) a.X[i + a.ldI * (j + (long long)a.J * k)] = m;
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
  const long long ti = ceil_div(a.I, kTile), tj = ceil_div(a.J, kTile);
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
  em_kernel<<<g, 256, 0, st>>>(a, kper);
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
