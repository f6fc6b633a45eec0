// em.cu - EM imputation of missing entries and the masked objective terms.
//
// Reference: functions/cmtf_fun_AOADMM.m
//   :408-441   after every sweep the missing entries (Z.miss == 0) of each object are overwritten by the current
//              low-rank model; f_rel_missing = sqrt(sum (new-old)^2 / sum old^2)
//   :1224-1226 objective of a CP object with a mask:   w (Znorm - 2 <X, M.*model> + ||M.*model||^2)
//   :1249-1252 objective of a PARAFAC2 object with a mask: w sum_k || M_k .* (X_k - A D_k B_k') ||^2
// One pass over the object does all of it: the model value of every element is a rank-R product
//     m(i,j,k) = sum_r Fi(i,r) Fj(j,r) Fk(k,r)
// evaluated by 64 x 64 output tiles (4 x 4 outputs per thread, factor tiles staged in shared memory), compared with
// the stored element and the mask, and five sums are reduced deterministically (per-CTA partials, fixed-order final
// sum):  [0] sum_missing (m-x)^2   [1] sum_missing x^2   [2] sum_observed x*m   [3] sum_observed m^2
//        [4] sum_observed (x-m)^2
// A PARAFAC2 object is the K = 1 case on the stacked I x Jtot matrix with Fj(j,:) = B(j,:) .* C(seg(j),:).
#include "em.cuh"

#include <algorithm>

namespace aoadmm {

namespace {

constexpr int kTile = 64;
constexpr int kRC = 32;  // rank chunk staged per pass

__global__ void __launch_bounds__(256) em_kernel(EmArgs a, int kper) {
  __shared__ double As[kRC][kTile];
  __shared__ double Bs[kRC][kTile];
  __shared__ double red[32];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long i0 = (long long)blockIdx.x * kTile, j0 = (long long)blockIdx.y * kTile;
  const int k0 = blockIdx.z * kper, k1 = min(a.K, k0 + kper);
  const int nchunk = (a.R + kRC - 1) / kRC;
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = k0; k < k1; ++k) {
    double acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      const int r0 = c * kRC;
      __syncthreads();
      for (int e = tid; e < kRC * kTile; e += 256) {
        const int ii = e % kTile, rr = e / kTile;
        const int r = r0 + rr;
        double va = 0.0;
        if (r < a.R && i0 + ii < a.I) {
          va = a.Fi[i0 + ii + (long long)r * a.ldFi];
          if (a.Fk != nullptr) va *= a.Fk[k + (long long)r * a.ldFk];
        }
        As[rr][ii] = va;
      }
      if (nchunk > 1 || k == k0) {
        for (int e = tid; e < kRC * kTile; e += 256) {
          const int jj = e % kTile, rr = e / kTile;
          const int r = r0 + rr;
          Bs[rr][jj] = (r < a.R && j0 + jj < a.J) ? a.Fj[j0 + jj + (long long)r * a.ldFj] : 0.0;
        }
      }
      __syncthreads();
#pragma unroll 8
      for (int rr = 0; rr < kRC; ++rr) {
        double av[4], bv[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) av[p] = As[rr][tx + 16 * p];
#pragma unroll
        for (int q = 0; q < 4; ++q) bv[q] = Bs[rr][ty + 16 * q];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fma(av[p], bv[q], acc[p][q]);
      }
    }
    // epilogue: all loads of the tile first (independent, so their DRAM latencies overlap), then compare / impute
    double xv[4][4];
    uint8_t mk[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long j = j0 + ty + 16 * q;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const long long i = i0 + tx + 16 * p;
        const bool in = (i < a.I) && (j < a.J);
        const long long idx = in ? i + a.ldI * (j + (long long)a.J * k) : 0;
        xv[p][q] = in ? a.X[idx] : 0.0;
        mk[p][q] = in ? (uint8_t)(a.mask[idx] != 0) : (uint8_t)2;   // any non-zero byte = observed (as in norm2_masked_kernel); 2 = outside the object
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long j = j0 + ty + 16 * q;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const long long i = i0 + tx + 16 * p;
        const double x = xv[p][q], m = acc[p][q];
        if (mk[p][q] == 1) {
          s[2] = fma(x, m, s[2]);
          s[3] = fma(m, m, s[3]);
          const double d = x - m;
          s[4] = fma(d, d, s[4]);
        } else if (mk[p][q] == 0) {
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
