// linalg.cu - see linalg.cuh
#include "linalg.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace aoadmm {

namespace {

constexpr int kTile = 32;
constexpr int kKStep = 16;

__global__ void __launch_bounds__(256) dgemm_small_kernel(int transA, int transB, long long M, long long N, long long K,
                                                           double alpha, const double* __restrict__ alpha_dev,
                                                           const double* __restrict__ A, long long lda,
                                                           const double* __restrict__ B, long long ldb, double beta,
                                                           double* __restrict__ C, long long ldc,
                                                           const int* __restrict__ skip, long long kchunk,
                                                           double* __restrict__ ws) {
  // ws != nullptr: split-K mode - CTA z reduces k in [z*kchunk, (z+1)*kchunk) and stores its raw partial tile into
  // ws[z] (M x N, ld = M); dgemm_splitk_reduce_kernel adds the partials in z order
  if (skip != nullptr && *skip != 0) return;
  __shared__ double As[kKStep][kTile + 1];
  __shared__ double Bs[kKStep][kTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * kTile, n0 = (long long)blockIdx.y * kTile;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  const long long kb = (ws != nullptr) ? (long long)blockIdx.z * kchunk : 0;
  if (ws != nullptr) K = min(K, kb + kchunk);
  for (long long k0 = kb; k0 < K; k0 += kKStep) {
    for (int e = tid; e < kKStep * kTile; e += 256) {
      int kk, mm;
      if (transA) {  // A stored K x M: consecutive threads walk k (contiguous)
        kk = e % kKStep;
        mm = e / kKStep;
      } else {       // A stored M x K: consecutive threads walk m
        mm = e % kTile;
        kk = e / kTile;
      }
      const long long gm = m0 + mm, gk = k0 + kk;
      double v = 0.0;
      if (gm < M && gk < K) v = transA ? A[gk + gm * lda] : A[gm + gk * lda];
      As[kk][mm] = v;
    }
    for (int e = tid; e < kKStep * kTile; e += 256) {
      int kk, nn;
      if (transB) {  // B stored N x K
        nn = e % kTile;
        kk = e / kTile;
      } else {       // B stored K x N
        kk = e % kKStep;
        nn = e / kKStep;
      }
      const long long gn = n0 + nn, gk = k0 + kk;
      double v = 0.0;
      if (gn < N && gk < K) v = transB ? B[gn + gk * ldb] : B[gk + gn * ldb];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kKStep; ++kk) {
      const double a0 = As[kk][tx], a1 = As[kk][tx + 16];
      const double b0 = Bs[kk][ty], b1 = Bs[kk][ty + 16];
      acc[0][0] = fma(a0, b0, acc[0][0]);
      acc[0][1] = fma(a0, b1, acc[0][1]);
      acc[1][0] = fma(a1, b0, acc[1][0]);
      acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  const double al = alpha * (alpha_dev != nullptr ? *alpha_dev : 1.0);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long gm = m0 + tx + 16 * i, gn = n0 + ty + 16 * j;
      if (gm < M && gn < N && ws != nullptr) {
        ws[(long long)blockIdx.z * M * N + gm + gn * M] = acc[i][j];
      } else if (gm < M && gn < N) {
        double* c = C + gm + gn * ldc;
        *c = (beta != 0.0) ? al * acc[i][j] + beta * (*c) : al * acc[i][j];
      }
    }
}

__global__ void dgemm_splitk_reduce_kernel(const double* __restrict__ ws, int splits, long long M, long long N, double alpha,
                                           double beta, double* __restrict__ C, long long ldc,
                                           const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long mn = M * N;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < mn; e += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int z = 0; z < splits; ++z) v += ws[(long long)z * mn + e];
    double* c = C + e % M + (e / M) * ldc;
    *c = (beta != 0.0) ? alpha * v + beta * (*c) : alpha * v;
  }
}

struct LinArgs {
  const double* x[5];
  double coef[5];
  const double* coef_dev[5];
  int n;
};

__global__ void lincomb_kernel(double* __restrict__ out, long long n, LinArgs a, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  double c[5];
  for (int t = 0; t < a.n; ++t) c[t] = a.coef[t] * (a.coef_dev[t] != nullptr ? *a.coef_dev[t] : 1.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int t = 0; t < a.n; ++t) v += c[t] * a.x[t][i];
    out[i] = v;
  }
}

__global__ void transpose_kernel(const double* __restrict__ in, long long rows, long long cols, double* __restrict__ out,
                                 const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long n = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx % cols, r = idx / cols;  // out index = c + r*cols
    out[idx] = in[r + c * rows];
  }
}

__global__ void __launch_bounds__(512) jacobi_onesided_kernel(double* S, long long m, int n, double* V,
                                                               double* __restrict__ sig, const int* __restrict__ skip,
                                                               int use_smem) {
  if (skip != nullptr && *skip != 0) return;
  __shared__ int s_rot;
  extern __shared__ double jsm[];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  // small problems (the Rayleigh-Ritz / Gram matrices of the eigen-iteration, the coupling matrices H) are iterated in
  // shared memory: every element access of a round is then a 30-cycle instead of a global / L2 round trip
  double* const Sg = S;
  double* const Vg = V;
  if (use_smem) {
    S = jsm;
    V = jsm + m * n;
    for (long long e = tid; e < m * n; e += nt) S[e] = Sg[e];
  }
  for (long long e = tid; e < (long long)n * n; e += nt) V[e] = (e % (n + 1) == 0) ? 1.0 : 0.0;
  __syncthreads();
  const int ne = (n + 1) & ~1, npairs = ne / 2;
  const double tol = 2.220446049250313e-16 * sqrt((double)(m > 1 ? m : 1));
  for (int sweep = 0; sweep < 60 && n > 1; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int rd = 0; rd < ne - 1; ++rd) {
      for (int q = warp; q < npairs; q += nw) {
        int p1, p2;
        if (q == 0) {
          p1 = ne - 1;
          p2 = rd;
        } else {
          p1 = (rd + q) % (ne - 1);
          p2 = (rd - q + ne - 1) % (ne - 1);
        }
        if (p1 > p2) {
          const int t = p1;
          p1 = p2;
          p2 = t;
        }
        if (p2 >= n) continue;
        double* x = S + (long long)p1 * m;
        double* y = S + (long long)p2 * m;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (long long j = lane; j < m; j += 32) {
          const double xv = x[j], yv = y[j];
          al = fma(xv, xv, al);
          be = fma(yv, yv, be);
          ga = fma(xv, yv, ga);
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_sum(ga);
        if (ga * ga > tol * tol * (al * be)) {
          // tan(theta) = 2 ga / (tau + sign(tau) sqrt(tau^2 + 4 ga^2)), tau = be - al - the same rotation as
          // sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = tau / (2 ga), without the two divisions and two
          // square roots whose latency was most of a round (x * rsqrt(x), one reciprocal; see par2.cu)
          const double tau = be - al, g2 = 2.0 * ga;
          const double x2 = fma(tau, tau, g2 * g2);
          const double t = g2 * __drcp_rn(tau + copysign(x2 * rsqrt(x2), tau));
          const double cs = rsqrt(fma(t, t, 1.0)), sn = cs * t;
          for (long long j = lane; j < m; j += 32) {
            const double xv = x[j], yv = y[j];
            x[j] = cs * xv - sn * yv;
            y[j] = sn * xv + cs * yv;
          }
          double* vx = V + (long long)p1 * n;
          double* vy = V + (long long)p2 * n;
          for (int i = lane; i < n; i += 32) {
            const double xv = vx[i], yv = vy[i];
            vx[i] = cs * xv - sn * yv;
            vy[i] = sn * xv + cs * yv;
          }
          if (lane == 0) s_rot = 1;
        }
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (rot == 0) break;
  }
  for (int c = warp; c < n; c += nw) {
    const double* x = S + (long long)c * m;
    double al = 0.0;
    for (long long j = lane; j < m; j += 32) al = fma(x[j], x[j], al);
    al = warp_sum(al);
    if (lane == 0) sig[c] = sqrt(al);
  }
  if (use_smem) {
    __syncthreads();
    for (long long e = tid; e < m * n; e += nt) Sg[e] = S[e];
    for (long long e = tid; e < (long long)n * n; e += nt) Vg[e] = V[e];
  }
}

__global__ void sylvester_scale_kernel(double* __restrict__ X, const double* __restrict__ At, long long rows, int cols,
                                       const double* __restrict__ lam, double shift, const double* __restrict__ mu,
                                       const double* __restrict__ rho_dev, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const double half = *rho_dev / 2.0;
  const long long n = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % rows, j = idx / rows;
    X[idx] = At[idx] / (half * (lam[i] + shift) + mu[j]);
  }
}

__global__ void scale_cols_inv_kernel(double* __restrict__ S, long long rows, int cols, const double* __restrict__ sig,
                                      const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long n = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const double s = sig[idx / rows];
    S[idx] = (s > 0.0) ? S[idx] / s : 0.0;
  }
}

__global__ void quad_scale_rows_kernel(double* __restrict__ Y, long long rows, int cols, const double* __restrict__ lam,
                                       double eta, const double* __restrict__ rho_dev, double rho_host,
                                       const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const double rho = (rho_dev != nullptr) ? *rho_dev : rho_host;
  const double g = 2.0 * eta / rho;
  const long long n = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x)
    Y[idx] = Y[idx] / (g * lam[idx % rows] + 1.0);
}

__global__ void sum_recip_kernel(double* __restrict__ out, LinArgs a) {
  double s = 0.0;
  for (int t = 0; t < a.n; ++t) s += a.coef[t] * (*a.coef_dev[t]);
  out[0] = s;
  out[1] = 1.0 / s;
}

__global__ void lin_finalize_kernel(LinFin fin, const double* __restrict__ red, InnerTol tol, InnerCtl* ctl) {
  if (ctl->done != 0) return;
  double rpk = 0.0, rdk = 0.0, rpc = 0.0, rdc = 0.0;
  int nc = 0;
  for (int i = 0; i < fin.nmodes; ++i) {
    const LinFinMode& f = fin.m[i];
    rpk += sqrt(red[f.i_pr_num]) / sqrt(red[f.i_pr_den]);
    const double sc = sqrt(red[f.i_mu]), dn = sqrt(red[f.i_du_num]);
    rdk += (sc > 0.0) ? dn / sc : dn;
    if (f.constrained) {
      ++nc;
      rpc += sqrt(red[f.i_fz]) / sqrt(red[f.i_fn]);
      const double sz = sqrt(red[f.i_muz]), zz = sqrt(red[f.i_zz]);
      rdc += (sz > 0.0) ? zz / sz : zz;
    }
  }
  rpk /= (double)fin.nmodes;
  rdk /= (double)fin.nmodes;
  if (nc > 0) {
    rpc /= (double)nc;
    rdc /= (double)nc;
  }
  ctl->res[0] = rpk;
  ctl->res[1] = rdk;
  ctl->res[2] = rpc;
  ctl->res[3] = rdc;
  ctl->iters += 1;
  const bool cont = (rpk > tol.pr_coupl) || (rpc > tol.pr_constr) || (rdk > tol.du_coupl) || (rdc > tol.du_constr);
  if (!cont) ctl->done = 1;
  // a NaN ratio (0/0 of a factor driven to zero) makes its comparison false and an Inf keeps the loop going, exactly
  // as in the reference's while-test (:600, :633, :519); the run goes on and the event is only recorded
  if (!isfinite(rpk + rdk + rpc + rdc)) ctl->warn = 4;
}

unsigned flat_grid(long long n) { return (unsigned)std::min<long long>(ceil_div(std::max<long long>(n, 1), 256), 148 * 8); }

}  // namespace

int dgemm_small(int transA, int transB, long long M, long long N, long long K, double alpha, const double* alpha_dev,
                const double* A, long long lda, const double* B, long long ldb, double beta, double* C, long long ldc,
                cudaStream_t st, const int* skip) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((unsigned)ceil_div(M, kTile), (unsigned)ceil_div(N, kTile));
  dgemm_small_kernel<<<grid, 256, 0, st>>>(transA, transB, M, N, K, alpha, alpha_dev, A, lda, B, ldb, beta, C, ldc, skip, 0,
                                           nullptr);
  AO_CHECK_LAUNCH();
  return 1;
}

int dgemm_splitk_count(long long M, long long N, long long K) {
  // enough CTAs for the whole GPU, at least 512 reduction steps each
  const long long tiles = ceil_div(M, kTile) * ceil_div(N, kTile);
  long long s = std::min<long long>(ceil_div(2 * 148, tiles), ceil_div(K, 512));
  return (int)std::max<long long>(1, std::min<long long>(s, 256));
}

int dgemm_small_splitk(int transA, int transB, long long M, long long N, long long K, double alpha, const double* A,
                       long long lda, const double* B, long long ldb, double beta, double* C, long long ldc, double* ws,
                       cudaStream_t st, const int* skip) {
  if (M <= 0 || N <= 0) return 0;
  const int splits = dgemm_splitk_count(M, N, K);
  if (splits <= 1 || ws == nullptr)
    return dgemm_small(transA, transB, M, N, K, alpha, nullptr, A, lda, B, ldb, beta, C, ldc, st, skip);
  const long long kchunk = ceil_div(ceil_div(K, splits), kKStep) * kKStep;
  dim3 grid((unsigned)ceil_div(M, kTile), (unsigned)ceil_div(N, kTile), (unsigned)ceil_div(K, kchunk));
  dgemm_small_kernel<<<grid, 256, 0, st>>>(transA, transB, M, N, K, 1.0, nullptr, A, lda, B, ldb, 0.0, C, ldc, skip, kchunk, ws);
  AO_CHECK_LAUNCH();
  dgemm_splitk_reduce_kernel<<<flat_grid(M * N), 256, 0, st>>>(ws, (int)grid.z, M, N, alpha, beta, C, ldc, skip);
  AO_CHECK_LAUNCH();
  return 2;
}

int lincomb(double* out, long long n, const LinTerm* terms, int nterms, cudaStream_t st, const int* skip) {
  if (nterms < 1 || nterms > 5) throw CudaError(1, "lincomb: 1..5 terms");
  LinArgs a{};
  a.n = nterms;
  for (int t = 0; t < nterms; ++t) {
    a.x[t] = terms[t].x;
    a.coef[t] = terms[t].coef;
    a.coef_dev[t] = terms[t].coef_dev;
  }
  lincomb_kernel<<<flat_grid(n), 256, 0, st>>>(out, n, a, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int transpose_small(const double* in, long long rows, long long cols, double* out, cudaStream_t st, const int* skip) {
  transpose_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(in, rows, cols, out, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int scale_cols_inv(double* S, long long rows, int cols, const double* sig, cudaStream_t st, const int* skip) {
  scale_cols_inv_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(S, rows, cols, sig, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int quad_scale_rows(double* Y, long long rows, int cols, const double* lam, double eta, const double* rho_dev,
                    double rho_host, cudaStream_t st, const int* skip) {
  quad_scale_rows_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(Y, rows, cols, lam, eta, rho_dev, rho_host, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int jacobi_onesided(double* S, long long m, int n, double* V, double* sig, cudaStream_t st, const int* skip) {
  const size_t bytes = ((size_t)m * n + (size_t)n * n) * sizeof(double);
  const int use_smem = (bytes <= 200 * 1024) ? 1 : 0;
  if (use_smem) ensure_dynamic_smem(reinterpret_cast<const void*>(jacobi_onesided_kernel), bytes);
  jacobi_onesided_kernel<<<1, 512, use_smem ? bytes : 0, st>>>(S, m, n, V, sig, skip, use_smem);
  AO_CHECK_LAUNCH();
  return 1;
}

int sylvester_scale(double* X, const double* At, long long rows, int cols, const double* lam, double shift,
                    const double* mu, const double* rho_dev, cudaStream_t st, const int* skip) {
  sylvester_scale_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(X, At, rows, cols, lam, shift, mu, rho_dev, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

void quad_prox_setup(QuadProx& q, const double* L_host, long long n, double eta, int maxcols, cudaStream_t st) {
  if (L_host == nullptr || n <= 0) throw CudaError(1, "quadratic regularization: matrix L is required");
  if (n > 4096) throw CudaError(2, "quadratic regularization: at most 4096 rows on device");
  double amax = 0.0, asym = 0.0;
  for (long long j = 0; j < n; ++j)
    for (long long i = 0; i < n; ++i) {
      amax = std::max(amax, std::fabs(L_host[i + j * n]));
      asym = std::max(asym, std::fabs(L_host[i + j * n] - L_host[j + i * n]));
    }
  if (asym > 1e-12 * std::max(amax, 1e-300)) throw CudaError(2, "quadratic regularization: L must be symmetric on device");
  q.n = n;
  q.eta = eta;
  q.maxcols = maxcols;
  const size_t nn = (size_t)n * n;
  AO_CUDA(cudaMalloc(&q.L, nn * sizeof(double)));
  AO_CUDA(cudaMalloc(&q.Q, nn * sizeof(double)));
  AO_CUDA(cudaMalloc(&q.lam, (size_t)n * sizeof(double)));
  AO_CUDA(cudaMalloc(&q.tmp, std::max(nn, (size_t)n * maxcols) * sizeof(double)));
  AO_CUDA(cudaMemcpy(q.L, L_host, nn * sizeof(double), cudaMemcpyHostToDevice));
  // eigenvectors: one-sided Jacobi on (a copy of) the symmetric L orthogonalises the columns of L*Q
  AO_CUDA(cudaMemcpy(q.tmp, L_host, nn * sizeof(double), cudaMemcpyHostToDevice));
  jacobi_onesided(q.tmp, n, (int)n, q.Q, q.lam, st, nullptr);
  // eigenvalues with their sign: Rayleigh quotients q_i' L q_i  (q.tmp now holds L*Q)
  std::vector<double> Qh(nn), Th(nn), lam((size_t)n);
  AO_CUDA(cudaMemcpyAsync(Qh.data(), q.Q, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
  AO_CUDA(cudaMemcpyAsync(Th.data(), q.tmp, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
  AO_CUDA(cudaStreamSynchronize(st));
  for (long long i = 0; i < n; ++i) {
    double s = 0.0;
    for (long long j = 0; j < n; ++j) s += Qh[(size_t)(j + i * n)] * Th[(size_t)(j + i * n)];
    lam[(size_t)i] = s;
  }
  AO_CUDA(cudaMemcpy(q.lam, lam.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
}

void quad_prox_free(QuadProx& q) {
  for (double** p : {&q.L, &q.Q, &q.lam, &q.tmp}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
}

int quad_prox_apply(const QuadProx& q, const double* X, long long ldx, double* out, long long ldo, int cols,
                    const double* rho_dev, double rho_host, cudaStream_t st, const int* skip) {
  if (cols > q.maxcols) throw CudaError(1, "quadratic regularization: too many columns");
  int n = dgemm_small(1, 0, q.n, cols, q.n, 1.0, nullptr, q.Q, q.n, X, ldx, 0.0, q.tmp, q.n, st, skip);
  n += quad_scale_rows(q.tmp, q.n, cols, q.lam, q.eta, rho_dev, rho_host, st, skip);
  n += dgemm_small(0, 0, q.n, cols, q.n, 1.0, nullptr, q.Q, q.n, q.tmp, q.n, 0.0, out, ldo, st, skip);
  return n;
}

int sum_recip(double* out, const LinTerm* terms, int nterms, cudaStream_t st) {
  if (nterms < 1 || nterms > 5) throw CudaError(2, "more than 5 modes in a linearly coupled group");
  LinArgs a{};
  a.n = nterms;
  for (int t = 0; t < nterms; ++t) {
    if (terms[t].coef_dev == nullptr) throw CudaError(1, "sum_recip: device scalar missing");
    a.coef[t] = terms[t].coef;
    a.coef_dev[t] = terms[t].coef_dev;
  }
  sum_recip_kernel<<<1, 1, 0, st>>>(out, a);
  AO_CHECK_LAUNCH();
  return 1;
}

int lin_finalize(const LinFin& fin, const double* red, const InnerTol& tol, InnerCtl* ctl, cudaStream_t st) {
  lin_finalize_kernel<<<1, 1, 0, st>>>(fin, red, tol, ctl);
  AO_CHECK_LAUNCH();
  return 1;
}

}  // namespace aoadmm
