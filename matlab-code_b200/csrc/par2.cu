// par2.cu - PARAFAC2 block kernels (see par2.cuh for the layout and the reference lines each entry point replaces).
#include "par2.cuh"

#include <algorithm>
#include <map>

namespace aoadmm {

namespace {

constexpr int kP2Threads = 256;

// ---------------------------------------------------------------------------------------------------------
// small dense helpers on an R x R column-major matrix held in shared memory (whole CTA cooperates)
// ---------------------------------------------------------------------------------------------------------
// right-looking Cholesky of the lower triangle of W; returns false (uniformly) when a pivot is not positive
__device__ bool cta_cholesky(double* W, int R) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = 0; j < R; ++j) {
    const double d = W[j * R + j];
    if (!(d > 0.0) || !isfinite(d)) return false;  // every thread reads the same value
    const double s = sqrt(d);
    __syncthreads();
    for (int i = j + 1 + tid; i < R; i += nt) W[j * R + i] = W[j * R + i] / s;
    if (tid == 0) W[j * R + j] = s;
    __syncthreads();
    const int n = R - j - 1;
    for (int e = tid; e < n * n; e += nt) {
      const int ci = e / n, ri = e % n;
      if (ri >= ci) W[(j + 1 + ci) * R + (j + 1 + ri)] -= W[j * R + (j + 1 + ri)] * W[j * R + (j + 1 + ci)];
    }
    __syncthreads();
  }
  return true;
}

// V (row i, column c at V[i*R + c]) = inv(L) for the lower factor held in W; then out = inv(L)' * inv(L)
__device__ void cta_inverse_from_chol(const double* W, double* V, int R, double* out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // forward substitution on all columns at once (see prep_system_kernel): same FMAs in the same order per entry as a
  // column-by-column substitution, spread over the whole CTA
  for (int e = tid; e < R * R; e += nt) V[e] = (e / R == e % R) ? 1.0 : 0.0;
  __syncthreads();
  for (int j = 0; j < R; ++j) {
    const double d = W[j * R + j];
    for (int c = tid; c <= j; c += nt) V[j * R + c] = V[j * R + c] / d;
    __syncthreads();
    const int ncol = j + 1, nent = (R - j - 1) * ncol;
    for (int e = tid; e < nent; e += nt) {
      const int i = j + 1 + e / ncol, c = e % ncol;
      V[i * R + c] = fma(-W[j * R + i], V[j * R + c], V[i * R + c]);
    }
    __syncthreads();
  }
  for (int e = tid; e < R * R; e += nt) {
    const int ca = e / R, cb = e % R;
    const int lo = ca > cb ? ca : cb;
    double acc = 0.0;
    for (int i = lo; i < R; ++i) acc = fma(V[i * R + ca], V[i * R + cb], acc);
    out[e] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void par2_scale_rows_kernel(Par2Layout L, double* __restrict__ out, const double* __restrict__ in,
                                       const double* __restrict__ C, long long ldc, double scale,
                                       const double* __restrict__ addend, double add_scale) {
  const long long nloc = L.jhi - L.jlo, n = nloc * L.R;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / nloc, j = L.jlo + e % nloc;
    const long long idx = j + r * L.Jtot;
    double v = scale * in[idx] * C[L.seg[j] + r * ldc];
    if (addend != nullptr) v += add_scale * addend[idx];
    out[idx] = v;
  }
}

__global__ void __launch_bounds__(kP2Threads) par2_batched_gram_kernel(Par2Layout L, const double* __restrict__ Bst,
                                                                        double* __restrict__ G2) {
  const int k = blockIdx.x + L.k0, R = L.R;
  const long long j0 = L.joff[k];
  const int Jk = (int)(L.joff[k + 1] - j0);
  double* out = G2 + (size_t)k * R * R;
  for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
    const int a = e % R, b = e / R;
    if (a > b) continue;
    const double* xa = Bst + j0 + (long long)a * L.Jtot;
    const double* xb = Bst + j0 + (long long)b * L.Jtot;
    double acc = 0.0;
    for (int j = 0; j < Jk; ++j) acc = fma(xa[j], xb[j], acc);
    out[a + b * R] = acc;
    out[b + a * R] = acc;
  }
}

// one warp per element of the R x R result: the lanes stride over the slices (fixed order), then a butterfly sum - the
// K-long sum is latency bound when one thread walks it (138 us at K = 512), a warp finishes it in a few microseconds
__global__ void __launch_bounds__(256) par2_modeA_had_kernel(Par2Layout L, const double* __restrict__ G2,
                                                             const double* __restrict__ C, long long ldc,
                                                             double* __restrict__ Csum) {
  const int R = L.R, lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= R * R) return;   // whole warps leave together
  const int a = e % R, b = e / R;
  const double* ca = C + a * ldc;
  const double* cb = C + b * ldc;
  const double* g = G2 + e;
  const size_t RR = (size_t)R * R;
  double acc = 0.0;
  for (int k = L.k0 + lane; k < L.k1; k += 32) acc += (ca[k] * g[(size_t)k * RR]) * cb[k];
  acc = warp_sum(acc);
  if (lane == 0) Csum[e] = acc;
}

__global__ void __launch_bounds__(128) par2_sys_prep_kernel(Par2Layout L, Par2SysArgs a) {
  extern __shared__ double sm[];
  __shared__ double red[32];
  __shared__ double s_rho;
  const int k = blockIdx.x + L.k0, R = L.R, RR = R * R, tid = threadIdx.x, nt = blockDim.x;
  // R <= 64: everything in shared memory; larger ranks factor in the caller's global workspace (2 R^2 per slice)
  const bool big = R > 64;
  double* W = big ? a.gws + (size_t)k * 2 * RR : sm;   // system matrix / Cholesky factor
  double* V = W + RR;                                   // inverse of the factor
  double* rhs_s = big ? sm : sm + 2 * RR;               // R
  const long long j0 = L.joff[k];
  const int Jk = (int)(L.joff[k + 1] - j0);
  if (k == L.k0 && tid == 0 && a.ctl != nullptr) {  // a new inner loop starts (err is sticky)
    a.ctl->done = 0;
    a.ctl->iters = 0;
    a.ctl->res[0] = a.ctl->res[1] = a.ctl->res[2] = a.ctl->res[3] = 0.0;
  }
  double tr = 0.0;
  for (int e = tid; e < RR; e += nt) {
    const int r = e % R, c = e / R;
    double v;
    if (a.mode == 2)
      v = (a.C[k + r * a.ldc] * a.G1[e]) * a.C[k + c * a.ldc];   // diag(c_k) A'A diag(c_k)   (:194)
    else
      v = a.G1[e] * a.G2[(size_t)k * RR + e];                    // A'A .* B_k'B_k            (:222)
    V[e] = v;  // keep C_k for the system assembly
    if (r == c) tr += v;
  }
  tr = block_sum(tr, red);
  if (tid == 0) {
    s_rho = tr / (double)R * a.rho_scale;                         // :195-198, :223
    a.rho_k[k] = s_rho;
  }
  __syncthreads();
  const double half = s_rho / 2.0;
  for (int e = tid; e < RR; e += nt) {
    double b = a.weight * V[e];
    if (e % (R + 1) == 0) {
      for (int q = 0; q < a.n_rho_terms; ++q) b += half;
      if (a.ridge != 0.0) b += a.ridge;
      if (a.bsum_half != 0.0) b += a.bsum_half;
    }
    if (a.HHt != nullptr) b += half * a.HHt[e];
    W[e] = b;
    if (a.Bsys != nullptr) a.Bsys[(size_t)k * RR + e] = b;
  }
  if (a.mode == 3) {
    // a_k(r) = w * sum_j T(j,r) B_k(j,r)   (= w * diag(A' X_k B_k), :221)  [+ bsum/2 * C(k,r), :231]
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int r = warp; r < R; r += nw) {
      const double* t = a.T + j0 + (long long)r * L.Jtot;
      const double* b = a.Bst + j0 + (long long)r * L.Jtot;
      double acc = 0.0;
      for (int j = lane; j < Jk; j += 32) acc = fma(t[j], b[j], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        double v = a.weight * acc;
        if (a.bsum_half != 0.0) v += a.bsum_half * a.C[k + r * a.ldc];
        rhs_s[r] = v;
        a.rhs[k + (long long)r * L.K] = v;
      }
    }
  }
  __syncthreads();
  if (a.no_factor) return;
  const bool ok = cta_cholesky(W, R);
  if (!ok) {
    if (tid == 0 && a.ctl != nullptr) a.ctl->err = 3;
    return;
  }
  if (a.mode == 3 && a.ls_direct) {
    // C(k,:) = (B_k \ a_k)'  (:236): forward / backward substitution with the Cholesky factor
    if (tid == 0) {
      for (int i = 0; i < R; ++i) {
        double s = rhs_s[i];
        for (int q = 0; q < i; ++q) s = fma(-W[q * R + i], rhs_s[q], s);
        rhs_s[i] = s / W[i * R + i];
      }
      for (int i = R - 1; i >= 0; --i) {
        double s = rhs_s[i];
        for (int q = i + 1; q < R; ++q) s = fma(-W[i * R + q], rhs_s[q], s);
        rhs_s[i] = s / W[i * R + i];
      }
      for (int r = 0; r < R; ++r) a.fac_out[k + (long long)r * L.K] = rhs_s[r];
    }
    return;
  }
  cta_inverse_from_chol(W, V, R, a.Binv + (size_t)k * RR);
}

// ---- row-wise helpers for the third PARAFAC2 mode inside linear couplings 2..4 ------------------------------------
__global__ void par2_rows_ainner_kernel(double* __restrict__ out, const double* __restrict__ A, const double* __restrict__ X,
                                        const double* __restrict__ Z, const double* __restrict__ muZ,
                                        const double* __restrict__ rho_k, long long rows, int cols,
                                        const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long n = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const double half = rho_k[e % rows] / 2.0;
    double v = A[e] + half * X[e];
    if (Z != nullptr) v += half * (Z[e] - muZ[e]);
    out[e] = v;
  }
}

__global__ void par2_rows_apply_kernel(double* __restrict__ out, const double* __restrict__ in,
                                       const double* __restrict__ Minv, long long rows, int q,
                                       const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long n = rows * q;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e % rows;
    const int c = (int)(e / rows);
    const double* M = Minv + (size_t)k * q * q + (size_t)c * q;   // column c of M_k
    double acc = 0.0;
    for (int j = 0; j < q; ++j) acc = fma(in[k + rows * j], M[j], acc);
    out[e] = acc;
  }
}

__global__ void par2_rows_scale_kernel(double* __restrict__ out, const double* __restrict__ in,
                                       const double* __restrict__ rho_k, long long rows, int cols,
                                       const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long n = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    out[e] = rho_k[e % rows] * in[e];
}

__global__ void par2_rows_weighted_accum_kernel(double* __restrict__ D, double* __restrict__ wsum,
                                                const double* __restrict__ S, const double* __restrict__ mu,
                                                const double* __restrict__ rho_scalar, const double* __restrict__ rho_k,
                                                long long rows, int cols, int first, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long n = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e % rows;
    const double w = (rho_k != nullptr) ? rho_k[k] : *rho_scalar;
    const double v = w * (S[e] + mu[e]);
    D[e] = first ? v : D[e] + v;
    if (e < rows) wsum[k] = first ? w : wsum[k] + w;
  }
}

__global__ void par2_rows_divide_kernel(double* __restrict__ D, const double* __restrict__ wsum, long long rows, int cols,
                                        const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long n = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    D[e] = (1.0 / wsum[e % rows]) * D[e];
}

__global__ void __launch_bounds__(128) par2_rowsys_inverse_kernel(const double* __restrict__ AA, const double* __restrict__ AAA,
                                                                  const double* __restrict__ rho_k, int q,
                                                                  double* __restrict__ Minv, InnerCtl* ctl) {
  extern __shared__ double sm[];
  const int k = blockIdx.x, qq = q * q;
  double* W = sm;
  double* V = sm + qq;
  const double rk = rho_k[k];
  for (int e = threadIdx.x; e < qq; e += blockDim.x) W[e] = AA[e] + rk * AAA[e];
  __syncthreads();
  if (!cta_cholesky(W, q)) {
    if (threadIdx.x == 0 && ctl != nullptr) ctl->err = 3;
    return;
  }
  cta_inverse_from_chol(W, V, q, Minv + (size_t)k * qq);
}

__global__ void par2_rho_stats_kernel(const double* __restrict__ rho_k, int K, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) s += rho_k[k];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    out[0] = s / (double)K;
    out[1] = s;
  }
}

__global__ void par2_assemble_B2_kernel(const double* __restrict__ Bsys, const double* __restrict__ HtH,
                                        const double* __restrict__ rhoC_dev, int constrained, int K, int R,
                                        double* __restrict__ B2) {
  const long long n = (long long)K * R;
  const double half = *rhoC_dev / 2.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e % n, col = e / n;
    const int k = (int)(row / R), r = (int)(row % R), k2 = (int)(col / R), s = (int)(col % R);
    double v = 0.0;
    if (k == k2) v = Bsys[(size_t)k * R * R + r + (size_t)s * R];
    if (r == s) v += half * HtH[k + (size_t)k2 * K];
    if (constrained && row == col) v += half;
    B2[e] = v;
  }
}

// forward / backward substitution with the lower Cholesky factor L (n x n, column-major) for ONE right-hand side
__global__ void __launch_bounds__(256) par2_chol_solve_vec_kernel(const double* __restrict__ L, int K, int R,
                                                                   const double* __restrict__ a, double* __restrict__ x,
                                                                   const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  extern __shared__ double xs[];
  const int n = K * R, tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n; i += nt) xs[i] = a[(i / R) + (size_t)(i % R) * K];
  __syncthreads();
  for (int j = 0; j < n; ++j) {            // L y = a
    const double yj = xs[j] / L[j + (size_t)j * n];
    __syncthreads();
    if (tid == 0) xs[j] = yj;
    for (int i = j + 1 + tid; i < n; i += nt) xs[i] = fma(-L[i + (size_t)j * n], yj, xs[i]);
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; --j) {       // L' x = y
    const double xj = xs[j] / L[j + (size_t)j * n];
    __syncthreads();
    if (tid == 0) xs[j] = xj;
    for (int i = tid; i < j; i += nt) xs[i] = fma(-L[j + (size_t)i * n], xj, xs[i]);
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) x[(i / R) + (size_t)(i % R) * K] = xs[i];
}

__global__ void par2_rho_max_kernel(const double* __restrict__ rho_k, int K, double* __restrict__ out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) m = fmax(m, rho_k[k]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) r = fmax(r, red[w]);
    *out = r;
  }
}

// ---------------------------------------------------------------------------------------------------------
// ADMM_B_Parafac2, step 1 (:525-535): per slice  B_k = A_inner_k inv(Bsys_k);  P_k = polar((B_k+mu_k) DeltaB')
// The polar factor U V' of the thin SVD is computed with a one-sided (Hestenes) Jacobi iteration on the J_k x R
// matrix, round-robin pair ordering, one warp per column pair.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kP2Threads) par2_B_step1_kernel(Par2Layout L, Par2BArgs a, const InnerCtl* ctl,
                                                                   int use_gmem, int warm) {
  if (ctl != nullptr && ctl->done != 0) return;
  extern __shared__ double sm[];
  __shared__ int s_rot;
  const int k = blockIdx.x + L.k0, R = L.R, RR = R * R, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const long long j0 = L.joff[k], ld = L.Jtot;
  const int Jk = (int)(L.joff[k + 1] - j0);
  // R <= 64: DeltaB, inv(Bsys_k) (later DeltaB' W and 1/sigma) and V in shared memory.  Larger ranks: DeltaB and the
  // inverse are read from global memory, DeltaB' W / 1/sigma live in this slice's part of `contrib` (written with its
  // result only at the very end) and V is iterated in place in `Vprev`.
  const bool bigR = R > 64;
  double* nrm = sm;   // squared column norms
  double* base = sm + ((R + 1) & ~1);
  const double* dB = bigR ? a.DeltaB : base;
  double* Bi = bigR ? a.contrib + (size_t)k * RR : base + RR;
  const double* Binv_ro = bigR ? a.Binv + (size_t)k * RR : Bi;
  double* V = bigR ? a.Vprev + (size_t)k * RR : base + 2 * RR;
  double* M = use_gmem ? a.gM + j0 * R : base + (bigR ? 0 : 3 * RR);
  double* S = use_gmem ? a.gS + j0 * R : M + (size_t)L.Jmax * R;
  const double rho = a.rho_k[k], half = rho / 2.0;
  for (int e = tid; e < RR; e += nt) {
    if (!bigR) {
      base[e] = a.DeltaB[e];
      Bi[e] = a.Binv[(size_t)k * RR + e];
    }
    if (!bigR || !warm) V[e] = (e % (R + 1) == 0) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int nitems = Jk * R;
  // A_inner = A_k + rho_k/2 (P_k DeltaB - mu_k) [+ rho_k/2 (Z_k - muZ_k)]          (:526-529)
  for (int it = tid; it < nitems; it += nt) {
    const int c = it / Jk, j = it % Jk;
    const long long g = j0 + j + (long long)c * ld;
    double pd = 0.0;
    for (int r = 0; r < R; ++r) pd = fma(a.P[j0 + j + (long long)r * ld], dB[r + c * R], pd);
    a.PDold[g] = pd;
    double v = a.A[g] + half * (pd - a.mu[g]);
    if (a.con_active) v += half * (a.Z[g] - a.muZ[g]);
    S[j + c * Jk] = v;
  }
  __syncthreads();
  // B_k = A_inner inv(Bsys_k)   (:530) ;  M = B_k + mu_k
  for (int it = tid; it < nitems; it += nt) {
    const int c = it / Jk, j = it % Jk;
    const long long g = j0 + j + (long long)c * ld;
    double x = 0.0;
    for (int r = 0; r < R; ++r) x = fma(S[j + r * Jk], Binv_ro[r + c * R], x);
    a.B[g] = x;
    M[j + c * Jk] = x + a.mu[g];
  }
  __syncthreads();
  // S = M DeltaB' W  (:532) with W = I on a cold start, else the rotations found for this slice in the previous inner
  // iteration (the matrix changes little between inner iterations, so the Jacobi iteration below then needs 1-2 sweeps
  // instead of 6-8; the polar factor it converges to is the same).  Wm = DeltaB' W is formed in the Binv area.
  double* Wm = Bi;
  if (warm) {
    if (!bigR)
      for (int e = tid; e < RR; e += nt) V[e] = a.Vprev[(size_t)k * RR + e];
    __syncthreads();
    for (int e = tid; e < RR; e += nt) {
      const int r = e % R, c = e / R;
      double x = 0.0;
      for (int q = 0; q < R; ++q) x = fma(dB[q + r * R], V[q + c * R], x);
      Wm[e] = x;
    }
  } else {
    for (int e = tid; e < RR; e += nt) Wm[e] = dB[(e / R) + (e % R) * R];  // DeltaB'
  }
  __syncthreads();
  for (int it = tid; it < nitems; it += nt) {
    const int c = it / Jk, j = it % Jk;
    double x = 0.0;
    for (int r = 0; r < R; ++r) x = fma(M[j + r * Jk], Wm[r + c * R], x);
    S[j + c * Jk] = x;
  }
  __syncthreads();
  // one-sided Jacobi: S <- S*V with orthogonal columns.  The round is issue bound (eight warps, one column pair each), so
  // it carries as few instructions as possible: the squared column norms are kept in shared memory, refreshed exactly at
  // the start of every sweep and updated by the rotation identities al' = al - t*ga, be' = be + t*ga in between (one dot
  // product and one warp reduction per pair instead of three); the round-robin partner indices use a conditional subtract
  // instead of a modulo; tan(theta) = 2ga / (tau + sign(tau) sqrt(tau^2 + 4ga^2)) needs one division and one square root.
  const int Re = (R + 1) & ~1, npairs = Re / 2;
  const double tol2 = 2.220446049250313e-16 * 2.220446049250313e-16 * (double)Jk;
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (tid == 0) s_rot = 0;
    for (int r = warp; r < R; r += nw) {
      const double* x = S + (size_t)r * Jk;
      double al = 0.0;
      for (int j = lane; j < Jk; j += 32) al = fma(x[j], x[j], al);
      al = warp_sum(al);
      if (lane == 0) nrm[r] = al;
    }
    __syncthreads();
    for (int rd = 0; rd < Re - 1; ++rd) {
      for (int q = warp; q < npairs; q += nw) {
        int p1, p2;
        if (q == 0) {
          p1 = Re - 1;
          p2 = rd;
        } else {
          p1 = rd + q;                 // < 2 (Re - 1)
          p2 = rd - q + Re - 1;        // in [0, 2 (Re - 1))
          if (p1 >= Re - 1) p1 -= Re - 1;
          if (p2 >= Re - 1) p2 -= Re - 1;
        }
        if (p1 > p2) {
          const int t = p1;
          p1 = p2;
          p2 = t;
        }
        if (p2 >= R) continue;  // bye (odd R)
        double* x = S + (size_t)p1 * Jk;
        double* y = S + (size_t)p2 * Jk;
        const double al = nrm[p1], be = nrm[p2];
        double ga = 0.0;
        for (int j = lane; j < Jk; j += 32) ga = fma(x[j], y[j], ga);
        ga = warp_sum(ga);
        if (ga * ga > tol2 * (al * be)) {
          const double tau = be - al, g2 = 2.0 * ga;
          const double t = g2 / (tau + copysign(sqrt(fma(tau, tau, g2 * g2)), tau));
          const double cs = rsqrt(fma(t, t, 1.0)), sn = cs * t;
          for (int j = lane; j < Jk; j += 32) {
            const double xv = x[j], yv = y[j];
            x[j] = cs * xv - sn * yv;
            y[j] = sn * xv + cs * yv;
          }
          for (int i = lane; i < R; i += 32) {
            const double xv = V[i + p1 * R], yv = V[i + p2 * R];
            V[i + p1 * R] = cs * xv - sn * yv;
            V[i + p2 * R] = sn * xv + cs * yv;
          }
          if (lane == 0) {
            nrm[p1] = al - t * ga;
            nrm[p2] = be + t * ga;
            s_rot = 1;
          }
        }
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (rot == 0) break;
  }
  if (!bigR)
    for (int e = tid; e < RR; e += nt) a.Vprev[(size_t)k * RR + e] = V[e];
  // singular values -> 1/sigma (kept in the Bi area, no longer needed)
  for (int r = warp; r < R; r += nw) {
    const double* x = S + (size_t)r * Jk;
    double al = 0.0;
    for (int j = lane; j < Jk; j += 32) al = fma(x[j], x[j], al);
    al = warp_sum(al);
    if (lane == 0) Bi[r] = (al > 0.0) ? 1.0 / sqrt(al) : 0.0;
  }
  __syncthreads();
  // P_k = U V'   (:534)
  for (int it = tid; it < nitems; it += nt) {
    const int c = it / Jk, j = it % Jk;
    double x = 0.0;
    for (int r = 0; r < R; ++r) x = fma(S[j + r * Jk] * Bi[r], V[c + r * R], x);
    a.P[j0 + j + (long long)c * ld] = x;
  }
  __syncthreads();
  // contribution to DeltaB: rho_k P_k' (B_k + mu_k)   (:541)
  for (int e = tid; e < RR; e += nt) {
    const int ra = e % R, cb = e / R;
    const double* pc = a.P + j0 + (long long)ra * ld;
    const double* mc = M + (size_t)cb * Jk;
    double acc = 0.0;
    for (int j = 0; j < Jk; ++j) acc = fma(pc[j], mc[j], acc);
    a.contrib[(size_t)k * RR + e] = rho * acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// The same step for SMALL slices (R <= 16, J_k <= 128): one WARP per slice, the slice in REGISTERS.
// The CTA kernel above is issue bound on such slices: each of its eight warps executes the scalar rotation arithmetic
// (an FP64 division, square root and reciprocal square root, ~150 instructions) of its own column pair with all 32
// lanes, every column element is a shared-memory round trip, and every round ends in a CTA barrier.  Here
//   * lane l owns the rows l, l+32, .. of the slice: S[JP][RE] doubles in registers (RE = R rounded up to 4 / 8 / 16,
//     the padding columns are zero and never rotate), and row l of the rotation matrix V;
//   * the RE/2 column pairs of a round sit in FIXED register positions (2q, 2q+1); between rounds the columns move one
//     step around the ring of the round-robin tournament (Brent-Luk: position 0 stays, the other RE-1 rotate), a static
//     register permutation, so every register index is a compile-time constant and after a full sweep of RE-1 rounds
//     the columns are back in their own positions;
//   * the RE/2 dot products of a round are reduced with ONE transposing butterfly (RE/2-1 + 5-log2(RE/2) shuffles
//     instead of 5 per pair), which leaves the total of pair q in the lanes of group q; those lanes compute the
//     rotation of their pair - the scalar chain runs once per round - and broadcast (cos, sin) through shared memory;
//   * the rotations of a round are branch-free (cos = 1, sin = 0 is exact) and independent: full ILP, warp barriers only.
// Same rotation formula, tolerance and tracked column norms as the CTA kernel; the pair ORDER within a sweep differs
// (any cyclic ordering converges to the same polar factor).
// ---------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void reduce_transpose(double (&v)[N], int lane) {
  // sums v[i] over the 32 lanes for all i at once; afterwards v[0] of lane l is the total of index transpose_index<N>(l)
  int off = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = up ? v[i] : v[i + half];
      const double keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
template <int N>
__device__ __forceinline__ int transpose_index(int lane) {
  int idx = 0, off = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1, off >>= 1) idx |= (lane & off) ? half : 0;
  return idx;
}
// one step of the tournament ring: t_0 stays, b_0 -> t_1, t_q -> t_{q+1}, t_{n-1} -> b_{n-1}, b_q -> b_{q-1}
// (t_q = position 2q, b_q = position 2q+1)
template <int RE>
__device__ __forceinline__ void ring_step(double (&c)[RE]) {
  constexpr int n = RE / 2;
  double nt[n], nb[n];
  nt[0] = c[0];
  nt[1] = c[1];
#pragma unroll
  for (int q = 2; q < n; ++q) nt[q] = c[2 * (q - 1)];
#pragma unroll
  for (int q = 0; q < n - 1; ++q) nb[q] = c[2 * (q + 1) + 1];
  nb[n - 1] = c[2 * (n - 1)];
#pragma unroll
  for (int q = 0; q < n; ++q) {
    c[2 * q] = nt[q];
    c[2 * q + 1] = nb[q];
  }
}
template <int RE>
__device__ __forceinline__ int ring_newpos(int pos) {
  constexpr int n = RE / 2;
  const int q = pos >> 1;
  if ((pos & 1) == 0) return (q == 0) ? 0 : ((q == n - 1) ? 2 * (n - 1) + 1 : 2 * (q + 1));
  return (q == 0) ? 2 : 2 * (q - 1) + 1;
}

template <int RE>
__host__ __device__ inline size_t par2_step1_reg_doubles(long long Jmax) {
  const long long pitch = Jmax | 1;
  return (size_t)4 * RE * RE + (size_t)2 * pitch * RE + 2 * (RE / 2) + 2 * RE + RE;
}

template <int RE, int JP>
__global__ void __launch_bounds__(32) par2_B_step1_reg_kernel(Par2Layout L, Par2BArgs a, const InnerCtl* ctl, int warm) {
  constexpr int NP = RE / 2;
  if (ctl != nullptr && ctl->done != 0) return;
  extern __shared__ __align__(16) double sm[];
  const int k = blockIdx.x + L.k0, R = L.R, RR = R * R, lane = threadIdx.x;
  const long long j0 = L.joff[k], ld = L.Jtot;
  const int Jk = (int)(L.joff[k + 1] - j0);
  const int pitch = (int)(L.Jmax | 1);
  double* dB = sm;                       // RE x RE, zero padded, element (r, c) at r + c*RE
  double* Bi = dB + RE * RE;
  double* Wm = Bi + RE * RE;             // DeltaB' W; after the iteration: V
  double* V0 = Wm + RE * RE;             // rotations of the previous inner iteration (warm) or I
  double* Ms = V0 + RE * RE;             // B_k + mu_k           (pitch x RE)
  double* Ts = Ms + (size_t)pitch * RE;  // P_k                  (pitch x RE)
  double* par = Ts + (size_t)pitch * RE; // (cos, sin) of the pairs of a round
  double* nrm = par + 2 * NP;            // tracked squared column norms by position, double buffered
  double* isg = nrm + 2 * RE;            // 1 / sigma
  const double rho = a.rho_k[k], half = rho / 2.0;

  for (int e = lane; e < RE * RE; e += 32) {
    const int r = e % RE, c = e / RE;
    const bool in = r < R && c < R;
    dB[e] = in ? a.DeltaB[r + c * R] : 0.0;
    Bi[e] = in ? a.Binv[(size_t)k * RR + r + c * R] : 0.0;
    V0[e] = (in && warm) ? a.Vprev[(size_t)k * RR + r + c * R] : ((r == c) ? 1.0 : 0.0);
  }
  __syncwarp();
  for (int e = lane; e < RE * RE; e += 32) {   // Wm = DeltaB' V0
    const int r = e % RE, c = e / RE;
    double x = 0.0;
    if (warm) {
      for (int q = 0; q < R; ++q) x = fma(dB[q + r * RE], V0[q + c * RE], x);
    } else {
      x = dB[c + r * RE];
    }
    Wm[e] = x;
  }
  __syncwarp();

  double S[JP][RE];
#pragma unroll
  for (int jp = 0; jp < JP; ++jp) {
    const int j = lane + 32 * jp;
    const bool ok = j < Jk;
    const long long g0 = j0 + j;
    double p[RE], mu[RE], s[RE];
#pragma unroll
    for (int c = 0; c < RE; ++c) {
      const bool in = ok && c < R;
      p[c] = in ? a.P[g0 + (long long)c * ld] : 0.0;
      mu[c] = in ? a.mu[g0 + (long long)c * ld] : 0.0;
    }
    // A_inner = A_k + rho_k/2 (P_k DeltaB - mu_k) [+ rho_k/2 (Z_k - muZ_k)]          (:526-529)
#pragma unroll
    for (int c = 0; c < RE; ++c) {
      double pd = 0.0;
#pragma unroll
      for (int r = 0; r < RE; ++r) pd = fma(p[r], dB[r + c * RE], pd);
      double v = 0.0;
      if (ok && c < R) {
        const long long g = g0 + (long long)c * ld;
        a.PDold[g] = pd;
        v = a.A[g] + half * (pd - mu[c]);
        if (a.con_active) v += half * (a.Z[g] - a.muZ[g]);
      }
      s[c] = v;
    }
    // B_k = A_inner inv(Bsys_k)   (:530) ;  M = B_k + mu_k ;  S = M DeltaB' W   (:532)
    double m[RE];
#pragma unroll
    for (int c = 0; c < RE; ++c) {
      double x = 0.0;
#pragma unroll
      for (int r = 0; r < RE; ++r) x = fma(s[r], Bi[r + c * RE], x);
      m[c] = 0.0;
      if (ok && c < R) {
        a.B[g0 + (long long)c * ld] = x;
        m[c] = x + mu[c];
        Ms[j + c * pitch] = m[c];
      }
    }
#pragma unroll
    for (int c = 0; c < RE; ++c) {
      double x = 0.0;
#pragma unroll
      for (int r = 0; r < RE; ++r) x = fma(m[r], Wm[r + c * RE], x);
      S[jp][c] = x;
    }
  }
  // ---- one-sided Jacobi: S <- S V with orthogonal columns
  double v[RE];   // row `lane` of V
#pragma unroll
  for (int c = 0; c < RE; ++c) v[c] = (lane < RE) ? V0[lane + c * RE] : 0.0;
  const int myq = transpose_index<NP>(lane);
  const bool pair_writer = (lane & (32 / NP - 1)) == 0;
  const int myc = transpose_index<RE>(lane);
  const bool col_writer = (lane & (32 / RE - 1)) == 0;
  const int np_t = ring_newpos<RE>(2 * myq), np_b = ring_newpos<RE>(2 * myq + 1);
  const double tol2 = 2.220446049250313e-16 * 2.220446049250313e-16 * (double)Jk;
  int buf = 0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    {  // exact squared column norms at the start of every sweep
      double n2[RE];
#pragma unroll
      for (int c = 0; c < RE; ++c) {
        double x = 0.0;
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) x = fma(S[jp][c], S[jp][c], x);
        n2[c] = x;
      }
      reduce_transpose<RE>(n2, lane);
      if (col_writer) nrm[buf * RE + myc] = n2[0];
    }
    __syncwarp();
    int rot = 0;
    for (int rd = 0; rd < RE - 1; ++rd) {
      double ga[NP];
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        double x = 0.0;
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) x = fma(S[jp][2 * q], S[jp][2 * q + 1], x);
        ga[q] = x;
      }
      reduce_transpose<NP>(ga, lane);
      const double g = ga[0];
      const double al = nrm[buf * RE + 2 * myq], be = nrm[buf * RE + 2 * myq + 1];
      double cs = 1.0, sn = 0.0, nal = al, nbe = be;
      int r1 = 0;
      if (g * g > tol2 * (al * be)) {
        // tan(theta) = 2g / (tau + sign(tau) sqrt(tau^2 + 4g^2)); this chain is the critical path of a round, so the
        // square root is x * rsqrt(x) and the division a reciprocal (a rotation only has to be orthogonal to rounding,
        // which cos = rsqrt(1 + tan^2), sin = cos * tan is for any tan)
        const double tau = be - al, g2 = 2.0 * g;
        const double x2 = fma(tau, tau, g2 * g2);
        const double t = g2 * __drcp_rn(tau + copysign(x2 * rsqrt(x2), tau));
        cs = rsqrt(fma(t, t, 1.0));
        sn = cs * t;
        nal = al - t * g;
        nbe = be + t * g;
        r1 = 1;
      }
      if (pair_writer) {
        *reinterpret_cast<double2*>(par + 2 * myq) = make_double2(cs, sn);
        nrm[(buf ^ 1) * RE + np_t] = nal;   // the columns move to their next positions below
        nrm[(buf ^ 1) * RE + np_b] = nbe;
      }
      __syncwarp();   // par / nrm written, visible to every lane
      const int any = __any_sync(0xffffffffu, r1);
      rot |= any;
      if (any) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const double2 cssn = *reinterpret_cast<const double2*>(par + 2 * q);
#pragma unroll
          for (int jp = 0; jp < JP; ++jp) {
            const double xv = S[jp][2 * q], yv = S[jp][2 * q + 1];
            S[jp][2 * q] = cssn.x * xv - cssn.y * yv;
            S[jp][2 * q + 1] = cssn.y * xv + cssn.x * yv;
          }
          const double xv = v[2 * q], yv = v[2 * q + 1];
          v[2 * q] = cssn.x * xv - cssn.y * yv;
          v[2 * q + 1] = cssn.y * xv + cssn.x * yv;
        }
      }
      __syncwarp();   // par is rewritten in the next round
#pragma unroll
      for (int jp = 0; jp < JP; ++jp) ring_step<RE>(S[jp]);
      ring_step<RE>(v);
      buf ^= 1;
    }
    if (rot == 0) break;
  }
  // after whole sweeps every column is back in its own position
  if (lane < RE) {
#pragma unroll
    for (int c = 0; c < RE; ++c) Wm[lane + c * RE] = v[c];   // V (element (i, c) at i + c*RE)
    if (lane < R) {
#pragma unroll
      for (int c = 0; c < RE; ++c)
        if (c < R) a.Vprev[(size_t)k * RR + lane + c * R] = v[c];
    }
  }
  {  // 1 / sigma
    double n2[RE];
#pragma unroll
    for (int c = 0; c < RE; ++c) {
      double x = 0.0;
#pragma unroll
      for (int jp = 0; jp < JP; ++jp) x = fma(S[jp][c], S[jp][c], x);
      n2[c] = x;
    }
    reduce_transpose<RE>(n2, lane);
    if (col_writer) isg[myc] = (n2[0] > 0.0) ? 1.0 / sqrt(n2[0]) : 0.0;
  }
  __syncwarp();
  // P_k = U V'   (:534)
#pragma unroll
  for (int jp = 0; jp < JP; ++jp) {
    const int j = lane + 32 * jp;
    const bool ok = j < Jk;
    double u[RE];
#pragma unroll
    for (int r = 0; r < RE; ++r) u[r] = S[jp][r] * isg[r];
#pragma unroll
    for (int c = 0; c < RE; ++c) {
      double x = 0.0;
#pragma unroll
      for (int r = 0; r < RE; ++r) x = fma(u[r], Wm[c + r * RE], x);
      if (ok && c < R) {
        a.P[j0 + j + (long long)c * ld] = x;
        Ts[j + c * pitch] = x;
      }
    }
  }
  __syncwarp();
  // contribution to DeltaB: rho_k P_k' (B_k + mu_k)   (:541)
  for (int e = lane; e < RR; e += 32) {
    const int ra = e % R, cb = e / R;
    const double* pc = Ts + ra * pitch;
    const double* mc = Ms + cb * pitch;
    double acc = 0.0;
    for (int j = 0; j < Jk; ++j) acc = fma(pc[j], mc[j], acc);
    a.contrib[(size_t)k * RR + e] = rho * acc;
  }
}

// one warp per element of DeltaB (see par2_modeA_had_kernel): 8 elements per CTA
__global__ void __launch_bounds__(256) par2_B_deltaB_kernel(Par2Layout L, Par2BArgs a, const InnerCtl* ctl,
                                                            double* __restrict__ sums_out) {
  if (ctl != nullptr && ctl->done != 0) return;
  __shared__ double red[32];
  __shared__ double s_sum;
  const int R = L.R, RR = R * R, K0 = L.k0, K1 = L.k1;
  // sum_k rho_k (:542) in a fixed (strided + tree) order: deterministic, independent of the launch
  double s = 0.0;
  for (int k = K0 + threadIdx.x; k < K1; k += blockDim.x) s += a.rho_k[k];
  s = block_sum(s, red);
  if (threadIdx.x == 0) s_sum = s;
  __syncthreads();
  const double tot = s_sum;
  if (sums_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) sums_out[RR] = tot;
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= RR) return;
  const double* c = a.contrib + e;
  double acc = 0.0;
  for (int k = K0 + lane; k < K1; k += 32) acc += c[(size_t)k * RR];   // lanes stride over the slices, fixed order
  acc = warp_sum(acc);
  if (lane == 0) {
    if (sums_out != nullptr) sums_out[e] = acc;   // sharded slices: all-reduced, then divided
    else a.DeltaB[e] = acc / tot;                 // :541-544
  }
}

__global__ void par2_B_deltaB_finish_kernel(int RR, double* __restrict__ DeltaB, const double* __restrict__ sums,
                                            const InnerCtl* ctl) {
  if (ctl != nullptr && ctl->done != 0) return;
  const double tot = sums[RR];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < RR; e += gridDim.x * blockDim.x) DeltaB[e] = sums[e] / tot;
}

__global__ void zero_rows_outside_kernel(double* __restrict__ M, long long rows, int cols, long long lo, long long hi) {
  const long long n = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e % rows;
    if (r < lo || r >= hi) M[e] = 0.0;
  }
}

__global__ void __launch_bounds__(kP2Threads) par2_B_step2a_kernel(Par2Layout L, Par2BArgs a, const InnerCtl* ctl) {
  if (ctl != nullptr && ctl->done != 0) return;
  extern __shared__ double sm[];
  __shared__ double red[32];
  const int k = blockIdx.x + L.k0, R = L.R, RR = R * R, tid = threadIdx.x, nt = blockDim.x;
  const long long j0 = L.joff[k], ld = L.Jtot;
  const int Jk = (int)(L.joff[k + 1] - j0);
  double* dB = sm;
  for (int e = tid; e < RR; e += nt) dB[e] = a.DeltaB[e];
  __syncthreads();
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int it = tid; it < Jk * R; it += nt) {
    const int c = it / Jk, j = it % Jk;
    const long long g = j0 + j + (long long)c * ld;
    double pd = 0.0;
    for (int r = 0; r < R; ++r) pd = fma(a.P[j0 + j + (long long)r * ld], dB[r + c * R], pd);
    const double b = a.B[g];
    const double mun = a.mu[g] + b - pd;               // :546
    a.mu[g] = mun;
    const double d0 = b - pd, d2 = a.PDold[g] - pd;
    s0 = fma(d0, d0, s0);
    s1 = fma(b, b, s1);
    s2 = fma(d2, d2, s2);
    s3 = fma(mun, mun, s3);
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  s3 = block_sum(s3, red);
  if (tid == 0) {
    double* o = a.norms + (size_t)k * 8;
    o[0] = s0;
    o[1] = s1;
    o[2] = s2;
    o[3] = s3;
  }
}

__global__ void par2_B_form_prox_input_kernel(Par2Layout L, Par2BArgs a, double* __restrict__ V, const InnerCtl* ctl) {
  if (ctl != nullptr && ctl->done != 0) return;
  const long long nloc = L.jhi - L.jlo, n = nloc * L.R;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long idx = L.jlo + e % nloc + (e / nloc) * L.Jtot;
    V[idx] = a.B[idx] + a.muZ[idx];
  }
}

__device__ void par2_finalize_ctl(double rpk, double rdk, double rpc, double rdc, int K, const InnerTol& tol, InnerCtl* ctl) {
  const double invK = 1.0 / (double)K;
  rpk *= invK;
  rdk *= invK;
  rpc *= invK;
  rdc *= invK;
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

__global__ void par2_B_finalize_kernel(int K, const double* __restrict__ res_sums, InnerTol tol, InnerCtl* ctl) {
  if (ctl->done != 0) return;
  par2_finalize_ctl(res_sums[0], res_sums[1], res_sums[2], res_sums[3], K, tol, ctl);
}

__global__ void __launch_bounds__(kP2Threads) par2_B_step2b_kernel(Par2Layout L, Par2BArgs a, InnerTol tol,
                                                                    InnerCtl* ctl, unsigned* counter,
                                                                    double* __restrict__ res_out) {
  if (ctl->done != 0) return;
  __shared__ double red[32];
  __shared__ bool s_last;
  const int k = blockIdx.x + L.k0, R = L.R, tid = threadIdx.x, nt = blockDim.x;
  const long long j0 = L.joff[k], ld = L.Jtot;
  const int Jk = (int)(L.joff[k + 1] - j0);
  if (a.con_active) {
    const double rho = a.rho_k[k];
    double s4 = 0.0, s5 = 0.0, s6 = 0.0;
    for (int it = tid; it < Jk * R; it += nt) {
      const int c = it / Jk, j = it % Jk;
      const long long g = j0 + j + (long long)c * ld;
      const double b = a.B[g], muz = a.muZ[g], zold = a.Z[g];
      const double z = (a.Znew != nullptr) ? a.Znew[g] : prox_elem(a.prox_kind, b + muz, a.p0, a.p1, rho);  // :568
      const double mun = muz + b - z;                                                                       // :569
      a.Z[g] = z;
      a.muZ[g] = mun;
      const double d4 = b - z, d5 = zold - z;
      s4 = fma(d4, d4, s4);
      s5 = fma(d5, d5, s5);
      s6 = fma(mun, mun, s6);
    }
    s4 = block_sum(s4, red);
    s5 = block_sum(s5, red);
    s6 = block_sum(s6, red);
    if (tid == 0) {
      double* o = a.norms + (size_t)k * 8;
      o[4] = s4;
      o[5] = s5;
      o[6] = s6;
    }
  }
  __threadfence();
  if (tid == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // residuals of :558-585, averaged over the slices in slice order (deterministic)
  double rpk = 0.0, rdk = 0.0, rpc = 0.0, rdc = 0.0;
  for (int kk = L.k0 + tid; kk < L.k1; kk += nt) {
    const double* o = a.norms + (size_t)kk * 8;
    const double nB = sqrt(o[1]);
    rpk += sqrt(o[0]) / nB;
    rdk += sqrt(o[2]) / sqrt(o[3]);                   // not guarded in the reference (:584)
    if (a.con_active) {
      rpc += sqrt(o[4]) / nB;
      const double sc = sqrt(o[6]);
      rdc += (sc > 0.0) ? sqrt(o[5]) / sc : sqrt(o[5]);
    }
  }
  rpk = block_sum(rpk, red);
  rdk = block_sum(rdk, red);
  rpc = block_sum(rpc, red);
  rdc = block_sum(rdc, red);
  if (tid == 0) {
    if (res_out != nullptr) {   // sharded slices: local sums only, the exit test follows the all-reduce
      res_out[0] = rpk;
      res_out[1] = rdk;
      res_out[2] = rpc;
      res_out[3] = rdc;
    } else {
      par2_finalize_ctl(rpk, rdk, rpc, rdc, L.K, tol, ctl);
    }
    *counter = 0u;
  }
}

__global__ void par2_tsmooth_diag_kernel(Par2Layout L, const double* __restrict__ rho_k, double eta,
                                         double* __restrict__ dp, const InnerCtl* ctl) {
  if (ctl != nullptr && ctl->done != 0) return;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int K = L.K;
  const double c = -2.0 * eta;
  for (int k = 0; k < K; ++k) {  // t_smoothness_prox.m:23-37
    double a = 4.0 * eta + rho_k[k];
    if (k == 0) a -= 2.0 * eta;
    if (k == K - 1) a -= 2.0 * eta;
    if (k > 0) {
      const double m = c / dp[k - 1];  // :41-45
      a -= m * c;
    }
    dp[k] = a;
  }
}

__global__ void par2_tsmooth_solve_kernel(Par2Layout L, const double* __restrict__ V, const double* __restrict__ rho_k,
                                          double eta, const double* __restrict__ dp, double* __restrict__ out,
                                          const InnerCtl* ctl) {
  if (ctl != nullptr && ctl->done != 0) return;
  const int K = L.K;
  const long long J = L.joff[1] - L.joff[0];
  const long long n = J * L.R;
  const double c = -2.0 * eta;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e % J, r = e / J;
    const long long base = j + r * L.Jtot;
    double prev = rho_k[0] * V[base];
    out[base] = prev;
    for (int k = 1; k < K; ++k) {  // forward elimination of the right-hand sides (:41-45)
      const long long g = base + (long long)k * J;
      const double m = c / dp[k - 1];
      prev = rho_k[k] * V[g] - m * prev;
      out[g] = prev;
    }
    double q = out[base + (long long)(K - 1) * J] / dp[K - 1];  // back substitution (:48-54)
    out[base + (long long)(K - 1) * J] = q;
    for (int k = K - 2; k >= 0; --k) {
      const long long g = base + (long long)k * J;
      q = (out[g] - c * q) / dp[k];
      out[g] = q;
    }
  }
}

__global__ void __launch_bounds__(kP2Threads) par2_seg_norms_kernel(Par2Layout L, const double* __restrict__ Bst,
                                                                     const double* __restrict__ Z,
                                                                     const double* __restrict__ P,
                                                                     const double* __restrict__ DeltaB, int reg_kind,
                                                                     double* __restrict__ out) {
  extern __shared__ double sm[];
  __shared__ double red[32];
  const int k = blockIdx.x + L.k0, R = L.R, RR = R * R, tid = threadIdx.x, nt = blockDim.x;
  const long long j0 = L.joff[k], ld = L.Jtot;
  const int Jk = (int)(L.joff[k + 1] - j0);
  double* dB = sm;
  for (int e = tid; e < RR; e += nt) dB[e] = DeltaB[e];
  __syncthreads();
  double n2 = 0.0, dz = 0.0, dp = 0.0, rg = 0.0;
  for (int it = tid; it < Jk * R; it += nt) {
    const int c = it / Jk, j = it % Jk;
    const long long g = j0 + j + (long long)c * ld;
    const double b = Bst[g];
    double pd = 0.0;
    for (int r = 0; r < R; ++r) pd = fma(P[j0 + j + (long long)r * ld], dB[r + c * R], pd);
    n2 = fma(b, b, n2);
    dp = fma(b - pd, b - pd, dp);
    if (Z != nullptr) {
      const double d = b - Z[g];
      dz = fma(d, d, dz);
    }
    switch (reg_kind) {
      case RED_L1: rg += fabs(b); break;
      case RED_NNZ: rg += (b != 0.0) ? 1.0 : 0.0; break;
      case RED_NORM2: rg = fma(b, b, rg); break;
      case RED_TVSUM:
        if (j + 1 < Jk) rg += Bst[g + 1] - b;
        break;
      case RED_TSMOOTH:
        if (k > 0) {
          const double d = b - Bst[g - Jk];  // same element of the previous slice (equal J_k)
          rg = fma(d, d, rg);
        }
        break;
      case RED_GLQUAD: {
        double lx = b;
        if (Jk > 1) {
          if (j == 0) lx = b - Bst[g + 1];
          else if (j == Jk - 1) lx = b - Bst[g - 1];
          else lx = 2.0 * b - Bst[g - 1] - Bst[g + 1];
        }
        rg = fma(b, lx, rg);
        break;
      }
      default: break;
    }
  }
  n2 = block_sum(n2, red);
  dz = block_sum(dz, red);
  dp = block_sum(dp, red);
  rg = block_sum(rg, red);
  if (reg_kind == RED_COLNORM) {
    double tot = 0.0;
    for (int c = 0; c < R; ++c) {
      double s = 0.0;
      for (int j = tid; j < Jk; j += nt) {
        const double b = Bst[j0 + j + (long long)c * ld];
        s = fma(b, b, s);
      }
      s = block_sum(s, red);
      if (tid == 0) tot += sqrt(s);
    }
    rg = tot;
  }
  if (tid == 0) {
    out[(size_t)k * 4 + 0] = n2;
    out[(size_t)k * 4 + 1] = dz;
    out[(size_t)k * 4 + 2] = dp;
    out[(size_t)k * 4 + 3] = rg;
  }
}

__global__ void __launch_bounds__(256) par2_residual_kernel(Par2Layout L, const double* __restrict__ X, long long ldX,
                                                             long long I, const double* __restrict__ A, long long ldA,
                                                             const double* __restrict__ Bst,
                                                             const double* __restrict__ C, long long ldc,
                                                             double* __restrict__ partials, unsigned* counter,
                                                             double* __restrict__ res) {
  __shared__ double red[32];
  __shared__ bool s_last;
  const int R = L.R;
  const long long n = I * (L.jhi - L.jlo);
  double acc = 0.0;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx % I, j = L.jlo + idx / I;
    const int k = L.seg[j];
    double m = 0.0;
    for (int r = 0; r < R; ++r) m = fma(A[i + r * ldA] * C[k + r * ldc], Bst[j + r * L.Jtot], m);
    const double v = X[i + j * ldX] - m;
    acc = fma(v, v, acc);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
  __threadfence();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    const double v = warp_sum_partials(partials, gridDim.x, 1u);
    if (threadIdx.x == 0) {
      *res = v;
      *counter = 0u;
    }
  }
}

template <typename Kern>
void opt_in_smem(Kern kern, size_t smem) {
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);  // once per device, kernel and size
}

unsigned flat_grid(long long n) { return (unsigned)std::min<long long>(ceil_div(std::max<long long>(n, 1), 256), 148 * 8); }

constexpr size_t kSmemBudget = 200 * 1024;

}  // namespace

size_t par2_step1_smem_bytes(long long Jmax, int R) {
  const size_t small = ((size_t)((R + 1) & ~1) + (R <= 64 ? (size_t)3 * R * R : 0)) * sizeof(double);
  const size_t big = small + (size_t)2 * Jmax * R * sizeof(double);
  return big <= kSmemBudget ? big : small;
}

int par2_scale_rows(const Par2Layout& L, double* out, const double* in, const double* C, long long ldc, double scale,
                    const double* addend, double add_scale, cudaStream_t st) {
  par2_scale_rows_kernel<<<flat_grid((L.jhi - L.jlo) * L.R), 256, 0, st>>>(L, out, in, C, ldc, scale, addend, add_scale);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_batched_gram(const Par2Layout& L, const double* Bst, double* G2, cudaStream_t st) {
  par2_batched_gram_kernel<<<L.k1 - L.k0, kP2Threads, 0, st>>>(L, Bst, G2);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_modeA_had(const Par2Layout& L, const double* G2, const double* C, long long ldc, double* Csum, cudaStream_t st) {
  par2_modeA_had_kernel<<<(unsigned)ceil_div(L.R * L.R, 8), 256, 0, st>>>(L, G2, C, ldc, Csum);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_sys_prep(const Par2Layout& L, const Par2SysArgs& a, cudaStream_t st) {
  if (L.R > 64 && a.gws == nullptr) throw CudaError(1, "par2_sys_prep: workspace required for R > 64");
  const size_t smem = ((L.R > 64 ? 0 : (size_t)2 * L.R * L.R) + L.R) * sizeof(double);
  opt_in_smem(par2_sys_prep_kernel, smem);
  par2_sys_prep_kernel<<<L.k1 - L.k0, 128, smem, st>>>(L, a);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rho_stats(const double* rho_k, int K, double* out, cudaStream_t st) {
  par2_rho_stats_kernel<<<1, 256, 0, st>>>(rho_k, K, out);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_assemble_B2(const double* Bsys, const double* HtH, const double* rhoC_dev, int constrained, int K, int R,
                     double* B2, cudaStream_t st) {
  const long long n = (long long)K * R;
  par2_assemble_B2_kernel<<<flat_grid(n * n), 256, 0, st>>>(Bsys, HtH, rhoC_dev, constrained, K, R, B2);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_chol_solve_vec(const double* L, int K, int R, const double* a, double* x, cudaStream_t st, const int* skip) {
  const size_t smem = (size_t)K * R * sizeof(double);
  if (smem > 40 * 1024) throw CudaError(2, "coupling type 1 with a PARAFAC2 third mode: K*R too large");
  par2_chol_solve_vec_kernel<<<1, 256, smem, st>>>(L, K, R, a, x, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rows_ainner(double* out, const double* A, const double* X, const double* Z, const double* muZ,
                     const double* rho_k, long long rows, int cols, cudaStream_t st, const int* skip) {
  par2_rows_ainner_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(out, A, X, Z, muZ, rho_k, rows, cols, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rows_apply(double* out, const double* in, const double* Minv, long long rows, int q, cudaStream_t st,
                    const int* skip) {
  par2_rows_apply_kernel<<<flat_grid(rows * q), 256, 0, st>>>(out, in, Minv, rows, q, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rows_scale(double* out, const double* in, const double* rho_k, long long rows, int cols, cudaStream_t st,
                    const int* skip) {
  par2_rows_scale_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(out, in, rho_k, rows, cols, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rows_weighted_accum(double* D, double* wsum, const double* S, const double* mu, const double* rho_scalar,
                             const double* rho_k, long long rows, int cols, int first, cudaStream_t st, const int* skip) {
  par2_rows_weighted_accum_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(D, wsum, S, mu, rho_scalar, rho_k, rows, cols,
                                                                         first, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rows_divide(double* D, const double* wsum, long long rows, int cols, cudaStream_t st, const int* skip) {
  par2_rows_divide_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(D, wsum, rows, cols, skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rowsys_inverse(const double* AA, const double* AAA, const double* rho_k, int K, int q, double* Minv, InnerCtl* ctl,
                        cudaStream_t st) {
  if (q > 64) throw CudaError(2, "coupling type 4 with a PARAFAC2 third mode: at most 64 columns in the coupling factor");
  const size_t smem = (size_t)2 * q * q * sizeof(double);
  opt_in_smem(par2_rowsys_inverse_kernel, smem);
  par2_rowsys_inverse_kernel<<<K, 128, smem, st>>>(AA, AAA, rho_k, q, Minv, ctl);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_rho_max(const double* rho_k, int K, double* out, cudaStream_t st) {
  par2_rho_max_kernel<<<1, 256, 0, st>>>(rho_k, K, out);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_B_step1(const Par2Layout& L, const Par2BArgs& a, const InnerCtl* ctl, int warm, cudaStream_t st) {
  // small slices: one warp per slice with the slice in registers
  if (L.R <= 16 && L.Jmax <= 128) {
    const int nsl = L.k1 - L.k0;
    const int jp = (L.Jmax <= 32) ? 1 : (L.Jmax <= 64 ? 2 : 4);
    auto go = [&](auto kern, size_t doubles) {
      const size_t bytes = doubles * sizeof(double);
      opt_in_smem(kern, bytes);
      kern<<<nsl, 32, bytes, st>>>(L, a, ctl, warm);
    };
#define AO_P2_REG(RE_)                                                                                   \
  do {                                                                                                   \
    const size_t d = par2_step1_reg_doubles<RE_>(L.Jmax);                                                \
    if (jp == 1) go(par2_B_step1_reg_kernel<RE_, 1>, d);                                                 \
    else if (jp == 2) go(par2_B_step1_reg_kernel<RE_, 2>, d);                                            \
    else go(par2_B_step1_reg_kernel<RE_, 4>, d);                                                         \
  } while (0)
    if (L.R <= 4) AO_P2_REG(4);
    else if (L.R <= 8) AO_P2_REG(8);
    else AO_P2_REG(16);
#undef AO_P2_REG
    AO_CHECK_LAUNCH();
    return 1;
  }
  const size_t smem = par2_step1_smem_bytes(L.Jmax, L.R);
  const size_t head = ((size_t)((L.R + 1) & ~1) + (L.R <= 64 ? (size_t)3 * L.R * L.R : 0)) * sizeof(double);
  const int use_gmem = (smem == head) ? 1 : 0;
  opt_in_smem(par2_B_step1_kernel, smem);
  par2_B_step1_kernel<<<L.k1 - L.k0, kP2Threads, smem, st>>>(L, a, ctl, use_gmem, warm);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_B_deltaB(const Par2Layout& L, const Par2BArgs& a, const InnerCtl* ctl, cudaStream_t st, double* sums_out) {
  par2_B_deltaB_kernel<<<(unsigned)ceil_div(L.R * L.R, 8), 256, 0, st>>>(L, a, ctl, sums_out);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_B_deltaB_finish(const Par2Layout& L, const Par2BArgs& a, const double* sums, const InnerCtl* ctl, cudaStream_t st) {
  par2_B_deltaB_finish_kernel<<<(unsigned)ceil_div(L.R * L.R, 256), 256, 0, st>>>(L.R * L.R, a.DeltaB, sums, ctl);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_B_finalize(const Par2Layout& L, const double* res_sums, const InnerTol& tol, InnerCtl* ctl, cudaStream_t st) {
  par2_B_finalize_kernel<<<1, 1, 0, st>>>(L.K, res_sums, tol, ctl);
  AO_CHECK_LAUNCH();
  return 1;
}

int zero_rows_outside(double* M, long long rows, int cols, long long lo, long long hi, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  zero_rows_outside_kernel<<<flat_grid(rows * cols), 256, 0, st>>>(M, rows, cols, lo, hi);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_B_step2a(const Par2Layout& L, const Par2BArgs& a, const InnerCtl* ctl, cudaStream_t st) {
  const size_t smem = (size_t)L.R * L.R * sizeof(double);
  opt_in_smem(par2_B_step2a_kernel, smem);
  par2_B_step2a_kernel<<<L.k1 - L.k0, kP2Threads, smem, st>>>(L, a, ctl);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_B_form_prox_input(const Par2Layout& L, const Par2BArgs& a, double* V, const InnerCtl* ctl, cudaStream_t st) {
  par2_B_form_prox_input_kernel<<<flat_grid((L.jhi - L.jlo) * L.R), 256, 0, st>>>(L, a, V, ctl);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_B_step2b(const Par2Layout& L, const Par2BArgs& a, const InnerTol& tol, InnerCtl* ctl, unsigned* counter,
                  cudaStream_t st, double* res_out) {
  par2_B_step2b_kernel<<<L.k1 - L.k0, kP2Threads, 0, st>>>(L, a, tol, ctl, counter, res_out);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_tsmooth_prox(const Par2Layout& L, const double* V, const double* rho_k, double eta, double* dp, double* out,
                      const InnerCtl* ctl, cudaStream_t st) {
  par2_tsmooth_diag_kernel<<<1, 32, 0, st>>>(L, rho_k, eta, dp, ctl);
  AO_CHECK_LAUNCH();
  par2_tsmooth_solve_kernel<<<flat_grid((L.Jtot / L.K) * L.R), 256, 0, st>>>(L, V, rho_k, eta, dp, out, ctl);
  AO_CHECK_LAUNCH();
  return 2;
}

int par2_seg_norms(const Par2Layout& L, const double* Bst, const double* Z, const double* P, const double* DeltaB,
                   int reg_kind, double* out, cudaStream_t st) {
  const size_t smem = (size_t)L.R * L.R * sizeof(double);
  opt_in_smem(par2_seg_norms_kernel, smem);
  par2_seg_norms_kernel<<<L.k1 - L.k0, kP2Threads, smem, st>>>(L, Bst, Z, P, DeltaB, reg_kind, out);
  AO_CHECK_LAUNCH();
  return 1;
}

int par2_residual(const Par2Layout& L, const double* X, long long ldX, long long I, const double* A, long long ldA,
                  const double* Bst, const double* C, long long ldc, double* partials, unsigned* counter, double* res,
                  cudaStream_t st) {
  par2_residual_kernel<<<flat_grid(I * (L.jhi - L.jlo)), 256, 0, st>>>(L, X, ldX, I, A, ldA, Bst, C, ldc, partials, counter, res);
  AO_CHECK_LAUNCH();
  return 1;
}

}  // namespace aoadmm
