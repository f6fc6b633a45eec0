// prox.cu - proximal / projection operators on device (sm_100a).  Compiled with --fmad=false so that the
// data-dependent branches of the serial algorithms (Condat TV, Stout unimodal, PAVA) see the same
// roundings as a plain IEEE evaluation.
//
// Reference: functions/constraints_to_prox.m:13-91 (the named specs), prox_normalized_nonneg.m:3-11,
// prox_TV.m:6-8 (+ Condat's direct 1-D TV algorithm, un-vendored TV_Condat_v2), project_unimodal.m:10-14,
// project_unimodal_vector.m:10-88 (Stout's prefix isotonic regression), and the Proximity Operator
// Repository functions project_box / project_simplex / project_L1 / project_L2 / project_monotone /
// prox_abs / prox_zero / prox_L2 (un-vendored; implemented from their mathematical definitions).
#include "smallops.cuh"
#include "linalg.cuh"

#include <algorithm>

namespace aoadmm {

namespace {

__device__ __forceinline__ double load_rho(const double* rho_dev, double rho_host) {
  return rho_dev != nullptr ? *rho_dev : rho_host;
}

// Segmented application (PARAFAC2 B_k mode, cmtf_fun_AOADMM.m:567-568): the stacked Jtot x R matrix is K independent
// row segments [seg_off[k], seg_off[k+1]) with their own rho_k; blockIdx.y picks the segment, so the prox of every
// slice and every column is ONE launch instead of K.  seg_off == nullptr: the whole column, scalar rho.
struct SegInfo {
  const long long* seg_off;
  long long max_rows;   // longest segment (sizes the per-column workspace)
};
__device__ __forceinline__ void seg_select(const SegInfo& sg, long long& rows, long long& row0, const double*& rho_dev) {
  row0 = 0;
  if (sg.seg_off != nullptr) {
    row0 = sg.seg_off[blockIdx.y];
    rows = sg.seg_off[blockIdx.y + 1] - row0;
    if (rho_dev != nullptr) rho_dev += blockIdx.y;
  }
}

__device__ __forceinline__ double prox_elem2(int kind, double v, double p0, double p1, double rho) {
  switch (kind) {
    case PROX_NONNEG: return fmax(v, 0.0);
    case PROX_BOX: return fmin(fmax(v, p0), p1);
    case PROX_L1_REG: {
      const double g = p0 / rho;
      const double mag = fmax(fabs(v) - g, 0.0);
      return (v > 0.0) ? mag : ((v < 0.0) ? -mag : 0.0);
    }
    case PROX_L0_REG: {
      const double g = p0 / rho;
      return (fabs(v) > sqrt(2.0 * g)) ? v : 0.0;
    }
    case PROX_RIDGE: return 1.0 / (2.0 * (p0 / rho) + 1.0) * v;
    default: return v;
  }
}

__global__ void prox_elementwise_kernel(int kind, double p0, double p1, const double* __restrict__ X, long long ldx,
                                        double* __restrict__ out, long long ldo, long long rows, int cols,
                                        const double* rho_dev, double rho_host, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const double rho = load_rho(rho_dev, rho_host);
  const long long n = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / rows, i = idx % rows;
    out[c * ldo + i] = prox_elem2(kind, X[c * ldx + i], p0, p1, rho);
  }
}

// ---- column-norm based operators: one CTA per column ----------------------------------------------
__global__ void prox_colnorm_kernel(int kind, double p0, const double* __restrict__ X, long long ldx,
                                    double* __restrict__ out, long long ldo, long long rows,
                                    const double* rho_dev, double rho_host, const int* __restrict__ skip, SegInfo sg) {
  if (skip != nullptr && *skip != 0) return;
  __shared__ double red[32];
  __shared__ double s_val;
  __shared__ long long s_arg;
  long long row0;
  seg_select(sg, rows, row0, rho_dev);
  const double rho = load_rho(rho_dev, rho_host);
  const double* x = X + (long long)blockIdx.x * ldx + row0;
  double* o = out + (long long)blockIdx.x * ldo + row0;
  const bool nonneg = (kind == PROX_NONNEG_L2_BALL || kind == PROX_NONNEG_L2_SPHERE);
  double s = 0.0;
  for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
    double v = x[i];
    if (nonneg) v = fmax(v, 0.0);
    s += v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) s_val = sqrt(s);
  __syncthreads();
  const double nrm = s_val;
  if (kind == PROX_NONNEG_L2_SPHERE && nrm == 0.0) {
    // prox_normalized_nonneg.m:5-7: unit vector at the FIRST maximum coordinate of the unprojected column
    if (threadIdx.x == 0) {
      double best = x[0];
      long long arg = 0;
      for (long long i = 1; i < rows; ++i)
        if (x[i] > best) {
          best = x[i];
          arg = i;
        }
      s_arg = arg;
    }
    __syncthreads();
    for (long long i = threadIdx.x; i < rows; i += blockDim.x) o[i] = (i == s_arg) ? 1.0 : 0.0;
    return;
  }
  double scale = 1.0;
  bool divide = false;
  if (kind == PROX_L2_BALL || kind == PROX_NONNEG_L2_BALL) {
    scale = (nrm > p0) ? p0 / nrm : 1.0;
  } else if (kind == PROX_L2_REG) {
    const double g = p0 / rho;
    scale = (nrm > g) ? 1.0 - g / nrm : 0.0;
  } else if (kind == PROX_NONNEG_L2_SPHERE) {
    divide = true;
  }
  for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
    double v = x[i];
    if (nonneg) v = fmax(v, 0.0);
    o[i] = divide ? v / nrm : v * scale;
  }
}

// ---- simplex / l1-ball: Michelot's finite algorithm, one CTA per column -----------------------------
__global__ void prox_simplex_col_kernel(int kind, double eta, const double* __restrict__ X, long long ldx,
                                        double* __restrict__ out, long long ldo, long long rows,
                                        const int* __restrict__ skip, SegInfo sg) {
  if (skip != nullptr && *skip != 0) return;
  __shared__ double red[32];
  __shared__ double s_theta;
  __shared__ double s_cnt;
  long long row0;
  const double* no_rho = nullptr;
  seg_select(sg, rows, row0, no_rho);
  const double* x = X + (long long)blockIdx.x * ldx + row0;
  double* o = out + (long long)blockIdx.x * ldo + row0;
  const bool l1 = (kind == PROX_L1_BALL);
  if (l1) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x) s += fabs(x[i]);
    s = block_sum(s, red);
    if (threadIdx.x == 0) s_theta = s;
    __syncthreads();
    if (s_theta <= eta) {
      for (long long i = threadIdx.x; i < rows; i += blockDim.x) o[i] = x[i];
      return;
    }
    __syncthreads();
  }
  double theta = -INFINITY;
  double prev_cnt = -1.0;
  for (long long iter = 0; iter <= rows; ++iter) {
    double s = 0.0, c = 0.0;
    for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
      const double v = l1 ? fabs(x[i]) : x[i];
      if (v > theta) {
        s += v;
        c += 1.0;
      }
    }
    s = block_sum(s, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) {
      s_cnt = c;
      s_theta = (c > 0.0) ? (s - eta) / c : theta;
    }
    __syncthreads();
    const double cnt = s_cnt;
    const double th = s_theta;
    __syncthreads();
    if (cnt == prev_cnt || cnt == 0.0) break;
    prev_cnt = cnt;
    theta = th;
  }
  for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
    const double xi = x[i];
    const double v = l1 ? fabs(xi) : xi;
    const double w = fmax(v - theta, 0.0);
    o[i] = l1 ? ((xi > 0.0) ? w : ((xi < 0.0) ? -w : 0.0)) : w;
  }
}

// row-wise simplex: one thread per row
__global__ void prox_simplex_row_kernel(double eta, const double* __restrict__ X, long long ldx,
                                        double* __restrict__ out, long long ldo, long long rows, int cols,
                                        const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double theta = -INFINITY;
  int prev = -1;
  for (int iter = 0; iter <= cols; ++iter) {
    double s = 0.0;
    int c = 0;
    for (int r = 0; r < cols; ++r) {
      const double v = X[(long long)r * ldx + i];
      if (v > theta) {
        s += v;
        ++c;
      }
    }
    if (c == prev || c == 0) break;
    prev = c;
    theta = (s - eta) / (double)c;
  }
  for (int r = 0; r < cols; ++r) out[(long long)r * ldo + i] = fmax(X[(long long)r * ldx + i] - theta, 0.0);
}

// ---- serial per-column algorithms ------------------------------------------------------------------
// Condat 2013 direct 1-D TV: x = argmin 0.5||x-y||^2 + lam * sum |x[i+1]-x[i]|
// rc[c] = 1/c (correctly rounded) replaces the two divisions per step by multiplications (<= 1 ulp from x/c).
__device__ __forceinline__ void tv_condat_serial(const double* y, double* x, long long n, double lam, const double* rc) {
  if (n <= 0) return;
  if (!(lam > 0.0)) {
    for (long long i = 0; i < n; ++i) x[i] = y[i];
    return;
  }
  long long k = 0, k0 = 0, kplus = 0, kminus = 0;
  double umin = lam, umax = -lam;
  double vmin = y[0] - lam, vmax = y[0] + lam;
  const double twolam = 2.0 * lam, minlam = -lam;
  for (;;) {
    while (k == n - 1) {
      if (umin < 0.0) {
        do x[k0++] = vmin; while (k0 <= kminus);
        kminus = k = k0;
        vmin = y[k];
        umin = lam;
        umax = vmin + umin - vmax;
      } else if (umax > 0.0) {
        do x[k0++] = vmax; while (k0 <= kplus);
        kplus = k = k0;
        vmax = y[k];
        umax = minlam;
        umin = vmax + umax - vmin;
      } else {
        vmin += umin * rc[k - k0 + 1];
        do x[k0++] = vmin; while (k0 <= k);
        return;
      }
    }
    umin += y[k + 1] - vmin;
    if (umin < minlam) {
      do x[k0++] = vmin; while (k0 <= kminus);
      kplus = kminus = k = k0;
      vmin = y[k];
      vmax = vmin + twolam;
      umin = lam;
      umax = minlam;
    } else {
      umax += y[k + 1] - vmax;
      if (umax > lam) {
        do x[k0++] = vmax; while (k0 <= kplus);
        kplus = kminus = k = k0;
        vmax = y[k];
        vmin = vmax - twolam;
        umin = lam;
        umax = minlam;
      } else {
        k++;
        if (umin >= lam) {
          kminus = k;
          vmin += (umin - lam) * rc[kminus - k0 + 1];
          umin = lam;
        }
        if (umax <= minlam) {
          kplus = k;
          vmax += (umax + lam) * rc[kplus - k0 + 1];
          umax = minlam;
        }
      }
    }
  }
}

// The same minimiser by dynamic programming (N. Johnson, "A dynamic programming algorithm for the fused lasso and
// L0-segmentation", 2013): one forward pass that keeps the knots of the piecewise-linear derivative of the message
// function in a double-ended queue (amortised two queue steps per element, no rescans), one backward pass through the
// clipping bounds tm/tp.  About three times faster than the direct algorithm on noisy columns, where that one restarts
// often; needs 9n+1 doubles of workspace: knot positions xk and their slope / intercept increments ak, bk (2n each),
// tm, tp (n each) and the table rc[c] = 1/c: the slopes are integers (sums of +-1), so the two divisions per element
// become multiplications (<= 1 ulp from the quotient; an FP64 division is a ~200-cycle routine on the critical path).
// The minimiser is unique, so both algorithms agree to rounding.
__device__ __forceinline__ void tv_dp_serial(const double* y, double* beta, long long n, double lam, double* work) {
  if (n <= 0) return;
  if (n == 1 || !(lam > 0.0)) {
    for (long long i = 0; i < n; ++i) beta[i] = y[i];
    return;
  }
  double* xk = work;
  double* ak = work + 2 * n;
  double* bk = work + 4 * n;
  double* tm = work + 6 * n;
  double* tp = work + 7 * n;
  const double* rc = work + 8 * n;   // rc[c] = 1/c, c = 0..n (filled by the whole warp before the serial part)
  tm[0] = y[0] - lam;
  tp[0] = y[0] + lam;
  long long l = n - 1, r = n;
  xk[l] = tm[0];
  xk[r] = tp[0];
  ak[l] = 1.0;
  bk[l] = lam - y[0];
  ak[r] = -1.0;
  bk[r] = y[0] + lam;
  double bfirst = -lam - y[1], blast = -lam + y[1];   // afirst = 1, alast = -1
  for (long long k = 1; k < n - 1; ++k) {
    double alo = 1.0, blo = bfirst;
    long long lo = l;
    for (; lo <= r; ++lo) {
      if (alo * xk[lo] + blo > -lam) break;
      alo += ak[lo];
      blo += bk[lo];
    }
    const double tmk = (-lam - blo) * rc[(int)alo];
    l = lo - 1;
    xk[l] = tmk;
    tm[k] = tmk;
    double ahi = -1.0, bhi = blast;
    long long hi = r;
    for (; hi >= l; --hi) {
      if (-ahi * xk[hi] - bhi < lam) break;
      ahi += ak[hi];
      bhi += bk[hi];
    }
    const double tpk = (lam + bhi) * rc[(int)(-ahi)];
    r = hi + 1;
    xk[r] = tpk;
    tp[k] = tpk;
    ak[l] = alo;
    bk[l] = blo + lam;
    ak[r] = ahi;
    bk[r] = bhi + lam;
    const double yn = y[k + 1];
    bfirst = -lam - yn;
    blast = -lam + yn;
  }
  {
    double alo = 1.0, blo = bfirst;
    for (long long lo = l; lo <= r; ++lo) {
      if (alo * xk[lo] + blo > 0.0) break;
      alo += ak[lo];
      blo += bk[lo];
    }
    beta[n - 1] = -blo / alo;
  }
  for (long long k = n - 2; k >= 0; --k) {
    const double nx = beta[k + 1];
    beta[k] = (nx > tp[k]) ? tp[k] : ((nx < tm[k]) ? tm[k] : nx);
  }
}

// pool-adjacent-violators, non-decreasing fit of sign*y; writes sign*fit into x.
// work: level[n], weight[n] (doubles), start[n] (ints)
__device__ void pava_serial(const double* y, double* x, long long n, double sign, double* level, double* weight,
                            int* start) {
  long long nb = 0;
  for (long long i = 0; i < n; ++i) {
    level[nb] = sign * y[i];
    weight[nb] = 1.0;
    start[nb] = (int)i;
    ++nb;
    while (nb > 1 && level[nb - 2] > level[nb - 1]) {
      const double w = weight[nb - 2] + weight[nb - 1];
      level[nb - 2] = (weight[nb - 2] * level[nb - 2] + weight[nb - 1] * level[nb - 1]) / w;
      weight[nb - 2] = w;
      --nb;
    }
  }
  for (long long b = 0; b < nb; ++b) {
    const long long end = (b + 1 < nb) ? start[b + 1] : n;
    for (long long i = start[b]; i < end; ++i) x[i] = sign * level[b];
  }
}

// prefix isotonic regression (project_unimodal_vector.m:43-88) of y[0..len) read with stride `dir` from `y0`.
// arrays are 1-based with a sentinel at 0: level, sumwy, sumwy2 (doubles), range, sumw (ints); err optional.
__device__ void prefix_isotonic(const double* y0, long long dir, long long len, bool nonneg, double* level,
                                double* sumwy, double* sumwy2, int* range, int* sumw, double* err) {
  level[0] = -INFINITY;
  sumwy[0] = 0.0;
  sumwy2[0] = 0.0;
  sumw[0] = 0;
  range[0] = 0;
  if (err != nullptr) err[0] = 0.0;
  double cs = 0.0;  // cumsum of squares of the elements BEFORE the current one (cumsumwy2(i-1), :70)
  for (long long i = 1; i <= len; ++i) {
    const double yi = y0[(i - 1) * dir];
    level[i] = yi;
    sumwy[i] = yi;
    sumwy2[i] = yi * yi;
    sumw[i] = 1;
    range[i] = (int)i;
    while (level[i] <= level[range[i] - 1]) {
      const int mg = range[i] - 1;
      sumwy[i] = sumwy[i] + sumwy[mg];
      sumwy2[i] = sumwy2[i] + sumwy2[mg];
      sumw[i] = sumw[i] + sumw[mg];
      level[i] = sumwy[i] / (double)sumw[i];
      range[i] = range[mg];
    }
    if (err != nullptr) {
      const double levelerror = sumwy2[i] - (sumwy[i] * sumwy[i] / (double)sumw[i]);
      if (nonneg && level[i] < 0.0)
        err[i] = cs;
      else
        err[i] = levelerror + err[range[i] - 1];
    }
    cs += yi * yi;
  }
}

__device__ void unimodal_serial(const double* y, double* x, long long n, bool nonneg, double* errL, double* errR,
                                double* level, double* sumwy, double* sumwy2, int* range, int* sumw) {
  if (n <= 0) return;
  prefix_isotonic(y, 1, n, nonneg, level, sumwy, sumwy2, range, sumw, errL);
  prefix_isotonic(y + (n - 1), -1, n, nonneg, level, sumwy, sumwy2, range, sumw, errR);
  // get_best_unimodality_index (:21-32)
  double best_error = errR[n];
  long long best_idx = 1;
  for (long long i = 2; i <= n; ++i) {
    const double e = errL[i] + errR[n - (i - 1)];
    if (e < best_error) {
      best_error = e;
      best_idx = i;
    }
  }
  // left part: compute_isotonic_from_index(best_idx, iso_left) (:34-41)
  prefix_isotonic(y, 1, best_idx, nonneg, level, sumwy, sumwy2, range, sumw, nullptr);
  {
    long long idx = best_idx;
    while (idx >= 1) {
      const double v = (nonneg && level[idx] < 0.0) ? 0.0 : level[idx];
      for (long long t = range[idx]; t <= idx; ++t) x[t - 1] = v;
      idx = range[idx] - 1;
    }
  }
  const long long nr = n - best_idx;
  if (nr > 0) {
    prefix_isotonic(y + (n - 1), -1, nr, nonneg, level, sumwy, sumwy2, range, sumw, nullptr);
    long long idx = nr;
    while (idx >= 1) {
      const double v = (nonneg && level[idx] < 0.0) ? 0.0 : level[idx];
      for (long long t = range[idx]; t <= idx; ++t) x[n - t] = v;  // flipped position of right index t
      idx = range[idx] - 1;
    }
  }
}

// Thomas algorithm for (I + 2*g*Lgl) x = v, Lgl = graph Laplacian of constraints_to_prox.m:71-73
__device__ void gl_solve_serial(const double* v, double* x, long long n, double g, double* cp) {
  if (n == 1) {  // L = [1] (constraints_to_prox.m:71-73 with szm = 1)
    x[0] = v[0] / (1.0 + 2.0 * g);
    return;
  }
  const double offd = -2.0 * g;
  double d0 = 1.0 + 2.0 * g;  // corner
  cp[0] = offd / d0;
  x[0] = v[0] / d0;
  for (long long i = 1; i < n; ++i) {
    const double di = (i == n - 1) ? (1.0 + 2.0 * g) : (1.0 + 4.0 * g);
    const double den = di - offd * cp[i - 1];
    cp[i] = offd / den;
    x[i] = (v[i] - offd * x[i - 1]) / den;
  }
  for (long long i = n - 2; i >= 0; --i) x[i] = x[i] - cp[i] * x[i + 1];
}

// one CTA (32 threads) per column; column staged in shared memory when it fits (SMEM: the compiler then emits
// LDS/STS instead of generic loads), else in global scratch
template <bool SMEM>
__global__ void prox_serial_col_kernel(int kind, double p0, const double* __restrict__ X, long long ldx,
                                       double* __restrict__ out, long long ldo, long long rows,
                                       const double* rho_dev, double rho_host, double* gscratch,
                                       long long scratch_per_col, const int* __restrict__ skip, int tv_dp, SegInfo sg) {
  if (skip != nullptr && *skip != 0) return;
  extern __shared__ double sm[];
  long long row0;
  seg_select(sg, rows, row0, rho_dev);
  const double rho = load_rho(rho_dev, rho_host);
  double* base = SMEM ? sm : (gscratch + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * scratch_per_col);
  const double* x = X + (long long)blockIdx.x * ldx + row0;
  double* o = out + (long long)blockIdx.x * ldo + row0;
  const long long n = rows;
  double* y = base;          // n
  double* res = base + n;    // n
  double* work = base + 2 * n;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) y[i] = x[i];
  if (kind == PROX_TV) {
    double* rc = tv_dp ? work + 8 * n : work;
    for (long long i = threadIdx.x; i <= n; i += blockDim.x) rc[i] = (i > 0) ? 1.0 / (double)i : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (kind == PROX_TV && tv_dp) {
      tv_dp_serial(y, res, n, p0 / rho, work);
    } else if (kind == PROX_TV) {
      tv_condat_serial(y, res, n, p0 / rho, work);
    } else if (kind == PROX_NONDECREASING || kind == PROX_NONINCREASING) {
      double* level = work;
      double* weight = work + n;
      int* start = reinterpret_cast<int*>(work + 2 * n);
      pava_serial(y, res, n, kind == PROX_NONDECREASING ? 1.0 : -1.0, level, weight, start);
    } else if (kind == PROX_UNIMODAL) {
      const long long n1 = n + 1;
      double* errL = work;
      double* errR = work + n1;
      double* level = work + 2 * n1;
      double* sumwy = work + 3 * n1;
      double* sumwy2 = work + 4 * n1;
      int* range = reinterpret_cast<int*>(work + 5 * n1);
      int* sumw = range + n1 + (n1 & 1);
      unimodal_serial(y, res, n, p0 != 0.0, errL, errR, level, sumwy, sumwy2, range, sumw);
    } else if (kind == PROX_GL_SMOOTH) {
      gl_solve_serial(y, res, n, p0 / rho, work);
    }
  }
  __syncthreads();
  for (long long i = threadIdx.x; i < n; i += blockDim.x) o[i] = res[i];
}

long long serial_doubles_per_col(int kind, long long rows) {
  const long long n = rows, n1 = rows + 1;
  switch (kind) {
    case PROX_TV: return 3 * n + 1;
    case PROX_NONDECREASING:
    case PROX_NONINCREASING: return 2 * n + 2 * n + (n + 1) / 2 + 2;
    case PROX_UNIMODAL: return 2 * n + 5 * n1 + n1 + 4;
    case PROX_GL_SMOOTH: return 3 * n;
    default: return 0;
  }
}

bool is_serial_kind(int kind) {
  return kind == PROX_TV || kind == PROX_NONDECREASING || kind == PROX_NONINCREASING || kind == PROX_UNIMODAL ||
         kind == PROX_GL_SMOOTH;
}

constexpr size_t kSerialSmemLimit = 200 * 1024;

}  // namespace

size_t prox_scratch_bytes(int kind, long long rows, int cols) {
  if (kind == PROX_ORTHONORMAL)  // copy of X, rotations V, singular values
    return ((size_t)rows * cols + (size_t)cols * cols + (size_t)cols) * sizeof(double);
  if (!is_serial_kind(kind)) return 0;
  const size_t per = (size_t)serial_doubles_per_col(kind, rows) * sizeof(double);
  return (per > kSerialSmemLimit) ? per * (size_t)cols : 0;
}

size_t prox_segments_scratch_bytes(int kind, long long max_rows, int cols, int nseg) {
  if (!is_serial_kind(kind)) return 0;
  const size_t per = (size_t)serial_doubles_per_col(kind, max_rows) * sizeof(double);
  return (per > kSerialSmemLimit) ? per * (size_t)cols * (size_t)nseg : 0;
}

bool prox_supports_segments(int kind) {
  return kind == PROX_L2_BALL || kind == PROX_NONNEG_L2_BALL || kind == PROX_NONNEG_L2_SPHERE || kind == PROX_L2_REG ||
         kind == PROX_SIMPLEX_COL || kind == PROX_L1_BALL || kind == PROX_SIMPLEX_ROW || is_serial_kind(kind);
}

namespace {
int prox_apply_impl(int kind, double p0, double p1, const double* X, long long ldx, double* out, long long ldo,
                    long long rows, int cols, const double* rho_dev, double rho_host, void* scratch, cudaStream_t st,
                    const int* skip, SegInfo sg, int nseg, long long total_rows);
}

int prox_apply(int kind, double p0, double p1, const double* X, long long ldx, double* out, long long ldo,
               long long rows, int cols, const double* rho_dev, double rho_host, void* scratch, cudaStream_t st,
               const int* skip) {
  return prox_apply_impl(kind, p0, p1, X, ldx, out, ldo, rows, cols, rho_dev, rho_host, scratch, st, skip,
                         SegInfo{nullptr, rows}, 1, rows);
}

int prox_apply_segments(int kind, double p0, double p1, const double* X, long long ldx, double* out, long long ldo,
                        const long long* seg_off_dev, int nseg, long long max_rows, long long total_rows, int cols,
                        const double* rho_dev_per_seg, void* scratch, cudaStream_t st, const int* skip) {
  if (!prox_supports_segments(kind)) throw CudaError(2, "prox_apply_segments: kind " + std::to_string(kind) + " works on whole matrices");
  if (nseg <= 0) return 0;
  return prox_apply_impl(kind, p0, p1, X, ldx, out, ldo, max_rows, cols, rho_dev_per_seg, 0.0, scratch, st, skip,
                         SegInfo{seg_off_dev, max_rows}, nseg, total_rows);
}

namespace {
int prox_apply_impl(int kind, double p0, double p1, const double* X, long long ldx, double* out, long long ldo,
                    long long rows, int cols, const double* rho_dev, double rho_host, void* scratch, cudaStream_t st,
                    const int* skip, SegInfo sg, int nseg, long long total_rows) {

  if (rows <= 0 || cols <= 0) return 0;
  const dim3 colgrid((unsigned)cols, (unsigned)nseg);
  if (prox_is_elementwise(kind) || kind == PROX_NONE) {
    const long long n = rows * cols;
    const unsigned ctas = (unsigned)std::min<long long>(ceil_div(n, 256), 148 * 8);
    prox_elementwise_kernel<<<ctas, 256, 0, st>>>(kind, p0, p1, X, ldx, out, ldo, rows, cols, rho_dev, rho_host, skip);
    AO_CHECK_LAUNCH();
    return 1;
  }
  switch (kind) {
    case PROX_L2_BALL:
    case PROX_NONNEG_L2_BALL:
    case PROX_NONNEG_L2_SPHERE:
    case PROX_L2_REG:
      prox_colnorm_kernel<<<colgrid, 256, 0, st>>>(kind, p0, X, ldx, out, ldo, rows, rho_dev, rho_host, skip, sg);
      AO_CHECK_LAUNCH();
      return 1;
    case PROX_SIMPLEX_COL:
    case PROX_L1_BALL:
      prox_simplex_col_kernel<<<colgrid, 256, 0, st>>>(kind, p0, X, ldx, out, ldo, rows, skip, sg);
      AO_CHECK_LAUNCH();
      return 1;
    case PROX_SIMPLEX_ROW:
      // rows are independent and the projection does not involve rho: the stacked matrix is one problem
      prox_simplex_row_kernel<<<(unsigned)ceil_div(total_rows, 128), 128, 0, st>>>(p0, X, ldx, out, ldo, total_rows, cols, skip);
      AO_CHECK_LAUNCH();
      return 1;
    default: break;
  }
  if (kind == PROX_ORTHONORMAL) {
    // project_ortho.m:3-4: Z = U*V' of the thin SVD = polar factor, by one-sided Jacobi: X*W = U*diag(sig)
    if (rows < cols) throw CudaError(2, "'orthonormal' needs at least as many rows as columns on device");
    if (scratch == nullptr) throw CudaError(1, "prox_apply: scratch buffer required for 'orthonormal'");
    double* S = static_cast<double*>(scratch);
    double* W = S + (size_t)rows * cols;
    double* sig = W + (size_t)cols * cols;
    AO_CUDA(cudaMemcpy2DAsync(S, (size_t)rows * 8, X, (size_t)ldx * 8, (size_t)rows * 8, (size_t)cols,
                              cudaMemcpyDeviceToDevice, st));
    int n = jacobi_onesided(S, rows, cols, W, sig, st, skip);
    n += scale_cols_inv(S, rows, cols, sig, st, skip);
    n += dgemm_small(0, 1, rows, cols, cols, 1.0, nullptr, S, rows, W, cols, 0.0, out, ldo, st, skip);
    return n;
  }
  if (is_serial_kind(kind)) {
    // TV: the dynamic-programming algorithm when its 11n+1 doubles (column, result, 8n of queue / bounds, reciprocal
    // table) fit in shared memory, else the direct algorithm (3n+1 doubles; shared memory up to ~8500 rows, global
    // scratch beyond)
    const int tv_dp = (kind == PROX_TV && ((size_t)rows * 11 + 1) * sizeof(double) <= kSerialSmemLimit) ? 1 : 0;
    const long long per = tv_dp ? 11 * rows + 1 : serial_doubles_per_col(kind, rows);
    const size_t bytes = (size_t)per * sizeof(double);
    const int use_smem = bytes <= kSerialSmemLimit;
    if (!use_smem && scratch == nullptr) throw CudaError(1, "prox_apply: scratch buffer required for this size");
    if (use_smem && bytes > 48 * 1024)
      ensure_dynamic_smem(reinterpret_cast<const void*>(prox_serial_col_kernel<true>), kSerialSmemLimit);
    if (use_smem)
      prox_serial_col_kernel<true><<<colgrid, 32, bytes, st>>>(kind, p0, X, ldx, out, ldo, rows, rho_dev, rho_host,
                                                               static_cast<double*>(scratch), per, skip, tv_dp, sg);
    else
      prox_serial_col_kernel<false><<<colgrid, 32, 0, st>>>(kind, p0, X, ldx, out, ldo, rows, rho_dev, rho_host,
                                                            static_cast<double*>(scratch), per, skip, 0, sg);
    AO_CHECK_LAUNCH();
    return 1;
  }
  throw CudaError(2, "prox kind " + std::to_string(kind) + " is not implemented on device");
}
}  // namespace

}  // namespace aoadmm
