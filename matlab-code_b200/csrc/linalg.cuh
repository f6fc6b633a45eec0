// linalg.cuh - small dense FP64 building blocks for the linear-coupling ADMM loops
// (functions/cmtf_fun_AOADMM.m:278-389 precompute, :698-1075 ADMM_coupled_case1..5, :1118-1210 residuals).
// These loops work on the (small) transformation matrices H of Z.coupling.coupl_trafo_matrices and on factor-sized
// operands; they are launch-latency bound, so the kernels favour determinism and generality over peak rate.
#pragma once
#include "smallops.cuh"

namespace aoadmm {

// C (M x N, ldc) = alpha * (alpha_dev ? *alpha_dev : 1) * op(A) * op(B) + beta * C
//   op(A) is M x K: A stored M x K (transA = 0, lda >= M) or K x M (transA = 1, lda >= K); likewise op(B) is K x N.
// Fixed summation order over K => bit-reproducible.  `skip` (device int, may be null) != 0 turns the launch into a no-op.
int dgemm_small(int transA, int transB, long long M, long long N, long long K, double alpha, const double* alpha_dev,
                const double* A, long long lda, const double* B, long long ldb, double beta, double* C, long long ldc,
                cudaStream_t st, const int* skip);

// The same product for a LONG reduction and a small result (e.g. V'W with V, W of 4096 x 80): the reduction range is
// split over CTAs (dgemm_splitk_count(M, N, K) parts) into `ws` (>= count * M * N doubles) and the partial tiles are
// added in a fixed order - two launches, bit-reproducible.  Falls back to dgemm_small when one part is enough.
int dgemm_splitk_count(long long M, long long N, long long K);
int dgemm_small_splitk(int transA, int transB, long long M, long long N, long long K, double alpha, const double* A,
                       long long lda, const double* B, long long ldb, double beta, double* C, long long ldc, double* ws,
                       cudaStream_t st, const int* skip);

// out[i] = sum_t coef[t] * (coef_dev[t] ? *coef_dev[t] : 1) * x[t][i],  i < n   (up to 5 terms; out may alias any x[t])
struct LinTerm {
  const double* x;
  double coef;
  const double* coef_dev;
};
int lincomb(double* out, long long n, const LinTerm* terms, int nterms, cudaStream_t st, const int* skip);

// out (cols x rows, ld = cols) = in' (in: rows x cols, ld = rows)
int transpose_small(const double* in, long long rows, long long cols, double* out, cudaStream_t st, const int* skip);

// One-sided (Hestenes) Jacobi on an m x n column-major matrix S (ld = m), executed by ONE CTA:
//   S <- S*V with mutually orthogonal columns, V (n x n, ld = n) the accumulated rotations (V'V = I).
//   sig[j] = ||S(:,j)||_2 afterwards.
// For S = H (q x n) this yields the eigen-decomposition H'H = V diag(sig^2) V'; for a symmetric positive definite S
// it yields S = V diag(sig) V'.
int jacobi_onesided(double* S, long long m, int n, double* V, double* sig, cudaStream_t st, const int* skip = nullptr);

// S(:,j) *= 1/sig[j]  (0 when sig[j] == 0)
int scale_cols_inv(double* S, long long rows, int cols, const double* sig, cudaStream_t st, const int* skip);

// Y(i,:) /= (2*eta/rho * lam[i] + 1)   (prox of the quadratic regulariser in the eigen-basis of L,
// constraints_to_prox.m:62-66: (2*eta/rho*L + I) \ x)
int quad_scale_rows(double* Y, long long rows, int cols, const double* lam, double eta, const double* rho_dev,
                    double rho_host, cudaStream_t st, const int* skip);

// X(i,j) = At(i,j) / (half_rho * (lam[i] + shift) + mu[j])   with half_rho = *rho_dev / 2   (Sylvester in eigen-bases)
int sylvester_scale(double* X, const double* At, long long rows, int cols, const double* lam, double shift,
                    const double* mu, const double* rho_dev, cudaStream_t st, const int* skip);

// out[0] = sum_t coef[t] * *coef_dev[t] ;  out[1] = 1 / out[0]     (sum of the rho's of a coupling group, :663-675)
int sum_recip(double* out, const LinTerm* terms, int nterms, cudaStream_t st);

// Residuals of eval_res_ADMM_coupl_case1..5 (:1118-1210) and eval_res_ADMM_constr (:1079-1096) from reduced squared
// norms `red` (indices per mode below), exit test of the while loop (:705 ...), iteration count.
struct LinFinMode {
  int i_pr_num, i_pr_den;   // ||G(F) - D(Delta)||^2 , ||G(F)||^2 or ||F||^2
  int i_du_num, i_mu;       // ||D(Delta - Delta_old)||^2 , ||mu_Delta||^2
  int constrained;
  int i_fz, i_fn, i_zz, i_muz;  // ||F-Z||^2, ||F||^2, ||Z-Zold||^2, ||mu_Z||^2
};
struct LinFin {
  int nmodes;
  LinFinMode m[kMaxGroup];
};
int lin_finalize(const LinFin& fin, const double* red, const InnerTol& tol, InnerCtl* ctl, cudaStream_t st);

}  // namespace aoadmm
