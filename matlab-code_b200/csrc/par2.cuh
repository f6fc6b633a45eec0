// par2.cuh - device kernels of the PARAFAC2 block (functions/cmtf_fun_AOADMM.m:157-250 precompute,
// :509-589 ADMM_B_Parafac2, :602-606 / :638-645 row-wise mode-C solves, :1254-1265 + :1351-1362 objective terms).
//
// Layout: the K slices X_k (I x J_k, column-major) of one PARAFAC2 object are stored side by side as one
// I x Jtot column-major matrix (Jtot = sum_k J_k, slice k = columns joff[k] .. joff[k+1]-1), and every per-slice
// J_k x R matrix (B_k, Z_Bk, mu_Z_Bk, P_k, mu_DeltaB_k) is stored stacked as one Jtot x R column-major matrix
// (slice k = rows joff[k] .. joff[k+1]-1).  With this layout the two big products of the block are plain
// matrix-block products served by the DMMA kernels of mttkrp.cu:
//     mode A :  sum_k X_k B_k diag(c_k)   = Xall * W          W(j,:)  = B(j,:) .* C(seg(j),:)
//     mode B :  X_k' A diag(c_k)  (all k) = (Xall' * A) .* C(seg(j),:)
//     mode C :  diag(A' X_k B_k)          = segmented column dot of (Xall' * A) with B
// and everything that is per-slice R x R work (Hadamard, rho_k, Cholesky, inverse, polar factor P_k) is one CTA
// per slice.
#pragma once
#include "smallops.cuh"

namespace aoadmm {

struct Par2Layout {
  int K = 0;
  int R = 0;
  long long Jtot = 0;
  long long Jmax = 0;
  const long long* joff = nullptr;  // device, K+1 entries
  const int* seg = nullptr;         // device, Jtot entries: slice index of every stacked row
  // Multi-GPU: the slices are sharded over the ranks in contiguous ranges.  Every per-slice array keeps its GLOBAL
  // indexing (slice k at k, stacked row j at j), a rank only touches slices k0 .. k1-1 = stacked rows jlo .. jhi-1;
  // one GPU: k0 = 0, k1 = K, jlo = 0, jhi = Jtot.
  int k0 = 0, k1 = 0;
  long long jlo = 0, jhi = 0;
};

// out(j,r) = scale * in(j,r) * C(seg(j), r)  [+ add_scale * addend(j,r)]     (all stacked Jtot x R, ld = Jtot)
int par2_scale_rows(const Par2Layout& L, double* out, const double* in, const double* C, long long ldc, double scale,
                    const double* addend, double add_scale, cudaStream_t st);

// G2[k] (R x R) = B_k' * B_k for every slice                                   (:72, :217)
int par2_batched_gram(const Par2Layout& L, const double* Bst, double* G2, cudaStream_t st);

// Csum (R x R) = sum_k diag(c_k) * G2[k] * diag(c_k)                           (:164)
int par2_modeA_had(const Par2Layout& L, const double* G2, const double* C, long long ldc, double* Csum, cudaStream_t st);

struct Par2SysArgs {
  int mode;                // 2: B_k systems (:192-213), 3: C rows (:220-243)
  const double* G1;        // A'A (R x R)
  const double* G2;        // per-slice B_k'B_k (K x R x R)     (mode 3)
  const double* C;         // C factor (K x R), ld = ldc
  long long ldc;
  double weight, ridge, bsum_half, rho_scale;
  int n_rho_terms;         // number of rho_k/2*I terms
  // mode 3 only: right-hand side a_k = w * diag(A' X_k B_k) (+ bsum/2 * C(k,:)) from T = Xall'*A and B
  const double* T;         // Jtot x R
  const double* Bst;       // Jtot x R
  double* rhs;             // K x R (ld = K): row k = a_k
  int ls_direct;           // mode 3, unconstrained + uncoupled: C(k,:) = B_k \ a_k written to fac_out (:236)
  double* fac_out;         // K x R (ld = K)
  // outputs
  double* rho_k;           // K
  double* Binv;            // K x R x R
  double* Bsys;            // optional: the assembled system matrices themselves (K x R x R)
  int no_factor;           // only assemble (rho_k, Bsys, rhs): the caller factors a larger system (coupling type 1)
  const double* HHt;       // optional (mode 3, coupling type 2): B_k += rho_k/2 * H*H' (:307)
  InnerCtl* ctl;           // err = 3 when a system is not positive definite
  double* gws = nullptr;   // R > 64: global workspace, 2 R^2 doubles per slice (indexed by the global slice number)
};
int par2_sys_prep(const Par2Layout& L, const Par2SysArgs& a, cudaStream_t st);

// out[0] = mean_k rho_k, out[1] = sum_k rho_k  (the scalars the type-1 coupling uses for a vector rho, :712, :742)
int par2_rho_stats(const double* rho_k, int K, double* out, cudaStream_t st);
// B2 (n x n, n = K*R, index k*R + r) = blkdiag(Bsys_k) + rhoC/2 * kron(HtH, I_R) [+ rhoC/2 * I]   (:283-293)
int par2_assemble_B2(const double* Bsys, const double* HtH, const double* rhoC_dev, int constrained, int K, int R,
                     double* B2, cudaStream_t st);
// x = (L L') \ a for one right-hand side stored as a K x R column-major matrix read in the order k*R + r (:721-722)
int par2_chol_solve_vec(const double* L, int K, int R, const double* a, double* x, cudaStream_t st, const int* skip);

// rho_max = max_k rho_k (update_constraint uses max(rho) when rho is a vector, :1423-1424)
int par2_rho_max(const double* rho_k, int K, double* out, cudaStream_t st);

// ---- third PARAFAC2 mode inside the linear couplings 2, 3, 4: row-wise rho_k and row-wise systems (:783-790, :848-855,
// :914-921, :940-960).  All matrices are rows x cols column-major with leading dimension rows.
// out(k,:) = A(k,:) + rho_k/2 * ( X(k,:) [+ Z(k,:) - muZ(k,:)] )                      (A_inner of row k)
int par2_rows_ainner(double* out, const double* A, const double* X, const double* Z, const double* muZ,
                     const double* rho_k, long long rows, int cols, cudaStream_t st, const int* skip);
// out(k,:) = in(k,:) * M_k   with M_k the q x q matrix at Minv + k*q*q (symmetric)   ((A_inner/L')/L, BB(k,:)/(AA+AA_k))
int par2_rows_apply(double* out, const double* in, const double* Minv, long long rows, int q, cudaStream_t st,
                    const int* skip);
// out(k,:) = rho_k * in(k,:)
int par2_rows_scale(double* out, const double* in, const double* rho_k, long long rows, int cols, cudaStream_t st,
                    const int* skip);
// D(k,:) (+)= w_k * (S(k,:) + mu(k,:)),  wsum[k] (+)= w_k,  w_k = rho_k[k] (rho_k != nullptr) or *rho_scalar  (:808-813)
int par2_rows_weighted_accum(double* D, double* wsum, const double* S, const double* mu, const double* rho_scalar,
                             const double* rho_k, long long rows, int cols, int first, cudaStream_t st, const int* skip);
// D(k,:) /= wsum[k]                                                                    (:815)
int par2_rows_divide(double* D, const double* wsum, long long rows, int cols, cudaStream_t st, const int* skip);
// Minv_k = inv(AA + rho_k * AAA), k < K (q x q each, q <= 64); ctl->err = 3 when a matrix is not positive definite  (:957-960)
int par2_rowsys_inverse(const double* AA, const double* AAA, const double* rho_k, int K, int q, double* Minv, InnerCtl* ctl,
                        cudaStream_t st);

struct Par2BArgs {
  const double* A;         // stacked right-hand sides A_k (Jtot x R)
  const double* Binv;      // K x R x R
  const double* rho_k;     // K
  double* B;               // stacked B_k        (G.fac{m})
  double* P;               // stacked P_k        (G.P{p})
  double* mu;              // stacked mu_DeltaB  (G.mu_DeltaB{p})
  double* DeltaB;          // R x R              (G.DeltaB{p})
  double* Z;               // stacked constraint_fac or nullptr
  double* muZ;             // stacked constraint_dual_fac or nullptr
  int con_active;          // Z.constrained_modes(m) && iter >= iter_start_PAR2Bkconstraint
  int prox_kind;
  double p0, p1;
  double* PDold;           // scratch Jtot x R: P_k*DeltaB before the update
  double* contrib;         // scratch K x R x R: rho_k P_k'(B_k+mu_k)
  double* gM;              // scratch Jtot x R (used when a slice does not fit shared memory)
  double* gS;              // scratch Jtot x R
  double* norms;           // scratch K x 8
  double* Znew;            // deferred prox output (stacked) or nullptr
  double* Vprev;           // K x R x R: Jacobi rotations of the previous inner iteration (warm start)
};
// one inner iteration of ADMM_B_Parafac2 is  step1 -> deltaB -> step2a -> [prox of every slice] -> step2b
// warm != 0: start the Jacobi iteration from the rotations of the previous call (a.Vprev)
int par2_B_step1(const Par2Layout& L, const Par2BArgs& a, const InnerCtl* ctl, int warm, cudaStream_t st);
// sums_out == nullptr: DeltaB = sum_k contrib_k / sum_k rho_k over all slices (one GPU).  sums_out != nullptr (sharded
// slices): only the local sums are written, sums_out[0..R*R-1] = sum_k contrib_k, sums_out[R*R] = sum_k rho_k; the caller
// all-reduces them and finishes with par2_B_deltaB_finish.
int par2_B_deltaB(const Par2Layout& L, const Par2BArgs& a, const InnerCtl* ctl, cudaStream_t st, double* sums_out = nullptr);
int par2_B_deltaB_finish(const Par2Layout& L, const Par2BArgs& a, const double* sums, const InnerCtl* ctl, cudaStream_t st);
// coupling part: mu_k += B_k - P_k DeltaB ; per-slice norms 0..3
int par2_B_step2a(const Par2Layout& L, const Par2BArgs& a, const InnerCtl* ctl, cudaStream_t st);
// V = B + muZ (input of a non element-wise prox)
int par2_B_form_prox_input(const Par2Layout& L, const Par2BArgs& a, double* V, const InnerCtl* ctl, cudaStream_t st);
// constraint part (element-wise prox inline, or Z taken from a.Znew) + residuals + exit test
// res_out == nullptr: the last CTA averages the residual ratios over all slices and evaluates the exit test (one GPU).
// res_out != nullptr (sharded slices): res_out[0..3] = the local sums of the four ratios; the caller all-reduces them
// and par2_B_finalize evaluates the exit test (identically on every rank).
int par2_B_step2b(const Par2Layout& L, const Par2BArgs& a, const InnerTol& tol, InnerCtl* ctl, unsigned* counter,
                  cudaStream_t st, double* res_out = nullptr);
int par2_B_finalize(const Par2Layout& L, const double* res_sums, const InnerTol& tol, InnerCtl* ctl, cudaStream_t st);
// zero every row outside [lo, hi) of a rows x cols column-major matrix (ld = rows): followed by an all-reduce this is the
// all-gather of row ranges owned by different ranks
int zero_rows_outside(double* M, long long rows, int cols, long long lo, long long hi, cudaStream_t st);

// tPARAFAC2 prox over all slices at once (t_smoothness_prox.m:23-56): for every element (j,r) the K values solve the
// tridiagonal system  diag(4l+rho_k; ends 2l+rho_k), off-diagonals -2l, right-hand side rho_k*V_k(j,r)  by the Thomas
// algorithm.  Needs equal J_k.  dp (K doubles) is scratch for the eliminated diagonal.
int par2_tsmooth_prox(const Par2Layout& L, const double* V, const double* rho_k, double eta, double* dp, double* out,
                      const InnerCtl* ctl, cudaStream_t st);
constexpr int RED_TSMOOTH = 100;  // reg kind for par2_seg_norms: ||B_k - B_{k-1}||^2 (t_smoothness_penalty.m:5-9)

// per-slice objective terms: out[k*4 + {0,1,2,3}] = ||B_k||^2, ||B_k - Z_k||^2, ||B_k - P_k DeltaB||^2, reg(B_k)
int par2_seg_norms(const Par2Layout& L, const double* Bst, const double* Z, const double* P, const double* DeltaB,
                   int reg_kind, double* out, cudaStream_t st);

// res = sum_k || X_k - A diag(c_k) B_k' ||_F^2  (explicit residual of :1254-1265); partials >= 148*8 doubles
// (sharded slices: the sum over this rank's slices only; X is addressed with global column indices)
int par2_residual(const Par2Layout& L, const double* X, long long ldX, long long I, const double* A, long long ldA,
                  const double* Bst, const double* C, long long ldc, double* partials, unsigned* counter, double* res,
                  cudaStream_t st);

size_t par2_step1_smem_bytes(long long Jmax, int R);

}  // namespace aoadmm
