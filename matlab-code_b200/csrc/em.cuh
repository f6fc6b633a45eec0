// em.cuh - EM imputation pass (functions/cmtf_fun_AOADMM.m:408-441) + masked objective sums (:1224-1226, :1249-1252).
#pragma once
#include "common.cuh"

namespace aoadmm {

struct EmArgs {
  double* X;               // object data, element (i,j,k) at i + ldI*(j + J*k); missing entries are overwritten
  const uint8_t* mask;     // same indexing; 1 = observed, 0 = missing (Z.miss)
  long long ldI;
  long long I, J;
  int K;
  const double* Fi;        // I x R column-major
  long long ldFi;
  const double* Fj;        // J x R
  long long ldFj;
  const double* Fk;        // K x R, or nullptr (matrices / PARAFAC2: K = 1, no third factor)
  long long ldFk;
  int R;
  int impute;              // 0: only the sums (iteration-0 objective)
  double* partials;        // >= em_partials_doubles(args)
  double* fkT = nullptr;   // optional scratch, >= em_fkT_doubles(K, R): enables the pipelined kernel for K >= 8
};
inline size_t em_fkT_doubles(long long K, int R) { return (size_t)K * (size_t)((R + 3) / 4 * 4); }

// Khatri-Rao product of the trailing factors of an N-way (N > 3) object, first factor fastest:
//   out(k, r) = prod_q F_q(k_q, r),  k = k_0 + d_0*(k_1 + d_1*(...)),  out: K x R (ld = K), K = prod d_q
// so that the object can be treated as I x J x K by em_pass (the merged index is exactly the memory order).
struct KrArgs {
  int n;                   // number of factors (<= 6)
  const double* F[6];
  long long ld[6];
  long long d[6];
};
int em_khatri_rao(const KrArgs& a, double* out, long long K, int R, cudaStream_t st);

size_t em_partials_doubles(const EmArgs& a);
// sums_out[0..4] = sum_missing (m-x)^2, sum_missing x^2, sum_observed x*m, sum_observed m^2, sum_observed (x-m)^2
// ([4] is the PARAFAC2 objective (:1249-1252) and is only evaluated for objects without a third factor, Fk == nullptr;
// objects with a third factor report 0 there - their objective uses [2] and [3], :1224-1226)
int em_pass(const EmArgs& a, double* sums_out, cudaStream_t st);

// *out = sum of squares of the elements (i < I of every column of length ld, `slab` columns) with mask != 0
// (mask == nullptr: all);  partials: >= 148*8 doubles.  Znorm_const of cmtf_AOADMM.m:124-156 on device.
int object_norm2(const double* X, const uint8_t* mask, long long I, long long ld, long long slab, double* partials,
                 double* out, cudaStream_t st);

}  // namespace aoadmm
