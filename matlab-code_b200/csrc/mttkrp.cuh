// mttkrp.cuh - dense MTTKRP on FP64 tensor cores (DMMA.8x8x4) fed by TMA, sm_100a.
//
// Replaces Tensor Toolbox mttkrp(X,U,n) as called at functions/cmtf_fun_AOADMM.m:97 (and cp_func.m:47)
// and the matrix-block products at cmtf_fun_AOADMM.m:106-113.  The Khatri-Rao operand is never
// materialised: its rows are generated from the (small) factor matrices while the tensor tile sits in
// shared memory.
#pragma once
#include "common.cuh"

namespace aoadmm {

// Transposed, zero-padded, column-chunked copy of a factor matrix F (rows x R):
//   data[(c * rows_pad + row) * ldc + col]  = F(row, c*NC + col)     (0 outside)
// Row-contiguous so that a tile of consecutive rows is ONE 1-D bulk copy, and ldc = NC+4 (ldc*8 a
// multiple of 16 bytes, ldc mod 16 in {4,12}) so that DMMA B-fragment reads are bank-conflict free.
struct PackedFactor {
  double* data = nullptr;
  int64_t rows = 0, rows_pad = 0;
  int R = 0, NC = 0, nchunk = 0, ldc = 0;
  size_t bytes() const { return (size_t)nchunk * rows_pad * ldc * sizeof(double); }
};

// chunk width used for rank R (columns handled by one CTA)
inline int mttkrp_chunk_cols(int R) { return R <= 8 ? 8 : (R <= 16 ? 16 : (R <= 32 ? 32 : 64)); }

// A dense column-major tensor viewed as 3-way I x J x K (leading dimension ldI, even), with the two
// TMA descriptors used by the kernels.
struct Tensor3 {
  const double* X = nullptr;
  int64_t I = 0, J = 0, K = 0, ldI = 0;
  CUtensorMap map_lead;   // box 16 x 32 x 1  (i, j, k)
  CUtensorMap map_inner;  // box 16 x 128 x 1
};

// Scratch owned by the caller (engine): partial-result workspace
struct MttkrpWorkspace {
  double* ws = nullptr;
  size_t ws_bytes = 0;
};

void make_tensor3(Tensor3& t, const double* X, int64_t I, int64_t J, int64_t K, int64_t ldI);
// generic 3-D FP64 tensor map: dims (dims[0] contiguous), byte strides of dims 1 and 2, box; with the 128-byte swizzle
// (box[0] = 16 doubles) or as a plain dense box
void encode_map3(CUtensorMap* map, const double* X, const uint64_t dims[3], const uint64_t strides_bytes[2],
                 const uint32_t box[3], bool swizzle128 = true);

// allocate + describe (rows_pad = rows rounded up to 128, plus one extra tile)
void packed_factor_alloc(PackedFactor& p, int64_t rows, int R);
void packed_factor_free(PackedFactor& p);
// pack F (rows x R, leading dimension ld) into p; if F == nullptr fills ones (used for matrices: K = 1)
void packed_factor_pack(const PackedFactor& p, const double* F, int64_t ld, cudaStream_t st, const int* skip);
// In-place conversion of a freshly packed factor into the operand format of the opt-in TF32 kernels (precision = 1):
// high word of every 8-byte slot = TF32-rounded float, low word 0.  Only operand 0 (f0) of mttkrp3 takes this format.
void packed_factor_to_tf32(const PackedFactor& p, cudaStream_t st, const int* skip);
// pack the Khatri-Rao product of two factors (first varies fastest): KR(a + Ra*b, r) = Fa(a,r)*Fb(b,r)
void packed_factor_pack_kr(const PackedFactor& p, const double* Fa, int64_t rows_a, int64_t lda, const double* Fb,
                           int64_t rows_b, int64_t ldb, cudaStream_t st, const int* skip);

size_t mttkrp_workspace_bytes(const Tensor3& t, int R);

// out (rows x R, leading dimension ldout) = scale * MTTKRP
//   pos = 0: rows = I, out(i,r) = sum_{j,k} X(i,j,k) Fj(j,r) Fk(k,r)      (lead kernel)
//   pos = 1: rows = J, out(j,r) = sum_{i,k} X(i,j,k) Fi(i,r) Fk(k,r)      (inner kernel, epilogue 0)
//   pos = 2: rows = K, out(k,r) = sum_{i,j} X(i,j,k) Fi(i,r) Fj(j,r)      (inner kernel, epilogue 1)
// The two factors are passed in packed form in natural order (for pos=0: Fj,Fk; pos=1: Fi,Fk; pos=2: Fi,Fj).
// `accumulate` adds into out instead of overwriting (used by nobody on the reference path).
// Returns the number of kernel launches issued.
// Tbuf != nullptr with pos == 1 additionally emits the partial contraction T(j,k,r) = sum_i X(i,j,k) F0(i,r)
// (mttkrp_T_bytes bytes) for a later mttkrp3_from_T (dimension-tree reuse: one tensor pass serves modes 2 and 3).
int mttkrp3(const Tensor3& t, int pos, const PackedFactor& f0, const PackedFactor& f1, int R, double scale,
            double* out, int64_t ldout, const MttkrpWorkspace& w, cudaStream_t st, const int* skip,
            double* Tbuf = nullptr, int precision = 0);  // precision 1: opt-in TF32 tensor-core path (FP32 accumulate)
size_t mttkrp_T_bytes(const Tensor3& t, int R);

// ---- opt-in reduced-precision path on tcgen05 / TMEM (mttkrp_tc.cu) --------------------------------------------------
// Scratch for the low-precision copy of the contracted factor (packed per call in the swizzled UMMA operand layout).
struct TcOperand {
  uint8_t* data = nullptr;
  size_t bytes = 0;
};
size_t tc_operand_bytes(int64_t rows, int R);
void tc_operand_alloc(TcOperand& op, int64_t max_rows, int R);
void tc_operand_free(TcOperand& op);
// Same contract as mttkrp3 for a 3-way view, from the plain column-major FP64 factors:
//   pos 0: out(i,:) = sum_k Fe(k,:) .* sum_j X(i,j,k) F0(j,:)      F0 = Fj (J x R), Fe = Fk (K x R)
//   pos 1: out(j,:) = sum_k Fe(k,:) .* sum_i X(i,j,k) F0(i,:)      F0 = Fi (I x R), Fe = Fk (K x R); Tbuf != nullptr also
//          emits T(j,k,:) = sum_i X(i,j,k) F0(i,:) for mttkrp3_from_T
//   pos 2: out(k,:) = sum_j Fe(j,:) .* sum_i X(i,j,k) F0(i,:)      F0 = Fi (I x R), Fe = Fj (J x R)
// precision 1: TF32 operands, 2: BF16 operands; FP32 accumulation in TMEM per slab, FP64 across slabs.
int mttkrp3_tc(const Tensor3& t, int pos, const TcOperand& op, const double* F0, int64_t ld0, const double* Fe, int64_t ldfe,
               int R, double scale, double* out, int64_t ldout, const MttkrpWorkspace& w, cudaStream_t st, const int* skip,
               double* Tbuf, int precision);
// out(k,r) = scale * sum_j T(j,k,r) * Fj(j,r)   (Fj: J x R column-major, leading dimension ldf)
int mttkrp3_from_T(const Tensor3& t, const double* Tbuf, int R, const double* Fj, int64_t ldf, double scale,
                   double* out, int64_t ldout, cudaStream_t st, const int* skip);

}  // namespace aoadmm
