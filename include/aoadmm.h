/*
 * aoadmm.h - C ABI of the B200-native AO-ADMM engine.
 *
 * Drop-in boundary: this library replaces the single call
 *     [Fac,out] = cmtf_fun_AOADMM(Z,Znorm_const,G,fh,gh,lscalar,uscalar,options)
 * made at functions/cmtf_AOADMM.m:193 of the reference (signature: functions/cmtf_fun_AOADMM.m:1).
 * The caller (MATLAB via the MEX gateway in matlab-code_b200/matlab/, or Python via ctypes)
 * still builds Z / options / G exactly as before (init_coupled_AOADMM_CMTF.m, constraints_to_prox.m);
 * the gateway flattens those structs into the plain-C structs below.
 *
 * Conventions
 *  - every matrix / tensor is IEEE double, column-major (MATLAB layout), dense;
 *  - mode ids, coupling ids are 1-based exactly as in Z.modes / Z.coupling.lin_coupled_modes;
 *  - the caller owns every host buffer; the library owns all device memory;
 *  - every entry point returns an aoadmm_status (0 = ok); no C++ exception crosses the ABI;
 *  - a handle is not thread-safe; one host thread drives one handle.  Multi-GPU comes in two forms:
 *      aoadmm_create_multi : ONE caller process / thread (a MATLAB session) hands over the whole problem and the
 *                            library drives n_gpus devices itself (the form the MEX gateway uses);
 *      aoadmm_create + aoadmm_dist : one process per GPU (torchrun / MPI launchers), each passing its own slab.
 */
#ifndef AOADMM_H_
#define AOADMM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 1: first release; 2: Z.miss masks, device Znorm_const, engine options (dimtree, fuse_inner, graph,
 * mttkrp_precision), out.f_rel_missing; 3: aoadmm_nvecs, aoadmm_object_mttkrp (no struct changed since 2);
 * 4: aoadmm_create_multi / aoadmm_gpu_count / aoadmm_comm_release, communicators cached per process,
 *    aoadmm_out.non_finite_mode appended (a non-finite inner residual no longer aborts the run). */
#define AOADMM_ABI_VERSION 4

typedef struct aoadmm_handle aoadmm_handle;

/* Error convention: replaces MATLAB error() (cmtf_AOADMM.m:37,:48,:61; chol failure; the
 * 'MATLAB:nearlySingularMatrix' warning-as-error at cmtf_fun_AOADMM.m:83). */
typedef enum {
  AOADMM_OK = 0,
  AOADMM_ERR_INVALID_ARG = 1,
  AOADMM_ERR_UNSUPPORTED = 2,          /* 'custom' constraint, non-Frobenius loss, ... */
  AOADMM_ERR_NOT_POSITIVE_DEFINITE = 3, /* chol() would have thrown (cmtf_fun_AOADMM.m:142 ...) */
  AOADMM_ERR_NON_FINITE = 4,            /* kept for ABI stability; no entry point returns it since version 4 */
  AOADMM_ERR_CUDA = 5,
  AOADMM_ERR_NCCL = 6,
  AOADMM_ERR_OOM = 7,
  AOADMM_ERR_NO_DEVICE = 8
} aoadmm_status;

/* Z.model{p} */
typedef enum { AOADMM_MODEL_CP = 0, AOADMM_MODEL_PAR2 = 1 } aoadmm_model;

/* Z.constraints{m}{1}: the named specs of functions/constraints_to_prox.m:13-91
 * ("List of constraints and regularizations.txt" 1-21). */
typedef enum {
  AOADMM_CON_NONE = 0,
  AOADMM_CON_NONNEG = 1,            /* 'non-negativity'                      :13 */
  AOADMM_CON_BOX = 2,               /* 'box', l=p0, u=p1                     :15 */
  AOADMM_CON_SIMPLEX_COL = 3,       /* 'simplex column-wise', eta=p0         :19 */
  AOADMM_CON_SIMPLEX_ROW = 4,       /* 'simplex row-wise', eta=p0            :22 */
  AOADMM_CON_NONDECREASING = 5,     /* 'non-decreasing'                      :25 */
  AOADMM_CON_NONINCREASING = 6,     /* 'non-increasing'                      :27 */
  AOADMM_CON_UNIMODAL = 7,          /* 'unimodality', nonneg flag=p0         :29 */
  AOADMM_CON_L1_BALL = 8,           /* 'l1-ball', eta=p0                     :32 */
  AOADMM_CON_L2_BALL = 9,           /* 'l2-ball', eta=p0                     :35 */
  AOADMM_CON_NONNEG_L2_BALL = 10,   /* 'non-negative l2-ball', eta=p0        :38 */
  AOADMM_CON_NONNEG_L2_SPHERE = 11, /* 'non-negative l2-sphere'              :41 */
  AOADMM_CON_ORTHONORMAL = 12,      /* 'orthonormal'                         :44 */
  AOADMM_CON_L1_REG = 13,           /* 'l1 regularization', eta=p0           :46 */
  AOADMM_CON_L0_REG = 14,           /* 'l0 regularization', eta=p0           :50 */
  AOADMM_CON_L2_REG = 15,           /* 'l2 regularization', eta=p0           :54 */
  AOADMM_CON_RIDGE = 16,            /* 'ridge', eta=p0                       :58 */
  AOADMM_CON_QUADRATIC = 17,        /* 'quadratic regularization', eta=p0, L=matrix :62 */
  AOADMM_CON_GL_SMOOTH = 18,        /* 'GL smoothness', eta=p0               :68 */
  AOADMM_CON_TV = 19,               /* 'TV regularization', eta=p0           :78 */
  AOADMM_CON_TPARAFAC2 = 20,        /* 'tPARAFAC2', eta=p0                   :82 */
  AOADMM_CON_CUSTOM = 21            /* 'custom' (function handles) -> AOADMM_ERR_UNSUPPORTED :86 */
} aoadmm_constraint_kind;

typedef struct {
  int32_t kind;          /* aoadmm_constraint_kind */
  double p0, p1;         /* numeric parameters, see above */
  const double *matrix;  /* QUADRATIC: n x n col-major L; else NULL */
  int64_t matrix_n;
} aoadmm_constraint;

/* One entry of Z.object / Z.model / Z.modes / Z.weights + Znorm_const{p} (cmtf_AOADMM.m:124-156). */
typedef struct {
  int32_t model;          /* aoadmm_model */
  int32_t order;          /* number of modes of this object: 2 (matrix), >=3 (tensor); PAR2: 3 */
  const int32_t *modes;   /* `order` global 1-based mode ids (Z.modes{p}) */
  double weight;          /* Z.weights(p) */
  double znorm_const;     /* Znorm_const{p} = ||X_p||_F^2 (observed entries only with Z.miss); NaN = compute it on
                             the device from the data (cmtf_AOADMM.m:124-156) */
  /* CP: dense column-major data of THIS RANK's slab: extents size(modes[0..order-2]) x shard_extent.
   * With one GPU shard_offset = 0 and shard_extent = size of the last mode. */
  const double *data;
  int64_t shard_offset;   /* first index (0-based) of the last mode held by this rank */
  int64_t shard_extent;   /* number of last-mode indices held by this rank */
  /* PAR2: K slices X_k (I x J_k col-major); every rank passes all slices (small objects) */
  const double *const *slices;
  int32_t n_slices;
  /* Z.miss{p} (cmtf_AOADMM.m:68-121): NULL = no missing data; else one byte per element with the layout of
   * `data` (CP, this rank's slab) / of each slice (PAR2): 1 = observed, 0 = missing.  Missing entries are
   * re-imputed from the model after every outer iteration (cmtf_fun_AOADMM.m:408-441). */
  const uint8_t *miss;
  const uint8_t *const *miss_slices;
} aoadmm_object;

/* The problem struct Z (example_script6_matrix_matrix_CP_nonneg.m:84-92) flattened. */
typedef struct {
  int32_t nb_modes;
  const int64_t *mode_rows;          /* rows of fac{m}; for the PAR2 B_k mode: 0 (see slice_rows) */
  const int32_t *mode_rank;          /* columns of fac{m} */
  const int64_t *const *slice_rows;  /* per mode: NULL, or K values J_k for the PAR2 B_k mode (Z.size{m}) */
  const int32_t *n_slices;           /* per mode: 0, or K */
  int32_t n_objects;
  const aoadmm_object *objects;
  const int32_t *lin_coupled_modes;  /* nb_modes entries: 0 = uncoupled, else coupling id (1-based) */
  int32_t n_couplings;
  const int32_t *coupling_type;      /* per coupling id: 0 exact,1 HC=D,2 CH=D,3 C=HD,4 C=DH,5 H1C=DH2 */
  const double *const *trafo;        /* per mode: Z.coupling.coupl_trafo_matrices{m} or NULL */
  const int64_t *trafo_rows, *trafo_cols;
  const double *const *trafo2;       /* per mode: coupl_trafo_matrices2{m} or NULL (type 5) */
  const int64_t *trafo2_rows, *trafo2_cols;
  const int64_t *coupling_rows, *coupling_cols; /* per coupling id: shape of G.coupling_fac{c} (Delta) */
  const int32_t *constrained_modes;  /* Z.constrained_modes */
  const aoadmm_constraint *constraints; /* nb_modes entries (kind NONE where unconstrained) */
  const double *ridge;               /* Z.ridge (nb_modes) or NULL */
} aoadmm_problem;

/* Multi-GPU: one process per GPU.  The tensor objects are sharded along their LAST mode
 * (contiguous slabs in column-major order); rank r passes its slab (aoadmm_object.data,
 * shard_offset, shard_extent).  `nccl_unique_id` is the 128-byte ncclUniqueId created by rank 0
 * with aoadmm_nccl_unique_id() and distributed by the host (MPI / torch.distributed / files). */
typedef struct {
  int32_t rank, world_size;
  int32_t device;                 /* CUDA device ordinal for this process */
  uint8_t nccl_unique_id[128];
} aoadmm_dist;

/* options struct (example_script6...m:120-132, cmtf_fun_AOADMM.m:7-9, :196-198) */
typedef struct {
  int32_t MaxOuterIters, MaxInnerIters;
  double AbsFuncTol, OuterRelTol;
  double innerRelPrTol_coupl, innerRelPrTol_constr, innerRelDualTol_coupl, innerRelDualTol_constr;
  int32_t bsum;
  double bsum_weight;
  int32_t iter_start_PAR2Bkconstraint;
  int32_t has_increase_factor_rhoBk;
  double increase_factor_rhoBk;
  /* engine-only knobs (0 = default) */
  int32_t mttkrp_precision;   /* 0: FP64 DMMA (default, the parity mode).  Opt-in reduced precision for the MTTKRPs of
                                 3-way tensors (the tensor stays FP64 in HBM, the pass becomes HBM bound; everything else
                                 stays FP64): 1 = TF32 operands, 2 = BF16 operands, both on the 5th-generation tensor cores
                                 (tcgen05.mma, FP32 accumulation in TMEM per slab, FP64 across slabs; operand rounding
                                 2^-11 / 2^-8); 3 = TF32 on the mma.sync variant of the FP64 kernels (kept for comparison) */
  int32_t dimtree;            /* 0: three independent MTTKRP passes (reference flop/byte count) */
  int32_t fuse_inner;         /* 0 (default): run the whole inner ADMM loop of a group in one cooperative launch when
                                 possible; -1: one launch per inner iteration; results are identical */
  int32_t graph;              /* CUDA-graph replay of the outer iteration: 0 auto (launch-bound problems on one GPU),
                                 1 on, -1 off; results are identical */
} aoadmm_options;

/* out struct (cmtf_fun_AOADMM.m:480-494).  History arrays are caller-allocated with
 * MaxOuterIters+1 entries; inner_iters is nb_modes x MaxOuterIters column-major (out.innerIters). */
typedef struct {
  double f_tensors, f_couplings, f_constraints, f_PAR2_couplings;
  int32_t OuterIterations;
  int32_t exit_flag;          /* 0 'maxIterations'; else bit mask of the make_exit_flag.m:9-28 struct:
                                 bit0 f_tensors<AbsFuncTol, bit1 f_couplings, bit2 f_constraints,
                                 bit3 f_PAR2 (bit set = "AbsFuncTol", clear = "RelFuncTol"), bit8 = stopped */
  double *func_val_conv, *func_coupl_conv, *func_constr_conv, *func_PAR2_coupl, *time_at_it;
  int32_t *inner_iters;
  int32_t error_mode;         /* mode id that raised NOT_POSITIVE_DEFINITE / NON_FINITE (0 if none) */
  double f_rel_missing;       /* out.f_rel_missing (NaN without Z.miss), cmtf_fun_AOADMM.m:436-440 */
  double *func_rel_missing;   /* out.func_rel_missing, MaxOuterIters+1 entries, or NULL */
  int32_t non_finite_mode;    /* 0, or the id of the first mode whose inner ADMM loop saw a NaN/Inf residual ratio
                                 (||fac|| = 0 or ||mu_DeltaB|| = 0, cmtf_fun_AOADMM.m:584, :1086-1112).  Like the
                                 reference the run goes on (NaN > tol is false, Inf > tol is true); this is a warning. */
} aoadmm_out;

/* State fields of G (init_coupled_AOADMM_CMTF.m:41-45, :62-80, :133-169) */
typedef enum {
  AOADMM_FIELD_FAC = 0,             /* G.fac{m}               index=m, slice=k (PAR2 B_k) or 0 */
  AOADMM_FIELD_CONSTRAINT_FAC = 1,  /* G.constraint_fac{m}                                      */
  AOADMM_FIELD_CONSTRAINT_DUAL = 2, /* G.constraint_dual_fac{m}                                 */
  AOADMM_FIELD_COUPLING_FAC = 3,    /* G.coupling_fac{c}      index=c (coupling id)             */
  AOADMM_FIELD_COUPLING_DUAL = 4,   /* G.coupling_dual_fac{m}                                   */
  AOADMM_FIELD_PAR2_P = 5,          /* G.P{p}{k}              index=p (1-based object), slice=k */
  AOADMM_FIELD_PAR2_DELTAB = 6,     /* G.DeltaB{p}                                              */
  AOADMM_FIELD_PAR2_MU_DELTAB = 7   /* G.mu_DeltaB{p}{k}                                        */
} aoadmm_field;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int aoadmm_abi_version(void);
int aoadmm_device_count(int *count);
/* rank 0 fills the 128-byte id; replaces nothing in the reference (no distributed layer exists) */
int aoadmm_nccl_unique_id(uint8_t id[128]);
/* copies / partitions the data to HBM once.  dist may be NULL (single GPU, device 0).  With dist->world_size > 1 the
 * NCCL communicator of (nccl_unique_id, rank, world_size) is created on first use and cached for the life of the
 * process: passing the same id again (every rank alike) reuses it, so repeated solves pay the bootstrap once. */
int aoadmm_create(const aoadmm_problem *problem, const aoadmm_dist *dist, aoadmm_handle **out);
/* Multi-GPU from ONE caller (functions/cmtf_AOADMM.m:193 is one MATLAB process): `problem` describes the WHOLE
 * objects (shard_offset = 0, shard_extent = size of the last mode, `data` / `miss` = the full arrays).  The library
 * cuts every tensor of order >= 3 into n_gpus contiguous slabs of its last mode (slab r = indices
 * [extent*r/n, extent*(r+1)/n)), uploads slab r to devices[r] (NULL = devices 0..n_gpus-1) from its own worker
 * thread, creates the communicators with ncclCommInitAll (cached per device list) and returns ONE handle.  Every
 * entry point below then acts on all devices: set_state replicates, run drives one worker thread per GPU,
 * get_state / out come from device 0 (all replicas are bit-identical), aoadmm_get_object_data gathers the slabs.
 * n_gpus = 1 is the same as aoadmm_create(problem, {0,1,devices[0]}). */
int aoadmm_create_multi(const aoadmm_problem *problem, int32_t n_gpus, const int32_t *devices, aoadmm_handle **out);
/* number of GPUs driven by this handle (1 for aoadmm_create) */
int aoadmm_gpu_count(const aoadmm_handle *h, int32_t *n_gpus);
/* destroys every cached NCCL communicator of this process (no handle may be alive) */
int aoadmm_comm_release(void);
int aoadmm_destroy(aoadmm_handle *h);
const char *aoadmm_last_error(const aoadmm_handle *h); /* h may be NULL: last creation error */

/* ---- state in / out (G is both input and output of cmtf_fun_AOADMM.m:1) ------------------- */
int aoadmm_set_state(aoadmm_handle *h, int32_t field, int32_t index, int32_t slice,
                     const double *data, int64_t rows, int64_t cols);
int aoadmm_get_state(aoadmm_handle *h, int32_t field, int32_t index, int32_t slice,
                     double *data, int64_t rows, int64_t cols);

/* ---- the solver: the body of cmtf_fun_AOADMM.m:32-506 ------------------------------------- */
int aoadmm_run(aoadmm_handle *h, const aoadmm_options *options, aoadmm_out *out);

/* ---- operator-level entry points (the reference's L0/L1 calls, for tests / profiling) ------
 * aoadmm_mttkrp:   Tensor Toolbox mttkrp(X,U,n) as called at cmtf_fun_AOADMM.m:97 (n 1-based).
 *                  X: dims[0..order-1] col-major host buffer, factors[k]: dims[k] x R host buffers,
 *                  out: dims[n-1] x R host buffer.
 * aoadmm_prox:     feval(Z.prox_operators{m}, X, rho) (cmtf_fun_AOADMM.m:1424-1426) for a named spec.
 * aoadmm_chol_solve: X = (A / L') / L with L = chol(B','lower') (cmtf_fun_AOADMM.m:142, :609).
 * aoadmm_gram:     G = F'*F (cmtf_fun_AOADMM.m:66, :148).
 */
int aoadmm_mttkrp(const double *X, int32_t order, const int64_t *dims, const double *const *factors,
                  int32_t R, int32_t n, double *out, int32_t device);
int aoadmm_prox(const aoadmm_constraint *spec, const double *X, int64_t rows, int64_t cols, double rho,
                double *out, int32_t device);
int aoadmm_chol_solve(const double *B, int32_t R, const double *A, int64_t rows, double *X, int32_t device);
int aoadmm_gram(const double *F, int64_t rows, int32_t R, double *G, int32_t device);

/* ---- device-resident benchmark helpers ----------------------------------------------------
 * Replace the data of CP object `object` (1-based) by a synthetic tensor generated ON DEVICE:
 * X = [[lambda; U_1..U_N]] + sigma*N(0,1) with sigma = noise*||X0||/||N|| (create_coupled_data.m:158-162),
 * then normalised to ||X||_F = 1 (example_script6...m:101-102).  factors[k] are host buffers of the
 * full (unsharded) factor matrices.  Needed for configs whose tensor does not fit host memory
 * (SURVEY.md 8d C3).  Writes the new Znorm_const (=1) into the handle. */
int aoadmm_generate_cp_data(aoadmm_handle *h, int32_t object, const double *const *factors,
                            double noise, uint64_t seed);
/* Copy this rank's slab of CP object `object` (1-based) back to a caller buffer with the layout aoadmm_create
 * takes (dims of the leading modes x shard_extent, column-major, no padding).  bench.py uses it to obtain the
 * device-generated tensor as a HOST buffer for the end-to-end leg. */
int aoadmm_get_object_data(aoadmm_handle *h, int32_t object, double *out, int64_t n_elements);
/* Leading eigenvectors for the 'nvecs' initialisation (cmtf_nvecs.m:33-58; init_coupled_AOADMM_CMTF.m:50-69):
 * out (rows x r, column-major host buffer) = eigs(Y, r, 'LM') with Y = X_(mode) X_(mode)' of the first object that
 * contains global mode id `mode` (1-based), computed from the data resident in the handle.  PARAFAC2 objects: mode A uses
 * the slices side by side, a B_k mode needs `slice` (1-based; Y = X_k' X_k), mode C is ones in the reference and is
 * refused.  slice = 0 otherwise.  Columns come in descending eigenvalue order, each signed so that its entry of
 * largest magnitude is positive (eigs leaves the sign open).  info (may be NULL): [0] subspace iterations,
 * [1] max ||Y u - theta u|| / theta_1 over the returned pairs.  With more than one GPU every rank must make the call
 * (collective): partial Gram matrices of the slabs are all-reduced; for the sharded (last) mode the slabs are exchanged
 * chunk by chunk with NCCL point-to-point transfers. */
int aoadmm_nvecs(aoadmm_handle *h, int32_t mode, int32_t slice, int32_t r, double *out, int64_t rows, double *info);
/* MTTKRP of the resident CP object `object` in mode position `pos` (1-based) with the factors currently in the handle
 * (cmtf_fun_AOADMM.m:97), at `precision` (0 FP64, 1/2/3 reduced precision: see aoadmm_options.mttkrp_precision), summed over
 * ranks; out: rows(mode) x R host buffer. */
int aoadmm_object_mttkrp(aoadmm_handle *h, int32_t object, int32_t pos, int32_t precision, double *out);
/* one timed MTTKRP of object `object` in mode position `pos` (1-based position inside the object)
 * using the factors currently resident in the handle; returns device milliseconds (CUDA events). */
int aoadmm_time_mttkrp(aoadmm_handle *h, int32_t object, int32_t pos, int32_t reps, float *ms_out);
/* number of kernels launched by this handle so far, summed over its GPUs (bench.py "gpu_launches"). */
int aoadmm_launch_count(const aoadmm_handle *h, int64_t *count);
/* per-phase device time accumulated by aoadmm_run since creation, milliseconds:
 * [0] MTTKRP (tensor), [1] matrix-block products, [2] everything else */
int aoadmm_phase_ms(const aoadmm_handle *h, double ms[3]);
/* device time (CUDA events on the engine's stream, first to last kernel) of the last aoadmm_run, ms */
int aoadmm_last_run_ms(const aoadmm_handle *h, double *ms);
/* same, but only the outer iterations (cmtf_fun_AOADMM.m:87-476), i.e. without the one-off iteration-0 objective
 * of cmtf_fun_AOADMM.m:32 (one extra MTTKRP per tensor) */
int aoadmm_last_loop_ms(const aoadmm_handle *h, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* AOADMM_H_ */
