// mttkrp.cu - dense MTTKRP kernels: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ->
// FP64 tensor-core DMMA.8x8x4 (mma.sync.m8n8k4.f64), warp-specialised producer/consumer pipeline.
//
// Replaces Tensor Toolbox `mttkrp` (called at functions/cmtf_fun_AOADMM.m:97, cp_func.m:47) and the
// matrix products at cmtf_fun_AOADMM.m:108,:111.
//
// Two main loops cover every mode of a column-major tensor viewed as I x J x K:
//   LEAD  (output mode = the contiguous mode i):  M(i,:)  = sum_k Fk(k,:) .* sum_j X(i,j,k) Fj(j,:)
//         MMA:  M-index = i (contiguous in the smem row), reduction index = j (smem row), one k per stage.
//   INNER (reduction over the contiguous mode i): T(j,k,:) = sum_i X(i,j,k) Fi(i,:)
//         MMA:  M-index = j (smem row), reduction index = i (contiguous in the smem row);
//         epilogue 0 (mode 2):  M(j,:) += T(j,k,:) .* Fk(k,:)          (accumulated over the CTA's k range)
//         epilogue 1 (mode 3):  M(k,:)  = sum_j T(j,k,:) .* Fj(j,:)    (reduced over the CTA's j tile)
//
// Shared-memory tensor tile: rows of 16 doubles (128 B) written by TMA with CU_TENSOR_MAP_SWIZZLE_128B:
//   element (i_local in 0..15, row) lives at  row*128 + (((i_local>>1) ^ (row&7)) << 4) + (i_local&1)*8.
// Fragment index maps are chosen so that every 64-bit fragment load is bank-conflict free:
//   LEAD : an m-tile of 8 accumulator rows covers i_local = {0,1,8,9,2,3,10,11} (+4 for the odd m-tile)
//   INNER: an m-tile covers the even (or odd) rows of a 16-row block.
// The factor operand is a row-contiguous packed copy (PackedFactor) with leading dimension NC+4.
#include "mttkrp.cuh"

#include <mutex>
#include <vector>

namespace aoadmm {

int64_t mttkrp_T_ld(const Tensor3& t);

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kStages = 4;
constexpr int kXBytes = 32768;  // tensor tile per stage: LEAD 128(i) x 32(j); INNER 32(i) x 128(j)

template <int NT, int WARPS_N>
struct Cfg {
  static constexpr int WN = 8 * NT;
  static constexpr int NC = WN * WARPS_N;
  static constexpr int KSPLIT = 2 / WARPS_N;  // warp groups splitting the k4-steps of a stage
  static constexpr int LDC = NC + 4;
  static constexpr int FBYTES = 32 * LDC * 8;  // 32 factor rows per stage
  static constexpr int CBYTES = ((LDC * 8 + 15) / 16) * 16;
  static constexpr int SCRATCH = (KSPLIT == 2) ? 128 * 8 * NT * 8 : 0;  // cross-group reduction
  static constexpr int RED = 2 * kConsumerWarps * WN * 8;              // epilogue-1 cross-warp reduction
  // layout: [X tiles][F tiles][C rows][scratch][red][barriers]
  static constexpr int OFF_X = 0;
  static constexpr int OFF_F = kStages * kXBytes;
  static constexpr int OFF_C = OFF_F + kStages * FBYTES;
  static constexpr int OFF_S = OFF_C + kStages * CBYTES;
  static constexpr int OFF_R = OFF_S + SCRATCH;
  static constexpr int OFF_BAR = OFF_R + RED;
  static constexpr int SMEM = OFF_BAR + 2 * kStages * 8 + 1024;  // +1024 for manual alignment
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
  static_assert(OFF_F % 16 == 0 && OFF_C % 16 == 0 && OFF_S % 16 == 0 && OFF_R % 16 == 0 && OFF_BAR % 8 == 0, "align");
};

// ---------------------------------------------------------------------------------------------
// LEAD kernel
// grid: x = m tile (128 rows of i), y = split of the (k, j-tile) stage sequence, z = column chunk
// ---------------------------------------------------------------------------------------------
// PREC = 0: FP64 DMMA (the parity path).  PREC = 1 (opt-in, options.mttkrp_precision): the FP64 tiles are rounded to TF32
// fragment by fragment and multiplied with mma.m16n8k8.tf32 (FP32 accumulate, flushed into the FP64 accumulators once per
// stage): one TF32 MMA covers two 8-row m-tiles and two k4-steps of the FP64 fragment layout, so the data path is unchanged.
template <int NT, int WARPS_N, int PREC>
__global__ void __launch_bounds__(kThreads, 1)
mttkrp_lead_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ Ft_j,
                   const double* __restrict__ Ft_k, double* __restrict__ ws, int I, int J, int K, long long Jpad,
                   long long Kpad, int njt, int nsplit, long long ldo, int Rp_total, const int* __restrict__ skip) {
  using C = Cfg<NT, WARPS_N>;
  if (skip != nullptr && *skip != 0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = sbase + C::OFF_X, sF = sbase + C::OFF_F, sC = sbase + C::OFF_C, sS = sbase + C::OFF_S;
  const uint32_t sBar = sbase + C::OFF_BAR;  // full[kStages], empty[kStages]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, split = blockIdx.y, chunk = blockIdx.z;
  const long long S = (long long)K * njt;
  const long long q0 = S * split / nsplit, q1 = S * (split + 1) / nsplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(sBar + s * 8, 1);
      mbar_init(sBar + (kStages + s) * 8, kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      prefetch_tensormap(&tmap);
      const int i0 = mt * 128;
      for (long long q = q0; q < q1; ++q) {
        const long long ql = q - q0;
        const int s = (int)(ql % kStages);
        const uint32_t ph = (uint32_t)((ql / kStages) & 1);
        mbar_wait(sBar + (kStages + s) * 8, ph ^ 1u);
        const int k = (int)(q / njt), jt = (int)(q % njt);
        const uint32_t full = sBar + s * 8;
        mbar_expect_tx(full, kXBytes + C::FBYTES + C::CBYTES);
#pragma unroll
        for (int b = 0; b < 8; ++b) tma_load_3d(sX + s * kXBytes + b * 4096, &tmap, i0 + 16 * b, jt * 32, k, full);
        bulk_load_1d(sF + s * C::FBYTES, Ft_j + ((long long)chunk * Jpad + (long long)jt * 32) * C::LDC, C::FBYTES, full);
        bulk_load_1d(sC + s * C::CBYTES, Ft_k + ((long long)chunk * Kpad + k) * C::LDC, C::CBYTES, full);
      }
    }
    return;
  }

  // ===== consumers =====
  const int wm = warp & 3, hi = warp >> 2;
  const int wn = (WARPS_N == 2) ? hi : 0;
  const int t0 = (WARPS_N == 2) ? 0 : 4 * hi;
  constexpr int TCOUNT = (WARPS_N == 2) ? 8 : 4;
  const int m = lane >> 2, kk = lane & 3;
  const int cbase = ((m >> 1) & 1) * 4 + (m >> 2);  // 16-byte chunk of i_local within the 128-byte row
  const int off = m & 1;

  double acc[4][NT][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  uint32_t abase[4], axor[4];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int bx = wm * 2 + (mi >> 1);
    abase[mi] = bx * 4096 + off * 8 + kk * 128;
    axor[mi] = (uint32_t)(((cbase + 2 * (mi & 1)) ^ kk) << 4);
  }
  const uint32_t boff = (uint32_t)((kk * C::LDC + wn * C::WN + m) * 8);
  float fa[4][NT][2];  // PREC == 1: FP32 accumulators of the TF32 MMAs
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) fa[mi][nt][0] = fa[mi][nt][1] = 0.f;

  // PREC 0: the Khatri-Rao scale Fk(k,:) is NOT multiplied into the B fragments (those DMULs would share the FP64 pipe
  // with the DMMAs, 32 per stage and warp): the products of one slab k are accumulated unscaled in accp and folded into
  // acc with one FMA per accumulator when the stage sequence moves on to the next k (every njt stages).
  double accp[4][NT][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) accp[a][b][0] = accp[a][b][1] = 0.0;
  int jtc = (int)(q0 % njt);   // j-tile index of the current stage inside its slab

  for (long long q = q0; q < q1; ++q) {
    const long long ql = q - q0;
    const int s = (int)(ql % kStages);
    const uint32_t ph = (uint32_t)((ql / kStages) & 1);
    mbar_wait(sBar + s * 8, ph);
    const uint32_t xs = sX + s * kXBytes, fs = sF + s * C::FBYTES + boff, cs = sC + s * C::CBYTES;
    double ck[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) ck[nt] = (PREC == 0) ? 0.0 : lds_f64(cs + (wn * C::WN + 8 * nt + m) * 8);
    if (PREC == 0) {
#pragma unroll
      for (int tt = 0; tt < TCOUNT; ++tt) {
        const int t = t0 + tt;
        const uint32_t rowoff = (uint32_t)(t * 512);        // 4 rows of 128 B per k4-step
        const uint32_t flip = (uint32_t)((t & 1) << 6);     // (row & 4) toggles chunk bit 2
        double a[4], b[NT];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = lds_f64(xs + abase[mi] + rowoff + (axor[mi] ^ flip));
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) b[nt] = lds_f64(fs + (uint32_t)((t * 4 * C::LDC + 8 * nt) * 8));
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(accp[mi][nt][0], accp[mi][nt][1], a[mi], b[nt]);
      }
      // last stage of this slab (or of this CTA's range): fold the slab into acc with its scale Fk(k, :), read from the
      // stage's own copy of the row while the slot is still ours
      if (++jtc == njt || q + 1 == q1) {
        jtc = (jtc == njt) ? 0 : jtc;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double c0 = lds_f64(cs + (wn * C::WN + 8 * nt + 2 * kk) * 8);
          const double c1 = lds_f64(cs + (wn * C::WN + 8 * nt + 2 * kk + 1) * 8);
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            acc[mi][nt][0] = fma(accp[mi][nt][0], c0, acc[mi][nt][0]);
            acc[mi][nt][1] = fma(accp[mi][nt][1], c1, acc[mi][nt][1]);
            accp[mi][nt][0] = accp[mi][nt][1] = 0.0;
          }
        }
      }
    } else {
      // operand 0 arrives in the TF32 operand format (packed_factor_to_tf32: the high word of every 8-byte slot is the
      // TF32-rounded float), so the Khatri-Rao row is one 32-bit load and one FP32 multiply per fragment element
      float ckf[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) ckf[nt] = (float)ck[nt];
#pragma unroll
      for (int tt = 0; tt < TCOUNT; tt += 2) {
        uint32_t a[2][4], b[2][NT];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int t = t0 + tt + h;
          const uint32_t rowoff = (uint32_t)(t * 512);
          const uint32_t flip = (uint32_t)((t & 1) << 6);
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) a[h][mi] = f64_to_tf32(lds_f64(xs + abase[mi] + rowoff + (axor[mi] ^ flip)));
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
            b[h][nt] = f32_to_tf32(__uint_as_float(lds_u32(fs + (uint32_t)((t * 4 * C::LDC + 8 * nt) * 8 + 4))) * ckf[nt]);
        }
#pragma unroll
        for (int mp = 0; mp < 4; mp += 2)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
            mma_tf32_1688(fa[mp][nt][0], fa[mp][nt][1], fa[mp + 1][nt][0], fa[mp + 1][nt][1], a[0][mp], a[0][mp + 1],
                          a[1][mp], a[1][mp + 1], b[0][nt], b[1][nt]);
      }
      // FP32 partial sums are folded into the FP64 accumulators every 4 stages (128 terms) and at the end
      if ((ql & 3) == 3 || q + 1 == q1) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            acc[mi][nt][0] += (double)fa[mi][nt][0];
            acc[mi][nt][1] += (double)fa[mi][nt][1];
            fa[mi][nt][0] = fa[mi][nt][1] = 0.f;
          }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sBar + (kStages + s) * 8);
  }

  // ===== epilogue =====
  if (C::KSPLIT == 2) {
    const int slot = wm * 32 + lane;
    if (hi == 1) {
      double* sc = reinterpret_cast<double*>(smem_raw + (sS - smem_u32(smem_raw)));
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          sc[((mi * NT + nt) * 2 + 0) * 128 + slot] = acc[mi][nt][0];
          sc[((mi * NT + nt) * 2 + 1) * 128 + slot] = acc[mi][nt][1];
        }
    }
    named_bar_sync(1, kConsumerWarps * 32);
    if (hi == 1) return;
    const double* sc = reinterpret_cast<const double*>(smem_raw + (sS - smem_u32(smem_raw)));
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        acc[mi][nt][0] += sc[((mi * NT + nt) * 2 + 0) * 128 + slot];
        acc[mi][nt][1] += sc[((mi * NT + nt) * 2 + 1) * 128 + slot];
      }
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int i = mt * 128 + wm * 32 + (mi >> 1) * 16 + 2 * (cbase + 2 * (mi & 1)) + off;
    if (i < I) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int r = chunk * C::NC + wn * C::WN + 8 * nt + 2 * kk;
        ws[((long long)split * Rp_total + r) * ldo + i] = acc[mi][nt][0];
        ws[((long long)split * Rp_total + r + 1) * ldo + i] = acc[mi][nt][1];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// INNER kernel
// grid: x = j tile (128 rows), y = split of the k range, z = column chunk
// EPI = 0: ws[(split*Rp + r)*ldo + j]   = sum_{k in range} Fe(k,r) * T(j,k,r)      (Fe = packed k factor)
// EPI = 1: ws[(jt*Rp + r)*ldo + k]      = sum_{j in tile}  Fe(j,r) * T(j,k,r)      (Fe = packed j factor)
// ---------------------------------------------------------------------------------------------
// EMIT (epilogue 0 only): additionally stores the un-scaled partial contraction T(j,k,r) = sum_i X(i,j,k) Fi(i,r)
//          to Tbuf[(k*Rp + r)*ldt + j]; a later mode-3 MTTKRP is then a cheap pass over T (dimension tree).
template <int NT, int WARPS_N, int EPI, bool EMIT, int PREC>
__global__ void __launch_bounds__(kThreads, 1)
mttkrp_inner_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ Ft_i,
                    const double* __restrict__ Ft_e, double* __restrict__ ws, int I, int J, int K, long long Ipad,
                    long long Epad, int nit, int nsplit, long long ldo, int Rp_total, double* __restrict__ Tbuf,
                    long long ldt, const int* __restrict__ skip) {
  using C = Cfg<NT, WARPS_N>;
  if (skip != nullptr && *skip != 0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = sbase + C::OFF_X, sF = sbase + C::OFF_F, sS = sbase + C::OFF_S, sR = sbase + C::OFF_R;
  const uint32_t sBar = sbase + C::OFF_BAR;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, split = blockIdx.y, chunk = blockIdx.z;
  const int k0 = (int)((long long)K * split / nsplit), k1 = (int)((long long)K * (split + 1) / nsplit);
  const long long nstage = (long long)(k1 - k0) * nit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(sBar + s * 8, 1);
      mbar_init(sBar + (kStages + s) * 8, kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      prefetch_tensormap(&tmap);
      for (long long ql = 0; ql < nstage; ++ql) {
        const int s = (int)(ql % kStages);
        const uint32_t ph = (uint32_t)((ql / kStages) & 1);
        mbar_wait(sBar + (kStages + s) * 8, ph ^ 1u);
        const int k = k0 + (int)(ql / nit), it = (int)(ql % nit);
        const uint32_t full = sBar + s * 8;
        mbar_expect_tx(full, kXBytes + C::FBYTES);
        tma_load_3d(sX + s * kXBytes, &tmap, it * 32, jt * 128, k, full);
        tma_load_3d(sX + s * kXBytes + 16384, &tmap, it * 32 + 16, jt * 128, k, full);
        bulk_load_1d(sF + s * C::FBYTES, Ft_i + ((long long)chunk * Ipad + (long long)it * 32) * C::LDC, C::FBYTES, full);
      }
    }
    return;
  }

  const int wm = warp & 3, hi = warp >> 2;
  const int wn = (WARPS_N == 2) ? hi : 0;
  const int t0 = (WARPS_N == 2) ? 0 : 4 * hi;
  constexpr int TCOUNT = (WARPS_N == 2) ? 8 : 4;
  const int m = lane >> 2, kk = lane & 3;

  double T[4][NT][2];
  float Tf[4][NT][2];  // PREC == 1: FP32 accumulators of the TF32 MMAs for one (j tile, k) (I terms each)
  double O[4][NT][2];  // EPI 0: running output; EPI 1: the epilogue factor values Fe(j(mi,m), cols)
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) {
      T[a][b][0] = T[a][b][1] = O[a][b][0] = O[a][b][1] = 0.0;
      Tf[a][b][0] = Tf[a][b][1] = 0.f;
    }

  uint32_t abase[4], arm[4];
  int jrow[4];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int row = wm * 32 + 16 * (mi >> 1) + 2 * m + (mi & 1);
    jrow[mi] = row;
    abase[mi] = (uint32_t)(row * 128 + (kk & 1) * 8);
    arm[mi] = (uint32_t)(row & 7);
  }
  const uint32_t boff = (uint32_t)((kk * C::LDC + wn * C::WN + m) * 8);
  const int ccol = wn * C::WN + 2 * kk;  // first accumulator column of this lane inside the chunk

  if (EPI == 1) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const long long j = (long long)jt * 128 + jrow[mi];  // < Epad by construction (rows_pad covers whole tiles)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 v = *reinterpret_cast<const double2*>(Ft_e + ((long long)chunk * Epad + j) * C::LDC + ccol + 8 * nt);
        O[mi][nt][0] = v.x;
        O[mi][nt][1] = v.y;
      }
    }
  }

  double* red = reinterpret_cast<double*>(smem_raw + (sR - smem_u32(smem_raw)));

  for (int k = k0; k < k1; ++k) {
    for (int it = 0; it < nit; ++it) {
      const long long ql = (long long)(k - k0) * nit + it;
      const int s = (int)(ql % kStages);
      const uint32_t ph = (uint32_t)((ql / kStages) & 1);
      mbar_wait(sBar + s * 8, ph);
      const uint32_t xs = sX + s * kXBytes, fs = sF + s * C::FBYTES + boff;
      if (PREC == 0) {
#pragma unroll
        for (int tt = 0; tt < TCOUNT; ++tt) {
          const int t = t0 + tt;
          const uint32_t box = (uint32_t)((t >> 2) * 16384);
          const uint32_t ch = (uint32_t)(2 * (t & 3) + (kk >> 1));
          double a[4], b[NT];
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) a[mi] = lds_f64(xs + box + abase[mi] + ((ch ^ arm[mi]) << 4));
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) b[nt] = lds_f64(fs + (uint32_t)((t * 4 * C::LDC + 8 * nt) * 8));
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) dmma884(T[mi][nt][0], T[mi][nt][1], a[mi], b[nt]);
        }
      } else {
#pragma unroll
        for (int tt = 0; tt < TCOUNT; tt += 2) {
          uint32_t a[2][4], b[2][NT];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int t = t0 + tt + h;
            const uint32_t box = (uint32_t)((t >> 2) * 16384);
            const uint32_t ch = (uint32_t)(2 * (t & 3) + (kk >> 1));
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[h][mi] = f64_to_tf32(lds_f64(xs + box + abase[mi] + ((ch ^ arm[mi]) << 4)));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) b[h][nt] = lds_u32(fs + (uint32_t)((t * 4 * C::LDC + 8 * nt) * 8 + 4));
          }
#pragma unroll
          for (int mp = 0; mp < 4; mp += 2)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
              mma_tf32_1688(Tf[mp][nt][0], Tf[mp][nt][1], Tf[mp + 1][nt][0], Tf[mp + 1][nt][1], a[0][mp], a[0][mp + 1],
                            a[1][mp], a[1][mp + 1], b[0][nt], b[1][nt]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sBar + (kStages + s) * 8);
    }
    // ---- per-(j tile, k) epilogue
    if (PREC == 1) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          T[mi][nt][0] = (double)Tf[mi][nt][0];
          T[mi][nt][1] = (double)Tf[mi][nt][1];
          Tf[mi][nt][0] = Tf[mi][nt][1] = 0.f;
        }
    }
    if (EPI == 0) {
      if (EMIT) {
        double* sc = reinterpret_cast<double*>(smem_raw + (sS - smem_u32(smem_raw)));
        const int slot = wm * 32 + lane;
        if (C::KSPLIT == 2) {
          if (hi == 1) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                sc[((mi * NT + nt) * 2 + 0) * 128 + slot] = T[mi][nt][0];
                sc[((mi * NT + nt) * 2 + 1) * 128 + slot] = T[mi][nt][1];
              }
          }
          named_bar_sync(1, kConsumerWarps * 32);
        }
        if (C::KSPLIT == 1 || hi == 0) {
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            const int j = jt * 128 + jrow[mi];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              double t0 = T[mi][nt][0], t1 = T[mi][nt][1];
              if (C::KSPLIT == 2) {
                t0 += sc[((mi * NT + nt) * 2 + 0) * 128 + slot];
                t1 += sc[((mi * NT + nt) * 2 + 1) * 128 + slot];
              }
              if (j < J) {
                const long long r = (long long)chunk * C::NC + ccol + 8 * nt;
                Tbuf[((long long)k * Rp_total + r) * ldt + j] = t0;
                Tbuf[((long long)k * Rp_total + r + 1) * ldt + j] = t1;
              }
            }
          }
        }
        if (C::KSPLIT == 2) named_bar_sync(1, kConsumerWarps * 32);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 c = *reinterpret_cast<const double2*>(Ft_e + ((long long)chunk * Epad + k) * C::LDC + ccol + 8 * nt);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          O[mi][nt][0] = fma(T[mi][nt][0], c.x, O[mi][nt][0]);
          O[mi][nt][1] = fma(T[mi][nt][1], c.y, O[mi][nt][1]);
          T[mi][nt][0] = T[mi][nt][1] = 0.0;
        }
      }
    } else {
      double* rb = red + ((k - k0) & 1) * (kConsumerWarps * C::WN);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double v = 0.0;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            v = fma(T[mi][nt][e], O[mi][nt][e], v);
            T[mi][nt][e] = 0.0;
          }
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (m == 0) rb[warp * C::WN + 8 * nt + 2 * kk + e] = v;
        }
      }
      named_bar_sync(1, kConsumerWarps * 32);
      const int tid = threadIdx.x;
      if (tid < C::NC) {
        const int n_wn = tid / C::WN, col = tid % C::WN;
        double v = 0.0;
        if (WARPS_N == 2) {
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) v += rb[(n_wn * 4 + w4) * C::WN + col];
        } else {
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) v += rb[w8 * C::WN + col];
        }
        ws[((long long)jt * Rp_total + chunk * C::NC + tid) * ldo + k] = v;
      }
    }
  }

  if (EPI == 0) {
    if (C::KSPLIT == 2) {
      const int slot = wm * 32 + lane;
      double* sc = reinterpret_cast<double*>(smem_raw + (sS - smem_u32(smem_raw)));
      if (hi == 1) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            sc[((mi * NT + nt) * 2 + 0) * 128 + slot] = O[mi][nt][0];
            sc[((mi * NT + nt) * 2 + 1) * 128 + slot] = O[mi][nt][1];
          }
      }
      named_bar_sync(1, kConsumerWarps * 32);
      if (hi == 1) return;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          O[mi][nt][0] += sc[((mi * NT + nt) * 2 + 0) * 128 + slot];
          O[mi][nt][1] += sc[((mi * NT + nt) * 2 + 1) * 128 + slot];
        }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int j = jt * 128 + jrow[mi];
      if (j < J) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int r = chunk * C::NC + ccol + 8 * nt;
          ws[((long long)split * Rp_total + r) * ldo + j] = O[mi][nt][0];
          ws[((long long)split * Rp_total + r + 1) * ldo + j] = O[mi][nt][1];
        }
      }
    }
  }
}

// out(row, r) = scale * sum_s ws[(s*Rp + r)*ldo + row]   (fixed order => deterministic)
__global__ void mttkrp_reduce_kernel(const double* __restrict__ ws, int nsplit, int Rp_total, long long ldo,
                                     long long rows, int R, double scale, double* __restrict__ out, long long ldout,
                                     const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (row >= rows || r >= R) return;
  double v = 0.0;
  for (int s = 0; s < nsplit; ++s) v += ws[((long long)s * Rp_total + r) * ldo + row];
  out[(long long)r * ldout + row] = scale * v;
}

// mode-3 MTTKRP from the cached partial contraction: out(k,r) = scale * sum_j T[(k*Rp + r)*ldt + j] * Fj(j,r)
// one warp per (k, r); lanes stride over j (coalesced in T and in the column-major factor).
__global__ void mttkrp_from_T_kernel(const double* __restrict__ Tbuf, long long ldt, int Rp_total, int J, int K, int R,
                                     const double* __restrict__ Fj, long long ldf, double scale,
                                     double* __restrict__ out, long long ldout, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x;
  const int r = blockIdx.y * 8 + warp;
  if (r >= R) return;
  const double* trow = Tbuf + ((long long)k * Rp_total + r) * ldt;
  const double* frow = Fj + (long long)r * ldf;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int j = lane;
  for (; j + 96 < J; j += 128) {
    a0 = fma(trow[j], frow[j], a0);
    a1 = fma(trow[j + 32], frow[j + 32], a1);
    a2 = fma(trow[j + 64], frow[j + 64], a2);
    a3 = fma(trow[j + 96], frow[j + 96], a3);
  }
  for (; j < J; j += 32) a0 = fma(trow[j], frow[j], a0);
  const double v = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) out[(long long)r * ldout + k] = scale * v;
}

__global__ void pack_factor_kernel(double* __restrict__ dst, long long rows_pad, int ldc, int NC, int nchunk,
                                   const double* __restrict__ F, long long rows, long long ld, int R,
                                   const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long total = (long long)nchunk * rows_pad * ldc;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % ldc);
    const long long row = (idx / ldc) % rows_pad;
    const int c = (int)(idx / ((long long)ldc * rows_pad));
    const int r = c * NC + col;
    double v = 0.0;
    if (col < NC && r < R && row < rows) v = (F != nullptr) ? F[(long long)r * ld + row] : 1.0;
    dst[idx] = v;
  }
}

// in place: every 8-byte slot becomes {low word 0, high word = TF32-rounded float of the value} (operand 0 of the
// PREC == 1 kernels)
__global__ void packed_to_tf32_kernel(double* __restrict__ data, long long total, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const uint32_t hi = f64_to_tf32(data[idx]);
    reinterpret_cast<unsigned long long*>(data)[idx] = (unsigned long long)hi << 32;
  }
}

__global__ void pack_kr_kernel(double* __restrict__ dst, long long rows_pad, int ldc, int NC, int nchunk,
                               const double* __restrict__ Fa, long long rows_a, long long lda,
                               const double* __restrict__ Fb, long long rows_b, long long ldb, int R,
                               const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long total = (long long)nchunk * rows_pad * ldc;
  const long long rows = rows_a * rows_b;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % ldc);
    const long long row = (idx / ldc) % rows_pad;
    const int c = (int)(idx / ((long long)ldc * rows_pad));
    const int r = c * NC + col;
    double v = 0.0;
    if (col < NC && r < R && row < rows) v = Fa[(long long)r * lda + row % rows_a] * Fb[(long long)r * ldb + row / rows_a];
    dst[idx] = v;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (fn == nullptr) throw CudaError(5, "cuTensorMapEncodeTiled driver entry point not available");
  return fn;
}

void encode_map(CUtensorMap* map, const double* X, int64_t I, int64_t J, int64_t K, int64_t ldI, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)I, (cuuint64_t)J, (cuuint64_t)K};
  cuuint64_t strides[2] = {(cuuint64_t)ldI * 8, (cuuint64_t)ldI * (cuuint64_t)J * 8};
  cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(X), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw CudaError(5, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " for tensor " +
                           std::to_string(I) + "x" + std::to_string(J) + "x" + std::to_string(K) + " ld " +
                           std::to_string(ldI));
}

}  // namespace

// generic 3-D FP64 tensor map with 128-byte swizzle (box[0] must be 16 doubles); used by the unfolding Gram kernel
void encode_map3(CUtensorMap* map, const double* X, const uint64_t dims[3], const uint64_t strides_bytes[2],
                 const uint32_t box[3], bool swizzle128) {
  cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
  cuuint64_t sb[2] = {strides_bytes[0], strides_bytes[1]};
  cuuint32_t bx[3] = {box[0], box[1], box[2]};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(X), d, sb, bx, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw CudaError(5, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " for a " +
                           std::to_string(dims[0]) + "x" + std::to_string(dims[1]) + "x" + std::to_string(dims[2]) + " view");
}

namespace {

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    AO_CUDA(cudaGetDevice(&dev));
    AO_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

// number of splits so that tiles*splits fills whole waves of one CTA per SM
int choose_splits(long long tiles, long long max_splits) {
  const int sms = sm_count();
  int best = 1;
  double best_eff = 0.0;
  for (int w = 1; w <= 4; ++w) {
    long long ns = ((long long)sms * w) / tiles;
    if (ns < 1) ns = 1;
    if (ns > max_splits) ns = max_splits;
    const long long ctas = tiles * ns;
    const double eff = (double)ctas / (double)(ceil_div(ctas, sms) * sms);
    // prefer fewer waves unless efficiency improves noticeably
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = (int)ns;
    }
  }
  return best;
}

// INNER kernels split the slab range in WHOLE slabs, so a CTA's work is ceil(K / splits) slabs and the pass lasts
// ceil(K / splits) * waves slab-times: pick the split count that minimises that makespan (e.g. K = 128 on 32 tiles:
// 9 splits -> 15 slabs x 2 waves = 30, 32 splits -> 4 x 7 = 28, ideal 27.7).  Ties go to fewer splits (less to reduce).
int choose_splits_slabs(long long tiles, long long slabs, long long max_splits) {
  const int sms = sm_count();
  long long best = 1;
  double best_cost = 1e300;
  const long long hi = std::max<long long>(1, std::min<long long>(std::min(slabs, max_splits), 592));
  for (long long ns = 1; ns <= hi; ++ns) {
    const long long waves = ceil_div(tiles * ns, sms);
    // per-CTA fixed cost (pipeline fill, epilogue, partial-result traffic) ~ a quarter of a slab-time per wave
    const double cost = (double)ceil_div(slabs, ns) * (double)waves + 0.25 * (double)waves;
    if (cost < best_cost * (1.0 - 1e-9)) {
      best_cost = cost;
      best = ns;
    }
  }
  return (int)best;
}

template <int NT, int WARPS_N, int PREC>
int launch_lead(const Tensor3& t, const PackedFactor& fj, const PackedFactor& fk, int R, double scale, double* out,
                int64_t ldout, const MttkrpWorkspace& w, cudaStream_t st, const int* skip) {
  using C = Cfg<NT, WARPS_N>;
  const int mtiles = (int)ceil_div(t.I, 128), njt = (int)ceil_div(t.J, 32), nchunk = fj.nchunk;
  const long long S = t.K * njt;
  const int nsplit = choose_splits((long long)mtiles * nchunk, S);
  const int Rp_total = nchunk * C::NC;
  const long long ldo = round_up(t.I, 2);
  if ((size_t)nsplit * Rp_total * ldo * 8 > w.ws_bytes) throw CudaError(1, "mttkrp workspace too small (lead)");
  auto kern = mttkrp_lead_kernel<NT, WARPS_N, PREC>;
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::SMEM);
  dim3 grid(mtiles, nsplit, nchunk);
  kern<<<grid, kThreads, C::SMEM, st>>>(t.map_lead, fj.data, fk.data, w.ws, (int)t.I, (int)t.J, (int)t.K,
                                        (long long)fj.rows_pad, (long long)fk.rows_pad, njt, nsplit, ldo, Rp_total, skip);
  AO_CHECK_LAUNCH();
  dim3 rgrid((unsigned)ceil_div(t.I, 128), R);
  mttkrp_reduce_kernel<<<rgrid, 128, 0, st>>>(w.ws, nsplit, Rp_total, ldo, t.I, R, scale, out, ldout, skip);
  AO_CHECK_LAUNCH();
  return 2;
}

template <int NT, int WARPS_N, int EPI, bool EMIT, int PREC>
int launch_inner(const Tensor3& t, const PackedFactor& fi, const PackedFactor& fe, int R, double scale, double* out,
                 int64_t ldout, const MttkrpWorkspace& w, double* Tbuf, cudaStream_t st, const int* skip) {
  using C = Cfg<NT, WARPS_N>;
  const int jtiles = (int)ceil_div(t.J, 128), nit = (int)ceil_div(t.I, 32), nchunk = fi.nchunk;
  const int nsplit = choose_splits_slabs((long long)jtiles * nchunk, t.K, 592);
  const int Rp_total = nchunk * C::NC;
  const long long rows = (EPI == 0) ? t.J : t.K;
  const long long ldo = round_up(rows, 2);
  const int nparts = (EPI == 0) ? nsplit : jtiles;
  if ((size_t)nparts * Rp_total * ldo * 8 > w.ws_bytes) throw CudaError(1, "mttkrp workspace too small (inner)");
  auto kern = mttkrp_inner_kernel<NT, WARPS_N, EPI, EMIT, PREC>;
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::SMEM);
  dim3 grid(jtiles, nsplit, nchunk);
  kern<<<grid, kThreads, C::SMEM, st>>>(t.map_inner, fi.data, fe.data, w.ws, (int)t.I, (int)t.J, (int)t.K,
                                        (long long)fi.rows_pad, (long long)fe.rows_pad, nit, nsplit, ldo, Rp_total, Tbuf,
                                        (long long)mttkrp_T_ld(t), skip);
  AO_CHECK_LAUNCH();
  dim3 rgrid((unsigned)ceil_div(rows, 128), R);
  mttkrp_reduce_kernel<<<rgrid, 128, 0, st>>>(w.ws, nparts, Rp_total, ldo, rows, R, scale, out, ldout, skip);
  AO_CHECK_LAUNCH();
  return 2;
}

}  // namespace

void make_tensor3(Tensor3& t, const double* X, int64_t I, int64_t J, int64_t K, int64_t ldI) {
  if (ldI % 2 != 0 || (reinterpret_cast<uintptr_t>(X) & 15u) != 0)
    throw CudaError(1, "tensor must be 16-byte aligned with an even leading dimension");
  t.X = X;
  t.I = I;
  t.J = J;
  t.K = K;
  t.ldI = ldI;
  encode_map(&t.map_lead, X, I, J, K, ldI, 32);
  encode_map(&t.map_inner, X, I, J, K, ldI, 128);
}

void packed_factor_alloc(PackedFactor& p, int64_t rows, int R) {
  p.rows = rows;
  p.R = R;
  p.NC = mttkrp_chunk_cols(R);
  p.nchunk = (int)ceil_div(R, p.NC);
  p.ldc = p.NC + 4;
  p.rows_pad = round_up(rows, 128) + 128;
  AO_CUDA(cudaMalloc(&p.data, p.bytes()));
  AO_CUDA(cudaMemset(p.data, 0, p.bytes()));
  AO_CUDA(cudaStreamSynchronize(0));  // callers use non-blocking streams
}

void packed_factor_free(PackedFactor& p) {
  if (p.data) cudaFree(p.data);
  p.data = nullptr;
}

void packed_factor_pack(const PackedFactor& p, const double* F, int64_t ld, cudaStream_t st, const int* skip) {
  const long long total = (long long)p.nchunk * p.rows_pad * p.ldc;
  const int blocks = (int)std::min<long long>(ceil_div(total, 256), 148 * 8);
  pack_factor_kernel<<<blocks, 256, 0, st>>>(p.data, p.rows_pad, p.ldc, p.NC, p.nchunk, F, p.rows, ld, p.R, skip);
  AO_CHECK_LAUNCH();
}

void packed_factor_to_tf32(const PackedFactor& p, cudaStream_t st, const int* skip) {
  const long long total = (long long)p.nchunk * p.rows_pad * p.ldc;
  const int blocks = (int)std::min<long long>(ceil_div(total, 256), 148 * 8);
  packed_to_tf32_kernel<<<blocks, 256, 0, st>>>(p.data, total, skip);
  AO_CHECK_LAUNCH();
}

void packed_factor_pack_kr(const PackedFactor& p, const double* Fa, int64_t rows_a, int64_t lda, const double* Fb,
                           int64_t rows_b, int64_t ldb, cudaStream_t st, const int* skip) {
  const long long total = (long long)p.nchunk * p.rows_pad * p.ldc;
  const int blocks = (int)std::min<long long>(ceil_div(total, 256), 148 * 8);
  pack_kr_kernel<<<blocks, 256, 0, st>>>(p.data, p.rows_pad, p.ldc, p.NC, p.nchunk, Fa, rows_a, lda, Fb, rows_b, ldb,
                                         p.R, skip);
  AO_CHECK_LAUNCH();
}

size_t mttkrp_workspace_bytes(const Tensor3& t, int R) {
  const int NC = mttkrp_chunk_cols(R);
  const int nchunk = (int)ceil_div(R, NC);
  // (the reduced-precision kernels of mttkrp_tc.cu always work on 64-column chunks)
  const long long Rp = std::max<long long>((long long)nchunk * NC, ceil_div(R, 64) * 64);
  const long long maxdim = std::max(t.I, std::max(t.J, t.K)) + 2;
  // splits never exceed 4 waves of CTAs; epilogue-1 partials are one per j tile
  const long long parts = std::max<long long>(4LL * sm_count(), ceil_div(t.J, 128));
  return (size_t)parts * Rp * maxdim * 8;
}

int64_t mttkrp_T_ld(const Tensor3& t) { return round_up(t.J, 2); }

size_t mttkrp_T_bytes(const Tensor3& t, int R) {
  const int NC = mttkrp_chunk_cols(R);
  const long long Rp = ceil_div(R, NC) * NC;
  return (size_t)t.K * Rp * mttkrp_T_ld(t) * sizeof(double);
}

int mttkrp3_from_T(const Tensor3& t, const double* Tbuf, int R, const double* Fj, int64_t ldf, double scale,
                   double* out, int64_t ldout, cudaStream_t st, const int* skip) {
  const int NC = mttkrp_chunk_cols(R);
  const int Rp = (int)(ceil_div(R, NC) * NC);
  dim3 grid((unsigned)t.K, (unsigned)ceil_div(R, 8));
  mttkrp_from_T_kernel<<<grid, 256, 0, st>>>(Tbuf, mttkrp_T_ld(t), Rp, (int)t.J, (int)t.K, R, Fj, ldf, scale, out, ldout,
                                             skip);
  AO_CHECK_LAUNCH();
  return 1;
}

int mttkrp3(const Tensor3& t, int pos, const PackedFactor& f0, const PackedFactor& f1, int R, double scale,
            double* out, int64_t ldout, const MttkrpWorkspace& w, cudaStream_t st, const int* skip, double* Tbuf,
            int precision) {
  const int NC = mttkrp_chunk_cols(R);
  if (f0.NC != NC || f1.NC != NC || f0.R != R || f1.R != R) throw CudaError(1, "packed factor / rank mismatch");
#define AO_DISPATCH_P(NT, WNN, P)                                                                                     \
  do {                                                                                                                \
    if (pos == 0) return launch_lead<NT, WNN, P>(t, f0, f1, R, scale, out, ldout, w, st, skip);                       \
    if (pos == 1 && Tbuf != nullptr)                                                                                  \
      return launch_inner<NT, WNN, 0, true, P>(t, f0, f1, R, scale, out, ldout, w, Tbuf, st, skip);                   \
    if (pos == 1) return launch_inner<NT, WNN, 0, false, P>(t, f0, f1, R, scale, out, ldout, w, nullptr, st, skip);   \
    return launch_inner<NT, WNN, 1, false, P>(t, f0, f1, R, scale, out, ldout, w, nullptr, st, skip);                 \
  } while (0)
#define AO_DISPATCH(NT, WNN)                 \
  do {                                       \
    if (precision == 1) AO_DISPATCH_P(NT, WNN, 1); \
    AO_DISPATCH_P(NT, WNN, 0);               \
  } while (0)
  if (precision != 0 && precision != 1) throw CudaError(2, "mttkrp3: precision 0 (FP64 DMMA) or 1 (TF32 on mma.sync)");
  switch (NC) {
    case 8: AO_DISPATCH(1, 1);
    case 16: AO_DISPATCH(2, 1);
    case 32: AO_DISPATCH(4, 1);
    default: AO_DISPATCH(4, 2);
  }
#undef AO_DISPATCH_P
#undef AO_DISPATCH
}

}  // namespace aoadmm
