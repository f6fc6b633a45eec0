// mttkrp_tc.cu - opt-in reduced-precision MTTKRP on the 5th-generation tensor cores: tcgen05.mma with the accumulator
// in tensor memory (TMEM), operands described by shared-memory matrix descriptors, tensor tiles brought in by TMA.
//
//   options.mttkrp_precision = 1 : TF32 operands (tcgen05.mma.kind::tf32), FP32 accumulation in TMEM
//   options.mttkrp_precision = 2 : BF16 operands (tcgen05.mma.kind::f16),  FP32 accumulation in TMEM
// (north_star: "TF32/BF16 opt-in mode"; the default and parity mode stays FP64 DMMA, mttkrp.cu).  Same operation as
// Tensor Toolbox mttkrp(X,U,n) at functions/cmtf_fun_AOADMM.m:97 - only the operand rounding differs.
//
// The tensor stays FP64 in HBM (it is the user's data), so a pass is bound by the 8-byte-per-element read and the kernel
// is organised around that stream.  For one slab k the MTTKRP of a 3-way tensor is a GEMM
//     CONV 0 (reduction over the contiguous mode i):  T_k(j,:) = sum_i X(i,j,k) F0(i,:)      rows = j
//     CONV 1 (reduction over the middle mode j)     :  T_k(i,:) = sum_j X(i,j,k) F0(j,:)      rows = i
// with a 128-row CTA tile, 32 reduction indices per pipeline stage and 64 rank columns (one chunk) per CTA:
//   warp 0      TMA producer: FP64 tensor tile (32 KB, cp.async.bulk.tensor, 128B swizzle) + the pre-packed low-precision
//               F0 block of the stage (one 1-D bulk copy), two independent mbarrier rings
//   warps 1-4   converters: FP64 tile -> TF32 / BF16 A operand (CONV 1 transposes on the way).  TF32: the operand row goes
//               into TENSOR MEMORY (tcgen05.st, 32 columns per stage) and the MMA takes A from there - the shared-memory
//               port already carries the TMA writes, the converter reads and the B operand, and a 16 KB A tile going in
//               and out of it as well is what kept the first TF32 version at 0.80-0.93 of the HBM rate.  BF16 (half the
//               operand bytes): canonical K-major swizzled UMMA tile in shared memory, fence.proxy.async.
//   warp 5      one elected thread issues tcgen05.mma (M=128, N=64, K=32 bytes per instruction) into one of two TMEM
//               accumulator buffers; tcgen05.commit releases the operand buffers / publishes a finished slab
//   warps 6-9   epilogue: tcgen05.ld the 128 x 64 FP32 slab result, then in FP64 registers
//               EPI 0: O(row,:) += T_k(row,:) .* Fe(k,:)            (modes 1 and 2; accumulated over the CTA's k range)
//               EPI 1: out(k,:)  = sum_rows T_k(row,:) .* Fe(row,:)  (mode 3; reduced over the tile, one partial per tile)
// FP32 accumulation covers one slab of one tile (I or J terms); everything across slabs, splits and tiles is FP64 and
// summed in a fixed order, so the result is deterministic and the error stays at the operand rounding (2^-11 / 2^-8).
#include "mttkrp.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

namespace aoadmm {

int64_t mttkrp_T_ld(const Tensor3& t);

namespace {

constexpr int kNX = 4;               // FP64 tensor-tile ring
constexpr int kNA = 3;               // converted A-operand ring
constexpr int kNB = 4;               // packed B-operand ring
constexpr int kXB = 32768;           // FP64 tile bytes per stage (128 x 32 doubles)
constexpr int kTcThreads = 320;      // 10 warps, roles above
constexpr int kNCols = 64;           // rank columns per CTA (UMMA N)

template <int PREC>
struct Op;
template <>
struct Op<1> {                       // TF32: 32 elements = 128-byte rows, SWIZZLE_128B
  static constexpr int ELT = 4, ROWB = 128, LAYOUT = 2, SBO = 1024, FMT = 2;
};
template <>
struct Op<2> {                       // BF16: 32 elements = 64-byte rows, SWIZZLE_64B
  static constexpr int ELT = 2, ROWB = 64, LAYOUT = 4, SBO = 512, FMT = 1;
};

template <int PREC>
struct Smem {
  static constexpr int BBYTES = kNCols * Op<PREC>::ROWB;   // packed B block
  static constexpr int OFF_X = 0;
  static constexpr int OFF_A = OFF_X + kNX * kXB;
  static constexpr int OFF_B = OFF_A + kNA * 16384;        // (sized for TF32 so that every buffer stays 1024-aligned)
  static constexpr int OFF_RED = OFF_B + kNB * 8192;
  static constexpr int OFF_BAR = OFF_RED + 2 * 4 * kNCols * 8;
  static constexpr int NBAR = 2 * kNX + 2 * kNA + 2 * kNB + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
  static constexpr int BYTES = OFF_TMEM + 16 + 1024;        // + slack for the manual 1024-byte alignment
  static_assert(BYTES <= 227 * 1024, "shared memory budget exceeded");
};

// byte offset of 16-byte chunk `c` of row `row` inside a canonical K-major swizzled operand tile (8-row atoms)
template <int PREC>
__device__ __forceinline__ uint32_t op_chunk_off(int row, int c) {
  if (PREC == 1) return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
  return (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((c ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ uint32_t to_tf32(double x) { return f64_to_tf32(x); }
__device__ __forceinline__ uint32_t pack_bf16(double lo, double hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn((float)lo, (float)hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// shared-memory matrix descriptor (K-major, swizzled): start address, LBO (unused for swizzled K-major: 1), SBO = bytes
// between 8-row atoms, descriptor version 1 (Blackwell), layout type
template <int PREC>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((Op<PREC>::SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)Op<PREC>::LAYOUT << 61;
  return d;
}
// instruction descriptor: D = F32, A/B format, both K-major, N = 64, M = 128
template <int PREC>
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | ((uint32_t)Op<PREC>::FMT << 7) | ((uint32_t)Op<PREC>::FMT << 10) | ((uint32_t)(kNCols >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

template <int PREC>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (PREC == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// A operand taken from tensor memory (128 lanes x K columns of 32 bits), B from shared memory
__device__ __forceinline__ void umma_ts_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one row of 32 consecutive 32-bit columns per lane (lanes 32*(warp%4) .. +31 of tensor memory)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// grid: x = 128-row tile, y = split of the slab range, z = 64-column rank chunk
//   CONV 0: rows = j, reduction over i (nst = ceil(I/32) stages per slab), tensor map box 16(i) x 128(j)
//   CONV 1: rows = i, reduction over j (nst = ceil(J/32) stages per slab), tensor map box 16(i) x 32(j)
//   EPI 0 : ws[(split*Rp + r)*ldo + row] = sum_k T_k(row,r) Fe(k,r);  Tbuf != nullptr additionally stores T_k (dimension tree)
//   EPI 1 : ws[(tile *Rp + r)*ldo + k]   = sum_rows T_k(row,r) Fe(row,r)
//   ATMEM : (TF32) the converters write the A operand into TENSOR MEMORY (tcgen05.st, 3 x 32 columns behind the two
//           accumulator buffers) and the MMA takes A from there: the 16 KB operand tile then neither goes into nor comes
//           out of shared memory, whose bandwidth (TMA writes + converter reads + operand reads) is what bounds the TF32
//           variant with both operands in shared memory
template <int PREC, int CONV, int EPI, int ATMEM>
__global__ void __launch_bounds__(kTcThreads, 1)
mttkrp_tc_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ Bpack, int nblk,
                 const double* __restrict__ Fe, long long ldfe, double* __restrict__ ws, int rows_total, int K, int nst,
                 int nsplit, long long ldo, int Rp_total, int R, double* __restrict__ Tbuf, long long ldt, int RpT,
                 const int* __restrict__ skip) {
  using S = Smem<PREC>;
  if (skip != nullptr && *skip != 0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sX = sbase + S::OFF_X, sA = sbase + S::OFF_A, sB = sbase + S::OFF_B, sBar = sbase + S::OFF_BAR;
  const uint32_t x_full = sBar, x_empty = x_full + kNX * 8, a_full = x_empty + kNX * 8, a_empty = a_full + kNA * 8,
                 b_full = a_empty + kNA * 8, b_empty = b_full + kNB * 8, acc_full = b_empty + kNB * 8,
                 acc_empty = acc_full + 2 * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + S::OFF_TMEM);

  // TMEM columns: [0,128) two accumulator buffers; ATMEM: [128,224) three A-operand buffers of 32 columns
  constexpr int kTmemCols = ATMEM ? 256 : 128;
  constexpr int kACol0 = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y, chunk = blockIdx.z;
  const int k0 = (int)((long long)K * split / nsplit), k1 = (int)((long long)K * (split + 1) / nsplit);
  const int nk = k1 - k0;
  const long long total = (long long)nk * nst;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kNX; ++s) {
      mbar_init(x_full + s * 8, 1);
      mbar_init(x_empty + s * 8, 4);
    }
    for (int s = 0; s < kNA; ++s) {
      mbar_init(a_full + s * 8, 4);
      mbar_init(a_empty + s * 8, 1);
    }
    for (int s = 0; s < kNB; ++s) {
      mbar_init(b_full + s * 8, 1);
      mbar_init(b_empty + s * 8, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full + s * 8, 1);
      mbar_init(acc_empty + s * 8, 4);
    }
    mbar_fence_init();
  }
  if (warp == 0) {   // TMEM: two 64-column FP32 accumulator buffers (allocation granularity: power of two >= 32 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + S::OFF_TMEM), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && total > 0) {
      prefetch_tensormap(&tmap);
      for (long long n = 0; n < total; ++n) {
        const int k = k0 + (int)(n / nst), st = (int)(n % nst);
        const int sx = (int)(n % kNX), sb = (int)(n % kNB);
        mbar_wait(x_empty + sx * 8, (uint32_t)(((n / kNX) & 1) ^ 1));
        mbar_expect_tx(x_full + sx * 8, kXB);
        if (CONV == 0) {
          tma_load_3d(sX + sx * kXB, &tmap, st * 32, tile * 128, k, x_full + sx * 8);
          tma_load_3d(sX + sx * kXB + 16384, &tmap, st * 32 + 16, tile * 128, k, x_full + sx * 8);
        } else {
#pragma unroll
          for (int b = 0; b < 8; ++b) tma_load_3d(sX + sx * kXB + b * 4096, &tmap, tile * 128 + 16 * b, st * 32, k, x_full + sx * 8);
        }
        mbar_wait(b_empty + sb * 8, (uint32_t)(((n / kNB) & 1) ^ 1));
        mbar_expect_tx(b_full + sb * 8, S::BBYTES);
        bulk_load_1d(sB + sb * 8192, Bpack + ((size_t)chunk * nblk + st) * S::BBYTES, S::BBYTES, b_full + sb * 8);
      }
    }
  } else if (warp <= 4) {
    // ===== converters: thread owns GEMM row t of the tile (ATMEM: the row inside the warp's own TMEM lane quarter) =====
    const int t = ATMEM ? ((warp & 3) * 32 + lane) : (threadIdx.x - 32);
    for (long long n = 0; n < total; ++n) {
      const int sx = (int)(n % kNX), ca = (int)(n % kNA);
      mbar_wait(x_full + sx * 8, (uint32_t)((n / kNX) & 1));
      mbar_wait(a_empty + ca * 8, (uint32_t)(((n / kNA) & 1) ^ 1));
      const uint32_t xs = sX + sx * kXB, as = sA + ca * 16384;
      if (CONV == 0) {
        // row j = t of the FP64 tile: 32 consecutive i in two boxes; pair q of box b at ((q ^ (j&7)) << 4)
        const uint32_t rbase = xs + (uint32_t)(t * 128);
        const int sw = t & 7;
        if (PREC == 1 && ATMEM) {
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 16; ++c) {         // pair c = reduction indices 2c, 2c+1
            double x0, x1;
            lds128(rbase + (c >> 3) * 16384 + (((c & 7) ^ sw) << 4), x0, x1);
            v[2 * c] = to_tf32(x0);
            v[2 * c + 1] = to_tf32(x1);
          }
          tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kACol0 + ca * 32), v);
        } else if (PREC == 1) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {          // output chunk c = reduction indices 4c .. 4c+3
            const int b = c >> 2, q = 2 * (c & 3);
            double x0, x1, x2, x3;
            lds128(rbase + b * 16384 + ((q ^ sw) << 4), x0, x1);
            lds128(rbase + b * 16384 + (((q + 1) ^ sw) << 4), x2, x3);
            sts128(as + op_chunk_off<PREC>(t, c), to_tf32(x0), to_tf32(x1), to_tf32(x2), to_tf32(x3));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {          // output chunk c = reduction indices 8c .. 8c+7
            const int b = c >> 1, q = 4 * (c & 1);
            double x[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) lds128(rbase + b * 16384 + (((q + u) ^ sw) << 4), x[2 * u], x[2 * u + 1]);
            sts128(as + op_chunk_off<PREC>(t, c), pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                   pack_bf16(x[6], x[7]));
          }
        }
      } else {
        // row i = t: box t/16, element i_l = t%16 of every tile row j (transposing read: one 8-byte load per j)
        const int il = t & 15;
        const uint32_t cbase = xs + (uint32_t)((t >> 4) * 4096 + (il & 1) * 8);
        const int half = il >> 1;
        if (PREC == 1 && ATMEM) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = to_tf32(lds_f64(cbase + (uint32_t)(j * 128 + ((half ^ (j & 7)) << 4))));
          tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kACol0 + ca * 32), v);
        } else if (PREC == 1) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = 4 * c + u;
              x[u] = lds_f64(cbase + (uint32_t)(j * 128 + ((half ^ (j & 7)) << 4)));
            }
            sts128(as + op_chunk_off<PREC>(t, c), to_tf32(x[0]), to_tf32(x[1]), to_tf32(x[2]), to_tf32(x[3]));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = 8 * c + u;
              x[u] = lds_f64(cbase + (uint32_t)(j * 128 + ((half ^ (j & 7)) << 4)));
            }
            sts128(as + op_chunk_off<PREC>(t, c), pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                   pack_bf16(x[6], x[7]));
          }
        }
      }
      if (ATMEM) tc_fence_before();   // the tcgen05.st above has completed (wait::st); order it before the hand-over
      else fence_async_smem();        // the tensor core reads shared memory through the async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_full + ca * 8);
        mbar_arrive(x_empty + sx * 8);
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer: one thread =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc<PREC>();
      constexpr int NMMA = Op<PREC>::ROWB / 32;   // K = 32 bytes per instruction
      for (long long n = 0; n < total; ++n) {
        const int kl = (int)(n / nst), st = (int)(n % nst);
        const int ca = (int)(n % kNA), sb = (int)(n % kNB);
        const int buf = kl & 1;
        if (st == 0) mbar_wait(acc_empty + buf * 8, (uint32_t)(((kl >> 1) & 1) ^ 1));   // the epilogue has drained this buffer
        mbar_wait(a_full + ca * 8, (uint32_t)((n / kNA) & 1));
        mbar_wait(b_full + sb * 8, (uint32_t)((n / kNB) & 1));
        tc_fence_after();
        const uint64_t adesc = make_desc<PREC>(sA + ca * 16384), bdesc = make_desc<PREC>(sB + sb * 8192);
#pragma unroll
        for (int kk = 0; kk < NMMA; ++kk) {  // advancing 32 bytes along K inside the swizzle atom = +2 in the address field
          if (ATMEM)                         // (A in tensor memory: 8 TF32 columns per instruction)
            umma_ts_tf32(tmem_base + (uint32_t)(buf * kNCols), tmem_base + (uint32_t)(kACol0 + ca * 32 + kk * 8), bdesc + 2 * kk,
                         idesc, (st > 0 || kk > 0) ? 1u : 0u);
          else
            umma<PREC>(tmem_base + (uint32_t)(buf * kNCols), adesc + 2 * kk, bdesc + 2 * kk, idesc, (st > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(a_empty + ca * 8);
        umma_commit(b_empty + sb * 8);
        if (st == nst - 1) umma_commit(acc_full + buf * 8);
      }
    }
  } else {
    // ===== epilogue: warp w reads TMEM lanes 32*(w%4) .. +31; thread = one GEMM row =====
    const int qd = warp & 3;
    const int row_l = qd * 32 + lane;
    const long long row_g = (long long)tile * 128 + row_l;
    const bool row_ok = row_g < rows_total;
    const int col0 = chunk * kNCols;
    double acc[kNCols];   // EPI 0: running output row;  EPI 1: the epilogue factor row Fe(row, :)
#pragma unroll
    for (int c = 0; c < kNCols; ++c) {
      if (EPI == 1) acc[c] = (row_ok && col0 + c < R) ? Fe[row_g + (long long)(col0 + c) * ldfe] : 0.0;
      else acc[c] = 0.0;
    }
    double* red = reinterpret_cast<double*>(gbase + S::OFF_RED);
    for (int kl = 0; kl < nk; ++kl) {
      const int buf = kl & 1, k = k0 + kl;
      mbar_wait(acc_full + buf * 8, (uint32_t)((kl >> 1) & 1));
      tc_fence_after();
      double* rb = red + (kl & 1) * (4 * kNCols);
#pragma unroll
      for (int c16 = 0; c16 < kNCols / 16; ++c16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * kNCols + c16 * 16), v);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int c = c16 * 16 + e;
          const double tv = (double)__uint_as_float(v[e]);
          if (EPI == 0) {
            if (col0 + c < R) {
              acc[c] = fma(tv, Fe[k + (long long)(col0 + c) * ldfe], acc[c]);
              if (Tbuf != nullptr && row_ok) Tbuf[((long long)k * RpT + col0 + c) * ldt + row_g] = tv;
            }
          } else {
            double s = tv * acc[c];
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (lane == 0) rb[qd * kNCols + c] = s;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + buf * 8);
      if (EPI == 1) {
        named_bar_sync(1, 128);   // the four epilogue warps
        const int te = threadIdx.x - 6 * 32;
        if (te < kNCols) {
          const double s = (rb[te] + rb[kNCols + te]) + (rb[2 * kNCols + te] + rb[3 * kNCols + te]);
          ws[((long long)tile * Rp_total + col0 + te) * ldo + k] = s;
        }
      }
    }
    if (EPI == 0 && row_ok) {
#pragma unroll
      for (int c = 0; c < kNCols; ++c) ws[((long long)split * Rp_total + col0 + c) * ldo + row_g] = acc[c];
    }
  }

  // ===== teardown =====
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// out(row, r) = scale * sum_s ws[(s*Rp + r)*ldo + row]   (fixed order => deterministic)
__global__ void tc_reduce_kernel(const double* __restrict__ ws, int nparts, int Rp_total, long long ldo, long long rows, int R,
                                 double scale, double* __restrict__ out, long long ldout, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (row >= rows || r >= R) return;
  double v = 0.0;
  for (int s = 0; s < nparts; ++s) v += ws[((long long)s * Rp_total + r) * ldo + row];
  out[(long long)r * ldout + row] = scale * v;
}

// F (rows x R, leading dimension ld, FP64) -> per 64-column chunk and 32-row block one operand block in the canonical
// K-major swizzled UMMA layout: element (n = column inside the chunk, kk = row inside the block)
template <int PREC>
__global__ void tc_pack_kernel(uint8_t* __restrict__ dst, int nblk, int nchunk, const double* __restrict__ F, long long rows,
                               long long ld, int R, const int* __restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  constexpr int BB = Smem<PREC>::BBYTES, ELT = Op<PREC>::ELT;
  const long long nelem = (long long)nchunk * nblk * kNCols * 32;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(e % 32);                 // row inside the block: consecutive threads read consecutive rows
    const int n = (int)((e / 32) % kNCols);
    const long long blk = e / (32 * kNCols);      // chunk * nblk + block
    const int c = (int)(blk / nblk);
    const long long row = (blk % nblk) * 32 + kk;
    const int col = c * kNCols + n;
    const double v = (row < rows && col < R) ? F[(long long)col * ld + row] : 0.0;
    const uint32_t off = op_chunk_off<PREC>(n, (kk * ELT) >> 4) + (uint32_t)((kk * ELT) & 15);
    uint8_t* p = dst + (size_t)blk * BB + off;
    if (PREC == 1) *reinterpret_cast<uint32_t*>(p) = to_tf32(v);
    else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn((float)v);
  }
}

int tc_sm_count() {
  int dev = 0, n = 0;
  AO_CUDA(cudaGetDevice(&dev));
  AO_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

// The slab range is split in WHOLE slabs: a CTA's work is ceil(K / splits) slabs and the pass lasts ceil(K / splits) *
// waves slab-times; pick the split count that minimises that makespan (ties: fewer splits, less to reduce).
int tc_choose_splits(long long tiles, long long slabs, long long cap) {
  const int sms = tc_sm_count();
  long long best = 1;
  double best_cost = 1e300;
  const long long hi = std::max<long long>(1, std::min<long long>(std::min(slabs, cap), 592));
  for (long long ns = 1; ns <= hi; ++ns) {
    const long long waves = ceil_div(tiles * ns, sms);
    const double cost = (double)ceil_div(slabs, ns) * (double)waves + 0.25 * (double)waves;
    if (cost < best_cost * (1.0 - 1e-9)) {
      best_cost = cost;
      best = ns;
    }
  }
  return (int)best;
}

template <int PREC, int CONV, int EPI, int ATMEM>
int launch_tc(const Tensor3& t, const uint8_t* Bpack, int nblk, int nchunk, const double* Fe, int64_t ldfe, int R,
              double scale, double* out, int64_t ldout, const MttkrpWorkspace& w, double* Tbuf, cudaStream_t st,
              const int* skip) {
  using S = Smem<PREC>;
  const long long rows_total = (CONV == 0) ? t.J : t.I;
  const int nst = (int)ceil_div((CONV == 0) ? t.I : t.J, 32);
  const int ntiles = (int)ceil_div(rows_total, 128);
  const int Rp_total = nchunk * kNCols;
  const long long out_rows = (EPI == 0) ? rows_total : t.K;
  const long long ldo = round_up(out_rows, 2);
  const long long cap = (long long)(w.ws_bytes / ((size_t)Rp_total * ldo * 8));   // partial results the workspace can hold
  if (cap < 1 || (EPI == 1 && cap < ntiles)) throw CudaError(1, "mttkrp workspace too small (reduced-precision path)");
  const int nsplit = tc_choose_splits((long long)ntiles * nchunk, t.K, EPI == 0 ? cap : (long long)t.K);
  const int nparts = (EPI == 0) ? nsplit : ntiles;
  auto kern = mttkrp_tc_kernel<PREC, CONV, EPI, ATMEM>;
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), S::BYTES);
  int RpT = 0;
  long long ldt = 0;
  if (Tbuf != nullptr) {
    const int NC = mttkrp_chunk_cols(R);
    RpT = (int)(ceil_div(R, NC) * NC);
    ldt = mttkrp_T_ld(t);
  }
  dim3 grid(ntiles, nsplit, nchunk);
  kern<<<grid, kTcThreads, S::BYTES, st>>>((CONV == 0) ? t.map_inner : t.map_lead, Bpack, nblk, Fe, (long long)ldfe, w.ws,
                                          (int)rows_total, (int)t.K, nst, nsplit, ldo, Rp_total, R, Tbuf, ldt, RpT, skip);
  AO_CHECK_LAUNCH();
  dim3 rgrid((unsigned)ceil_div(out_rows, 128), R);
  tc_reduce_kernel<<<rgrid, 128, 0, st>>>(w.ws, nparts, Rp_total, ldo, out_rows, R, scale, out, ldout, skip);
  AO_CHECK_LAUNCH();
  return 2;
}

template <int PREC>
int dispatch_tc(const Tensor3& t, int pos, const TcOperand& op, const double* F0, int64_t ld0, const double* Fe, int64_t ldfe,
                int R, double scale, double* out, int64_t ldout, const MttkrpWorkspace& w, double* Tbuf, cudaStream_t st,
                const int* skip) {
  // pos 0: rows i, reduction over j with F0 = Fj, Fe = Fk;   pos 1: rows j, reduction over i with F0 = Fi, Fe = Fk;
  // pos 2: rows j, reduction over i with F0 = Fi, Fe = Fj (row factor), reduced over the tile
  const long long red_rows = (pos == 0) ? t.J : t.I;
  const int nblk = (int)ceil_div(red_rows, 32), nchunk = (int)ceil_div(R, kNCols);
  if ((size_t)nchunk * nblk * Smem<PREC>::BBYTES > op.bytes) throw CudaError(1, "reduced-precision operand buffer too small");
  const long long nelem = (long long)nchunk * nblk * kNCols * 32;
  tc_pack_kernel<PREC><<<(unsigned)std::min<long long>(ceil_div(nelem, 256), 148 * 8), 256, 0, st>>>(
      op.data, nblk, nchunk, F0, red_rows, ld0, R, skip);
  AO_CHECK_LAUNCH();
  int n = 1;
  // TF32: A operand through tensor memory (AOADMM_TC_A_SMEM=1 keeps it in shared memory, for comparison); BF16: the
  // operand tile is half the size and the pass is already HBM bound with both operands in shared memory
  static const bool a_smem = (std::getenv("AOADMM_TC_A_SMEM") != nullptr);
  constexpr int AT = (PREC == 1) ? 1 : 0;
  if (PREC == 1 && a_smem) {
    if (pos == 0) n += launch_tc<PREC, 1, 0, 0>(t, op.data, nblk, nchunk, Fe, ldfe, R, scale, out, ldout, w, nullptr, st, skip);
    else if (pos == 1) n += launch_tc<PREC, 0, 0, 0>(t, op.data, nblk, nchunk, Fe, ldfe, R, scale, out, ldout, w, Tbuf, st, skip);
    else n += launch_tc<PREC, 0, 1, 0>(t, op.data, nblk, nchunk, Fe, ldfe, R, scale, out, ldout, w, nullptr, st, skip);
    return n;
  }
  if (pos == 0) n += launch_tc<PREC, 1, 0, AT>(t, op.data, nblk, nchunk, Fe, ldfe, R, scale, out, ldout, w, nullptr, st, skip);
  else if (pos == 1) n += launch_tc<PREC, 0, 0, AT>(t, op.data, nblk, nchunk, Fe, ldfe, R, scale, out, ldout, w, Tbuf, st, skip);
  else n += launch_tc<PREC, 0, 1, AT>(t, op.data, nblk, nchunk, Fe, ldfe, R, scale, out, ldout, w, nullptr, st, skip);
  return n;
}

}  // namespace

size_t tc_operand_bytes(int64_t rows, int R) {
  return (size_t)ceil_div(R, kNCols) * (size_t)ceil_div(std::max<int64_t>(rows, 1), 32) * Smem<1>::BBYTES;
}

void tc_operand_alloc(TcOperand& op, int64_t max_rows, int R) {
  op.bytes = tc_operand_bytes(max_rows, R);
  AO_CUDA(cudaMalloc(&op.data, op.bytes));
}

void tc_operand_free(TcOperand& op) {
  if (op.data) cudaFree(op.data);
  op.data = nullptr;
  op.bytes = 0;
}

int mttkrp3_tc(const Tensor3& t, int pos, const TcOperand& op, const double* F0, int64_t ld0, const double* Fe, int64_t ldfe,
               int R, double scale, double* out, int64_t ldout, const MttkrpWorkspace& w, cudaStream_t st, const int* skip,
               double* Tbuf, int precision) {
  if (pos < 0 || pos > 2) throw CudaError(1, "mttkrp3_tc: mode position out of range");
  if (Tbuf != nullptr && pos != 1) throw CudaError(1, "mttkrp3_tc: the partial contraction is emitted by the mode-2 pass only");
  if (precision == 1) return dispatch_tc<1>(t, pos, op, F0, ld0, Fe, ldfe, R, scale, out, ldout, w, Tbuf, st, skip);
  if (precision == 2) return dispatch_tc<2>(t, pos, op, F0, ld0, Fe, ldfe, R, scale, out, ldout, w, Tbuf, st, skip);
  throw CudaError(2, "mttkrp3_tc: precision must be 1 (TF32) or 2 (BF16)");
}

}  // namespace aoadmm
