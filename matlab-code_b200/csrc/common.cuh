// common.cuh - shared device/host helpers for the AO-ADMM B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <stdexcept>
#include <utility>

namespace aoadmm {

struct CudaError : public std::runtime_error {
  int code;
  CudaError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define AO_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      throw ::aoadmm::CudaError((_e == cudaErrorMemoryAllocation) ? 7 : 5,                     \
                                std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                                    __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
    }                                                                                          \
  } while (0)

#define AO_CHECK_LAUNCH() AO_CUDA(cudaGetLastError())

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) acts on the CURRENT device only, and one process may drive several
// GPUs from several host threads (aoadmm_create_multi): remember the largest size configured per (device, kernel).
// No attribute call happens once a size is configured, so the steady state stays legal inside CUDA-graph capture.
inline void ensure_dynamic_smem(const void* kern, size_t bytes, size_t threshold = 48 * 1024) {
  if (bytes <= threshold) return;
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> configured;
  int dev = 0;
  AO_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = configured[std::make_pair(dev, kern)];
  if (bytes > cur) {
    AO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
  }
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA (cp.async.bulk[.tensor]), FP64 tensor-core MMA (DMMA.8x8x4)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 3-D tiled TMA load (tensor map in kernel param space), completes on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned src/dst/size), completes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// D(8x8) += A(8x4) * B(4x8), FP64 tensor core.  SASS: DMMA.8x8x4
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// D(16x8) += A(16x8) * B(8x8), TF32 inputs, FP32 accumulate (opt-in reduced-precision MTTKRP).  SASS: HMMA.1688.F32.TF32
__device__ __forceinline__ void mma_tf32_1688(float& c0, float& c1, float& c2, float& c3, uint32_t a0, uint32_t a1,
                                              uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};"
               : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t f64_to_tf32(double x) {
  uint32_t r;
  const float f = (float)x;
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ uint32_t f32_to_tf32(float f) {
  uint32_t r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// deterministic block sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= 32 doubles */) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = (lane < nw) ? scratch[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}
// Final step of a two-level deterministic reduction, executed by ONE WARP of the last CTA: sum of the n per-CTA
// partials p[0], p[stride], p[2*stride], ... in a fixed order (lane l takes l, l+32, ...; then the shuffle tree).  A
// single thread walking the partials pays one L2 round trip per handful of loads - 20 us for 256 CTAs, which was most of
// an inner ADMM iteration; a warp issues all its loads at once.  Result valid in every lane.
__device__ __forceinline__ double warp_sum_partials(const double* p, unsigned n, unsigned stride) {
  const unsigned lane = threadIdx.x & 31;
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  unsigned b = lane;
  // __ldcg: the partials were written by other SMs (and this SM may hold last iteration's lines in L1)
  for (; b + 96 < n; b += 128) {
    v0 += __ldcg(p + (size_t)b * stride);
    v1 += __ldcg(p + (size_t)(b + 32) * stride);
    v2 += __ldcg(p + (size_t)(b + 64) * stride);
    v3 += __ldcg(p + (size_t)(b + 96) * stride);
  }
  for (; b < n; b += 32) v0 += __ldcg(p + (size_t)b * stride);
  return warp_sum((v0 + v1) + (v2 + v3));
}
#endif

}  // namespace aoadmm
