// FP64 peak probe for B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate, cuBLAS DGEMM
// (library peak probe only - not on the product path), HBM read bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_probe tools/fp64_probe.cu -lcublas
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

// every launch is followed by CK(cudaGetLastError()): a launch that fails (e.g. too many resources for 32 warps x 16
// accumulators) must not be timed - an unlaunched kernel "runs" in microseconds and prints an absurd rate
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void dmma_rate(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template<int NACC>
__global__ void dfma_rate(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = 1e-9;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void read_bw(const double2* __restrict__ x, size_t n2, double* out) {
  double s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 v = x[i]; s += v.x + v.y;
  }
  if (s == 12345.678) out[0] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; CK(cudaMalloc(&out, 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = p.multiProcessorCount;
  // DMMA
  for (int warps : {4, 8, 16, 32}) {
    int iters = 20000;
    auto run = [&](int nacc) {
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (nacc == 4) dmma_rate<4><<<sms, warps * 32>>>(out, iters, 1.0); CK(cudaGetLastError());
        if (nacc == 8) dmma_rate<8><<<sms, warps * 32>>>(out, iters, 1.0); CK(cudaGetLastError());
        if (nacc == 16) dmma_rate<16><<<sms, warps * 32>>>(out, iters, 1.0); CK(cudaGetLastError());
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      }
      float ms = time_ms(e0, e1);
      double flops = 2.0 * 256 * nacc * (double)iters * warps * sms;
      printf("DMMA.884 warps/SM=%2d nacc=%2d : %.3f ms  %.2f TFLOP/s\n", warps, nacc, ms, flops / ms * 1e-9);
    };
    run(4); run(8); run(16);
  }
  for (int warps : {4, 8, 16, 32}) {
    int iters = 20000;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      dfma_rate<16><<<sms, warps * 32>>>(out, iters, 1.0); CK(cudaGetLastError());
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    }
    float ms = time_ms(e0, e1);
    double flops = 2.0 * 32 * 16 * (double)iters * warps * sms;
    printf("DFMA warps/SM=%2d nacc=16 : %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  // sustained DMMA for ~2 s
  {
    int iters = 2000000;
    cudaEventRecord(e0);
    dmma_rate<16><<<sms, 16 * 32>>>(out, iters, 1.0); CK(cudaGetLastError());
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1);
    double flops = 2.0 * 256 * 16 * (double)iters * 16 * sms;
    printf("DMMA.884 sustained warps/SM=16 nacc=16 : %.1f ms  %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  // HBM read
  {
    size_t bytes = (size_t)8 << 30;
    double2* x; CK(cudaMalloc(&x, bytes)); CK(cudaMemset(x, 0, bytes));
    for (int blocks : {sms * 4, sms * 8, sms * 16}) {
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        read_bw<<<blocks, 512>>>(x, bytes / 16, out); CK(cudaGetLastError());
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      }
      float ms = time_ms(e0, e1);
      printf("HBM read 8GiB blocks=%d : %.3f ms  %.1f GB/s\n", blocks, ms, bytes / ms * 1e-6);
    }
    CK(cudaFree(x));
  }
  // cuBLAS DGEMM
  {
    cublasHandle_t h; cublasCreate(&h);
    struct Shape { int m, n, k; const char* name; };
    std::vector<Shape> shapes = {{8192, 8192, 8192, "square"}, {1000000, 32, 1000, "C2 mode-1 shape (IJ x R x K)"},
                                 {1000, 32, 1000000, "C2 X(1) x KR (I x R x JK)"}, {4096, 64, 4194304 / 4, "C3-ish X(1) x KR"}};
    for (auto s : shapes) {
      double *A, *B, *C;
      CK(cudaMalloc(&A, (size_t)s.m * s.k * 8)); CK(cudaMalloc(&B, (size_t)s.k * s.n * 8)); CK(cudaMalloc(&C, (size_t)s.m * s.n * 8));
      CK(cudaMemset(A, 0, (size_t)s.m * s.k * 8)); CK(cudaMemset(B, 0, (size_t)s.k * s.n * 8));
      double one = 1, zero = 0;
      float best = 1e30f;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, s.m, s.n, s.k, &one, A, s.m, B, s.k, &zero, C, s.m);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms = time_ms(e0, e1); if (rep > 0 && ms < best) best = ms;
      }
      double flops = 2.0 * s.m * s.n * (double)s.k;
      printf("cuBLAS DGEMM %s m=%d n=%d k=%d : %.3f ms  %.2f TFLOP/s  A-read %.1f GB/s\n", s.name, s.m, s.n, s.k, best,
             flops / best * 1e-9, (double)s.m * s.k * 8 / best * 1e-6);
      cudaFree(A); cudaFree(B); cudaFree(C);
    }
    cublasDestroy(h);
  }
  return 0;
}
