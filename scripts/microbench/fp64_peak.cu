// FP64 issue-rate microbenchmark for sm_100a: DFMA (SIMT) vs DMMA (mma.sync.m8n8k4.f64) on register operands.
// Answers: which instruction should the Schur-complement GEMM of the multifrontal LU be built on, and what is
// the achievable FP64 ceiling that `roofline.lu_factor` should be read against?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c0[i] = threadIdx.x * 1e-3 + i, c1[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int bps : {1, 2, 4}) {
      if (threads * bps > 2048) continue;
      const int blocks = 148 * bps;
      double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 16 * iters * (double)blocks * threads;
      printf("DFMA  threads %4d blocks/SM %d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms / 1e9);
      ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      fl = 512.0 * 8 * iters * (double)blocks * (threads / 32);
      printf("DMMA8 threads %4d blocks/SM %d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms / 1e9);
      ms = time_ms([&] { k_dmma<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      fl = 512.0 * 16 * iters * (double)blocks * (threads / 32);
      printf("DMMA16 threads %4d blocks/SM %d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms / 1e9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
