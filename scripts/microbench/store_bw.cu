// Micro-benchmark: pure-store bandwidth with the K1 output pattern (SoA planes, one thread per triangle).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_store_planes(int ne, int planes, double* __restrict__ out, double seed) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  double v = seed + e * 1e-9;
  for (int k = blockIdx.y; k < planes; k += gridDim.y) out[(size_t)k * ne + e] = v + k;
}
__global__ void k_store_linear(size_t n, double* __restrict__ out, double seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = seed + i * 1e-9;
}
__global__ void k_copy(size_t n, const double* __restrict__ in, double* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
int main() {
  const int ne = 500000, planes = 260;
  size_t n = (size_t)ne * planes;
  double *a, *b;
  cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int gy : {1, 14, 28, 260}) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_store_planes<<<dim3((ne + 127) / 128, gy), 128>>>(ne, planes, a, 1.0 + rep);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("store planes gy=%3d: %.3f ms  %.1f GB/s\n", gy, ms, n * 8 / ms / 1e6);
  }
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k_store_linear<<<148 * 16, 256>>>(n, a, 2.0 + rep); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("store linear      : %.3f ms  %.1f GB/s\n", ms, n * 8 / ms / 1e6);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k_copy<<<148 * 16, 256>>>(n, a, b); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("copy (r+w bytes)  : %.3f ms  %.1f GB/s\n", ms, 2 * n * 8 / ms / 1e6);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); cudaMemsetAsync(a, 0, n * 8); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("memset zero       : %.3f ms  %.1f GB/s\n", ms, n * 8 / ms / 1e6);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
