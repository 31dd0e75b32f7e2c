// Microbenchmark behind femwell_b200/csrc/lu_solve.cu row_dot_sub: how must the strided factor loads of the
// substitution kernels be written so that ptxas really keeps a batch of them in flight?  Mirrors k_fwd_update
// (128 rows x 4 column groups per CTA, 256-column chunks, x in shared memory).  Prints GB/s per variant.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int ROWS = 128, G = 4, CHUNK = 256;
__device__ __forceinline__ double ldnc(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
template <int VAR, int B>
__device__ __forceinline__ void rowdot(double& acc, const double* __restrict__ col, long long m, int cnt,
                                       const double* xs, int x0, int stride, double* stage) {
  int j = 0;
  const double* p = col;
  const long long step = (long long)stride * m;
  for (; j + B <= cnt; j += B) {
    double l[B];
    if (VAR == 0) {
#pragma unroll
      for (int q = 0; q < B; ++q) l[q] = col[(long long)(j + q) * stride * m];
    } else if (VAR == 1) {
#pragma unroll
      for (int q = 0; q < B; ++q) l[q] = __ldg(p + q * step);
    } else if (VAR == 2) {
#pragma unroll
      for (int q = 0; q < B; ++q) l[q] = ldnc(p + q * step);
      asm volatile("fence.acq_rel.cta;" ::: "memory");
    } else if (VAR == 3) {
#pragma unroll
      for (int q = 0; q < B; ++q) l[q] = ldnc(p + q * step);
      asm volatile("membar.cta;" ::: "memory");
    } else if (VAR == 4) {
      // LDGSTS into this thread's own staging slots
#pragma unroll
      for (int q = 0; q < B; ++q) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(stage + q * (ROWS * G) + threadIdx.x);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(p + q * step) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
      for (int q = 0; q < B; ++q) l[q] = stage[q * (ROWS * G) + threadIdx.x];
    }
    p += B * step;
#pragma unroll
    for (int q = 0; q < B; ++q) acc = fma(-l[q], xs[x0 + (j + q) * stride], acc);
  }
}
template <int VAR, int B>
__global__ void __launch_bounds__(ROWS* G) k_upd(const double* __restrict__ F, long long front_entries, int m, int ns,
                                                 double* w) {
  extern __shared__ double stage[];
  __shared__ double xs[CHUNK];
  __shared__ double part[G][ROWS];
  const double* Fn = F + blockIdx.x * front_entries;
  double* wn = w + (size_t)blockIdx.x * m;
  const int t = threadIdx.x, lr = t % ROWS, g = t / ROWS;
  const int row = ns + blockIdx.y * ROWS + lr;
  const bool valid = row < m;
  double acc = 0;
  for (int j0 = blockIdx.z * CHUNK; j0 < ns; j0 += CHUNK * gridDim.z) {
    const int cn = min(CHUNK, ns - j0);
    for (int j = t; j < cn; j += ROWS * G) xs[j] = wn[j0 + j];
    __syncthreads();
    if (valid && g < cn) {
      const int cnt = (cn - g + G - 1) / G;
      rowdot<VAR, B>(acc, Fn + row + (long long)(j0 + g) * m, m, cnt, xs, g, G, stage);
    }
    __syncthreads();
  }
  part[g][lr] = acc;
  __syncthreads();
  if (g == 0 && valid) atomicAdd(&wn[row], part[0][lr] + part[1][lr] + part[2][lr] + part[3][lr]);
}
template <int VAR, int B>
void run(const char* name, const double* F, long long fe, int nf, int m, int ns, double* w, int nz) {
  const int tiles = (m - ns + ROWS - 1) / ROWS;
  const size_t sm = VAR == 4 ? (size_t)B * ROWS * G * 8 : 0;
  if (sm) cudaFuncSetAttribute(k_upd<VAR, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int it = 0; it < 3; ++it) k_upd<VAR, B><<<dim3(nf, tiles, nz), ROWS * G, sm>>>(F, fe, m, ns, w);
  cudaEventRecord(a);
  const int reps = 10;
  for (int it = 0; it < reps; ++it) k_upd<VAR, B><<<dim3(nf, tiles, nz), ROWS * G, sm>>>(F, fe, m, ns, w);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)nf * (m - ns) * (double)ns * 8.0;
  printf("%-28s nf=%d m=%d ns=%d nz=%d  %8.3f ms  %8.1f GB/s  (%s)\n", name, nf, m, ns, nz, ms / reps,
         bytes / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char** argv) {
  struct Case { int nf, m, ns, nz; } cases[] = {{16, 3000, 1500, 4}, {64, 1200, 600, 2}, {512, 400, 200, 1}};
  for (auto c : cases) {
    const long long fe = (long long)c.m * c.m;
    double *F, *w;
    cudaMalloc(&F, fe * c.nf * 8);
    cudaMalloc(&w, (size_t)c.nf * c.m * 8);
    cudaMemset(F, 0, fe * c.nf * 8);
    cudaMemset(w, 0, (size_t)c.nf * c.m * 8);
    run<0, 16>("index (current) B16", F, fe, c.nf, c.m, c.ns, w, c.nz);
    run<1, 16>("pointer-step B16", F, fe, c.nf, c.m, c.ns, w, c.nz);
    run<2, 8>("fence.acq_rel.cta B8", F, fe, c.nf, c.m, c.ns, w, c.nz);
    run<2, 16>("fence.acq_rel.cta B16", F, fe, c.nf, c.m, c.ns, w, c.nz);
    run<2, 32>("fence.acq_rel.cta B32", F, fe, c.nf, c.m, c.ns, w, c.nz);
    run<3, 16>("membar.cta B16", F, fe, c.nf, c.m, c.ns, w, c.nz);
    run<4, 8>("cp.async stage B8", F, fe, c.nf, c.m, c.ns, w, c.nz);
    run<4, 16>("cp.async stage B16", F, fe, c.nf, c.m, c.ns, w, c.nz);
    cudaFree(F);
    cudaFree(w);
  }
  return 0;
}
