// ptxas scheduling probe: which source form keeps 16 loads in flight before the dependent DFMA chain?
#include <cstdint>
__device__ __forceinline__ double ldnc(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
template <int VAR>
__global__ void __launch_bounds__(512) k(const double* __restrict__ col, long long m, int cnt, int stride,
                                         double* out) {
  __shared__ double xs[512];
  xs[threadIdx.x] = out[threadIdx.x];
  __syncthreads();
  double acc = 0;
  const double* p = col + threadIdx.x;
  const long long step = (long long)stride * m;
  int j = 0;
  for (; j + 16 <= cnt; j += 16) {
    double l[16];
    if (VAR == 0) {
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = col[threadIdx.x + (long long)(j + q) * stride * m];
    } else if (VAR == 1) {
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = __ldg(p + q * step);
      p += 16 * step;
    } else if (VAR == 2) {
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = ldnc(p + q * step);
      p += 16 * step;
      asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
    } else if (VAR == 3) {
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = ldnc(p + q * step);
      p += 16 * step;
      asm volatile("fence.acq_rel.cta;" ::: "memory");
    } else if (VAR == 5) {
      const unsigned mask = __activemask();
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = ldnc(p + q * step);
      p += 16 * step;
      asm volatile("bar.warp.sync %0;" ::"r"(mask) : "memory");
    } else if (VAR == 6) {
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = ldnc(p + q * step);
      p += 16 * step;
      asm volatile("fence.acq_rel.cta;" ::: "memory");
      asm volatile("" ::: "memory");
    } else if (VAR == 7) {
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = ldnc(p + q * step);
      p += 16 * step;
      __nanosleep(0);
    } else if (VAR == 8) {
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = ldnc(p + q * step);
      p += 16 * step;
      asm volatile("membar.cta;" ::: "memory");
    } else if (VAR == 4) {
      double s[4] = {0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < 16; ++q) l[q] = __ldg(p + q * step);
      p += 16 * step;
#pragma unroll
      for (int q = 0; q < 16; ++q) s[q & 3] = fma(l[q], xs[(j + q) * stride], s[q & 3]);
      acc -= (s[0] + s[1]) + (s[2] + s[3]);
      continue;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) acc = fma(-l[q], xs[(j + q) * stride], acc);
  }
  out[threadIdx.x + blockIdx.x * 512] = acc;
}
template __global__ void k<0>(const double*, long long, int, int, double*);
template __global__ void k<1>(const double*, long long, int, int, double*);
template __global__ void k<2>(const double*, long long, int, int, double*);
template __global__ void k<3>(const double*, long long, int, int, double*);
template __global__ void k<4>(const double*, long long, int, int, double*);
template __global__ void k<5>(const double*, long long, int, int, double*);
template __global__ void k<6>(const double*, long long, int, int, double*);
template __global__ void k<7>(const double*, long long, int, int, double*);
template __global__ void k<8>(const double*, long long, int, int, double*);
