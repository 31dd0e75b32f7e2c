// Multi-right-hand-side Jacobi-preconditioned CG for the L2 projection of calculate_hfield
// (femwell/maxwell/waveguide.py:666-677: scipy.sparse.linalg.spsolve(bform.assemble(basis), ...)).
//
// The mass matrix (block diagonal: Nedelec mass + Lagrange mass) is real SPD and mode independent, the
// right-hand sides are the k complex vectors C(beta_i) e_i.  All k systems advance together so that the
// matrix is streamed once per iteration for all modes (SpMV with k vectors), and every CG scalar
// (alpha, beta, residual norms) lives in device memory -- the host only looks at the residuals every
// few iterations.  HBM-bound: nnz*(8+4) + k*N*16*~8 bytes per iteration.
#include <algorithm>

#include "solver.h"

namespace fw {

constexpr int CG_MAXK = 8;

// y[r][row] = sum_c M[row,c] x[r][c], 8 lanes per row, all k vectors per matrix pass
__global__ void __launch_bounds__(256) k_spmv_multi(int n, int k, const int* __restrict__ indptr,
                                                    const int* __restrict__ indices, const double* __restrict__ vals,
                                                    const cplx* __restrict__ x, cplx* __restrict__ y) {
  constexpr int TPR = 8;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(gt / TPR), l = (int)(gt % TPR);
  cplx acc[CG_MAXK];
#pragma unroll
  for (int r = 0; r < CG_MAXK; ++r) acc[r] = cplx{0.0, 0.0};
  if (row < n) {
    const int a = indptr[row], b = indptr[row + 1];
    for (int e = a + l; e < b; e += TPR) {
      const double v = vals[e];
      if (v == 0.0) continue;  // structural zeros of the cross blocks on the shared pattern
      const int c = indices[e];
#pragma unroll
      for (int r = 0; r < CG_MAXK; ++r)
        if (r < k) fma_(acc[r], v, x[(size_t)r * n + c]);
    }
  }
#pragma unroll
  for (int r = 0; r < CG_MAXK; ++r) {
    if (r < k) {
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) {
        acc[r].re += __shfl_xor_sync(0xffffffffu, acc[r].re, o);
        acc[r].im += __shfl_xor_sync(0xffffffffu, acc[r].im, o);
      }
      if (l == 0 && row < n) y[(size_t)r * n + row] = acc[r];
    }
  }
}

__global__ void k_cg_diag(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                          const double* __restrict__ vals, double* __restrict__ idiag) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0;
  for (int e = indptr[r]; e < indptr[r + 1]; ++e)
    if (indices[e] == r) d = vals[e];
  idiag[r] = d != 0.0 ? 1.0 / d : 0.0;
}

__device__ inline double block_sum_256(double s, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double q = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) q += red[i];
  __syncthreads();
  return q;  // valid in thread 0
}

// partial[(r*2+0)*nblk + b] = Re<a_r, b_r> over the rows of block b (slot `which` of 2)
__global__ void __launch_bounds__(256) k_cg_dot(int n, int k, const cplx* __restrict__ a, const cplx* __restrict__ b,
                                                double* __restrict__ partial, int nblk) {
  __shared__ double red[8];
  const int64_t base = (int64_t)blockIdx.x * 1024;
  for (int r = 0; r < k; ++r) {
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int64_t i = base + q * 256 + threadIdx.x;
      if (i < n) {
        const cplx u = a[(size_t)r * n + i], v = b[(size_t)r * n + i];
        s += u.re * v.re + u.im * v.im;
      }
    }
    double t = block_sum_256(s, red);
    if (threadIdx.x == 0) partial[(size_t)r * nblk + blockIdx.x] = t;
  }
}

// out[r] = sum_b partial[r*nblk + b]  (one block per (r, slot))
__global__ void __launch_bounds__(256) k_cg_reduce(const double* __restrict__ partial, int nblk,
                                                   double* __restrict__ out) {
  __shared__ double sh[256];
  const int r = blockIdx.x, t = threadIdx.x;
  double s = 0;
  for (int b = t; b < nblk; b += 256) s += partial[(size_t)r * nblk + b];
  sh[t] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  if (t == 0) out[r] = sh[0];
}

// x += alpha p ; r -= alpha q ; z = r / diag ; partials of <r,r> and <r,z>     (alpha_r = rz_r / pq_r)
__global__ void __launch_bounds__(256) k_cg_step(int n, int k, cplx* __restrict__ x, cplx* __restrict__ rv,
                                                 const cplx* __restrict__ p, const cplx* __restrict__ q,
                                                 cplx* __restrict__ z, const double* __restrict__ idiag,
                                                 const double* __restrict__ rz, const double* __restrict__ pq,
                                                 double* __restrict__ part_rr, double* __restrict__ part_rz,
                                                 int nblk) {
  __shared__ double red[8];
  const int64_t base = (int64_t)blockIdx.x * 1024;
  for (int r = 0; r < k; ++r) {
    const double den = pq[r];
    const double alpha = (den != 0.0 && isfinite(den)) ? rz[r] / den : 0.0;
    double s_rr = 0, s_rz = 0;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      int64_t i = base + qq * 256 + threadIdx.x;
      if (i < n) {
        const size_t o = (size_t)r * n + i;
        x[o] = x[o] + p[o] * alpha;
        const cplx res = rv[o] - q[o] * alpha;
        rv[o] = res;
        const cplx zz = res * idiag[i];
        z[o] = zz;
        s_rr += res.re * res.re + res.im * res.im;
        s_rz += res.re * zz.re + res.im * zz.im;
      }
    }
    double t1 = block_sum_256(s_rr, red);
    double t2 = block_sum_256(s_rz, red);
    if (threadIdx.x == 0) {
      part_rr[(size_t)r * nblk + blockIdx.x] = t1;
      part_rz[(size_t)r * nblk + blockIdx.x] = t2;
    }
  }
}

// p = z + (rz_new / rz_old) p
__global__ void k_cg_direction(int n, int k, cplx* __restrict__ p, const cplx* __restrict__ z,
                               const double* __restrict__ rz_new, const double* __restrict__ rz_old) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int r = 0; r < k; ++r) {
    const double den = rz_old[r];
    const double beta = (den != 0.0 && isfinite(den)) ? rz_new[r] / den : 0.0;
    const size_t o = (size_t)r * n + i;
    p[o] = z[o] + p[o] * beta;
  }
}

// z = r / diag, partials of <r,r>, <r,z>  (initialisation)
__global__ void __launch_bounds__(256) k_cg_init(int n, int k, const cplx* __restrict__ rv, cplx* __restrict__ z,
                                                 const double* __restrict__ idiag, double* __restrict__ part_rr,
                                                 double* __restrict__ part_rz, int nblk) {
  __shared__ double red[8];
  const int64_t base = (int64_t)blockIdx.x * 1024;
  for (int r = 0; r < k; ++r) {
    double s_rr = 0, s_rz = 0;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      int64_t i = base + qq * 256 + threadIdx.x;
      if (i < n) {
        const size_t o = (size_t)r * n + i;
        const cplx res = rv[o];
        const cplx zz = res * idiag[i];
        z[o] = zz;
        s_rr += res.re * res.re + res.im * res.im;
        s_rz += res.re * zz.re + res.im * zz.im;
      }
    }
    double t1 = block_sum_256(s_rr, red);
    double t2 = block_sum_256(s_rz, red);
    if (threadIdx.x == 0) {
      part_rr[(size_t)r * nblk + blockIdx.x] = t1;
      part_rz[(size_t)r * nblk + blockIdx.x] = t2;
    }
  }
}

// Mass X = B for k <= CG_MAXK complex right-hand sides ([k][n] each); X overwritten.
void mass_solve_multi(Handle& h, int k, const cplx* B, cplx* X) {
  FW_REQUIRE(k >= 1 && k <= CG_MAXK, "mass_solve_multi: 1 <= k <= 8");
  FW_REQUIRE(h.Mass.valid, "mass matrix missing");
  const int n = h.sp.n;
  cudaStream_t s = h.stream;
  const Pattern& pat = h.pat;
  const int nblk = cdiv(n, 1024);
  const size_t vec = (size_t)k * n * 2;
  DevBuf<double> idiag, part, scal, r, z, p, q;
  idiag.alloc(n);
  part.alloc((size_t)3 * k * nblk);
  scal.alloc((size_t)5 * CG_MAXK);
  r.alloc(vec), z.alloc(vec), p.alloc(vec), q.alloc(vec);
  cplx *rv = (cplx*)r.p, *zv = (cplx*)z.p, *pv = (cplx*)p.p, *qv = (cplx*)q.p;
  double* part_rr = part.p;
  double* part_rz = part.p + (size_t)k * nblk;
  double* part_pq = part.p + (size_t)2 * k * nblk;
  double* d_rr = scal.p;                 // [k]
  double* d_rz[2] = {scal.p + CG_MAXK, scal.p + 2 * CG_MAXK};
  double* d_pq = scal.p + 3 * CG_MAXK;
  const int g256 = cdiv(n, 256);
  k_cg_diag<<<g256, 256, 0, s>>>(n, pat.indptr.p, pat.indices.p, h.Mass.v.p, idiag.p);
  FW_CHECK_CUDA(cudaMemsetAsync(X, 0, vec * sizeof(double), s));
  FW_CHECK_CUDA(cudaMemcpyAsync(rv, B, vec * sizeof(double), cudaMemcpyDeviceToDevice, s));
  k_cg_init<<<nblk, 256, 0, s>>>(n, k, rv, zv, idiag.p, part_rr, part_rz, nblk);
  k_cg_reduce<<<k, 256, 0, s>>>(part_rr, nblk, d_rr);
  k_cg_reduce<<<k, 256, 0, s>>>(part_rz, nblk, d_rz[0]);
  FW_CHECK_CUDA(cudaMemcpyAsync(pv, zv, vec * sizeof(double), cudaMemcpyDeviceToDevice, s));
  double bnorm2[CG_MAXK], rr[CG_MAXK];
  FW_CHECK_CUDA(cudaMemcpyAsync(bnorm2, d_rr, sizeof(double) * k, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  int64_t launches = 4;
  int cur = 0;
  const int check_every = 4;
  for (int it = 0; it < 600; ++it) {
    k_spmv_multi<<<cdiv((int64_t)n * 8, 256), 256, 0, s>>>(n, k, pat.indptr.p, pat.indices.p, h.Mass.v.p, pv, qv);
    k_cg_dot<<<nblk, 256, 0, s>>>(n, k, pv, qv, part_pq, nblk);
    k_cg_reduce<<<k, 256, 0, s>>>(part_pq, nblk, d_pq);
    k_cg_step<<<nblk, 256, 0, s>>>(n, k, X, rv, pv, qv, zv, idiag.p, d_rz[cur], d_pq, part_rr, part_rz, nblk);
    k_cg_reduce<<<k, 256, 0, s>>>(part_rr, nblk, d_rr);
    k_cg_reduce<<<k, 256, 0, s>>>(part_rz, nblk, d_rz[cur ^ 1]);
    k_cg_direction<<<g256, 256, 0, s>>>(n, k, pv, zv, d_rz[cur ^ 1], d_rz[cur]);
    cur ^= 1;
    launches += 7;
    if ((it + 1) % check_every == 0) {
      FW_CHECK_CUDA(cudaMemcpyAsync(rr, d_rr, sizeof(double) * k, cudaMemcpyDeviceToHost, s));
      FW_CHECK_CUDA(cudaStreamSynchronize(s));
      bool done = true;
      for (int i = 0; i < k; ++i)
        if (!(rr[i] <= 1e-26 * bnorm2[i])) done = false;  // ||r|| <= 1e-13 ||b|| (the reference's H carries complex64 rounding, 1e-7)
      if (done) break;
    }
  }
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += launches;
  if (getenv("FW_TRACE")) fprintf(stderr, "[fw-trace]   mass PCG: %d rhs, %lld kernel launches (7 per iteration)\n", k, (long long)launches);
}

}  // namespace fw
