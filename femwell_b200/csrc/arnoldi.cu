// Shift-invert Arnoldi with thick (Krylov-Schur style) restarts on OP = (K - sigma M)^-1 M,
// K = -A, M = -B.  Replaces scipy.sparse.linalg.eigs(K, M=M, k, sigma) -- ARPACK mode 3 --
// as called through skfem's solver_eigen_scipy (femwell/maxwell/waveguide.py:611-625;
// scipy arpack.py:1184-1442).  Hand-written kernels: masked SpMV (K4), CGS2 multi-dot /
// multi-axpy (K5), basis update V <- V Q and Ritz vector extraction (K6).
#include <algorithm>
#include <complex>

#include "lu.h"
#include "solver.h"

namespace fw {

using cd = std::complex<double>;
constexpr int MAXV = 128;

// ------------------------------------------------------------------------------ operator kernels
template <class V, class S>
__device__ inline S cvt_(V v);
template <>
__device__ inline double cvt_<double, double>(double v) { return v; }
template <>
__device__ inline cplx cvt_<double, cplx>(double v) { return {v, 0.0}; }
template <>
__device__ inline cplx cvt_<cplx, cplx>(cplx v) { return v; }

__device__ inline double shfl_xor_(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ inline cplx shfl_xor_(cplx v, int o) {
  return {__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}

template <class V, class S>
__global__ void __launch_bounds__(256, 4) k_op_apply(int n, const int* __restrict__ indptr,
                                                  const int* __restrict__ indices, const V* __restrict__ A,
                                                  const V* __restrict__ B, S alphaA, S alphaB, bool useA,
                                                  const uint8_t* __restrict__ active, const S* __restrict__ x,
                                                  S* __restrict__ y) {
  constexpr int TPR = 8;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(gt / TPR), l = (int)(gt % TPR);
  S accA = scalar_traits<S>::zero(), accB = scalar_traits<S>::zero();
  const bool live = row < n && active[row];
  if (live) {
    const int a = indptr[row], b = indptr[row + 1];
    for (int k = a + l; k < b; k += TPR) {
      const int c = indices[k];
      if (!active[c]) continue;
      const S xv = x[c];
      if (useA) fma_(accA, cvt_<V, S>(A[k]), xv);
      fma_(accB, cvt_<V, S>(B[k]), xv);
    }
  }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) {
    accA = accA + shfl_xor_(accA, o);
    accB = accB + shfl_xor_(accB, o);
  }
  if (l == 0 && row < n) y[row] = live ? (alphaA * accA + alphaB * accB) : scalar_traits<S>::zero();
}

template <class V, class S>
static void op_apply_impl(Handle& h, const uint8_t* active, S alphaA, S alphaB, bool useA, const S* x, S* y) {
  const Pattern& pat = h.pat;
  k_op_apply<V, S><<<cdiv((int64_t)pat.n * 8, 256), 256, 0, h.stream>>>(
      pat.n, pat.indptr.p, pat.indices.p, (const V*)h.A.v.p, (const V*)h.B.v.p, alphaA, alphaB, useA, active, x, y);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 1;
}

// y = alpha * Bc x on the compact copy of B (rows/columns of inactive dofs masked; active == nullptr: none)
// (explicit occupancy hints and batches of independent loads: see the note on upd_min_blocks in lu_solve.cu)
template <class V, class S, int TPR>
__global__ void __launch_bounds__(256, 4) k_spmv_bc(int n, const int* __restrict__ indptr,
                                                    const int* __restrict__ indices, const V* __restrict__ B, S alpha,
                                                    const uint8_t* __restrict__ active, const S* __restrict__ x,
                                                    S* __restrict__ y) {
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(gt / TPR), l = (int)(gt % TPR);
  S acc = scalar_traits<S>::zero();
  const bool live = row < n && (!active || active[row]);
  if (live) {
    const int a = indptr[row], b = indptr[row + 1];
    int k = a + l;
    for (; k + 3 * TPR < b; k += 4 * TPR) {  // four entries per pass: indices and values first, then the gathers
      int c[4];
      V bv[4];
      S xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = indices[k + u * TPR];
        bv[u] = B[k + u * TPR];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = (active && !active[c[u]]) ? scalar_traits<S>::zero() : x[c[u]];
#pragma unroll
      for (int u = 0; u < 4; ++u) fma_(acc, cvt_<V, S>(bv[u]), xv[u]);
    }
    for (; k < b; k += TPR) {
      const int c = indices[k];
      if (active && !active[c]) continue;
      fma_(acc, cvt_<V, S>(B[k]), x[c]);
    }
  }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) acc = acc + shfl_xor_(acc, o);
  if (l == 0 && row < n) y[row] = live ? alpha * acc : scalar_traits<S>::zero();
}

// ------------------------------------------------------------------------------ CGS2 kernels
// partial[b*nvec + i] = sum over the rows of block b of conj(V[i][r]) * w[r]
template <class T>
__global__ void __launch_bounds__(256, 4) k_multidot(const T* __restrict__ V, int nvec, int64_t ld,
                                                     const T* __restrict__ w, int n, T* __restrict__ partial) {
  __shared__ T part[8][MAXV];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  T wr[4];
  bool ok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t r = base + k * 256 + t;
    ok[k] = r < n;
    wr[k] = ok[k] ? w[r] : scalar_traits<T>::zero();
  }
  // IB basis vectors per pass: 4 * IB independent loads in flight, then the warp reductions
  constexpr int IB = scalar_traits<T>::is_complex ? 2 : 4;
  int i = 0;
  for (; i + IB <= nvec; i += IB) {
    T v[IB][4];
#pragma unroll
    for (int q = 0; q < IB; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[q][k] = ok[k] ? V[(int64_t)(i + q) * ld + base + k * 256 + t] : scalar_traits<T>::zero();
#pragma unroll
    for (int q = 0; q < IB; ++q) {
      T s = scalar_traits<T>::zero();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) fma_(s, conj_(v[q][k]), wr[k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s = s + shfl_xor_(s, o);
      if (lane == 0) part[warp][i + q] = s;
    }
  }
  for (; i < nvec; ++i) {
    const T* vi = V + (int64_t)i * ld;
    T s = scalar_traits<T>::zero();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int64_t r = base + k * 256 + t;
      if (r < n) fma_(s, conj_(vi[r]), wr[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = s + shfl_xor_(s, o);
    if (lane == 0) part[warp][i] = s;
  }
  __syncthreads();
  for (int i2 = t; i2 < nvec; i2 += 256) {
    T s = part[0][i2];
#pragma unroll
    for (int q = 1; q < 8; ++q) s = s + part[q][i2];
    partial[(int64_t)blockIdx.x * nvec + i2] = s;
  }
}

// out[i] = sum_b partial[b*nvec + i]   (one block per i, fixed order => deterministic)
template <class T>
__global__ void __launch_bounds__(256) k_reduce_partials(const T* __restrict__ partial, int nblocks, int nvec,
                                                         T* __restrict__ out) {
  __shared__ T sh[256];
  const int i = blockIdx.x, t = threadIdx.x;
  T s = scalar_traits<T>::zero();
  for (int b = t; b < nblocks; b += 256) s = s + partial[(int64_t)b * nvec + i];
  sh[t] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sh[t] = sh[t] + sh[t + o];
    __syncthreads();
  }
  if (t == 0) out[i] = sh[0];
}

// w[r] -= sum_i V[i][r] * hc[i] ; optionally accumulates |w|^2 per block into nrm_partial
template <class T>
__global__ void __launch_bounds__(256, 4) k_multiaxpy(T* __restrict__ w, const T* __restrict__ V, int nvec, int64_t ld,
                                                      int n, const T* __restrict__ hc,
                                                      double* __restrict__ nrm_partial) {
  __shared__ T sh[MAXV];
  __shared__ double red[8];
  const int t = threadIdx.x;
  for (int i = t; i < nvec; i += 256) sh[i] = hc[i];
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * 1024;
  double nrm = 0.0;
  // the four rows of this thread advance together, IB basis vectors per pass: 4 * IB independent loads in flight
  // (row after row with a run-time loop over the vectors kept ONE load in flight); same summation order per row
  constexpr int IB = scalar_traits<T>::is_complex ? 2 : 4;
  T acc[4];
  bool ok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = base + k * 256 + t;
    ok[k] = r < n;
    acc[k] = ok[k] ? w[r] : scalar_traits<T>::zero();
  }
  int i = 0;
  for (; i + IB <= nvec; i += IB) {
    T v[IB][4];
#pragma unroll
    for (int q = 0; q < IB; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[q][k] = ok[k] ? V[(int64_t)(i + q) * ld + base + k * 256 + t] : scalar_traits<T>::zero();
#pragma unroll
    for (int q = 0; q < IB; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k) fnma_(acc[k], v[q][k], sh[i + q]);
  }
  for (; i < nvec; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) fnma_(acc[k], V[(int64_t)i * ld + base + k * 256 + t], sh[i]);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (ok[k]) {
      w[base + k * 256 + t] = acc[k];
      nrm += abs2_(acc[k]);
    }
  if (nrm_partial) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if ((t & 31) == 0) red[t >> 5] = nrm;
    __syncthreads();
    if (t == 0) {
      double s2 = 0;
      for (int q = 0; q < 8; ++q) s2 += red[q];
      nrm_partial[blockIdx.x] = s2;
    }
  }
}

__global__ void __launch_bounds__(256) k_reduce_double(const double* __restrict__ partial, int nblocks,
                                                       double* __restrict__ out) {
  __shared__ double sh[256];
  const int t = threadIdx.x;
  double s = 0;
  for (int b = t; b < nblocks; b += 256) s += partial[b];
  sh[t] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  if (t == 0) out[0] = sh[0];
}

template <class T>
__global__ void k_scale_copy(const T* __restrict__ src, T* __restrict__ dst, int n, double scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * scale;
}
template <class T>
__global__ void k_axpy1(T* __restrict__ y, const T* __restrict__ x, int n, double a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = y[i] + x[i] * a;
}
template <class T>
__global__ void k_sub(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - b[i];
}
template <class T>
__global__ void k_sqnorm_partial(const T* __restrict__ v, int n, double* __restrict__ partial) {
  __shared__ double red[8];
  const int t = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t r = base + k * 256 + t;
    if (r < n) s += abs2_(v[r]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((t & 31) == 0) red[t >> 5] = s;
  __syncthreads();
  if (t == 0) {
    double q = 0;
    for (int i = 0; i < 8; ++i) q += red[i];
    partial[blockIdx.x] = q;
  }
}

// V[0..p) <- V[0..m) Q (Q: m x p column-major) ; V[p] <- V[m].  One thread per row.
template <class T>
__global__ void __launch_bounds__(128, 8) k_basis_update(T* __restrict__ V, int64_t ld, int n, int m, int p,
                                                      const T* __restrict__ Q) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  T v[MAXV];
  for (int l = 0; l < m; ++l) v[l] = V[(int64_t)l * ld + r];
  const T last = V[(int64_t)m * ld + r];
  for (int i = 0; i < p; ++i) {
    T acc = scalar_traits<T>::zero();
    for (int l = 0; l < m; ++l) fma_(acc, v[l], Q[(size_t)i * m + l]);
    V[(int64_t)i * ld + r] = acc;
  }
  V[(int64_t)p * ld + r] = last;
}

__device__ inline cplx mulc_(double a, cplx b) { return {a * b.re, a * b.im}; }
__device__ inline cplx mulc_(cplx a, cplx b) { return a * b; }

// X[i][r] = sum_l V[l][r] * Y[l + i*m]   (Y complex, X complex)
template <class T>
__global__ void __launch_bounds__(128, 8) k_ritz_vectors(const T* __restrict__ V, int64_t ld, int n, int m, int k,
                                                      const cplx* __restrict__ Y, cplx* __restrict__ X) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int i = 0; i < k; ++i) {
    cplx acc{0.0, 0.0};
    for (int l = 0; l < m; ++l) acc += mulc_(V[(int64_t)l * ld + r], Y[(size_t)i * m + l]);
    X[(int64_t)i * n + r] = acc;
  }
}

__global__ void k_mask_real_to(const double* __restrict__ src, const uint8_t* __restrict__ active, int n,
                               double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = active[i] ? src[i] : 0.0;
}
__global__ void k_mask_real_to(const double* __restrict__ src, const uint8_t* __restrict__ active, int n,
                               cplx* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = cplx{active[i] ? src[i] : 0.0, 0.0};
}

// default start vector: uniform(-1, 1) from a splitmix64 stream indexed by the dof (the same numbers on every run
// and every device; ARPACK's default is a uniform(-1, 1) vector as well, scipy arpack.py:361-364)
__device__ __host__ inline double start_value(int i) {
  uint64_t x = 0x9E3779B97F4A7C15ull * (uint64_t)(i + 2);
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
__global__ void k_start_vector(const uint8_t* __restrict__ active, int n, double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = active[i] ? start_value(i) : 0.0;
}
__global__ void k_start_vector(const uint8_t* __restrict__ active, int n, cplx* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = cplx{active[i] ? start_value(i) : 0.0, 0.0};
}

// ------------------------------------------------------------------------------ host helpers
namespace {

template <class T>
struct HostScalar;
template <>
struct HostScalar<double> {
  static cd to_cd(double v) { return cd(v, 0.0); }
  static double from_cd(cd v) { return v.real(); }
};
template <>
struct HostScalar<cplx> {
  static cd to_cd(cplx v) { return cd(v.re, v.im); }
  static cplx from_cd(cd v) { return cplx{v.real(), v.imag()}; }
};

struct Work {
  DevBuf<double> V, w, t1, t2, partial, hdev, nrmp, nrm, Qd, Yd;
};

template <class T>
double dev_norm(Handle& h, const T* v, int n, Work& W) {
  const int nb = cdiv(n, 1024);
  W.nrmp.alloc((size_t)nb);
  W.nrm.alloc(1);
  k_sqnorm_partial<T><<<nb, 256, 0, h.stream>>>(v, n, W.nrmp.p);
  k_reduce_double<<<1, 256, 0, h.stream>>>(W.nrmp.p, nb, W.nrm.p);
  double r = 0;
  W.nrm.download(&r, 1, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  h.tim.kernel_launches += 2;
  return std::sqrt(r);
}

}  // namespace

// ------------------------------------------------------------------------------ the solver
template <class T, class V>
static void eigs_typed(Handle& h, const fw_eigs_params& prm, int refine, const std::vector<uint8_t>& active_host,
                       cd sigma, cd* lams, DevBuf<double>& X_dev, fw_eigs_stats* st) {
  using HS = HostScalar<T>;
  constexpr size_t wsc = sizeof(T) / sizeof(double);
  cudaStream_t s = h.stream;
  LU& lu = *h.lu;
  const int n = h.sp.n;
  const int k = prm.k;
  int m = prm.ncv > 0 ? prm.ncv : std::max(2 * k + 1, 20);
  const int n_active = lu.sym.n_active;
  m = std::min(m, std::min(n_active - 1, MAXV - 2));
  FW_REQUIRE(k >= 1 && k < m, "need 1 <= k < ncv <= n_active - 1");
  const int maxrestart = prm.maxiter > 0 ? prm.maxiter : 300;
  const double tol = prm.tol > 0 ? prm.tol : 1e-14;

  DevBuf<uint8_t> act;
  act.upload(active_host.data(), active_host.size(), s);
  const bool any_inactive = std::find(active_host.begin(), active_host.end(), (uint8_t)0) != active_host.end();
  Work W;
  W.V.alloc((size_t)(m + 1) * n * wsc);
  W.w.alloc((size_t)n * wsc);
  W.t1.alloc((size_t)n * wsc);
  W.t2.alloc((size_t)n * wsc);
  const int nblk = cdiv(n, 1024);
  W.partial.alloc((size_t)nblk * MAXV * wsc);
  W.hdev.alloc((size_t)2 * MAXV * wsc);
  W.nrmp.alloc((size_t)nblk);
  W.nrm.alloc(1);
  T* Vb = (T*)W.V.p;
  T* w = (T*)W.w.p;
  T* t1 = (T*)W.t1.p;
  T* t2 = (T*)W.t2.p;
  const int64_t ld = n;
  int n_op = 0;
  double solve_ms = 0;
  struct EventList {
    std::vector<cudaEvent_t> ev;  // (start, stop) pairs
    ~EventList() {
      for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
  } solve_events;
  const T sigT = HS::from_cd(sigma);
  const T minus1 = HS::from_cd(cd(-1.0, 0.0));
  const T zeroT = HS::from_cd(cd(0.0, 0.0));

  // lanes per row of the M x product on the compact B (rows hold ~12 entries of the Nedelec block): 4 lanes measured
  // 4 ms faster per mode solve than 8 (FW_SPMV_BC_TPR=8 selects the old mapping)
  static const bool spmv_tpr4 = !(getenv("FW_SPMV_BC_TPR") && atoi(getenv("FW_SPMV_BC_TPR")) == 8);
  // y = OP x = (K - sigma M)^-1 M x, with `refine` steps of iterative refinement on the solve
  auto apply_op = [&](const T* x, T* y) {
    // t1 = M x = -(B x)
    if (h.Bc.valid && !h.generic.valid) {
      if (spmv_tpr4)
        k_spmv_bc<V, T, 4><<<cdiv((int64_t)n * 4, 256), 256, 0, s>>>(n, h.Bc.indptr.p, h.Bc.indices.p, (const V*)h.Bc.vals.p,
                                                               minus1, any_inactive ? act.p : nullptr, x, t1);
      else
        k_spmv_bc<V, T, 8><<<cdiv((int64_t)n * 8, 256), 256, 0, s>>>(n, h.Bc.indptr.p, h.Bc.indices.p, (const V*)h.Bc.vals.p,
                                                               minus1, any_inactive ? act.p : nullptr, x, t1);
      h.tim.kernel_launches += 1;
    } else {
      op_apply_impl<V, T>(h, act.p, zeroT, minus1, false, x, t1);
    }
    // event pair per application, read back once at the end: a synchronising timer here would stall the host (and leave
    // the GPU idle until the CGS2 launches arrive) after every one of the ~50 solves
    cudaEvent_t ea = nullptr, eb = nullptr;
    FW_CHECK_CUDA(cudaEventCreate(&ea));
    FW_CHECK_CUDA(cudaEventCreate(&eb));
    solve_events.ev.push_back(ea);
    solve_events.ev.push_back(eb);
    FW_CHECK_CUDA(cudaEventRecord(ea, s));
    lu_solve_dev(h, t1, y, 1);
    for (int it = 0; it < refine; ++it) {
      // r = t1 - (K - sigma M) y = t1 - (-A + sigma B) y
      op_apply_impl<V, T>(h, act.p, minus1, sigT, true, y, t2);
      k_sub<T><<<cdiv(n, 256), 256, 0, s>>>(t1, t2, t2, n);
      lu_solve_dev(h, t2, t2, 1);
      k_axpy1<T><<<cdiv(n, 256), 256, 0, s>>>(y, t2, n, 1.0);
      h.tim.kernel_launches += 2;
    }
    FW_CHECK_CUDA(cudaEventRecord(eb, s));
    if (getenv("FW_TRACE")) FW_CHECK_CUDA(cudaEventSynchronize(eb));  // the host wall-clock split of the trace needs it
    ++n_op;
  };

  // ---- start vector (ARPACK forces it into range(OP) with one application)
  {
    DevBuf<double> v0d;
    if (prm.v0) {
      v0d.upload(prm.v0, (size_t)n, s);
      k_mask_real_to<<<cdiv(n, 256), 256, 0, s>>>(v0d.p, act.p, n, w);
    } else if (h.v0_dev_valid && h.v0_dev.n >= (size_t)n) {
      k_mask_real_to<<<cdiv(n, 256), 256, 0, s>>>(h.v0_dev.p, act.p, n, w);  // smooth default (start_vector.cu)
    } else {
      k_start_vector<<<cdiv(n, 256), 256, 0, s>>>(act.p, n, w);
    }
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    apply_op(w, Vb);
    double nv = dev_norm<T>(h, Vb, n, W);
    FW_REQUIRE(nv > 0 && std::isfinite(nv), "start vector has no component in range(OP)");
    k_scale_copy<T><<<cdiv(n, 256), 256, 0, s>>>(Vb, Vb, n, 1.0 / nv);
  }

  std::vector<cd> H((size_t)(m + 1) * m, cd(0));  // column-major, (m+1) x m
  auto Hat = [&](int i, int j) -> cd& { return H[(size_t)j * (m + 1) + i]; };
  std::vector<cd> theta(m), Y((size_t)m * m), Hm((size_t)m * m);
  std::vector<int> order(m);
  std::vector<double> resid(m);
  std::vector<T> hbuf(2 * MAXV);
  int j0 = 0, restarts = 0, iterations = 0, nconv = 0, m_eff = m;
  bool done = false;

  double tw_op = 0, tw_cgs = 0, tw_ritz = 0, tw_restart = 0;  // host wall-clock (FW_TRACE)
  while (!done) {
    for (int j = j0; j < m_eff; ++j) {
      double tw0 = Trace::now();
      apply_op(Vb + (int64_t)j * ld, w);
      double tw1 = Trace::now();
      tw_op += tw1 - tw0;
      const int nvec = j + 1;
      T* h1 = (T*)W.hdev.p;
      T* h2 = h1 + MAXV;
      // CGS2: two passes of (h = V^H w ; w -= V h)
      k_multidot<T><<<nblk, 256, 0, s>>>(Vb, nvec, ld, w, n, (T*)W.partial.p);
      k_reduce_partials<T><<<nvec, 256, 0, s>>>((const T*)W.partial.p, nblk, nvec, h1);
      k_multiaxpy<T><<<nblk, 256, 0, s>>>(w, Vb, nvec, ld, n, h1, nullptr);
      k_multidot<T><<<nblk, 256, 0, s>>>(Vb, nvec, ld, w, n, (T*)W.partial.p);
      k_reduce_partials<T><<<nvec, 256, 0, s>>>((const T*)W.partial.p, nblk, nvec, h2);
      k_multiaxpy<T><<<nblk, 256, 0, s>>>(w, Vb, nvec, ld, n, h2, W.nrmp.p);
      k_reduce_double<<<1, 256, 0, s>>>(W.nrmp.p, nblk, W.nrm.p);
      h.tim.kernel_launches += 7;
      double nrm2 = 0;
      FW_CHECK_CUDA(cudaMemcpyAsync(hbuf.data(), W.hdev.p, (size_t)2 * MAXV * sizeof(T), cudaMemcpyDeviceToHost, s));
      W.nrm.download(&nrm2, 1, s);
      FW_CHECK_CUDA(cudaStreamSynchronize(s));
      double hn = 0;
      for (int i = 0; i < nvec; ++i) {
        cd hv = HS::to_cd(hbuf[i]) + HS::to_cd(hbuf[MAXV + i]);
        Hat(i, j) += hv;  // += : after a restart column j0.. are fresh zeros
        hn += std::norm(hv);
      }
      const double beta = std::sqrt(nrm2);
      FW_REQUIRE(std::isfinite(beta), "Arnoldi produced a non-finite vector (factorisation failed?)");
      Hat(j + 1, j) = beta;
      ++iterations;
      if (beta <= 1e-14 * std::sqrt(hn) || beta == 0.0) {
        // invariant subspace: the Ritz pairs of the leading (j+1) block are exact
        m_eff = j + 1;
        Hat(j + 1, j) = 0.0;
        break;
      }
      k_scale_copy<T><<<cdiv(n, 256), 256, 0, s>>>(w, Vb + (int64_t)(j + 1) * ld, n, 1.0 / beta);
      h.tim.kernel_launches += 1;
      tw_cgs += Trace::now() - tw1;
      // ---- early exit: test the Ritz estimates of the current (j+1)-dimensional decomposition every step
      // (ARPACK only looks at a full basis; for k = 1..2 -- sweeps -- that wastes up to half of the OP calls).
      // The last row of the decomposition is beta e_{j+1}^T here, so the estimate is beta |y_i[j]|.
      const int mc = j + 1;
      if (mc < m_eff && mc >= k + 2 && mc - j0 >= 2) {
        for (int c = 0; c < mc; ++c)
          for (int i = 0; i < mc; ++i) Hm[(size_t)c * mc + i] = Hat(i, c);
        if (dense_eig(mc, Hm.data(), theta.data(), Y.data())) {
          for (int i = 0; i < mc; ++i) order[i] = i;
          std::sort(order.begin(), order.begin() + mc,
                    [&](int a, int b) { return std::abs(theta[a]) > std::abs(theta[b]); });
          int okc = 0;
          for (int i = 0; i < k; ++i) {
            const int o = order[i];
            const double est = beta * std::abs(Y[(size_t)o * mc + (mc - 1)]);
            if (est <= tol * std::max(3.666852862501036e-11, std::abs(theta[o]))) ++okc;
          }
          if (okc >= k) {
            m_eff = mc;  // the Ritz section below re-evaluates this block and finishes
            break;
          }
        }
      }
    }
    // ---- Ritz pairs of the m_eff x m_eff projected matrix
    const double tw2 = Trace::now();
    const int me = m_eff;
    for (int j = 0; j < me; ++j)
      for (int i = 0; i < me; ++i) Hm[(size_t)j * me + i] = Hat(i, j);
    if (!dense_eig(me, Hm.data(), theta.data(), Y.data()))
      throw Error(FW_ERR_NOCONV, "QR iteration on the projected matrix failed");
    for (int i = 0; i < me; ++i) {
      order[i] = i;
      cd r = 0;
      for (int l = 0; l < me; ++l) r += Hat(me, l) * Y[(size_t)i * me + l];
      resid[i] = std::abs(r);
    }
    std::sort(order.begin(), order.begin() + me, [&](int a, int b) {
      double aa = std::abs(theta[a]), bb = std::abs(theta[b]);
      if (aa != bb) return aa > bb;
      return theta[a].imag() > theta[b].imag();
    });
    nconv = 0;
    const double eps23 = 3.666852862501036e-11;  // eps^(2/3), ARPACK's floor for the Ritz value scale
    const int kk = std::min(k, me);
    for (int i = 0; i < kk; ++i)
      if (resid[order[i]] <= tol * std::max(eps23, std::abs(theta[order[i]]))) ++nconv;
    if (getenv("FW_TRACE")) {
      fprintf(stderr, "[fw-trace] arnoldi cycle %d (ops %d): rel. residual estimates", restarts, n_op);
      for (int i = 0; i < kk; ++i) fprintf(stderr, " %.2e", resid[order[i]] / std::abs(theta[order[i]]));
      fprintf(stderr, "\n");
    }
    if (nconv >= kk || me < m || restarts >= maxrestart) {
      done = true;
      break;
    }
    // ---- thick restart: keep p Ritz directions, V <- V Q, H <- Q^H H Q, last row b^T Q
    const double tw3 = Trace::now();
    tw_ritz += tw3 - tw2;
    ++restarts;
    int p = std::min(me - 1, k + std::max(1, (me - k) / 2));
    std::vector<std::vector<cd>> basis;
    const bool real_arith = !scalar_traits<T>::is_complex;
    std::vector<char> used(me, 0);
    for (int ii = 0; ii < me && (int)basis.size() < p; ++ii) {
      const int i = order[ii];
      if (used[i]) continue;
      used[i] = 1;
      std::vector<cd> y(Y.begin() + (size_t)i * me, Y.begin() + (size_t)(i + 1) * me);
      if (!real_arith) {
        basis.push_back(y);
        continue;
      }
      const bool is_real = std::abs(theta[i].imag()) <= 1e-12 * std::abs(theta[i]);
      if (is_real) {
        int big = 0;
        for (int l = 1; l < me; ++l)
          if (std::abs(y[l]) > std::abs(y[big])) big = l;
        cd ph = std::conj(y[big]) / std::abs(y[big]);
        for (auto& v : y) v = cd((v * ph).real(), 0.0);
        basis.push_back(y);
      } else {
        // complex pair: keep Re and Im, and mark the conjugate partner as used
        int partner = -1;
        double best = 1e300;
        for (int l = 0; l < me; ++l)
          if (!used[l]) {
            double dd = std::abs(theta[l] - std::conj(theta[i]));
            if (dd < best) best = dd, partner = l;
          }
        if (partner >= 0 && best <= 1e-8 * std::abs(theta[i])) used[partner] = 1;
        std::vector<cd> yr(me), yi(me);
        for (int l = 0; l < me; ++l) yr[l] = cd(y[l].real(), 0.0), yi[l] = cd(y[l].imag(), 0.0);
        basis.push_back(yr);
        basis.push_back(yi);
      }
    }
    // orthonormalise (MGS, two sweeps), dropping dependent directions
    std::vector<std::vector<cd>> Q;
    for (auto& b : basis) {
      std::vector<cd> v = b;
      for (int sweep = 0; sweep < 2; ++sweep)
        for (auto& q : Q) {
          cd d = 0;
          for (int l = 0; l < me; ++l) d += std::conj(q[l]) * v[l];
          for (int l = 0; l < me; ++l) v[l] -= d * q[l];
        }
      double nv = 0;
      for (auto& x : v) nv += std::norm(x);
      nv = std::sqrt(nv);
      if (nv < 1e-10) continue;
      for (auto& x : v) x /= nv;
      Q.push_back(v);
    }
    p = (int)Q.size();
    if (p >= me) p = me - 1, Q.resize(p);
    // new projected matrix
    std::vector<cd> HQ((size_t)me * p, cd(0));
    for (int c = 0; c < p; ++c)
      for (int l = 0; l < me; ++l) {
        const cd q = Q[c][l];
        if (q == cd(0)) continue;
        for (int i = 0; i < me; ++i) HQ[(size_t)c * me + i] += Hat(i, l) * q;
      }
    std::vector<cd> Hn((size_t)(p + 1) * p, cd(0));
    for (int c = 0; c < p; ++c) {
      for (int r = 0; r < p; ++r) {
        cd sacc = 0;
        for (int l = 0; l < me; ++l) sacc += std::conj(Q[r][l]) * HQ[(size_t)c * me + l];
        Hn[(size_t)c * (p + 1) + r] = sacc;
      }
      cd bacc = 0;
      for (int l = 0; l < me; ++l) bacc += Hat(me, l) * Q[c][l];
      Hn[(size_t)c * (p + 1) + p] = bacc;
    }
    std::fill(H.begin(), H.end(), cd(0));
    for (int c = 0; c < p; ++c)
      for (int r = 0; r <= p; ++r) Hat(r, c) = Hn[(size_t)c * (p + 1) + r];
    // device basis update
    std::vector<T> Qh((size_t)me * p);
    for (int c = 0; c < p; ++c)
      for (int l = 0; l < me; ++l) Qh[(size_t)c * me + l] = HS::from_cd(Q[c][l]);
    W.Qd.upload((const double*)Qh.data(), Qh.size() * wsc, s);
    k_basis_update<T><<<cdiv(n, 128), 128, 0, s>>>(Vb, ld, n, me, p, (const T*)W.Qd.p);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    h.tim.kernel_launches += 1;
    j0 = p;
    tw_restart += Trace::now() - tw3;
  }
  if (getenv("FW_TRACE"))
    fprintf(stderr, "[fw-trace] arnoldi host wall: op %.1f ms  cgs2 %.1f ms  ritz %.1f ms  restart %.1f ms\n", tw_op,
            tw_cgs, tw_ritz, tw_restart);

  // ---- extract the k wanted Ritz pairs.  Without convergence only the converged ones among them are returned
  // (first st->nconv entries of lams / rows of X): scipy's ArpackNoConvergence carries exactly those
  // (arpack.py:303-329); the caller turns st->nconv < k into FW_ERR_NOCONV after the stats are complete.
  const int me = m_eff;
  const int kk = std::min(k, me);
  const bool noconv = nconv < kk || kk < k;
  std::vector<int> sel;
  for (int i = 0; i < kk; ++i) {
    const int o = order[i];
    if (!noconv || resid[o] <= tol * std::max(3.666852862501036e-11, std::abs(theta[o]))) sel.push_back(o);
  }
  const int kout = (int)sel.size();
  for (int i = 0; i < k; ++i) lams[i] = cd(0.0, 0.0);
  X_dev.alloc((size_t)k * n * 2);
  if (kout < k) X_dev.zero(s, (size_t)k * n * 2);
  std::vector<cplx> Yh((size_t)me * std::max(1, kout));
  double maxres = 0;
  for (int i = 0; i < kout; ++i) {
    const int o = sel[i];
    // (ARPACK's real driver returns exactly real values for real Ritz values, scipy arpack.py:898-915; the complex QR
    // used on the projected matrix here leaves an imaginary part of a few ulp)
    const bool real_ritz = !scalar_traits<T>::is_complex && std::abs(theta[o].imag()) <= 1e-12 * std::abs(theta[o]);
    lams[i] = sigma + 1.0 / (real_ritz ? cd(theta[o].real(), 0.0) : theta[o]);
    maxres = std::max(maxres, resid[o] / std::abs(theta[o]));
    // real arithmetic and a real Ritz value: the eigenvector of the real projected matrix is real up to a phase.
    // Remove the phase, so that the returned eigenvector is exactly real (the H-field recovery then solves real
    // mass systems, mass_cg.cu)
    cd ph(1.0, 0.0);
    const bool make_real = real_ritz;
    if (make_real) {
      int big = 0;
      for (int l = 1; l < me; ++l)
        if (std::abs(Y[(size_t)o * me + l]) > std::abs(Y[(size_t)o * me + big])) big = l;
      const cd yb = Y[(size_t)o * me + big];
      if (std::abs(yb) > 0) ph = std::conj(yb) / std::abs(yb);
    }
    for (int l = 0; l < me; ++l) {
      const cd y = Y[(size_t)o * me + l] * ph;
      Yh[(size_t)i * me + l] = cplx{y.real(), make_real ? 0.0 : y.imag()};
    }
  }
  if (kout > 0) {
    W.Yd.upload((const double*)Yh.data(), (size_t)me * kout * 2, s);
    k_ritz_vectors<T><<<cdiv(n, 128), 128, 0, s>>>(Vb, ld, n, me, kout, (const cplx*)W.Yd.p, (cplx*)X_dev.p);
    h.tim.kernel_launches += 1;
  }
  for (int i = 0; i < kout; ++i) {
    cplx* xi = (cplx*)X_dev.p + (int64_t)i * n;
    double nv = dev_norm<cplx>(h, xi, n, W);
    if (nv > 0) k_scale_copy<cplx><<<cdiv(n, 256), 256, 0, s>>>(xi, xi, n, 1.0 / nv);
  }
  // ---- true residuals ||K x - lam M x|| / ||K x|| of the returned pairs (K = -A, M = -B)
  {
    DevBuf<double> rbuf;
    rbuf.alloc((size_t)2 * n);
    cplx* rv = (cplx*)rbuf.p;
    maxres = 0;
    for (int i = 0; i < kout; ++i) {
      const cplx* xi = (const cplx*)X_dev.p + (int64_t)i * n;
      op_apply_impl<V, cplx>(h, act.p, cplx{-1.0, 0.0}, cplx{0.0, 0.0}, true, xi, rv);
      const double kx = dev_norm<cplx>(h, rv, n, W);
      op_apply_impl<V, cplx>(h, act.p, cplx{-1.0, 0.0}, cplx{lams[i].real(), lams[i].imag()}, true, xi, rv);
      const double rr = dev_norm<cplx>(h, rv, n, W);
      const double rel = kx > 0 ? rr / kx : rr;
      if (!(rel <= maxres)) maxres = rel;  // a NaN sticks
    }
  }
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  for (size_t i = 0; i + 1 < solve_events.ev.size(); i += 2) {
    float ms = 0;
    FW_CHECK_CUDA(cudaEventElapsedTime(&ms, solve_events.ev[i], solve_events.ev[i + 1]));
    solve_ms += ms;
  }
  if (st) {
    st->nconv = noconv ? kout : nconv;
    st->iterations = iterations;
    st->restarts = restarts;
    st->n_op = n_op;
    st->t_solve_ms_total = solve_ms;
    st->max_residual = maxres;
  }
}

bool eigs_shift_invert(Handle& h, const fw_eigs_params& prm, cd* lams, DevBuf<double>& X_dev, fw_eigs_stats* st,
                       std::string* why) {
  FW_REQUIRE((h.sp.valid || h.generic.valid) && h.has_symbolic && h.A.valid && h.B.valid, "assemble A and B first");
  FW_REQUIRE(prm.k >= 1, "k must be >= 1");
  const int n = h.sp.n;
  std::vector<uint8_t> active((size_t)n, 1);
  for (int i = 0; i < prm.n_dirichlet; ++i) {
    int d = prm.dirichlet[i];
    FW_REQUIRE(d >= 0 && d < n, "dirichlet dof out of range");
    active[d] = 0;
  }
  cudaStream_t s = h.stream;
  if (st) memset(st, 0, sizeof(*st));
  {
    Timer t(s, "fw:LU analysis");
    lu_analyse(h, active, prm.leaf_elements);
    double ms = t.stop();
    h.tim.lu_symbolic_ms = ms;
    if (st) st->t_symbolic_ms = ms;
  }
  const cd sigma(prm.sigma_re, prm.sigma_im);
  const bool cplx_arith = h.A.is_complex || prm.sigma_im != 0.0;
  {
    Timer t(s, "fw:LU factor");
    // K - sigma M = -A + sigma B
    lu_factor(h, -1.0, 0.0, prm.sigma_re, prm.sigma_im, cplx_arith);
    double ms = t.stop();
    h.tim.lu_factor_ms = ms;
    if (st) st->t_factor_ms = ms;
  }
  if (st) {
    st->pivots_perturbed = h.lu->perturbed;
    st->lu_levels = h.lu->sym.n_levels;
    st->lu_fronts = h.lu->sym.n_nodes;
    st->lu_factor_entries = h.lu->sym.factor_entries;
  }
  h.tim.lu_solve_entries = h.lu->sym.solve_entries;
  h.tim.lu_factor_flops = h.lu->sym.factor_flops * (cplx_arith ? 4.0 : 1.0);  // one complex FMA = 4 real ones
  // refine == 0: no iterative refinement unless the returned pairs miss 1e-10 (then retried with one, then two
  // refinement steps per solve); refine > 0: that many steps; refine < 0: never.  Pairs that still miss
  // RESID_FAIL after the last attempt are an error, not a result.
  constexpr double RESID_RETRY = 1e-10, RESID_FAIL = 1e-8;
  fw_eigs_stats tmp{}, acc{};
  int refine = prm.refine > 0 ? prm.refine : 0;
  bool converged = true;
  {
    Timer t(s, "fw:Arnoldi (shift-invert)");
    int refine_retries = 0;
    for (int attempt = 0; attempt < 4 && refine_retries < 3; ++attempt) {
      tmp = fw_eigs_stats{};
      if (!cplx_arith)
        eigs_typed<double, double>(h, prm, refine, active, sigma, lams, X_dev, &tmp);
      else if (!h.A.is_complex)
        eigs_typed<cplx, double>(h, prm, refine, active, sigma, lams, X_dev, &tmp);
      else
        eigs_typed<cplx, cplx>(h, prm, refine, active, sigma, lams, X_dev, &tmp);
      acc.iterations += tmp.iterations;
      acc.restarts += tmp.restarts;
      acc.n_op += tmp.n_op;
      acc.t_solve_ms_total += tmp.t_solve_ms_total;
      if (tmp.nconv < prm.k) {
        converged = false;
        if (why)
          *why = "ARPACK-style error: no convergence (" + std::to_string(tmp.nconv) + "/" + std::to_string(prm.k) +
                 " eigenvalues converged after " + std::to_string(tmp.restarts) + " restarts)";
        break;
      }
      if (prm.refine != 0 || tmp.max_residual <= RESID_RETRY) break;
      if (attempt == 0 && h.lu && !h.lu->no_tinv && !cplx_arith) {
        // first retry: same solver without the inverse pivot blocks (lu_tinv.cu) -- substitution chains are backward
        // stable where an explicit inverse of an ill-conditioned pivot block is not
        h.lu->no_tinv = true;
        h.lu->drop_graphs();
        continue;
      }
      refine = ++refine_retries;
    }
    double ms = t.stop();
    h.tim.arnoldi_ms = ms;
    if (st) st->t_arnoldi_ms = ms;
  }
  if (st) {
    st->nconv = tmp.nconv;
    st->iterations = acc.iterations;
    st->restarts = acc.restarts;
    st->n_op = acc.n_op;
    st->t_solve_ms_total = acc.t_solve_ms_total;
    st->max_residual = tmp.max_residual;
  }
  if (converged && !(tmp.max_residual <= RESID_FAIL)) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "eigenpairs of the shift-invert Arnoldi miss the residual bound: ||K x - lam M x|| / ||K x|| = %.3e "
             "(limit %.0e, %d pivots perturbed) -- choose another n_guess / sigma",
             tmp.max_residual, RESID_FAIL, h.lu->perturbed);
    throw Error(FW_ERR_NOCONV, buf);
  }
  return converged;
}

}  // namespace fw
