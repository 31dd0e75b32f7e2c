// Default start vector of the shift-invert Arnoldi iteration inside fw_compute_modes.
//
// scipy's eigs (the reference path, arpack.py:361-364) starts from uniform(-1, 1) noise: the wanted guided modes have
// components of size ~N^-1/2 in it and the first Krylov steps are spent amplifying them.  Here the default is a smooth
// field that lives where the modes live: the Whitney (lowest-order Nedelec) moments of a low-order polynomial vector
// field weighted by the local index contrast (eps - eps_min) / (eps_max - eps_min), with even and odd terms in x and y
// so that every symmetry class of the guided modes is present.  Measured on BASELINE config 2 (scripts/v0_probe.py):
// 55 -> 51 operator applications for 4 modes, 33 -> 30 for 1 mode; the eigenpairs are the same to round-off.
// Everything is integer-atomic or idempotent: the vector is bit-reproducible from run to run.
// FW_OPT_START_VECTOR = 0 restores the uniform(-1, 1) stream; an explicit v0 always wins.
#include "handle.h"

namespace fw {

namespace {

// monotone map double -> uint64 (atomicMin / atomicMax on the image order the doubles)
__device__ inline unsigned long long enc_(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double dec_(unsigned long long e) {
  const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  double v;
  memcpy(&v, &b, sizeof(v));
  return v;
}

template <int EPSKIND, bool CPLX>
__device__ inline double eps_re_(const double* __restrict__ eps, int e) {
  constexpr int ST = CPLX ? 2 : 1;
  if (EPSKIND == FW_EPS_P0) return eps[(size_t)e * ST];
  return (eps[(size_t)(3 * e) * ST] + eps[(size_t)(3 * e + 1) * ST] + eps[(size_t)(3 * e + 2) * ST]) * (1.0 / 3.0);
}

// range[0] = min, range[1] = max of Re(eps) over the elements (encoded)
template <int EPSKIND, bool CPLX>
__global__ void k_sv_eps_range(int ne, const double* __restrict__ eps, unsigned long long* __restrict__ range) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  // one atomic per warp (every lane takes part in the shuffles: lanes past the end carry neutral values)
  unsigned long long lo = ~0ull, hi = 0ull;
  if (e < ne) lo = hi = enc_(eps_re_<EPSKIND, CPLX>(eps, e));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(range, lo);
    atomicMax(range + 1, hi);
  }
}

// per vertex: largest contrast weight of the elements around it (float bits, non-negative: integer order = float
// order); range[2..5]: bounding box (x min, x max, y min, y max) of the vertices of elements with weight > 1/2
template <int EPSKIND, bool CPLX>
__global__ void k_sv_vertex_weights(int ne, int nv, const double* __restrict__ p, const int* __restrict__ t,
                                    const double* __restrict__ eps, unsigned long long* __restrict__ range,
                                    unsigned int* __restrict__ wv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const double lo = dec_(range[0]), hi = dec_(range[1]);
  const float wgt = (float)((eps_re_<EPSKIND, CPLX>(eps, e) - lo) / (hi - lo));
  for (int k = 0; k < 3; ++k) {
    const int v = t[(size_t)k * ne + e];
    atomicMax(wv + v, __float_as_uint(wgt));
    if (wgt > 0.5f) {
      const unsigned long long x = enc_(p[v]), y = enc_(p[(size_t)nv + v]);
      atomicMin(range + 2, x);
      atomicMax(range + 3, x);
      atomicMin(range + 4, y);
      atomicMax(range + 5, y);
    }
  }
}

// Whitney moment of the field on every edge: dof(first Nedelec function of the edge) = F(midpoint) . (p_b - p_a);
// the two triangles of an interior edge write the same value
__global__ void k_sv_edge_moments(int ne, int nv, const double* __restrict__ p, const int* __restrict__ t,
                                  const int* __restrict__ edofs, int row0, int row_stride,
                                  const unsigned long long* __restrict__ range, const unsigned int* __restrict__ wv,
                                  double* __restrict__ v0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const double x0 = dec_(range[2]), x1 = dec_(range[3]), y0 = dec_(range[4]), y1 = dec_(range[5]);
  const double xc = 0.5 * (x0 + x1), yc = 0.5 * (y0 + y1);
  const double sx = fmax(0.5 * (x1 - x0), 1e-300), sy = fmax(0.5 * (y1 - y0), 1e-300);
  int v[3];
  for (int k = 0; k < 3; ++k) v[k] = t[(size_t)k * ne + e];
  const int ea[3] = {0, 1, 0}, eb[3] = {1, 2, 2};  // local edges (0,1), (1,2), (0,2); t is sorted: a < b
  for (int k = 0; k < 3; ++k) {
    const int a = v[ea[k]], b = v[eb[k]];
    const double ax = p[a], ay = p[(size_t)nv + a], bx = p[b], by = p[(size_t)nv + b];
    const double wf = 0.5 * ((double)__uint_as_float(wv[a]) + (double)__uint_as_float(wv[b]));
    const double xn = (0.5 * (ax + bx) - xc) / sx, yn = (0.5 * (ay + by) - yc) / sy;
    const double fx = wf * (1.0 + 0.9 * xn + 0.8 * (xn * xn - 0.3) + 0.7 * xn * xn * xn) * (1.0 + 0.6 * yn);
    const double fy = wf * (0.95 + 0.85 * xn + 0.75 * (xn * xn - 0.3) + 0.65 * xn * xn * xn) * (1.0 + 0.5 * yn);
    v0[edofs[(size_t)(row0 + k * row_stride) * ne + e]] = fx * (bx - ax) + fy * (by - ay);
  }
}

template <int EPSKIND, bool CPLX>
bool build(Handle& h, const double* eps_dev, DevBuf<double>& v0) {
  Space& sp = h.sp;
  cudaStream_t s = h.stream;
  DevBuf<unsigned long long> range;
  DevBuf<unsigned int> wv;
  range.alloc(6);
  wv.alloc((size_t)sp.nv);
  const unsigned long long init[6] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
  range.upload(init, 6, s);
  wv.zero(s);
  const int grid = cdiv(sp.ne, 256);
  k_sv_eps_range<EPSKIND, CPLX><<<grid, 256, 0, s>>>(sp.ne, eps_dev, range.p);
  unsigned long long r[2];
  range.download(r, 2, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  const double lo = dec_(r[0]), hi = dec_(r[1]);
  if (!(hi > lo) || !std::isfinite(hi - lo)) return false;  // homogeneous cross-section: no preferred region
  k_sv_vertex_weights<EPSKIND, CPLX><<<grid, 256, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, eps_dev, range.p, wv.p);
  v0.alloc((size_t)sp.n);
  v0.zero(s);
  // local layout: order 1: [3 vertex P1, 3 edge N1]; order 2: [3 vertex P2, per edge (N2, N2, P2), 2 interior]
  const int row0 = 3, stride = sp.order == 1 ? 1 : 3;
  k_sv_edge_moments<<<grid, 256, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, row0, stride, range.p, wv.p, v0.p);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 3;
  if (getenv("FW_TRACE")) {
    unsigned long long rr[6];
    range.download(rr, 6, s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    fprintf(stderr, "[fw-trace] smooth start vector: eps in [%g, %g], high-index box x [%g, %g] y [%g, %g]\n", dec_(rr[0]),
            dec_(rr[1]), dec_(rr[2]), dec_(rr[3]), dec_(rr[4]), dec_(rr[5]));
  }
  return true;
}

}  // namespace

// eps_host: the caller's epsilon (FW_EPS_* layout, f64 or c128).  Returns false when no smooth start vector applies.
bool smooth_start_vector(Handle& h, int eps_kind, bool eps_complex, const void* eps_host, DevBuf<double>& v0) {
  Space& sp = h.sp;
  if (!(sp.order == 1 || sp.order == 2) || sp.ne <= 0) return false;
  const size_t neps = (eps_kind == FW_EPS_P0 ? (size_t)sp.ne : (size_t)3 * sp.ne) * (eps_complex ? 2 : 1);
  DevBuf<double> eps_dev;
  eps_dev.upload((const double*)eps_host, neps, h.stream);
  const double* e = eps_dev.p;
  switch (eps_kind) {
    case FW_EPS_P0:
      return eps_complex ? build<FW_EPS_P0, true>(h, e, v0) : build<FW_EPS_P0, false>(h, e, v0);
    case FW_EPS_DG_P1:
      return eps_complex ? build<FW_EPS_DG_P1, true>(h, e, v0) : build<FW_EPS_DG_P1, false>(h, e, v0);
    case FW_EPS_VEC_P0:
      return eps_complex ? build<FW_EPS_VEC_P0, true>(h, e, v0) : build<FW_EPS_VEC_P0, false>(h, e, v0);
  }
  return false;
}

}  // namespace fw
