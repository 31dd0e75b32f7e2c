// K7-K9: E_z un-scaling, H-field recovery, power normalisation and the interpolate + pointwise +
// reduce functionals of Mode.  Replaces femwell/maxwell/waveguide.py:627-639 (un-scaling, H,
// normalisation loop), :657-677 (calculate_hfield), :699-736 (calculate_overlap, same-basis
// branch) and the Functionals at :112-258.
//
// H-field: the reference assembles C(beta) and the mass matrix per mode and calls a fresh sparse LU
// (spsolve).  Here the mode-independent real operators Mass, C0, C1 (C = C0 + i beta C1) live on the
// structural CSR pattern, rhs = C0 e + i beta C1 e is two SpMVs and the SPD mass system is solved by
// Jacobi-preconditioned CG (condition number O(1)) with hand-written kernels.
#include <algorithm>
#include <complex>

#include "solver.h"

namespace fw {

using cd = std::complex<double>;

struct TablesP {
  int nb, nq;
  int kind[MAXNB];
  double vx[MAXNB][MAXQ], vy[MAXNB][MAXQ], sc[MAXNB][MAXQ];
  double X[MAXQ], Y[MAXQ], W[MAXQ];
};
__constant__ TablesP c_ptab;

static void upload_ptab(Handle& h) {
  const Space& sp = h.sp;
  TablesP t;
  memset(&t, 0, sizeof(t));
  t.nb = sp.nb, t.nq = sp.nq;
  for (int l = 0; l < sp.nb; ++l) {
    t.kind[l] = sp.kind[l];
    for (int q = 0; q < sp.nq; ++q) {
      t.vx[l][q] = sp.vec[((size_t)l * 2 + 0) * sp.nq + q];
      t.vy[l][q] = sp.vec[((size_t)l * 2 + 1) * sp.nq + q];
      t.sc[l][q] = sp.sc[(size_t)l * sp.nq + q];
    }
  }
  for (int q = 0; q < sp.nq; ++q) t.X[q] = sp.X[q], t.Y[q] = sp.X[sp.nq + q], t.W[q] = sp.W[q];
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_ptab, &t, sizeof(t)));
}

// ------------------------------------------------------------------------------ small vector kernels
__global__ void k_unscale_ez(cplx* __restrict__ x, const uint8_t* __restrict__ is_z, int n, cplx inv_factor) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && is_z[i]) x[i] = x[i] * inv_factor;
}
__global__ void k_cscale(cplx* __restrict__ x, int64_t n, cplx f) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i] * f;
}
__global__ void k_rhs_combine(const cplx* __restrict__ a, const cplx* __restrict__ b, cplx ibeta, int n,
                              cplx* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + ibeta * b[i];
}
__global__ void k_csr_diag(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                           const double* __restrict__ vals, double* __restrict__ diag) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0;
  for (int k = indptr[r]; k < indptr[r + 1]; ++k)
    if (indices[k] == r) d = vals[k];
  diag[r] = d;
}
// z = r / diag ; partial sums of conj(r).z (real for SPD) per block
__global__ void __launch_bounds__(256) k_precond_dot(const cplx* __restrict__ r, const double* __restrict__ diag,
                                                     int n, cplx* __restrict__ z, double* __restrict__ partial) {
  __shared__ double red[8];
  const int t = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t i = base + k * 256 + t;
    if (i < n) {
      cplx rv = r[i];
      cplx zv = rv / diag[i];
      z[i] = zv;
      s += rv.re * zv.re + rv.im * zv.im;
    }
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
// partial sums of Re(conj(a).b)
__global__ void __launch_bounds__(256) k_rdot(const cplx* __restrict__ a, const cplx* __restrict__ b, int n,
                                              double* __restrict__ partial) {
  __shared__ double red[8];
  const int t = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t i = base + k * 256 + t;
    if (i < n) s += a[i].re * b[i].re + a[i].im * b[i].im;
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
__global__ void __launch_bounds__(256) k_sum_double(const double* __restrict__ partial, int nblocks,
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
// x += alpha p ; r -= alpha q
__global__ void k_cg_update(cplx* __restrict__ x, cplx* __restrict__ r, const cplx* __restrict__ p,
                            const cplx* __restrict__ q, int n, double alpha) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    x[i] = x[i] + p[i] * alpha;
    r[i] = r[i] - q[i] * alpha;
  }
}
// p = z + beta p
__global__ void k_cg_dir(cplx* __restrict__ p, const cplx* __restrict__ z, int n, double beta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = z[i] + p[i] * beta;
}

namespace {
struct CG {
  DevBuf<double> diag, partial, scal;
  DevBuf<double> r, z, p, q;
};
double reduce_to_host(Handle& h, CG& w, int nblk) {
  k_sum_double<<<1, 256, 0, h.stream>>>(w.partial.p, nblk, w.scal.p);
  double v = 0;
  w.scal.download(&v, 1, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  h.tim.kernel_launches += 1;
  return v;
}
}  // namespace

// Mass x = b  (Mass real SPD on the structural pattern, b complex); x overwritten
static void mass_solve(Handle& h, const cplx* b, cplx* x) {
  const int n = h.sp.n;
  cudaStream_t s = h.stream;
  const Pattern& pat = h.pat;
  CG w;
  const int nblk = cdiv(n, 1024);
  w.diag.alloc(n);
  w.partial.alloc(nblk);
  w.scal.alloc(1);
  w.r.alloc((size_t)2 * n);
  w.z.alloc((size_t)2 * n);
  w.p.alloc((size_t)2 * n);
  w.q.alloc((size_t)2 * n);
  cplx *r = (cplx*)w.r.p, *z = (cplx*)w.z.p, *p = (cplx*)w.p.p, *q = (cplx*)w.q.p;
  const int g = cdiv(n, 256);
  k_csr_diag<<<g, 256, 0, s>>>(n, pat.indptr.p, pat.indices.p, h.Mass.v.p, w.diag.p);
  FW_CHECK_CUDA(cudaMemsetAsync(x, 0, (size_t)n * sizeof(cplx), s));
  FW_CHECK_CUDA(cudaMemcpyAsync(r, b, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  k_rdot<<<nblk, 256, 0, s>>>(r, r, n, w.partial.p);
  const double bnorm2 = reduce_to_host(h, w, nblk);
  if (!(bnorm2 > 0)) return;
  k_precond_dot<<<nblk, 256, 0, s>>>(r, w.diag.p, n, z, w.partial.p);
  double rz = reduce_to_host(h, w, nblk);
  FW_CHECK_CUDA(cudaMemcpyAsync(p, z, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
  const double target = 1e-28 * bnorm2;  // ||r|| <= 1e-14 ||b||
  for (int it = 0; it < 500; ++it) {
    spmv<double, cplx>(n, pat.indptr.p, pat.indices.p, h.Mass.v.p, p, q, s);
    k_rdot<<<nblk, 256, 0, s>>>(p, q, n, w.partial.p);
    const double pq = reduce_to_host(h, w, nblk);
    if (!(pq > 0)) break;
    const double alpha = rz / pq;
    k_cg_update<<<g, 256, 0, s>>>(x, r, p, q, n, alpha);
    k_rdot<<<nblk, 256, 0, s>>>(r, r, n, w.partial.p);
    const double rr = reduce_to_host(h, w, nblk);
    h.tim.kernel_launches += 4;
    if (rr <= target) break;
    k_precond_dot<<<nblk, 256, 0, s>>>(r, w.diag.p, n, z, w.partial.p);
    const double rz_new = reduce_to_host(h, w, nblk);
    k_cg_dir<<<g, 256, 0, s>>>(p, z, n, rz_new / rz);
    rz = rz_new;
    h.tim.kernel_launches += 2;
  }
  FW_CHECK_CUDA(cudaGetLastError());
}

// H = (-i / (mu0 omega)) Mass^-1 (C0 + i beta C1) E        (waveguide.py:657-677)
// all k modes at once: rhs_i = C0 e_i + i beta_i C1 e_i, one block PCG for the k mass solves
void hfield_multi_dev(Handle& h, int k, const cplx* E_dev, const cplx* betas, double omega, double mu0, cplx* H_dev) {
  Trace tr(h.stream);
  Timer th(h.stream, "fw:H field (mass solves)");
  h.tim.pcg_iterations = 0;
  h.tim.pcg_rel_residual = 0.0;
  h.tim.mass_direct_blocks = 0;
  assemble_aux(h);
  tr.mark("  hfield: aux operators");
  const int n = h.sp.n;
  cudaStream_t s = h.stream;
  const Pattern& pat = h.pat;
  DevBuf<double> t0, t1, rhs;
  t0.alloc((size_t)2 * n);
  t1.alloc((size_t)2 * n);
  rhs.alloc((size_t)2 * n * k);
  const cplx f = cplx{0.0, -1.0 / (mu0 * omega)};
  for (int k0 = 0; k0 < k; k0 += 8) {
    const int kc = std::min(8, k - k0);
    for (int i = k0; i < k0 + kc; ++i) {
      const cplx* E = E_dev + (size_t)i * n;
      spmv<double, cplx>(n, pat.indptr.p, pat.indices.p, h.C0.v.p, E, (cplx*)t0.p, s);
      spmv<double, cplx>(n, pat.indptr.p, pat.indices.p, h.C1.v.p, E, (cplx*)t1.p, s);
      const cplx ibeta = cplx{0.0, 1.0} * betas[i];
      k_rhs_combine<<<cdiv(n, 256), 256, 0, s>>>((const cplx*)t0.p, (const cplx*)t1.p, ibeta, n,
                                                (cplx*)rhs.p + (size_t)i * n);
      h.tim.kernel_launches += 3;
    }
    mass_solve_multi(h, kc, (const cplx*)rhs.p + (size_t)k0 * n, H_dev + (size_t)k0 * n);
  }
  k_cscale<<<cdiv((int64_t)n * k, 256), 256, 0, s>>>(H_dev, n * k, f);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 1;
  h.tim.hfield_ms = th.stop();
}

void hfield_dev(Handle& h, const cplx* E_dev, cplx beta, double omega, double mu0, cplx* H_dev) {
  hfield_multi_dev(h, 1, E_dev, &beta, omega, mu0, H_dev);
}

// ------------------------------------------------------------------------------ functionals
__device__ inline cplx csqrt_(cplx z) {
  double r = hypot(z.re, z.im);
  if (r == 0.0) return {0.0, 0.0};
  double a = sqrt(0.5 * (r + fabs(z.re)));
  double b = 0.5 * z.im / a;
  if (z.re >= 0) return {a, b};
  return {fabs(b), z.im >= 0 ? a : -a};
}

template <int NB, class AUX>
__global__ void __launch_bounds__(128) k_functional(int kind, int nsel, const int* __restrict__ sel, int ne, int nv,
                                                    const double* __restrict__ p, const int* __restrict__ t,
                                                    const int* __restrict__ edofs, const cplx* __restrict__ f0,
                                                    const cplx* __restrict__ f1, const cplx* __restrict__ f2,
                                                    const cplx* __restrict__ f3, int aux_kind,
                                                    const AUX* __restrict__ aux, int option,
                                                    cplx* __restrict__ partial) {
  __shared__ cplx red[4][3];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  cplx out[3] = {{0, 0}, {0, 0}, {0, 0}};
  if (tid < nsel) {
    const int e = sel ? sel[tid] : tid;
    const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
    const double x0 = p[v0], y0 = p[nv + v0];
    const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
    const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
    const double det = a11 * a22 - a12 * a21;
    const double idet = 1.0 / det, adet = fabs(det);
    const cplx* fields[4] = {f0, f1, f2, f3};
    const int nf = (kind == FW_FUN_OVERLAP) ? 4 : (kind == FW_FUN_COUPLING ? 2 : 1);
    cplx c[4][NB];
    for (int f = 0; f < nf; ++f)
#pragma unroll
      for (int l = 0; l < NB; ++l) c[f][l] = fields[f][edofs[(size_t)l * ne + e]];
    AUX a0{}, a1{}, a2{};
    if (aux) {
      if (aux_kind == FW_EPS_P0)
        a0 = a1 = a2 = aux[e];
      else
        a0 = aux[3 * (size_t)e], a1 = aux[3 * (size_t)e + 1], a2 = aux[3 * (size_t)e + 2];
    }
    const int nq = c_ptab.nq;
    for (int q = 0; q < nq; ++q) {
      const double w = c_ptab.W[q] * adet;
      cplx ex[4], ey[4], ez[4];
      for (int f = 0; f < nf; ++f) {
        cplx sx{0, 0}, sy{0, 0}, sz{0, 0};
#pragma unroll
        for (int l = 0; l < NB; ++l) {
          if (c_ptab.kind[l] == 0) {
            fma_(sx, c[f][l], c_ptab.vx[l][q]);
            fma_(sy, c[f][l], c_ptab.vy[l][q]);
          } else {
            fma_(sz, c[f][l], c_ptab.sc[l][q]);
          }
        }
        // covariant Piola: E_t = A^-T (sx, sy)
        ex[f] = (sx * a22 - sy * a21) * idet;
        ey[f] = (sy * a11 - sx * a12) * idet;
        ez[f] = sz;
      }
      AUX ax = a0, ay = a0, az = a0;
      if (aux) {
        if (aux_kind == FW_EPS_DG_P1) {
          const double Xq = c_ptab.X[q], Yq = c_ptab.Y[q];
          ax = a0 * (1.0 - Xq - Yq) + a1 * Xq + a2 * Yq;
          ay = az = ax;
        } else if (aux_kind == FW_EPS_VEC_P0) {
          ax = a0, ay = a1, az = a2;
        }
      }
      if (kind == FW_FUN_OVERLAP) {
        // fields: E_i, H_i, E_j, H_j ; cross(conj(E_i), H_j) + cross(E_j, conj(H_i))
        cplx v = conj_(ex[0]) * ey[3] - conj_(ey[0]) * ex[3] + ex[2] * conj_(ey[1]) - ey[2] * conj_(ex[1]);
        out[0] += v * (0.5 * w);
      } else if (kind == FW_FUN_EX2_EY2 || kind == FW_FUN_EX2_EY2_EALL) {
        const double x2 = abs2_(ex[0]), y2 = abs2_(ey[0]);
        out[0].re += x2 * w;
        out[1].re += y2 * w;
        if (kind == FW_FUN_EX2_EY2_EALL) out[2].re += (x2 + y2 + abs2_(ez[0])) * w;
      } else if (kind == FW_FUN_COUPLING) {
        cplx v = conj_(ex[0]) * ex[1] + conj_(ey[0]) * ey[1] + conj_(ez[0]) * ez[1];
        out[0] += (ax * v) * w;
      } else if (kind == FW_FUN_E2_E4) {
        double e2;
        if (option == 0)
          e2 = abs2_(ex[0]) + abs2_(ey[0]);
        else if (option == 1)
          e2 = abs2_(ex[0]);
        else
          e2 = abs2_(ey[0]);
        out[0].re += e2 * w;
        out[1].re += e2 * e2 * w;
      } else if (kind == FW_FUN_CONFINEMENT) {
        cplx sxr, syr, szr;
        if constexpr (scalar_traits<AUX>::is_complex) {
          sxr = csqrt_(ax), syr = csqrt_(ay), szr = csqrt_(az);
        } else {
          sxr = cplx{sqrt(ax), 0.0}, syr = cplx{sqrt(ay), 0.0}, szr = cplx{sqrt(az), 0.0};
        }
        cplx v = sxr * abs2_(ex[0]) + syr * abs2_(ey[0]) + szr * abs2_(ez[0]);
        out[0] += v * w;
      }
    }
  }
  // block reduce (4 warps)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      out[k].re += __shfl_xor_sync(0xffffffffu, out[k].re, o);
      out[k].im += __shfl_xor_sync(0xffffffffu, out[k].im, o);
    }
    if (lane == 0) red[warp][k] = out[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    cplx s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    partial[(size_t)blockIdx.x * 3 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) k_sum_cplx3(const cplx* __restrict__ partial, int nblocks,
                                                   cplx* __restrict__ out) {
  __shared__ cplx sh[256];
  const int t = threadIdx.x, k = blockIdx.x;
  cplx s{0, 0};
  for (int b = t; b < nblocks; b += 256) s += partial[(size_t)b * 3 + k];
  sh[t] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  if (t == 0) out[k] = sh[0];
}

void functional_dev(Handle& h, int kind, const cplx* f0, const cplx* f1, const cplx* f2, const cplx* f3, int aux_kind,
                    bool aux_complex, const void* aux_dev, int n_sel, const int* elements_dev, int option, cd* out3) {
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "space missing");
  FW_REQUIRE(kind >= FW_FUN_OVERLAP && kind <= FW_FUN_CONFINEMENT, "unknown functional");
  FW_REQUIRE(f0 != nullptr, "field 0 missing");
  if (kind == FW_FUN_OVERLAP) FW_REQUIRE(f1 && f2 && f3, "overlap needs 4 fields");
  if (kind == FW_FUN_COUPLING) FW_REQUIRE(f1 && aux_dev, "coupling needs 2 fields and delta_epsilon");
  if (kind == FW_FUN_CONFINEMENT) FW_REQUIRE(aux_dev, "confinement needs epsilon");
  cudaStream_t s = h.stream;
  upload_ptab(h);
  const int nsel = elements_dev ? n_sel : sp.ne;
  const int nblk = std::max(1, cdiv(nsel, 128));
  DevBuf<double> partial, res;
  partial.alloc((size_t)nblk * 3 * 2);
  res.alloc(6);
#define FW_LAUNCH_FUN(NBV, AUXT)                                                                               \
  k_functional<NBV, AUXT><<<nblk, 128, 0, s>>>(kind, nsel, elements_dev, sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, \
                                               f0, f1, f2, f3, aux_kind, (const AUXT*)aux_dev, option,         \
                                               (cplx*)partial.p)
  if (sp.nb == 6) {
    if (aux_complex)
      FW_LAUNCH_FUN(6, cplx);
    else
      FW_LAUNCH_FUN(6, double);
  } else {
    if (aux_complex)
      FW_LAUNCH_FUN(14, cplx);
    else
      FW_LAUNCH_FUN(14, double);
  }
#undef FW_LAUNCH_FUN
  k_sum_cplx3<<<3, 256, 0, s>>>((const cplx*)partial.p, nblk, (cplx*)res.p);
  FW_CHECK_CUDA(cudaGetLastError());
  double r[6];
  res.download(r, 6, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  for (int k = 0; k < 3; ++k) out3[k] = cd(r[2 * k], r[2 * k + 1]);
  h.tim.kernel_launches += 2;
}

// ------------------------------------------------------------------------------ intensity (a16)
// Mode.calculate_intensity (waveguide.py:260-280): I = 0.5 Re(Ex Hy* - Ey Hx*) at the quadrature points,
// L2-projected onto discontinuous P1 (one 3x3 solve per triangle, dof = 3 e + a).
template <int NB>
__global__ void __launch_bounds__(128) k_intensity(int ne, int nv, const double* __restrict__ p,
                                                   const int* __restrict__ t, const int* __restrict__ edofs,
                                                   const cplx* __restrict__ E, const cplx* __restrict__ Hf,
                                                   double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
  const double x0 = p[v0], y0 = p[nv + v0];
  const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
  const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
  const double det = a11 * a22 - a12 * a21;
  const double idet = 1.0 / det, adet = fabs(det);
  cplx ce[NB], ch[NB];
#pragma unroll
  for (int l = 0; l < NB; ++l) {
    const int d = edofs[(size_t)l * ne + e];
    ce[l] = E[d];
    ch[l] = Hf[d];
  }
  double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3] = {0, 0, 0};
  const int nq = c_ptab.nq;
  for (int q = 0; q < nq; ++q) {
    const double w = c_ptab.W[q] * adet;
    cplx sx{0, 0}, sy{0, 0}, hx{0, 0}, hy{0, 0};
#pragma unroll
    for (int l = 0; l < NB; ++l) {
      if (c_ptab.kind[l] == 0) {
        fma_(sx, ce[l], c_ptab.vx[l][q]);
        fma_(sy, ce[l], c_ptab.vy[l][q]);
        fma_(hx, ch[l], c_ptab.vx[l][q]);
        fma_(hy, ch[l], c_ptab.vy[l][q]);
      }
    }
    const cplx ex = (sx * a22 - sy * a21) * idet, ey = (sy * a11 - sx * a12) * idet;
    const cplx bx = (hx * a22 - hy * a21) * idet, by = (hy * a11 - hx * a12) * idet;
    // 0.5 Re(Ex conj(Hy) - Ey conj(Hx))
    const double f = 0.5 * ((ex.re * by.re + ex.im * by.im) - (ey.re * bx.re + ey.im * bx.im));
    const double lam[3] = {1.0 - c_ptab.X[q] - c_ptab.Y[q], c_ptab.X[q], c_ptab.Y[q]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      b[a] += w * lam[a] * f;
#pragma unroll
      for (int c = 0; c < 3; ++c) M[a][c] += w * lam[a] * lam[c];
    }
  }
  // 3x3 solve (Cramer; the local mass matrix is SPD)
  const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
  const double c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
  const double c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
  const double dd = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02;
  const double id = 1.0 / dd;
  const double x = (b[0] * c00 + b[1] * (M[0][2] * M[2][1] - M[0][1] * M[2][2]) +
                    b[2] * (M[0][1] * M[1][2] - M[0][2] * M[1][1])) * id;
  const double y = (b[0] * c01 + b[1] * (M[0][0] * M[2][2] - M[0][2] * M[2][0]) +
                    b[2] * (M[0][2] * M[1][0] - M[0][0] * M[1][2])) * id;
  const double z = (b[0] * c02 + b[1] * (M[0][1] * M[2][0] - M[0][0] * M[2][1]) +
                    b[2] * (M[0][0] * M[1][1] - M[0][1] * M[1][0])) * id;
  out[3 * (size_t)e] = x;
  out[3 * (size_t)e + 1] = y;
  out[3 * (size_t)e + 2] = z;
}

void intensity_dev(Handle& h, const cplx* E, const cplx* Hf, double* out) {
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "space missing");
  upload_ptab(h);
  if (sp.nb == 6)
    k_intensity<6><<<cdiv(sp.ne, 128), 128, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, E, Hf, out);
  else
    k_intensity<14><<<cdiv(sp.ne, 128), 128, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, E, Hf, out);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 1;
}

// ------------------------------------------------------------------------------ Poynting vector (a16)
// Mode.poynting (waveguide.py:74-95): P = E x H (no conjugation) at the quadrature points, every component
// L2-projected onto the discontinuous Lagrange space of the E_z field.  The local mass matrix is |det| times a
// reference matrix, so the projection is u = Mref^-1 sum_q W phi_a P(q): Mref^-1 (nl x nl, built from the
// basis quadrature like basis.project does) sits in constant memory.
__constant__ double c_minv[6][6];

template <int NB>
__global__ void __launch_bounds__(128) k_poynting(int ne, int nv, const double* __restrict__ p,
                                                  const int* __restrict__ t, const int* __restrict__ edofs,
                                                  const cplx* __restrict__ E, const cplx* __restrict__ Hf,
                                                  cplx* __restrict__ out) {
  constexpr int NL = NB == 6 ? 3 : 6;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
  const double x0 = p[v0], y0 = p[nv + v0];
  const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
  const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
  const double det = a11 * a22 - a12 * a21;
  const double idet = 1.0 / det;
  cplx ce[NB], ch[NB];
#pragma unroll
  for (int l = 0; l < NB; ++l) {
    const int d = edofs[(size_t)l * ne + e];
    ce[l] = E[d];
    ch[l] = Hf[d];
  }
  cplx rhs[NL][3];
#pragma unroll
  for (int a = 0; a < NL; ++a) rhs[a][0] = rhs[a][1] = rhs[a][2] = cplx{0, 0};
  const int nq = c_ptab.nq;
  for (int q = 0; q < nq; ++q) {
    cplx sx{0, 0}, sy{0, 0}, ez{0, 0}, tx{0, 0}, ty{0, 0}, hz{0, 0};
#pragma unroll
    for (int l = 0; l < NB; ++l) {
      if (c_ptab.kind[l] == 0) {
        fma_(sx, ce[l], c_ptab.vx[l][q]);
        fma_(sy, ce[l], c_ptab.vy[l][q]);
        fma_(tx, ch[l], c_ptab.vx[l][q]);
        fma_(ty, ch[l], c_ptab.vy[l][q]);
      } else {
        fma_(ez, ce[l], c_ptab.sc[l][q]);
        fma_(hz, ch[l], c_ptab.sc[l][q]);
      }
    }
    const cplx ex = (sx * a22 - sy * a21) * idet, ey = (sy * a11 - sx * a12) * idet;
    const cplx hx = (tx * a22 - ty * a21) * idet, hy = (ty * a11 - tx * a12) * idet;
    const cplx P[3] = {ey * hz - ez * hy, ez * hx - ex * hz, ex * hy - ey * hx};
    int a = 0;
#pragma unroll
    for (int l = 0; l < NB; ++l) {
      if (c_ptab.kind[l] == 1) {
        const double wphi = c_ptab.W[q] * c_ptab.sc[l][q];
#pragma unroll
        for (int c = 0; c < 3; ++c) fma_(rhs[a][c], P[c], wphi);
        ++a;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < NL; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cplx u{0, 0};
#pragma unroll
      for (int b = 0; b < NL; ++b) fma_(u, rhs[b][c], c_minv[a][b]);
      out[(size_t)e * (3 * NL) + a * 3 + c] = u;
    }
}

void poynting_dev(Handle& h, const cplx* E, const cplx* Hf, cplx* out) {
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "space missing");
  upload_ptab(h);
  // reference mass matrix of the Lagrange functions with the basis quadrature, inverted by Gauss-Jordan
  const int nl = sp.nb == 6 ? 3 : 6;
  std::vector<int> lag;
  for (int l = 0; l < sp.nb; ++l)
    if (sp.kind[l] == 1) lag.push_back(l);
  double M[6][12];
  for (int a = 0; a < nl; ++a)
    for (int b = 0; b < 2 * nl; ++b) M[a][b] = 0.0;
  for (int a = 0; a < nl; ++a) {
    for (int b = 0; b < nl; ++b)
      for (int q = 0; q < sp.nq; ++q)
        M[a][b] += sp.W[q] * sp.sc[(size_t)lag[a] * sp.nq + q] * sp.sc[(size_t)lag[b] * sp.nq + q];
    M[a][nl + a] = 1.0;
  }
  for (int c = 0; c < nl; ++c) {
    int piv = c;
    for (int r = c + 1; r < nl; ++r)
      if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
    FW_REQUIRE(fabs(M[piv][c]) > 1e-14, "the basis quadrature cannot resolve the Lagrange mass matrix (singular projection)");
    if (piv != c)
      for (int b = 0; b < 2 * nl; ++b) std::swap(M[c][b], M[piv][b]);
    const double inv = 1.0 / M[c][c];
    for (int b = 0; b < 2 * nl; ++b) M[c][b] *= inv;
    for (int r = 0; r < nl; ++r) {
      if (r == c) continue;
      const double f = M[r][c];
      if (f != 0.0)
        for (int b = 0; b < 2 * nl; ++b) M[r][b] -= f * M[c][b];
    }
  }
  double minv[6][6];
  memset(minv, 0, sizeof(minv));
  for (int a = 0; a < nl; ++a)
    for (int b = 0; b < nl; ++b) minv[a][b] = M[a][nl + b];
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_minv, minv, sizeof(minv)));
  if (sp.nb == 6)
    k_poynting<6><<<cdiv(sp.ne, 128), 128, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, E, Hf, out);
  else
    k_poynting<14><<<cdiv(sp.ne, 128), 128, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, E, Hf, out);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 1;
}

// ------------------------------------------------------------------------------ energy current density (a16)
// calculate_energy_current_density (waveguide.py:680-696): P0 projection of |e_x|^2 + |e_y|^2 + |e_z| (sic: the
// reference does not square e_z); the P0 mass matrix is diag(area), so the value is sum_q W f / sum_q W.
template <int NB>
__global__ void __launch_bounds__(128) k_energy_density(int ne, int nv, const double* __restrict__ p,
                                                        const int* __restrict__ t, const int* __restrict__ edofs,
                                                        const cplx* __restrict__ E, cplx* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
  const double x0 = p[v0], y0 = p[nv + v0];
  const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
  const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
  const double idet = 1.0 / (a11 * a22 - a12 * a21);
  cplx ce[NB];
#pragma unroll
  for (int l = 0; l < NB; ++l) ce[l] = E[edofs[(size_t)l * ne + e]];
  double num = 0.0, den = 0.0;
  const int nq = c_ptab.nq;
  for (int q = 0; q < nq; ++q) {
    cplx sx{0, 0}, sy{0, 0}, ez{0, 0};
#pragma unroll
    for (int l = 0; l < NB; ++l) {
      if (c_ptab.kind[l] == 0) {
        fma_(sx, ce[l], c_ptab.vx[l][q]);
        fma_(sy, ce[l], c_ptab.vy[l][q]);
      } else {
        fma_(ez, ce[l], c_ptab.sc[l][q]);
      }
    }
    const cplx ex = (sx * a22 - sy * a21) * idet, ey = (sy * a11 - sx * a12) * idet;
    num += c_ptab.W[q] * (abs2_(ex) + abs2_(ey) + absv_(ez));
    den += c_ptab.W[q];
  }
  out[e] = cplx{num / den, 0.0};
}

void energy_density_dev(Handle& h, const cplx* E, cplx* out) {
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "space missing");
  upload_ptab(h);
  if (sp.nb == 6)
    k_energy_density<6><<<cdiv(sp.ne, 128), 128, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, E, out);
  else
    k_energy_density<14><<<cdiv(sp.ne, 128), 128, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, E, out);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 1;
}

// waveguide.py:627-639: un-scale E_z, recover H, power-normalise
void postprocess_modes_dev(Handle& h, int k, const cd* lams, const double* X_dev, double k0, double omega, double mu0,
                           double* E_dev, double* H_dev, cd* powers) {
  const int n = h.sp.n;
  cudaStream_t s = h.stream;
  if ((const void*)E_dev != (const void*)X_dev)
    FW_CHECK_CUDA(cudaMemcpyAsync(E_dev, X_dev, (size_t)k * n * 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  Trace tr(s);
  std::vector<cplx> betas(k);
  for (int i = 0; i < k; ++i) {
    cplx* E = (cplx*)E_dev + (size_t)i * n;
    const cd beta = std::sqrt(lams[i]);
    betas[i] = cplx{beta.real(), beta.imag()};
    // xs[z] /= 1j * sqrt(lam / k0^4)
    const cd denom = cd(0, 1) * std::sqrt(lams[i] / (k0 * k0 * k0 * k0));
    const cd inv = 1.0 / denom;
    k_unscale_ez<<<cdiv(n, 256), 256, 0, s>>>(E, h.sp.is_z.p, n, cplx{inv.real(), inv.imag()});
  }
  tr.mark("post: unscale");
  hfield_multi_dev(h, k, (const cplx*)E_dev, betas.data(), omega, mu0, (cplx*)H_dev);
  tr.mark("post: hfield (aux assembly + rhs + PCG)");
  for (int i = 0; i < k; ++i) {
    cplx* E = (cplx*)E_dev + (size_t)i * n;
    cplx* Hf = (cplx*)H_dev + (size_t)i * n;
    cd o3[3];
    functional_dev(h, FW_FUN_OVERLAP, E, Hf, E, Hf, 0, false, nullptr, 0, nullptr, 0, o3);
    const cd power = o3[0];
    powers[i] = power;
    const cd isq = 1.0 / std::sqrt(power);
    k_cscale<<<cdiv(n, 256), 256, 0, s>>>(E, n, cplx{isq.real(), isq.imag()});
    k_cscale<<<cdiv(n, 256), 256, 0, s>>>(Hf, n, cplx{isq.real(), isq.imag()});
    h.tim.kernel_launches += 4;
  }
  tr.mark("post: power normalisation");
  FW_CHECK_CUDA(cudaGetLastError());
}

}  // namespace fw
