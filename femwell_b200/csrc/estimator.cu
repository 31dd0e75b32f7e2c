// Mode.eval_error_estimator (femwell/maxwell/waveguide.py:473-502): interior-facet jump estimator that drives the
// adaptive refinement loop of docs/photonics/examples/refinement.py:58-105.
//
//   tmp[f]  = int_f h ( |[grad E_z] . n|^2 + |[eps E_t] . n|^2 ) ds     for every interior facet f (h = |f|),
//   eta[e]  = 1/2 sum over the three facets of e of tmp[f]               (boundary facets contribute 0)
//
// The reference evaluates both sides with scikit-fem InteriorFacetBasis objects (side 0 / 1) on the Gauss-Legendre rule
// of degree 2 * maxdeg.  Here: one thread per facet evaluates the two traces from the reference tables of the local
// basis AT THE FACET POINTS (table driven like every other kernel: [nb][3 local edges][2][nqf], built by the host
// from the same local basis), one thread per element gathers.  With column-sorted `t` the parametrisation of an
// edge (low -> high global vertex) is the same in both neighbours, so the quadrature points coincide.
#include "solver.h"

namespace fw {

constexpr int MAXQF = 8;

struct FacetTables {
  int nb, nqf;
  int kind[MAXNB];
  double vx[MAXNB][3][MAXQF], vy[MAXNB][3][MAXQF];  // reference vector (Nedelec value | Lagrange gradient)
  double s[MAXQF], w[MAXQF];                        // Gauss-Legendre points / weights on [0, 1]
};
__constant__ FacetTables c_ftab;

template <int NB, class ET>
__device__ inline void facet_trace(int e, int k, int ne, int nv, const double* __restrict__ p,
                                   const int* __restrict__ t, const int* __restrict__ edofs,
                                   const cplx* __restrict__ E, int eps_kind, const ET* __restrict__ eps, double nx,
                                   double ny, cplx (&g)[MAXQF], cplx (&d)[MAXQF]) {
  const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
  const double x0 = p[v0], y0 = p[nv + v0];
  const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
  const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
  const double idet = 1.0 / (a11 * a22 - a12 * a21);
  cplx ce[NB];
#pragma unroll
  for (int l = 0; l < NB; ++l) ce[l] = E[edofs[(size_t)l * ne + e]];
  // reference coordinates of the end points of local edge k: (0,1) (1,2) (0,2)
  const double Xa = k == 1 ? 1.0 : 0.0, Ya = 0.0;
  const double Xb = k == 0 ? 1.0 : 0.0, Yb = k == 0 ? 0.0 : 1.0;
  for (int q = 0; q < c_ftab.nqf; ++q) {
    cplx tx{0, 0}, ty{0, 0}, gx{0, 0}, gy{0, 0};
#pragma unroll
    for (int l = 0; l < NB; ++l) {
      if (c_ftab.kind[l] == 0) {
        fma_(tx, ce[l], c_ftab.vx[l][k][q]);
        fma_(ty, ce[l], c_ftab.vy[l][k][q]);
      } else {
        fma_(gx, ce[l], c_ftab.vx[l][k][q]);
        fma_(gy, ce[l], c_ftab.vy[l][k][q]);
      }
    }
    // covariant Piola / gradient map: A^-T v_hat
    const cplx ex = (tx * a22 - ty * a21) * idet, ey = (ty * a11 - tx * a12) * idet;
    const cplx zx = (gx * a22 - gy * a21) * idet, zy = (gy * a11 - gx * a12) * idet;
    ET ev;
    if (eps_kind == FW_EPS_P0) {
      ev = eps[e];
    } else {  // DG-P1: dof (e, a) sits on vertex a
      const double sq = c_ftab.s[q];
      const double X = Xa + sq * (Xb - Xa), Y = Ya + sq * (Yb - Ya);
      ev = eps[3 * (size_t)e] * (1.0 - X - Y) + eps[3 * (size_t)e + 1] * X + eps[3 * (size_t)e + 2] * Y;
    }
    g[q] = zx * nx + zy * ny;
    d[q] = (ex * nx + ey * ny) * ev;
  }
}

template <int NB, class ET>
__global__ void __launch_bounds__(128) k_facet_jump(int nf, int ne, int nv, const double* __restrict__ p,
                                                    const int* __restrict__ t, const int* __restrict__ edofs,
                                                    const int* __restrict__ t2f, const int* __restrict__ f2t,
                                                    const cplx* __restrict__ E, int eps_kind,
                                                    const ET* __restrict__ eps, double* __restrict__ tmp) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int e0 = f2t[f], e1 = f2t[nf + f];
  if (e0 < 0 || e1 < 0) {
    tmp[f] = 0.0;
    return;
  }
  int k0 = 0, k1 = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (t2f[(size_t)k * ne + e0] == f) k0 = k;
    if (t2f[(size_t)k * ne + e1] == f) k1 = k;
  }
  // end points of the facet (local edge k0 of e0: vertices (0,1) (1,2) (0,2))
  const int ia = k0 == 1 ? 1 : 0, ib = k0 == 0 ? 1 : 2;
  const int va = t[(size_t)ia * ne + e0], vb = t[(size_t)ib * ne + e0];
  const double dx = p[vb] - p[va], dy = p[nv + vb] - p[nv + va];
  const double len = sqrt(dx * dx + dy * dy);
  const double nx = dy / len, ny = -dx / len;
  cplx g0[MAXQF], d0[MAXQF], g1[MAXQF], d1[MAXQF];
  facet_trace<NB, ET>(e0, k0, ne, nv, p, t, edofs, E, eps_kind, eps, nx, ny, g0, d0);
  facet_trace<NB, ET>(e1, k1, ne, nv, p, t, edofs, E, eps_kind, eps, nx, ny, g1, d1);
  double acc = 0.0;
  for (int q = 0; q < c_ftab.nqf; ++q) acc += c_ftab.w[q] * (abs2_(g0[q] - g1[q]) + abs2_(d0[q] - d1[q]));
  tmp[f] = acc * len * len;  // dx = |f| w_q and w.h = |f|
}

__global__ void __launch_bounds__(256) k_facet_gather(int ne, const int* __restrict__ t2f,
                                                      const double* __restrict__ tmp, double* __restrict__ eta) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  eta[e] = 0.5 * tmp[t2f[e]] + 0.5 * tmp[t2f[(size_t)ne + e]] + 0.5 * tmp[t2f[2 * (size_t)ne + e]];
}

void error_estimator_dev(Handle& h, const cplx* E, int eps_kind, bool eps_complex, const void* eps_dev, int nf,
                         const int* t2f_dev, const int* f2t_dev, int nqf, const double* s, const double* w,
                         const double* vec /*[nb][3][2][nqf]*/, double* eta_dev) {
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "space missing");
  FW_REQUIRE(nqf >= 1 && nqf <= MAXQF, "facet quadrature: 1..8 points");
  FW_REQUIRE(eps_kind == FW_EPS_P0 || eps_kind == FW_EPS_DG_P1,
             "the estimator multiplies E_t by a scalar epsilon (P0 or DG-P1), as the reference does");
  FacetTables ft;
  memset(&ft, 0, sizeof(ft));
  ft.nb = sp.nb, ft.nqf = nqf;
  for (int l = 0; l < sp.nb; ++l) {
    ft.kind[l] = sp.kind[l];
    for (int k = 0; k < 3; ++k)
      for (int q = 0; q < nqf; ++q) {
        ft.vx[l][k][q] = vec[(((size_t)l * 3 + k) * 2 + 0) * nqf + q];
        ft.vy[l][k][q] = vec[(((size_t)l * 3 + k) * 2 + 1) * nqf + q];
      }
  }
  for (int q = 0; q < nqf; ++q) ft.s[q] = s[q], ft.w[q] = w[q];
  cudaStream_t st = h.stream;
  FW_CHECK_CUDA(cudaStreamSynchronize(st));
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_ftab, &ft, sizeof(ft)));
  DevBuf<double> tmp;
  tmp.alloc((size_t)nf);
  const int g = cdiv(nf, 128);
#define FW_LAUNCH_JUMP(NBV, ET)                                                                                     \
  k_facet_jump<NBV, ET><<<g, 128, 0, st>>>(nf, sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, t2f_dev, f2t_dev, E, eps_kind, \
                                           (const ET*)eps_dev, tmp.p)
  if (sp.nb == 6) {
    if (eps_complex)
      FW_LAUNCH_JUMP(6, cplx);
    else
      FW_LAUNCH_JUMP(6, double);
  } else {
    if (eps_complex)
      FW_LAUNCH_JUMP(14, cplx);
    else
      FW_LAUNCH_JUMP(14, double);
  }
#undef FW_LAUNCH_JUMP
  k_facet_gather<<<cdiv(sp.ne, 256), 256, 0, st>>>(sp.ne, t2f_dev, tmp.p, eta_dev);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 2;
}

}  // namespace fw
