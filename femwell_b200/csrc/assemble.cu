// K1: fused per-element FP64 kernel for the mixed Nedelec x Lagrange forms of the waveguide
// eigenproblem.  Replaces, in one pass per element:
//   * affine map / covariant Piola (scikit-fem MappingAffine)
//   * basis_epsilon_r.interpolate(epsilon_r)      femwell/maxwell/waveguide.py:602
//   * aform  (curl-curl, eps-mass, grad couplings)  waveguide.py:577-594
//   * bform  (1/mu_t mass on E_t)                   waveguide.py:596-600
//   * the two forms of calculate_hfield           waveguide.py:658-668 (Mass, C0 + i*beta*C1)
//
// Formulation (differs from scikit-fem's pair-by-pair evaluation of physical basis arrays):
// every term of every form is bilinear in the REFERENCE tables,
//     a(u_j, v_i)|_q = vec_j(q) . alpha + sc_j(q) * beta,
// where (alpha, beta) depend only on the test function i, the point q and per-element metric
// tensors  G = A^-1 A^-T,  G_eps = A^-1 diag(eps_x, eps_y) A^-T  (A = affine Jacobian).
//
// Mapping: thread = (triangle, RPT test functions); blockIdx.y selects the row group; one launch per kind of
// test function (template KI: Nedelec rows / Lagrange rows), so the structurally-zero terms of every
// (test kind, trial kind) pair vanish at compile time and there is no divergence.  A warp covers 32
// consecutive triangles.  The thread keeps its rows of the local block (RPT x NB accumulators) in registers.
// The quadrature loop and the trial loop are fully unrolled (NQ and NB are template parameters); the
// reference tables are warp-uniform constant-bank loads (LDCU) feeding DFMAs with uniform-register operands.
// Stores are SoA, out[(i*NB+j)*Ne + e]: every store instruction of a warp is one contiguous 256 B (f64) /
// 512 B (c128) segment.  The kernel sits at the FP64 ridge: the HBM-store floor and the FP64-pipe floor are
// both ~0.16 ms at 500k triangles (profiles/).
#include <algorithm>

#include "handle.h"

namespace fw {


struct Tables {
  int nb, nq;
  int kind[MAXNB];
  int rows[2][MAXNB];  // local dofs of kind 0 (Nedelec) / 1 (Lagrange), in ascending order
  double vx[MAXNB][MAXQ], vy[MAXNB][MAXQ], sc[MAXNB][MAXQ];
  double X[MAXQ], Y[MAXQ], W[MAXQ];
  int nrank[MAXNB];  // rank of a Nedelec dof among the Nedelec dofs
};
__constant__ Tables c_tab;

// Quadrature sums of products of reference tables (constant-coefficient fast path, see k_element_blocks_const):
//   T11 = sum_q W vx_i vx_j,  T12 = sum_q W (vx_i vy_j + vy_i vx_j),  T22 = sum_q W vy_i vy_j,  CC = sum_q W sc_i sc_j
struct PairTables {
  double T11[MAXNB][MAXNB], T12[MAXNB][MAXNB], T22[MAXNB][MAXNB], CC[MAXNB][MAXNB];
};
__constant__ PairTables c_pair;

void upload_tables(Handle& h) {
  const Space& sp = h.sp;
  Tables t;
  memset(&t, 0, sizeof(t));
  t.nb = sp.nb;
  t.nq = sp.nq;
  int cnt[2] = {0, 0};
  for (int l = 0; l < sp.nb; ++l) {
    t.kind[l] = sp.kind[l];
    t.rows[sp.kind[l]][cnt[sp.kind[l]]++] = l;
    for (int q = 0; q < sp.nq; ++q) {
      t.vx[l][q] = sp.vec[((size_t)l * 2 + 0) * sp.nq + q];
      t.vy[l][q] = sp.vec[((size_t)l * 2 + 1) * sp.nq + q];
      t.sc[l][q] = sp.sc[(size_t)l * sp.nq + q];
    }
  }
  for (int q = 0; q < sp.nq; ++q) {
    t.X[q] = sp.X[q];
    t.Y[q] = sp.X[sp.nq + q];
    t.W[q] = sp.W[q];
  }
  for (int l = 0, r = 0; l < sp.nb; ++l) t.nrank[l] = sp.kind[l] == 0 ? r++ : 0;
  PairTables pt;
  memset(&pt, 0, sizeof(pt));
  for (int i = 0; i < sp.nb; ++i)
    for (int j = 0; j < sp.nb; ++j) {
      double s11 = 0, s12 = 0, s22 = 0, scc = 0;
      for (int q = 0; q < sp.nq; ++q) {
        const double w = t.W[q];
        s11 += w * t.vx[i][q] * t.vx[j][q];
        s12 += w * (t.vx[i][q] * t.vy[j][q] + t.vy[i][q] * t.vx[j][q]);
        s22 += w * t.vy[i][q] * t.vy[j][q];
        scc += w * t.sc[i][q] * t.sc[j][q];
      }
      pt.T11[i][j] = s11, pt.T12[i][j] = s12, pt.T22[i][j] = s22, pt.CC[i][j] = scc;
    }
  // synchronous copy from a stack object: the stream is idle-ordered afterwards
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_tab, &t, sizeof(t)));
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_pair, &pt, sizeof(pt)));
}

enum { FORM_A = 0, FORM_B = 1, FORM_MASS = 2, FORM_C0 = 3, FORM_C1 = 4 };

// composite local layout is fixed per order (element.py local_layout / SURVEY.md A.3):
// order 1: P P P N N N ; order 2: P P P (N N P)x3 N N      (0 = Nedelec, 1 = Lagrange)
template <int NB>
__host__ __device__ constexpr int kind_of(int l) {
  return NB == 6 ? (l < 3 ? 1 : 0) : (l < 3 ? 1 : (l >= 12 ? 0 : (((l - 3) % 3) == 2 ? 1 : 0)));
}
// rank of local dof l among the Nedelec dofs (compact tt-block index used by FORM_B)
template <int NB>
__host__ __device__ constexpr int nrank_of(int l) {
  int r = 0;
  for (int m = 0; m < l; ++m) r += kind_of<NB>(m) == 0 ? 1 : 0;
  return r;
}

// per-triangle state kept in registers
template <class S>
struct Elem {
  double x0, a11, a12, adet, idet, g11, g12, g22, c_curl;
  S e0, e1, e2, ge11, ge12, ge22;
};

// NQ > 0: compile-time number of quadrature points (loops fully unrolled); NQ == 0: run-time nq.
// KI: kind of the test functions handled by this launch (0 = Nedelec rows, 1 = Lagrange rows).
// RPT: test functions (rows of the local block) a thread computes SIMULTANEOUSLY: blockIdx.y selects a group
//      of RPT rows of kind KI.  Every reference-table value fetched (uniform constant load) feeds RPT DFMAs
//      and the per-triangle prologue (dependent t -> p loads, metric tensors) is amortised over RPT rows --
//      the ncu source page of the one-row variant showed 19 % of the stall samples on the constant loads
//      and 22 % on the prologue.
template <int NB, int NQ, class S, int EPSKIND, int FORM, int KI, int RPT>
__global__ void __launch_bounds__(128) k_element_blocks(int ne, int nv, const double* __restrict__ p,
                                                        const int* __restrict__ t, const S* __restrict__ eps,
                                                        double k0, double mu_r, double inv_radius,
                                                        S* __restrict__ out) {
  constexpr int NT = (NB == 6) ? 3 : 8;
  constexpr int NROWS = KI == 0 ? NT : NB - NT;
  using T = scalar_traits<S>;
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= ne) return;
  const double ik02 = 1.0 / (k0 * k0);
  const double imu = 1.0 / mu_r;
  const bool bend = inv_radius != 0.0;  // warp-uniform: radius=inf makes the bend factor exactly 1
  const int nq = NQ > 0 ? NQ : c_tab.nq;

  Elem<S> g;
  {
    const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
    const double x0 = p[v0], y0 = p[nv + v0];
    const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
    const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
    const double det = a11 * a22 - a12 * a21;
    const double idet = 1.0 / det;
    const double id2 = idet * idet;
    g.x0 = x0, g.a11 = a11, g.a12 = a12;
    g.adet = fabs(det), g.idet = idet;
    // G = A^-1 A^-T (symmetric)
    g.g11 = id2 * (a22 * a22 + a12 * a12);
    g.g12 = -id2 * (a22 * a21 + a12 * a11);
    g.g22 = id2 * (a21 * a21 + a11 * a11);
    g.c_curl = imu * ik02 * id2;  // 1/(mu k0^2 det^2): curl-curl coefficient
    g.e0 = g.e1 = g.e2 = g.ge11 = g.ge12 = g.ge22 = T::zero();
    if (FORM == FORM_A) {
      if (EPSKIND == FW_EPS_P0) {
        g.e0 = eps[e];
      } else {
        g.e0 = eps[3 * (size_t)e], g.e1 = eps[3 * (size_t)e + 1], g.e2 = eps[3 * (size_t)e + 2];
      }
      if (EPSKIND == FW_EPS_VEC_P0) {
        // G_eps = A^-1 diag(eps_x, eps_y) A^-T for diagonal anisotropy (constant per element)
        g.ge11 = (g.e0 * (a22 * a22) + g.e1 * (a12 * a12)) * id2;
        g.ge12 = (g.e0 * (a22 * a21) + g.e1 * (a12 * a11)) * (-id2);
        g.ge22 = (g.e0 * (a21 * a21) + g.e1 * (a11 * a11)) * id2;
      }
    }
  }

  // which of (alpha, beta) are structurally non-zero for this (form, test kind)
  constexpr bool useAN = (FORM == FORM_A) || (FORM == FORM_B) || (FORM == FORM_MASS && KI == 0) || (FORM == FORM_C1 && KI == 0);
  constexpr bool useBN = (FORM == FORM_A && KI == 0) || (FORM == FORM_C0 && KI == 1);
  constexpr bool useAP = (FORM == FORM_A && KI == 0) || (FORM == FORM_C0 && KI == 0);
  constexpr bool useBP = (FORM == FORM_A && KI == 1) || (FORM == FORM_MASS && KI == 1);

  // the RPT rows of this thread (block-uniform); groups at the end may be short
  int irow[RPT], il[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    irow[r] = blockIdx.y * RPT + r;
    il[r] = c_tab.rows[KI][irow[r] < NROWS ? irow[r] : NROWS - 1];
  }

  S acc[RPT][NB];
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[r][j] = T::zero();

#pragma unroll
  for (int q = 0; q < (NQ > 0 ? NQ : MAXQ); ++q) {
    if (NQ == 0 && q >= nq) break;
    const double Xq = c_tab.X[q], Yq = c_tab.Y[q], Wq = c_tab.W[q];
    const double w = Wq * g.adet;
    // bend factor f = 1 + x/R (waveguide.py:580-586); wf = w f, wrf = w / f
    double wf = w, wrf = w;
    if (bend) {
      const double f = 1.0 + (g.x0 + g.a11 * Xq + g.a12 * Yq) * inv_radius;
      wf = w * f;
      wrf = w / f;
    }
    S ezq = g.e0;  // eps_z at the point (isotropic kinds)
    if (FORM == FORM_A && EPSKIND == FW_EPS_DG_P1) ezq = g.e0 * (1.0 - Xq - Yq) + g.e1 * Xq + g.e2 * Yq;
    if (FORM == FORM_A && EPSKIND == FW_EPS_VEC_P0) ezq = g.e2;
    S aNx[RPT], aNy[RPT], aPx[RPT], aPy[RPT], bN[RPT], bP[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int i = il[r];
      const double vxi = c_tab.vx[i][q], vyi = c_tab.vy[i][q], sci = c_tab.sc[i][q];
      // G v_i (geometry only)
      const double gvx = g.g11 * vxi + g.g12 * vyi, gvy = g.g12 * vxi + g.g22 * vyi;
      aNx[r] = aNy[r] = aPx[r] = aPy[r] = bN[r] = bP[r] = T::zero();
      if (FORM == FORM_A) {
        // G_eps v_i
        S gex, gey;
        if (EPSKIND == FW_EPS_VEC_P0) {
          gex = g.ge11 * vxi + g.ge12 * vyi, gey = g.ge12 * vxi + g.ge22 * vyi;
        } else {
          gex = ezq * gvx, gey = ezq * gvy;
        }
        if (KI == 0) {
          // test v_t: trial N -> curl curl / (mu_z k0^2) - eps_t e_t.v_t ; trial P -> grad e_z . v_t / mu_t
          aNx[r] = gex * (-wf), aNy[r] = gey * (-wf);
          bN[r] = T::from(wf * g.c_curl * sci, 0.0);
          const double c = wrf * imu;
          aPx[r] = T::from(c * gvx, 0.0), aPy[r] = T::from(c * gvy, 0.0);
        } else {
          // test v_z: trial N -> eps_t e_t . grad v_z ; trial P -> -eps_z e_z v_z k0^2
          aNx[r] = gex * wf, aNy[r] = gey * wf;
          bP[r] = ezq * (-wrf * k0 * k0 * sci);
        }
      } else if (FORM == FORM_B) {
        const double c = -wrf * imu * ik02;
        aNx[r] = T::from(c * gvx, 0.0), aNy[r] = T::from(c * gvy, 0.0);
      } else if (FORM == FORM_MASS) {
        if (KI == 0) {
          aNx[r] = T::from(w * gvx, 0.0), aNy[r] = T::from(w * gvy, 0.0);
        } else {
          bP[r] = T::from(w * sci, 0.0);
        }
      } else if (FORM == FORM_C1) {
        // i*beta * (e_x v_y - e_y v_x) = i*beta * idet * (u_hat x v_hat)
        if (KI == 0) aNx[r] = T::from(w * g.idet * vyi, 0.0), aNy[r] = T::from(-w * g.idet * vxi, 0.0);
      } else {  // FORM_C0
        if (KI == 0) {
          // d_y e_z v_x - d_x e_z v_y = -idet * (g_hat x v_hat)
          aPx[r] = T::from(-w * g.idet * vyi, 0.0), aPy[r] = T::from(w * g.idet * vxi, 0.0);
        } else {
          // curl(e_t) v_z = idet * c_hat_j * psi_i
          bN[r] = T::from(w * g.idet * sci, 0.0);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      // one table fetch feeds RPT accumulators
      const double tvx = c_tab.vx[j][q], tvy = c_tab.vy[j][q], tsc = c_tab.sc[j][q];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        if (kind_of<NB>(j) == 0) {
          if (useAN) {
            fma_(acc[r][j], aNx[r], tvx);
            fma_(acc[r][j], aNy[r], tvy);
          }
          if (useBN) fma_(acc[r][j], bN[r], tsc);
        } else {
          if (useAP) {
            fma_(acc[r][j], aPx[r], tvx);
            fma_(acc[r][j], aPy[r], tvy);
          }
          if (useBP) fma_(acc[r][j], bP[r], tsc);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    if (irow[r] >= NROWS) continue;
    if (FORM == FORM_B) {
      // compact NT x NT block: entry (rank(i), rank(j))
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (kind_of<NB>(j) == 0) out[(size_t)(irow[r] * NT + nrank_of<NB>(j)) * ne + e] = acc[r][j];
    } else {
      const int i = il[r];
#pragma unroll
      for (int j = 0; j < NB; ++j) out[(size_t)(i * NB + j) * ne + e] = acc[r][j];
    }
  }
}

// Constant-coefficient fast path of aform + bform: straight waveguide (radius = inf) and element-wise constant
// permittivity (P0 or diagonal-anisotropic P0).  The weight of every term is then constant over the triangle and
// the quadrature sums collapse to reference constants (c_pair):
//     sum_q w_q (G v_i).v_j = |det| (G11 T11_ij + G12 T12_ij + G22 T22_ij) =: m_ij
//   (N,N): A = c_curl |det| CC_ij - m^eps_ij        B = -m_ij / (mu k0^2)
//   (N,P): A = m_ij / mu     (P,N): A = m^eps_ij    (P,P): A = -eps_z k0^2 |det| CC_ij
// ~5 DFMA per entry instead of 2-3 per entry AND quadrature point: the kernel is purely store bound.  No
// accumulators live across a loop, so a thread can write RPT rows of the block (grid.y = NB / RPT) with a
// handful of registers; A and B (compact Nedelec block) are written by the same thread.
template <int NB, class S, int EPSKIND, int RPT>
__global__ void __launch_bounds__(128) k_element_blocks_const(int ne, int nv, const double* __restrict__ p,
                                                              const int* __restrict__ t, const S* __restrict__ eps,
                                                              double k0, double mu_r, S* __restrict__ outA,
                                                              S* __restrict__ outB) {
  constexpr int NT = (NB == 6) ? 3 : 8;
  using T = scalar_traits<S>;
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= ne) return;
  const double ik02 = 1.0 / (k0 * k0);
  const double imu = 1.0 / mu_r;
  const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
  const double x0 = p[v0], y0 = p[nv + v0];
  const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
  const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
  const double det = a11 * a22 - a12 * a21;
  const double adet = fabs(det);
  const double id2a = adet / (det * det);
  // |det| G  (G = A^-1 A^-T)
  const double h11 = id2a * (a22 * a22 + a12 * a12);
  const double h12 = -id2a * (a22 * a21 + a12 * a11);
  const double h22 = id2a * (a21 * a21 + a11 * a11);
  const double ccurl = imu * ik02 * id2a;  // |det| / (mu k0^2 det^2)
  S ex, ey, ez;
  if (EPSKIND == FW_EPS_P0) {
    ex = ey = ez = eps[e];
  } else {
    ex = eps[3 * (size_t)e], ey = eps[3 * (size_t)e + 1], ez = eps[3 * (size_t)e + 2];
  }
  // |det| G_eps = |det| A^-1 diag(eps_x, eps_y) A^-T
  S he11, he12, he22;
  if (EPSKIND == FW_EPS_VEC_P0) {
    he11 = (ex * (a22 * a22) + ey * (a12 * a12)) * id2a;
    he12 = (ex * (a22 * a21) + ey * (a12 * a11)) * (-id2a);
    he22 = (ex * (a21 * a21) + ey * (a11 * a11)) * id2a;
  } else {
    he11 = ex * h11, he12 = ex * h12, he22 = ex * h22;
  }
  const S ezk = ez * (-k0 * k0 * adet);
  const double cb = -imu * ik02;
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int i = blockIdx.y * RPT + r;  // block-uniform local test dof
    if (i >= NB) break;
    const bool rowN = c_tab.kind[i] == 0;
    const int ri = c_tab.nrank[i];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const double t11 = c_pair.T11[i][j], t12 = c_pair.T12[i][j], t22 = c_pair.T22[i][j], cc = c_pair.CC[i][j];
      S a;
      if (rowN) {
        const double m = h11 * t11 + h12 * t12 + h22 * t22;
        if (kind_of<NB>(j) == 0) {
          const S me = he11 * t11 + he12 * t12 + he22 * t22;
          a = T::from(ccurl * cc, 0.0) - me;
          outB[(size_t)(ri * NT + nrank_of<NB>(j)) * ne + e] = T::from(cb * m, 0.0);
        } else {
          a = T::from(imu * m, 0.0);
        }
      } else {
        if (kind_of<NB>(j) == 0)
          a = he11 * t11 + he12 * t12 + he22 * t22;
        else
          a = ezk * cc;
      }
      outA[(size_t)(i * NB + j) * ne + e] = a;
    }
  }
}

template <int NB, class S, int EPSKIND, int RPT>
static void launch_const_rpt(Handle& h, const S* eps_dev, double k0, double mu_r, S* bA, S* bB) {
  const Space& sp = h.sp;
  k_element_blocks_const<NB, S, EPSKIND, RPT><<<dim3(cdiv(sp.ne, 128), (NB + RPT - 1) / RPT), 128, 0, h.stream>>>(
      sp.ne, sp.nv, sp.p.p, sp.t.p, eps_dev, k0, mu_r, bA, bB);
  FW_CHECK_CUDA(cudaGetLastError());
}

template <int NB, class S, int EPSKIND>
static void launch_const(Handle& h, const S* eps_dev, double k0, double mu_r, S* bA, S* bB) {
  static const int rpt = getenv("FW_K1_RPT") ? atoi(getenv("FW_K1_RPT")) : NB / 2;  // measured: 1: 0.200, 2: 0.199, 7: 0.186, 14: 0.19 ms
  switch (rpt) {
    case 1:
      launch_const_rpt<NB, S, EPSKIND, 1>(h, eps_dev, k0, mu_r, bA, bB);
      break;
    case NB / 2:
      launch_const_rpt<NB, S, EPSKIND, NB / 2>(h, eps_dev, k0, mu_r, bA, bB);
      break;
    case NB:
      launch_const_rpt<NB, S, EPSKIND, NB>(h, eps_dev, k0, mu_r, bA, bB);
      break;
    default:
      launch_const_rpt<NB, S, EPSKIND, 2>(h, eps_dev, k0, mu_r, bA, bB);
  }
}

template <int NB, int NQ, class S, int EPSKIND, int FORM>
static void launch_form(Handle& h, const S* eps_dev, double k0, double mu_r, double inv_radius, S* blocks) {
  const Space& sp = h.sp;
  constexpr int NT = (NB == 6) ? 3 : 8;
  // rows computed simultaneously per thread.  Measured on B200 (500k triangles, f64): 1 row/thread 0.335 ms
  // (76 registers), 4+3 rows/thread 0.358 ms (236 registers), 2 triangles/thread 0.36 ms: the kernel is
  // issue-latency bound (FP64 pipe 43-65 % busy), not bound by constant loads or by HBM stores (a pure
  // store kernel with the same pattern reaches 6.3-7.0 TB/s, scripts/microbench/store_bw.cu).
  constexpr int RN = 1;
  constexpr int RP = 1;
  dim3 block(128);
  const int gx = cdiv(sp.ne, 128);
  // one launch per test-function kind (bform has Nedelec rows only)
  k_element_blocks<NB, NQ, S, EPSKIND, FORM, 0, RN><<<dim3(gx, (NT + RN - 1) / RN), block, 0, h.stream>>>(
      sp.ne, sp.nv, sp.p.p, sp.t.p, eps_dev, k0, mu_r, inv_radius, blocks);
  if (FORM != FORM_B)
    k_element_blocks<NB, NQ, S, EPSKIND, FORM, 1, RP><<<dim3(gx, (NB - NT + RP - 1) / RP), block, 0, h.stream>>>(
        sp.ne, sp.nv, sp.p.p, sp.t.p, eps_dev, k0, mu_r, inv_radius, blocks);
}

template <int NB, class S, int EPSKIND, int FORM>
static void dispatch_nq(Handle& h, const S* eps_dev, double k0, double mu_r, double inv_radius, S* blocks) {
  switch (h.sp.nq) {
    case 3:
      launch_form<NB, 3, S, EPSKIND, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
      break;
    case 6:
      launch_form<NB, 6, S, EPSKIND, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
      break;
    case 7:
      launch_form<NB, 7, S, EPSKIND, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
      break;
    case 12:
      launch_form<NB, 12, S, EPSKIND, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
      break;
    default:
      launch_form<NB, 0, S, EPSKIND, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
  }
  FW_CHECK_CUDA(cudaGetLastError());
}

template <int NB, class S, int FORM>
static void dispatch_eps(Handle& h, int eps_kind, const S* eps_dev, double k0, double mu_r, double inv_radius,
                         S* blocks) {
  switch (eps_kind) {
    case FW_EPS_P0:
      dispatch_nq<NB, S, FW_EPS_P0, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
      break;
    case FW_EPS_DG_P1:
      dispatch_nq<NB, S, FW_EPS_DG_P1, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
      break;
    case FW_EPS_VEC_P0:
      dispatch_nq<NB, S, FW_EPS_VEC_P0, FORM>(h, eps_dev, k0, mu_r, inv_radius, blocks);
      break;
    default:
      throw Error(FW_ERR_INVALID, "unknown eps_kind");
  }
}

// Element subset of the basis (scikit-fem CellBasis(elements=...), kept by with_element, waveguide.py:574): the forms
// are assembled over the selected elements only.  The local blocks of the others are zeroed before the reduce, so
// their contributions vanish from the values and -- being exact zeros -- from the exported pattern.
template <class S>
__global__ void __launch_bounds__(256) k_zero_masked_blocks(int ne, int nplanes, const uint8_t* __restrict__ emask,
                                                            S* __restrict__ blocks) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || emask[e]) return;
  for (int pl = blockIdx.y; pl < nplanes; pl += gridDim.y) blocks[(size_t)pl * ne + e] = scalar_traits<S>::zero();
}
template <class S>
static void apply_element_mask(Handle& h, S* blocks, int nplanes) {
  const Space& sp = h.sp;
  if (!sp.has_emask) return;
  k_zero_masked_blocks<S><<<dim3(cdiv(sp.ne, 256), std::min(nplanes, 64)), 256, 0, h.stream>>>(sp.ne, nplanes, sp.emask.p,
                                                                                             blocks);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 1;
}

static void make_b_lut(const Space& sp, std::vector<int>& lut) {
  lut.assign((size_t)sp.nb * sp.nb, -1);
  std::vector<int> rank(sp.nb, -1);
  int r = 0;
  for (int l = 0; l < sp.nb; ++l)
    if (sp.kind[l] == 0) rank[l] = r++;
  for (int i = 0; i < sp.nb; ++i)
    for (int j = 0; j < sp.nb; ++j)
      if (rank[i] >= 0 && rank[j] >= 0) lut[(size_t)i * sp.nb + j] = rank[i] * sp.nt + rank[j];
}

template <class S>
static void assemble_ab_typed(Handle& h, int eps_kind, const void* eps_host, double k0, double mu_r, double radius) {
  Space& sp = h.sp;
  cudaStream_t s = h.stream;
  const size_t w = sizeof(S) / sizeof(double);
  const size_t neps = (eps_kind == FW_EPS_P0) ? (size_t)sp.ne : (size_t)3 * sp.ne;
  DevBuf<double> eps_dev;
  {
    Timer tu(s);
    eps_dev.upload((const double*)eps_host, neps * w, s);
    h.tim.upload_ms += tu.stop();
  }
  const size_t nA = (size_t)sp.nb * sp.nb * sp.ne, nB = (size_t)sp.nt * sp.nt * sp.ne;
  h.blocks.alloc((nA + nB) * w);
  S* bA = (S*)h.blocks.p;
  S* bB = bA + nA;
  const double inv_radius = std::isinf(radius) ? 0.0 : 1.0 / radius;
  upload_tables(h);
  {
    Timer tk(s, "fw:element kernel K1");
    static const bool no_fast = getenv("FW_K1_GENERIC") != nullptr;
    const bool fast = inv_radius == 0.0 && eps_kind != FW_EPS_DG_P1 && !no_fast;
    h.tim.element_kernel_launches = fast ? 1 : 3;
    if (fast) {
      // one launch writes the aform blocks and the compact bform blocks
      const S* ed = (const S*)eps_dev.p;
      if (sp.nb == 6) {
        if (eps_kind == FW_EPS_P0)
          launch_const<6, S, FW_EPS_P0>(h, ed, k0, mu_r, bA, bB);
        else
          launch_const<6, S, FW_EPS_VEC_P0>(h, ed, k0, mu_r, bA, bB);
      } else {
        if (eps_kind == FW_EPS_P0)
          launch_const<14, S, FW_EPS_P0>(h, ed, k0, mu_r, bA, bB);
        else
          launch_const<14, S, FW_EPS_VEC_P0>(h, ed, k0, mu_r, bA, bB);
      }
    } else if (sp.nb == 6) {
      dispatch_eps<6, S, FORM_A>(h, eps_kind, (const S*)eps_dev.p, k0, mu_r, inv_radius, bA);
      dispatch_nq<6, S, FW_EPS_P0, FORM_B>(h, (const S*)eps_dev.p, k0, mu_r, inv_radius, bB);
    } else {
      dispatch_eps<14, S, FORM_A>(h, eps_kind, (const S*)eps_dev.p, k0, mu_r, inv_radius, bA);
      dispatch_nq<14, S, FW_EPS_P0, FORM_B>(h, (const S*)eps_dev.p, k0, mu_r, inv_radius, bB);
    }
    h.tim.element_kernel_ms = tk.stop();
    // generic path: aform Nedelec rows + Lagrange rows, bform Nedelec rows
    h.tim.kernel_launches += h.tim.element_kernel_launches;
  }
  apply_element_mask<S>(h, bA, sp.nb * sp.nb);
  apply_element_mask<S>(h, bB, sp.nt * sp.nt);
  // segmented reduce onto the structural pattern
  std::vector<int> lut;
  make_b_lut(sp, lut);
  DevBuf<int> lut_dev;
  lut_dev.upload(lut.data(), lut.size(), s);
  build_slot_list(h, h.pat, lut_dev.p, sp.ne);  // once per symbolic pattern
  const Pattern& pat = h.pat;
  for (Values* v : {&h.A, &h.B}) {
    v->is_complex = scalar_traits<S>::is_complex;
    v->v.alloc((size_t)pat.nnz * w);
    v->flag.alloc((size_t)pat.nnz);
    v->nnz_kept = -1;
  }
  {
    Timer tr(s, "fw:segmented reduce K2b");
    reduce_onto_pattern<S>(pat, bA, nullptr, 0, sp.ne, (S*)h.A.v.p, h.A.flag.p, s);
    reduce_onto_pattern<S>(pat, bB, lut_dev.p, (int)lut.size(), sp.ne, (S*)h.B.v.p, h.B.flag.p, s, pat.tt_slots.p,
                           pat.n_tt);
    h.tim.reduce_ms = tr.stop();
    h.tim.kernel_launches += 2;
  }
  h.A.valid = h.B.valid = true;
  // compact copy of B (29 M of the 79.5 M slots at config 2) for the M x products of the Arnoldi iteration
  h.Bc.nnz = compact_csr_dev(h, pat, h.B, 2, h.Bc.indptr, h.Bc.indices, h.Bc.vals);
  h.Bc.valid = true;
  h.tim.kernel_launches += 4;
}

void assemble_waveguide(Handle& h, int eps_kind, bool eps_complex, const void* eps_host, double k0, double mu_r,
                        double radius) {
  FW_REQUIRE(h.sp.valid, "fw_set_space must be called first");
  FW_REQUIRE(h.has_symbolic, "fw_symbolic must be called first");
  FW_REQUIRE(eps_host != nullptr, "eps is NULL");
  if (eps_complex)
    assemble_ab_typed<cplx>(h, eps_kind, eps_host, k0, mu_r, radius);
  else
    assemble_ab_typed<double>(h, eps_kind, eps_host, k0, mu_r, radius);
}

// order 2 local layout P P P (N N P)x3 N N: the Nedelec pairs are the locals (3,4) (6,7) (9,10) (12,13)
__global__ void k_mass_pairs(int ne, const int* __restrict__ edofs, int* __restrict__ pair) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int la[4] = {3, 6, 9, 12};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int a = edofs[(size_t)la[q] * ne + e], b = edofs[(size_t)(la[q] + 1) * ne + e];
    pair[a] = b;  // every element sharing the edge writes the same values
    pair[b] = a;
  }
}

// ---- split of the compact mass matrix into its Nedelec and Lagrange diagonal blocks
__global__ void k_flag_z_u32(const uint8_t* __restrict__ is_z, int n, uint32_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = is_z[i] ? 1u : 0u;
}
// rank[d] = index of dof d inside its own block; idxN / idxP = the inverse maps
__global__ void k_block_index(const uint8_t* __restrict__ is_z, const uint32_t* __restrict__ zscan, int n,
                              int* __restrict__ rank, int* __restrict__ idxN, int* __restrict__ idxP) {
  int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const int rp = (int)zscan[d];
  if (is_z[d]) {
    rank[d] = rp;
    idxP[rp] = d;
  } else {
    rank[d] = d - rp;
    idxN[d - rp] = d;
  }
}
__global__ void k_block_rowlen(const int* __restrict__ idx, int nsub, const int* __restrict__ indptr,
                               uint32_t* __restrict__ len) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsub) len[i] = (uint32_t)(indptr[idx[i] + 1] - indptr[idx[i]]);
  if (i == nsub) len[i] = 0;
}
__global__ void k_block_fill(const int* __restrict__ idx, int nsub, const int* __restrict__ indptr,
                             const int* __restrict__ indices, const double* __restrict__ vals,
                             const int* __restrict__ rank, const uint8_t* __restrict__ is_z,
                             const int* __restrict__ sub_indptr, int* __restrict__ sub_indices,
                             double* __restrict__ sub_vals, const int* __restrict__ pair, int* __restrict__ sub_pair,
                             int* __restrict__ cross) {
  constexpr int TPR = 8;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(gt / TPR), l = (int)(gt % TPR);
  if (i >= nsub) return;
  const int d = idx[i];
  const int a = indptr[d], b = indptr[d + 1], o = sub_indptr[i];
  for (int e = a + l; e < b; e += TPR) {
    const int c = indices[e];
    if (is_z[c] != is_z[d]) atomicExch(cross, 1);  // cannot happen for a block-diagonal mass matrix
    sub_indices[o + (e - a)] = rank[c];
    sub_vals[o + (e - a)] = vals[e];
  }
  if (l == 0 && sub_pair) sub_pair[i] = (pair && pair[d] >= 0) ? rank[pair[d]] : -1;
}

static void build_mass_blocks(Handle& h) {
  const Space& sp = h.sp;
  cudaStream_t s = h.stream;
  const int n = sp.n;
  h.mass_split = false;
  if (getenv("FW_NO_MASS_SPLIT")) return;
  DevBuf<uint32_t> zs;
  DevBuf<int> rank, cross;
  zs.alloc((size_t)n + 1), rank.alloc((size_t)n), cross.alloc(1);
  cross.zero(s);
  k_flag_z_u32<<<cdiv(n, 256), 256, 0, s>>>(sp.is_z.p, n, zs.p);
  uint32_t last_flag = 0, last_scan = 0;
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_flag, zs.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
  exclusive_scan_u32(zs.p, zs.p, n, h.scan_tmp, s);
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_scan, zs.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  const int nP = (int)(last_scan + last_flag), nN = n - nP;
  if (nP == 0 || nN == 0) return;
  h.MassN.idx.alloc((size_t)nN), h.MassP.idx.alloc((size_t)nP);
  k_block_index<<<cdiv(n, 256), 256, 0, s>>>(sp.is_z.p, zs.p, n, rank.p, h.MassN.idx.p, h.MassP.idx.p);
  const int* pair = h.MassC.pair.n ? h.MassC.pair.p : nullptr;
  for (int blk = 0; blk < 2; ++blk) {
    Handle::MassBlock& B = blk ? h.MassP : h.MassN;
    B.n = blk ? nP : nN;
    DevBuf<uint32_t> len;
    len.alloc((size_t)B.n + 1);
    k_block_rowlen<<<cdiv(B.n + 1, 256), 256, 0, s>>>(B.idx.p, B.n, h.MassC.indptr.p, len.p);
    exclusive_scan_u32(len.p, len.p, (int64_t)B.n + 1, h.scan_tmp, s);
    uint32_t tot = 0;
    FW_CHECK_CUDA(cudaMemcpyAsync(&tot, len.p + B.n, 4, cudaMemcpyDeviceToHost, s));
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    B.nnz = tot;
    B.indptr.alloc((size_t)B.n + 1);
    FW_CHECK_CUDA(cudaMemcpyAsync(B.indptr.p, len.p, ((size_t)B.n + 1) * 4, cudaMemcpyDeviceToDevice, s));
    B.indices.alloc((size_t)std::max<int64_t>(1, B.nnz)), B.vals.alloc((size_t)std::max<int64_t>(1, B.nnz));
    const bool with_pair = !blk && pair;
    if (with_pair)
      B.pair.alloc((size_t)B.n);
    else
      B.pair.release();
    k_block_fill<<<cdiv((int64_t)B.n * 8, 256), 256, 0, s>>>(B.idx.p, B.n, h.MassC.indptr.p, h.MassC.indices.p,
                                                            h.MassC.vals.p, rank.p, sp.is_z.p, B.indptr.p, B.indices.p,
                                                            B.vals.p, pair, with_pair ? B.pair.p : nullptr, cross.p);
  }
  int hcross = 0;
  cross.download(&hcross, 1, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  FW_CHECK_CUDA(cudaGetLastError());
  h.mass_split = hcross == 0;
  h.tim.kernel_launches += 12;
}

// geometry-only operators of calculate_hfield on the structural pattern (always f64)
void assemble_aux(Handle& h) {
  FW_REQUIRE(h.sp.valid && h.has_symbolic, "space/symbolic missing");
  if (h.Mass.valid && h.C0.valid && h.C1.valid) return;
  Space& sp = h.sp;
  cudaStream_t s = h.stream;
  const size_t nA = (size_t)sp.nb * sp.nb * sp.ne;
  h.blocks.alloc(nA);
  upload_tables(h);
  const Pattern& pat = h.pat;
  Values* vals[3] = {&h.Mass, &h.C0, &h.C1};
  for (int f = 0; f < 3; ++f) {
    double* b = h.blocks.p;
    if (sp.nb == 6) {
      if (f == 0)
        dispatch_nq<6, double, FW_EPS_P0, FORM_MASS>(h, nullptr, 1.0, 1.0, 0.0, b);
      else if (f == 1)
        dispatch_nq<6, double, FW_EPS_P0, FORM_C0>(h, nullptr, 1.0, 1.0, 0.0, b);
      else
        dispatch_nq<6, double, FW_EPS_P0, FORM_C1>(h, nullptr, 1.0, 1.0, 0.0, b);
    } else {
      if (f == 0)
        dispatch_nq<14, double, FW_EPS_P0, FORM_MASS>(h, nullptr, 1.0, 1.0, 0.0, b);
      else if (f == 1)
        dispatch_nq<14, double, FW_EPS_P0, FORM_C0>(h, nullptr, 1.0, 1.0, 0.0, b);
      else
        dispatch_nq<14, double, FW_EPS_P0, FORM_C1>(h, nullptr, 1.0, 1.0, 0.0, b);
    }
    apply_element_mask<double>(h, b, sp.nb * sp.nb);
    Values& v = *vals[f];
    v.is_complex = false;
    v.v.alloc((size_t)pat.nnz);
    v.flag.alloc((size_t)pat.nnz);
    v.nnz_kept = -1;
    reduce_onto_pattern<double>(pat, b, nullptr, 0, sp.ne, v.v.p, v.flag.p, s);
    v.valid = true;
    h.tim.kernel_launches += 3;
  }
  // compact copy of the mass matrix (the cross blocks are structural zeros on the shared pattern)
  h.MassC.nnz = compact_csr_dev(h, pat, h.Mass, 2, h.MassC.indptr, h.MassC.indices, h.MassC.vals);
  // 2 x 2 blocks of the block-Jacobi preconditioner of the H-field PCG: consecutive Nedelec functions of one entity
  if (sp.nb == 14) {
    h.MassC.pair.alloc((size_t)sp.n);
    FW_CHECK_CUDA(cudaMemsetAsync(h.MassC.pair.p, 0xff, (size_t)sp.n * sizeof(int), s));
    k_mass_pairs<<<cdiv(sp.ne, 256), 256, 0, s>>>(sp.ne, sp.edofs.p, h.MassC.pair.p);
    h.tim.kernel_launches += 1;
  } else {
    h.MassC.pair.release();
  }
  build_mass_blocks(h);
  h.MassC.valid = true;
  h.tim.kernel_launches += 5;
}

}  // namespace fw
