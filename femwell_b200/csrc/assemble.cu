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
// tensors  G = A^-1 A^-T,  G_eps = A^-1 diag(eps_x, eps_y) A^-T  (A = affine Jacobian).  One thread
// owns one triangle, keeps one row of the local block in registers, reads the reference tables as
// warp-uniform constant-memory operands, and writes the row in SoA layout out[(i*NB+j)*Ne + e], so
// every store instruction of a warp is one contiguous 256 B (f64) / 512 B (c128) segment.
#include "handle.h"

namespace fw {

struct Tables {
  int nb, nq;
  int kind[MAXNB];
  double vx[MAXNB][MAXQ], vy[MAXNB][MAXQ], sc[MAXNB][MAXQ];
  double X[MAXQ], Y[MAXQ], W[MAXQ];
};
__constant__ Tables c_tab;

void upload_tables(Handle& h) {
  const Space& sp = h.sp;
  Tables t;
  memset(&t, 0, sizeof(t));
  t.nb = sp.nb;
  t.nq = sp.nq;
  for (int l = 0; l < sp.nb; ++l) {
    t.kind[l] = sp.kind[l];
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
  // synchronous copy from a stack object: the stream is idle-ordered afterwards
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_tab, &t, sizeof(t)));
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

template <class S>
struct Vec2 {
  S x, y;
};

template <int NB, class S, int EPSKIND, int FORM>
__global__ void __launch_bounds__(128) k_element_blocks(int ne, int nv, const double* __restrict__ p,
                                                        const int* __restrict__ t, const S* __restrict__ eps,
                                                        double k0, double mu_r, double inv_radius,
                                                        S* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  constexpr int NT = (NB == 6) ? 3 : 8;
  using T = scalar_traits<S>;
  const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
  const double x0 = p[v0], y0 = p[nv + v0];
  const double a11 = p[v1] - x0, a21 = p[nv + v1] - y0;
  const double a12 = p[v2] - x0, a22 = p[nv + v2] - y0;
  const double det = a11 * a22 - a12 * a21;
  const double idet = 1.0 / det;
  const double adet = fabs(det);
  const double id2 = idet * idet;
  // G = A^-1 A^-T (symmetric)
  const double g11 = id2 * (a22 * a22 + a12 * a12);
  const double g12 = -id2 * (a22 * a21 + a12 * a11);
  const double g22 = id2 * (a21 * a21 + a11 * a11);
  // epsilon degrees of freedom of this element
  S e0 = T::zero(), e1 = T::zero(), e2 = T::zero();
  if (FORM == FORM_A) {
    if (EPSKIND == FW_EPS_P0) {
      e0 = eps[e];
    } else {
      e0 = eps[3 * (size_t)e], e1 = eps[3 * (size_t)e + 1], e2 = eps[3 * (size_t)e + 2];
    }
  }
  const double ik02 = 1.0 / (k0 * k0);
  const double imu = 1.0 / mu_r;
  const int nq = c_tab.nq;

  for (int i = 0; i < NB; ++i) {
    const int ki = c_tab.kind[i];
    if (FORM == FORM_B && ki != 0) continue;  // only the tt block exists
    S acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = T::zero();
    for (int q = 0; q < nq; ++q) {
      const double Xq = c_tab.X[q], Yq = c_tab.Y[q];
      const double w = c_tab.W[q] * adet;
      const double f = 1.0 + (x0 + a11 * Xq + a12 * Yq) * inv_radius;  // bend factor (waveguide.py:580-586)
      const double vxi = c_tab.vx[i][q], vyi = c_tab.vy[i][q], sci = c_tab.sc[i][q];
      // G v_i (geometry only)
      const double gvx = g11 * vxi + g12 * vyi, gvy = g12 * vxi + g22 * vyi;
      Vec2<S> aN{T::zero(), T::zero()}, aP{T::zero(), T::zero()};
      S bN = T::zero(), bP = T::zero();
      if (FORM == FORM_A) {
        // epsilon at the point
        S ex, ey, ez;
        if (EPSKIND == FW_EPS_P0) {
          ex = ey = ez = e0;
        } else if (EPSKIND == FW_EPS_DG_P1) {
          ex = e0 * (1.0 - Xq - Yq) + e1 * Xq + e2 * Yq;
          ey = ez = ex;
        } else {
          ex = e0, ey = e1, ez = e2;
        }
        // G_eps v_i = A^-1 diag(ex,ey) A^-T v_i ;  A^-T v = idet*(a22 vx - a21 vy, -a12 vx + a11 vy)
        const double tx = idet * (a22 * vxi - a21 * vyi), ty = idet * (-a12 * vxi + a11 * vyi);
        const S sx = ex * tx, sy = ey * ty;
        // A^-1 s = idet*(a22 sx - a12 sy, -a21 sx + a11 sy)
        const S gex = (sx * a22 - sy * a12) * idet, gey = (sy * a11 - sx * a21) * idet;
        if (ki == 0) {
          // test v_t: trial N -> curl curl / (mu_z k0^2) - eps_t e_t.v_t ; trial P -> grad e_z . v_t / mu_t
          aN.x = gex * (-w * f), aN.y = gey * (-w * f);
          bN = T::from(w * f * imu * ik02 * id2 * sci, 0.0);
          aP.x = T::from(w * imu / f * gvx, 0.0), aP.y = T::from(w * imu / f * gvy, 0.0);
        } else {
          // test v_z: trial N -> eps_t e_t . grad v_z ; trial P -> -eps_z e_z v_z k0^2
          aN.x = gex * (w * f), aN.y = gey * (w * f);
          bP = ez * (-w / f * k0 * k0 * sci);
        }
      } else if (FORM == FORM_B) {
        const double c = -w * imu / f * ik02;
        aN.x = T::from(c * gvx, 0.0), aN.y = T::from(c * gvy, 0.0);
      } else if (FORM == FORM_MASS) {
        if (ki == 0) {
          aN.x = T::from(w * gvx, 0.0), aN.y = T::from(w * gvy, 0.0);
        } else {
          bP = T::from(w * sci, 0.0);
        }
      } else if (FORM == FORM_C1) {
        // i*beta * (e_x v_y - e_y v_x) = i*beta * idet * (u_hat x v_hat)
        if (ki == 0) aN.x = T::from(w * idet * vyi, 0.0), aN.y = T::from(-w * idet * vxi, 0.0);
      } else {  // FORM_C0
        if (ki == 0) {
          // d_y e_z v_x - d_x e_z v_y = -idet * (g_hat x v_hat)
          aP.x = T::from(-w * idet * vyi, 0.0), aP.y = T::from(w * idet * vxi, 0.0);
        } else {
          // curl(e_t) v_z = idet * c_hat_j * psi_i
          bN = T::from(w * idet * sci, 0.0);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (kind_of<NB>(j) == 0) {
          fma_(acc[j], aN.x, c_tab.vx[j][q]);
          fma_(acc[j], aN.y, c_tab.vy[j][q]);
          fma_(acc[j], bN, c_tab.sc[j][q]);
        } else {
          fma_(acc[j], aP.x, c_tab.vx[j][q]);
          fma_(acc[j], aP.y, c_tab.vy[j][q]);
          fma_(acc[j], bP, c_tab.sc[j][q]);
        }
      }
    }
    if (FORM == FORM_B) {
      // compact NT x NT block: entry (rank(i), rank(j))
      int ri = 0;
      for (int m = 0; m < i; ++m) ri += (c_tab.kind[m] == 0);
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (kind_of<NB>(j) == 0) out[(size_t)(ri * NT + nrank_of<NB>(j)) * ne + e] = acc[j];
    } else {
#pragma unroll
      for (int j = 0; j < NB; ++j) out[(size_t)(i * NB + j) * ne + e] = acc[j];
    }
  }
}

template <int NB, class S, int EPSKIND>
static void launch_ab(Handle& h, const S* eps_dev, double k0, double mu_r, double inv_radius, S* blocksA,
                      S* blocksB) {
  const Space& sp = h.sp;
  dim3 grid(cdiv(sp.ne, 128)), block(128);
  k_element_blocks<NB, S, EPSKIND, FORM_A>
      <<<grid, block, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, eps_dev, k0, mu_r, inv_radius, blocksA);
  k_element_blocks<NB, S, EPSKIND, FORM_B>
      <<<grid, block, 0, h.stream>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, eps_dev, k0, mu_r, inv_radius, blocksB);
  FW_CHECK_CUDA(cudaGetLastError());
}

template <int NB, class S>
static void dispatch_eps(Handle& h, int eps_kind, const S* eps_dev, double k0, double mu_r, double inv_radius,
                         S* bA, S* bB) {
  switch (eps_kind) {
    case FW_EPS_P0:
      launch_ab<NB, S, FW_EPS_P0>(h, eps_dev, k0, mu_r, inv_radius, bA, bB);
      break;
    case FW_EPS_DG_P1:
      launch_ab<NB, S, FW_EPS_DG_P1>(h, eps_dev, k0, mu_r, inv_radius, bA, bB);
      break;
    case FW_EPS_VEC_P0:
      launch_ab<NB, S, FW_EPS_VEC_P0>(h, eps_dev, k0, mu_r, inv_radius, bA, bB);
      break;
    default:
      throw Error(FW_ERR_INVALID, "unknown eps_kind");
  }
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
    Timer tk(s);
    if (sp.nb == 6)
      dispatch_eps<6, S>(h, eps_kind, (const S*)eps_dev.p, k0, mu_r, inv_radius, bA, bB);
    else
      dispatch_eps<14, S>(h, eps_kind, (const S*)eps_dev.p, k0, mu_r, inv_radius, bA, bB);
    h.tim.element_kernel_ms = tk.stop();
    h.tim.element_kernel_launches = 2;
    h.tim.kernel_launches += 2;
  }
  // segmented reduce onto the structural pattern
  std::vector<int> lut;
  make_b_lut(sp, lut);
  DevBuf<int> lut_dev;
  lut_dev.upload(lut.data(), lut.size(), s);
  const Pattern& pat = h.pat;
  for (Values* v : {&h.A, &h.B}) {
    v->is_complex = scalar_traits<S>::is_complex;
    v->v.alloc((size_t)pat.nnz * w);
    v->flag.alloc((size_t)pat.nnz);
    v->nnz_kept = -1;
  }
  {
    Timer tr(s);
    reduce_onto_pattern<S>(pat, bA, nullptr, 0, sp.ne, (S*)h.A.v.p, h.A.flag.p, s);
    reduce_onto_pattern<S>(pat, bB, lut_dev.p, (int)lut.size(), sp.ne, (S*)h.B.v.p, h.B.flag.p, s);
    h.tim.reduce_ms = tr.stop();
    h.tim.kernel_launches += 2;
  }
  h.A.valid = h.B.valid = true;
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

// geometry-only operators of calculate_hfield on the structural pattern (always f64)
void assemble_aux(Handle& h) {
  FW_REQUIRE(h.sp.valid && h.has_symbolic, "space/symbolic missing");
  if (h.Mass.valid && h.C0.valid && h.C1.valid) return;
  Space& sp = h.sp;
  cudaStream_t s = h.stream;
  const size_t nA = (size_t)sp.nb * sp.nb * sp.ne;
  h.blocks.alloc(nA);
  upload_tables(h);
  dim3 grid(cdiv(sp.ne, 128)), block(128);
  const Pattern& pat = h.pat;
  Values* vals[3] = {&h.Mass, &h.C0, &h.C1};
  for (int f = 0; f < 3; ++f) {
    double* b = h.blocks.p;
    if (sp.nb == 6) {
      if (f == 0)
        k_element_blocks<6, double, FW_EPS_P0, FORM_MASS><<<grid, block, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, nullptr, 1.0, 1.0, 0.0, b);
      else if (f == 1)
        k_element_blocks<6, double, FW_EPS_P0, FORM_C0><<<grid, block, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, nullptr, 1.0, 1.0, 0.0, b);
      else
        k_element_blocks<6, double, FW_EPS_P0, FORM_C1><<<grid, block, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, nullptr, 1.0, 1.0, 0.0, b);
    } else {
      if (f == 0)
        k_element_blocks<14, double, FW_EPS_P0, FORM_MASS><<<grid, block, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, nullptr, 1.0, 1.0, 0.0, b);
      else if (f == 1)
        k_element_blocks<14, double, FW_EPS_P0, FORM_C0><<<grid, block, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, nullptr, 1.0, 1.0, 0.0, b);
      else
        k_element_blocks<14, double, FW_EPS_P0, FORM_C1><<<grid, block, 0, s>>>(sp.ne, sp.nv, sp.p.p, sp.t.p, nullptr, 1.0, 1.0, 0.0, b);
    }
    FW_CHECK_CUDA(cudaGetLastError());
    Values& v = *vals[f];
    v.is_complex = false;
    v.v.alloc((size_t)pat.nnz);
    v.flag.alloc((size_t)pat.nnz);
    v.nnz_kept = -1;
    reduce_onto_pattern<double>(pat, b, nullptr, 0, sp.ne, v.v.p, v.flag.p, s);
    v.valid = true;
    h.tim.kernel_launches += 2;
  }
}

}  // namespace fw
