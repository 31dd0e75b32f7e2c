// Cross-mesh branch of calculate_overlap (femwell/maxwell/waveguide.py:737-759) and of
// calculate_scalar_product (:762-782): the two modes live on different meshes.
//
// Reference algorithm: on mesh j the transverse parts of E_j and H_j are L2-projected onto continuous
// vector-valued P1 (basis_j.with_element(ElementVector(ElementTriP1())).project(...)); the projected fields
// are then evaluated at the quadrature points of mesh i (interpolator(w.x): point location + barycentric
// interpolation) inside the Functional that is assembled on basis_i.
//
// Device pipeline:
//   1. k_p1_project_coo   per triangle of mesh j: P1 mass block (9 triplets) and the 3 x 4 load entries of the
//                         four complex fields (Ex, Ey, Hx, Hy) -- COO, summed deterministically by the radix-sort
//                         COO->CSR machinery of the assembly (sort.cu)
//   2. pcg_solve_multi    the four mass solves advance together (mass_cg.cu)
//   3. uniform cell grid  over mesh j: count / scan / fill of the triangles overlapping every cell
//   4. k_cross_overlap    per triangle of mesh i and quadrature point: locate the point (smallest containing
//                         triangle index of the cell: deterministic), interpolate, accumulate; fixed-order
//                         two-stage reduction.
// HBM/L2-bound gather work; nothing here is a dense contraction.
#include <algorithm>

#include "solver.h"

namespace fw {

struct CrossTab {
  int nb, nq;
  int kind[MAXNB];
  double vx[MAXNB][MAXQ], vy[MAXNB][MAXQ];
  double X[MAXQ], Y[MAXQ], W[MAXQ];
};
__constant__ CrossTab c_ti, c_tj;

static void fill_tab(CrossTab& t, int nb, int nq, const int* kind, const double* vec, const double* X, const double* W) {
  memset(&t, 0, sizeof(t));
  t.nb = nb, t.nq = nq;
  for (int l = 0; l < nb; ++l) {
    t.kind[l] = kind[l];
    for (int q = 0; q < nq; ++q) {
      t.vx[l][q] = vec[((size_t)l * 2 + 0) * nq + q];
      t.vy[l][q] = vec[((size_t)l * 2 + 1) * nq + q];
    }
  }
  for (int q = 0; q < nq; ++q) t.X[q] = X[q], t.Y[q] = X[nq + q], t.W[q] = W[q];
}

// ------------------------------------------------------------------------------ 1. projection triplets (mesh j)
template <int NB>
__global__ void __launch_bounds__(128) k_p1_project_coo(int ne, int nv, const double* __restrict__ p,
                                                        const int* __restrict__ t, const int* __restrict__ edofs,
                                                        const cplx* __restrict__ E, const cplx* __restrict__ Hf,
                                                        int* __restrict__ rows_m, int* __restrict__ cols_m,
                                                        double* __restrict__ vals_m, int* __restrict__ rows_r,
                                                        int* __restrict__ cols_r, cplx* __restrict__ vals_r) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int v[3] = {t[e], t[ne + e], t[2 * ne + e]};
  const double x0 = p[v[0]], y0 = p[nv + v[0]];
  const double a11 = p[v[1]] - x0, a21 = p[nv + v[1]] - y0;
  const double a12 = p[v[2]] - x0, a22 = p[nv + v[2]] - y0;
  const double det = a11 * a22 - a12 * a21;
  const double idet = 1.0 / det, adet = fabs(det);
  cplx ce[NB], ch[NB];
#pragma unroll
  for (int l = 0; l < NB; ++l) {
    const int d = edofs[(size_t)l * ne + e];
    ce[l] = E[d];
    ch[l] = Hf[d];
  }
  double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  cplx r[3][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) r[a][c] = cplx{0, 0};
  const int nq = c_tj.nq;
  for (int q = 0; q < nq; ++q) {
    const double w = c_tj.W[q] * adet;
    cplx sx{0, 0}, sy{0, 0}, hx{0, 0}, hy{0, 0};
#pragma unroll
    for (int l = 0; l < NB; ++l) {
      if (c_tj.kind[l] == 0) {
        fma_(sx, ce[l], c_tj.vx[l][q]);
        fma_(sy, ce[l], c_tj.vy[l][q]);
        fma_(hx, ch[l], c_tj.vx[l][q]);
        fma_(hy, ch[l], c_tj.vy[l][q]);
      }
    }
    const cplx F[4] = {(sx * a22 - sy * a21) * idet, (sy * a11 - sx * a12) * idet, (hx * a22 - hy * a21) * idet,
                       (hy * a11 - hx * a12) * idet};
    const double lam[3] = {1.0 - c_tj.X[q] - c_tj.Y[q], c_tj.X[q], c_tj.Y[q]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b) M[a][b] += w * lam[a] * lam[b];
#pragma unroll
      for (int c = 0; c < 4; ++c) fma_(r[a][c], F[c], w * lam[a]);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const size_t o = (size_t)(a * 3 + b) * ne + e;
      rows_m[o] = v[a], cols_m[o] = v[b], vals_m[o] = M[a][b];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const size_t o = (size_t)(a * 4 + c) * ne + e;
      rows_r[o] = v[a], cols_r[o] = c, vals_r[o] = r[a][c];
    }
  }
}

// CSR (nv rows, columns 0..3) -> dense [4][nv]
__global__ void k_rhs_to_dense(int nv, const int* __restrict__ indptr, const int* __restrict__ indices,
                               const cplx* __restrict__ vals, cplx* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  for (int e = indptr[v]; e < indptr[v + 1]; ++e) out[(size_t)indices[e] * nv + v] = vals[e];
}

// ------------------------------------------------------------------------------ 3. cell grid over mesh j
struct Grid {
  int gx, gy;
  double x0, y0, inv_dx, inv_dy;
};
__device__ inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <bool FILL>
__global__ void k_grid_insert(int ne, int nv, const double* __restrict__ p, const int* __restrict__ t, Grid g,
                              uint32_t* __restrict__ count, const uint32_t* __restrict__ offsets,
                              int* __restrict__ list) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int v0 = t[e], v1 = t[ne + e], v2 = t[2 * ne + e];
  const double xa = p[v0], xb = p[v1], xc = p[v2], ya = p[nv + v0], yb = p[nv + v1], yc = p[nv + v2];
  const int i0 = clampi((int)floor((fmin(xa, fmin(xb, xc)) - g.x0) * g.inv_dx), 0, g.gx - 1);
  const int i1 = clampi((int)floor((fmax(xa, fmax(xb, xc)) - g.x0) * g.inv_dx), 0, g.gx - 1);
  const int j0 = clampi((int)floor((fmin(ya, fmin(yb, yc)) - g.y0) * g.inv_dy), 0, g.gy - 1);
  const int j1 = clampi((int)floor((fmax(ya, fmax(yb, yc)) - g.y0) * g.inv_dy), 0, g.gy - 1);
  for (int j = j0; j <= j1; ++j)
    for (int i = i0; i <= i1; ++i) {
      const int c = j * g.gx + i;
      const uint32_t pos = atomicAdd(&count[c], 1u);
      if (FILL) list[offsets[c] + pos] = e;
    }
}

// ------------------------------------------------------------------------------ 4. evaluation on mesh i
template <int NB>
__global__ void __launch_bounds__(128) k_cross_overlap(int nsel, const int* __restrict__ sel, int ne, int nv,
                                                       const double* __restrict__ p, const int* __restrict__ t,
                                                       const int* __restrict__ edofs, const cplx* __restrict__ Ei,
                                                       const cplx* __restrict__ Hi, int nvj, int nej,
                                                       const double* __restrict__ pj, const int* __restrict__ tj,
                                                       const cplx* __restrict__ U, Grid g,
                                                       const uint32_t* __restrict__ offsets,
                                                       const int* __restrict__ list, int mode,
                                                       cplx* __restrict__ partial, int* __restrict__ err) {
  __shared__ cplx red[4];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  cplx acc{0, 0};
  if (tid < nsel) {
    const int e = sel ? sel[tid] : tid;
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
      ce[l] = Ei[d];
      ch[l] = Hi ? Hi[d] : cplx{0, 0};
    }
    const int nq = c_ti.nq;
    for (int q = 0; q < nq; ++q) {
      cplx sx{0, 0}, sy{0, 0}, hx{0, 0}, hy{0, 0};
#pragma unroll
      for (int l = 0; l < NB; ++l) {
        if (c_ti.kind[l] == 0) {
          fma_(sx, ce[l], c_ti.vx[l][q]);
          fma_(sy, ce[l], c_ti.vy[l][q]);
          fma_(hx, ch[l], c_ti.vx[l][q]);
          fma_(hy, ch[l], c_ti.vy[l][q]);
        }
      }
      const cplx ex = (sx * a22 - sy * a21) * idet, ey = (sy * a11 - sx * a12) * idet;
      const cplx bx = (hx * a22 - hy * a21) * idet, by = (hy * a11 - hx * a12) * idet;
      // physical point and its triangle in mesh j
      const double X = c_ti.X[q], Y = c_ti.Y[q];
      const double px = x0 + a11 * X + a12 * Y, py = y0 + a21 * X + a22 * Y;
      const int ci = clampi((int)floor((px - g.x0) * g.inv_dx), 0, g.gx - 1);
      const int cj = clampi((int)floor((py - g.y0) * g.inv_dy), 0, g.gy - 1);
      const int cell = cj * g.gx + ci;
      int best = 0x7fffffff;
      double bl1 = 0, bl2 = 0;
      for (uint32_t k = offsets[cell]; k < offsets[cell + 1]; ++k) {
        const int f = list[k];
        if (f >= best) continue;
        const int w0 = tj[f], w1 = tj[nej + f], w2 = tj[2 * nej + f];
        const double qx = pj[w0], qy = pj[nvj + w0];
        const double b11 = pj[w1] - qx, b21 = pj[nvj + w1] - qy;
        const double b12 = pj[w2] - qx, b22 = pj[nvj + w2] - qy;
        const double dj = 1.0 / (b11 * b22 - b12 * b21);
        const double l1 = ((px - qx) * b22 - (py - qy) * b12) * dj;
        const double l2 = ((py - qy) * b11 - (px - qx) * b21) * dj;
        const double tol = 1e-10;
        if (l1 >= -tol && l2 >= -tol && l1 + l2 <= 1.0 + tol) best = f, bl1 = l1, bl2 = l2;
      }
      if (best == 0x7fffffff) {
        atomicExch(err, 1);
        continue;
      }
      const int w0 = tj[best], w1 = tj[nej + best], w2 = tj[2 * nej + best];
      const double l0 = 1.0 - bl1 - bl2;
      cplx F[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const cplx* u = U + (size_t)c * nvj;
        F[c] = u[w0] * l0 + u[w1] * bl1 + u[w2] * bl2;
      }
      const double w = c_ti.W[q] * adet;
      // conj(E_i) x H_j
      cplx v = conj_(ex) * F[3] - conj_(ey) * F[2];
      if (mode == 0) {
        // + E_j x conj(H_i), all times 1/2
        v = (v + F[0] * conj_(by) - F[1] * conj_(bx)) * 0.5;
      }
      acc = acc + v * w;
    }
  }
  // fixed-order block reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc.re += __shfl_down_sync(0xffffffffu, acc.re, o);
    acc.im += __shfl_down_sync(0xffffffffu, acc.im, o);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) partial[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
}

__global__ void __launch_bounds__(256) k_sum_partials(const cplx* __restrict__ partial, int n, cplx* __restrict__ out) {
  __shared__ cplx sh[256];
  cplx s{0, 0};
  for (int i = threadIdx.x; i < n; i += 256) s = s + partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// mode 0: 0.5 int [conj(E_i) x H_j + E_j x conj(H_i)] ; mode 1: int conj(E_i) x H_j   (over the elements of basis_i)
std::complex<double> overlap_cross_dev(Handle& h, const fw_space& sj, const cplx* Ei, const cplx* Hi,
                                       const double* Ej_host, const double* Hj_host, int mode, int n_sel,
                                       const int* elements_dev) {
  const Space& sp = h.sp;
  cudaStream_t s = h.stream;
  FW_REQUIRE(sp.valid, "space missing");
  FW_REQUIRE(sj.order == 1 || sj.order == 2, "Only order 1 and 2 implemented by now.");
  const int nbj = sj.order == 1 ? 6 : 14;
  FW_REQUIRE(sj.nbfun == nbj && sj.nqp >= 1 && sj.nqp <= MAXQ, "bad foreign space");
  FW_REQUIRE(sj.p && sj.t && sj.element_dofs && sj.kind && sj.vec && sj.X && sj.W && Hj_host, "NULL array");
  const int nvj = sj.n_vertices, nej = sj.n_elements, nj = sj.n_dofs;
  // ---- tables
  CrossTab ti, tj;
  fill_tab(ti, sp.nb, sp.nq, sp.kind, sp.vec.data(), sp.X.data(), sp.W.data());
  fill_tab(tj, nbj, sj.nqp, sj.kind, sj.vec, sj.X, sj.W);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_ti, &ti, sizeof(ti)));
  FW_CHECK_CUDA(cudaMemcpyToSymbol(c_tj, &tj, sizeof(tj)));
  // ---- foreign mesh and fields
  DevBuf<double> pj, dEj, dHj;
  DevBuf<int> tjd, edj;
  pj.upload(sj.p, (size_t)2 * nvj, s);
  tjd.upload(sj.t, (size_t)3 * nej, s);
  edj.upload(sj.element_dofs, (size_t)nbj * nej, s);
  dHj.upload(Hj_host, (size_t)2 * nj, s);
  if (Ej_host) {
    dEj.upload(Ej_host, (size_t)2 * nj, s);
  } else {
    dEj.alloc((size_t)2 * nj);
    dEj.zero(s);
  }
  // ---- 1. projection triplets
  DevBuf<int> rm, cm, rr, cr;
  DevBuf<double> vm, vr;
  rm.alloc((size_t)9 * nej), cm.alloc((size_t)9 * nej), vm.alloc((size_t)9 * nej);
  rr.alloc((size_t)12 * nej), cr.alloc((size_t)12 * nej), vr.alloc((size_t)24 * nej);
  if (nbj == 6)
    k_p1_project_coo<6><<<cdiv(nej, 128), 128, 0, s>>>(nej, nvj, pj.p, tjd.p, edj.p, (const cplx*)dEj.p,
                                                       (const cplx*)dHj.p, rm.p, cm.p, vm.p, rr.p, cr.p, (cplx*)vr.p);
  else
    k_p1_project_coo<14><<<cdiv(nej, 128), 128, 0, s>>>(nej, nvj, pj.p, tjd.p, edj.p, (const cplx*)dEj.p,
                                                        (const cplx*)dHj.p, rm.p, cm.p, vm.p, rr.p, cr.p, (cplx*)vr.p);
  FW_CHECK_CUDA(cudaGetLastError());
  // P1 mass matrix (CSR)
  DevBuf<int> mip, mind;
  DevBuf<double> mval;
  coo_to_csr_dev(h, nvj, (int64_t)9 * nej, rm, cm, vm, false);
  const int64_t mnnz = compact_csr_dev(h, h.coo_pat, h.coo_val, 3, mip, mind, mval);
  // loads: CSR with the field index as column -> dense [4][nvj]
  DevBuf<double> rhs, U;
  rhs.alloc((size_t)8 * nvj), U.alloc((size_t)8 * nvj);
  rhs.zero(s);
  coo_to_csr_dev(h, nvj, (int64_t)12 * nej, rr, cr, vr, true);
  k_rhs_to_dense<<<cdiv(nvj, 256), 256, 0, s>>>(nvj, h.coo_pat.indptr.p, h.coo_pat.indices.p,
                                                (const cplx*)h.coo_val.v.p, (cplx*)rhs.p);
  // ---- 2. the four mass solves
  pcg_solve_multi(h, nvj, mip.p, mind.p, mval.p, mnnz, 4, (const cplx*)rhs.p, (cplx*)U.p);
  // ---- 3. cell grid (about two triangles per cell)
  Grid g;
  {
    double xa = 1e300, xb = -1e300, ya = 1e300, yb = -1e300;
    for (int v = 0; v < nvj; ++v) {
      xa = std::min(xa, sj.p[v]), xb = std::max(xb, sj.p[v]);
      ya = std::min(ya, sj.p[nvj + v]), yb = std::max(yb, sj.p[nvj + v]);
    }
    const double ar = (xb - xa) / std::max(1e-300, yb - ya);
    const double cells = std::max(1.0, nej / 2.0);
    g.gx = std::max(1, std::min(4096, (int)std::ceil(std::sqrt(cells * ar))));
    g.gy = std::max(1, std::min(4096, (int)std::ceil(std::sqrt(cells / ar))));
    g.x0 = xa, g.y0 = ya;
    g.inv_dx = g.gx / std::max(1e-300, xb - xa);
    g.inv_dy = g.gy / std::max(1e-300, yb - ya);
  }
  const int ncell = g.gx * g.gy;
  DevBuf<uint32_t> count, offsets;
  DevBuf<int> list;
  count.alloc((size_t)ncell + 1), offsets.alloc((size_t)ncell + 1);
  count.zero(s);
  k_grid_insert<false><<<cdiv(nej, 128), 128, 0, s>>>(nej, nvj, pj.p, tjd.p, g, count.p, nullptr, nullptr);
  exclusive_scan_u32(count.p, offsets.p, (int64_t)ncell + 1, h.scan_tmp, s);
  uint32_t total = 0;
  FW_CHECK_CUDA(cudaMemcpyAsync(&total, offsets.p + ncell, 4, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  list.alloc((size_t)std::max<uint32_t>(1, total));
  count.zero(s);
  k_grid_insert<true><<<cdiv(nej, 128), 128, 0, s>>>(nej, nvj, pj.p, tjd.p, g, count.p, offsets.p, list.p);
  // ---- 4. evaluation on basis_i
  const int nsel = elements_dev ? n_sel : sp.ne;
  const int nblk = std::max(1, cdiv(nsel, 128));
  DevBuf<double> partial, res;
  DevBuf<int> err;
  partial.alloc((size_t)2 * nblk), res.alloc(2), err.alloc(1);
  err.zero(s);
  if (sp.nb == 6)
    k_cross_overlap<6><<<nblk, 128, 0, s>>>(nsel, elements_dev, sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, Ei, Hi, nvj, nej,
                                            pj.p, tjd.p, (const cplx*)U.p, g, offsets.p, list.p, mode, (cplx*)partial.p,
                                            err.p);
  else
    k_cross_overlap<14><<<nblk, 128, 0, s>>>(nsel, elements_dev, sp.ne, sp.nv, sp.p.p, sp.t.p, sp.edofs.p, Ei, Hi, nvj, nej,
                                             pj.p, tjd.p, (const cplx*)U.p, g, offsets.p, list.p, mode, (cplx*)partial.p,
                                             err.p);
  k_sum_partials<<<1, 256, 0, s>>>((const cplx*)partial.p, nblk, (cplx*)res.p);
  FW_CHECK_CUDA(cudaGetLastError());
  double out[2];
  int herr = 0;
  res.download(out, 2, s);
  err.download(&herr, 1, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  h.tim.kernel_launches += 12;
  if (herr) throw Error(FW_ERR_INVALID, "Point is outside of the mesh.");
  return {out[0], out[1]};
}

}  // namespace fw
