// Numeric phase of the multifrontal LU + level-batched triangular solves (device).  See lu.h.
//
// Fronts are dense column-major m x m blocks in one arena (own dofs first, then boundary dofs).
// Per tree level (deepest first), for every panel step k (32 columns):
//   k_lu_panel      one CTA per front: LU of the pivot window (<= 256 fully-summed rows x 32 cols)
//                   in shared memory with partial pivoting inside the window
//   k_lu_swap_trsm  row interchanges on the other columns, unit-lower TRSM for the U row block,
//                   and U_kk^-1 TRSM for the rows below the window
//   k_lu_gemm       rank-32 trailing update, 64x64 tiles, 4x4 register micro-tiles
// followed by the extend-add of the Schur complements into the parents.
#include <algorithm>

#include "lu.h"

namespace fw {

constexpr int PW = 32;    // panel width
constexpr int HWIN = 256; // pivot window height (rows eligible as pivots per panel)

struct FrontTab {
  const int *m, *ns, *own_start, *parent, *child0, *child1;
  const long long *front_off, *work_off, *blist_off;
  const int *blist, *rel;
};

struct LUDevice {
  DevBuf<int> iperm, perm, node_of;
  DevBuf<int> parent, child0, child1, m, ns, own_start;
  DevBuf<long long> front_off, work_off, blist_off;
  DevBuf<int> blist, rel, level_nodes;
  DevBuf<int> piv, rowperm;
  DevBuf<double> factors, work, xperm;
  DevBuf<int> counters;
  DevBuf<uint8_t> active;
  FrontTab tab() const {
    return FrontTab{m.p, ns.p, own_start.p, parent.p, child0.p, child1.p, front_off.p, work_off.p, blist_off.p,
                    blist.p, rel.p};
  }
};

LU::LU() {}
LU::~LU() {}

template <class S, class V>
__device__ inline S to_(V v);
template <>
__device__ inline double to_<double, double>(double v) { return v; }
template <>
__device__ inline cplx to_<cplx, double>(double v) { return {v, 0.0}; }
template <>
__device__ inline cplx to_<cplx, cplx>(cplx v) { return v; }

// ------------------------------------------------------------------------------ analysis upload
template <class T>
static void up(DevBuf<T>& d, const std::vector<T>& v, cudaStream_t s) {
  d.upload(v.data(), v.size(), s);
}
static void up64(DevBuf<long long>& d, const std::vector<int64_t>& v, cudaStream_t s) {
  static_assert(sizeof(long long) == sizeof(int64_t), "");
  d.upload((const long long*)v.data(), v.size(), s);
}

void lu_analyse(Handle& h, const std::vector<uint8_t>& active, int leaf_elements) {
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "space missing");
  h.lu.reset(new LU());
  LU& lu = *h.lu;
  std::vector<double> cx(sp.ne), cy(sp.ne);
  for (int e = 0; e < sp.ne; ++e) {
    int a = sp.h_t[e], b = sp.h_t[sp.ne + e], c = sp.h_t[2 * (size_t)sp.ne + e];
    cx[e] = (sp.h_p[a] + sp.h_p[b] + sp.h_p[c]) / 3.0;
    cy[e] = (sp.h_p[sp.nv + a] + sp.h_p[sp.nv + b] + sp.h_p[sp.nv + c]) / 3.0;
  }
  Trace tr(h.stream);
  lu_symbolic_build(lu.sym, sp.n, sp.nb, sp.ne, sp.h_edofs.data(), cx.data(), cy.data(), active.data(),
                    leaf_elements);
  tr.mark("lu analyse: host symbolic");
  lu.dev.reset(new LUDevice());
  LUDevice& d = *lu.dev;
  const LUSymbolic& S = lu.sym;
  cudaStream_t s = h.stream;
  up(d.iperm, S.iperm, s);
  up(d.perm, S.perm, s);
  up(d.node_of, S.node_of, s);
  up(d.parent, S.parent, s);
  up(d.child0, S.child0, s);
  up(d.child1, S.child1, s);
  up(d.m, S.m, s);
  up(d.ns, S.ns, s);
  up(d.own_start, S.own_start, s);
  up64(d.front_off, S.front_off, s);
  up64(d.work_off, S.work_off, s);
  up64(d.blist_off, S.blist_off, s);
  if (!S.blist.empty()) {
    up(d.blist, S.blist, s);
    up(d.rel, S.rel, s);
  } else {
    d.blist.alloc(1);
    d.rel.alloc(1);
  }
  up(d.level_nodes, S.level_nodes, s);
  d.active.upload(active.data(), active.size(), s);
  d.piv.alloc((size_t)std::max(1, S.n_active));
  d.rowperm.alloc((size_t)std::max(1, S.n_active));
  d.counters.alloc(4);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  tr.mark("lu analyse: upload");
}

// ------------------------------------------------------------------------------ scatter of OP into fronts
template <class S, class V>
__global__ void __launch_bounds__(256) k_lu_scatter(int n, const int* __restrict__ indptr,
                                                    const int* __restrict__ indices, const V* __restrict__ A,
                                                    const V* __restrict__ B, S alphaA, S alphaB, FrontTab T,
                                                    const int* __restrict__ iperm, const int* __restrict__ node_of,
                                                    S* __restrict__ F) {
  constexpr int TPR = 8;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(gt / TPR), l = (int)(gt % TPR);
  if (row >= n) return;
  const int pr = iperm[row];
  if (pr < 0) return;
  const int a = indptr[row], b = indptr[row + 1];
  for (int k = a + l; k < b; k += TPR) {
    const int pc = iperm[indices[k]];
    if (pc < 0) continue;
    const int mn = pr < pc ? pr : pc;
    const int node = node_of[mn];
    const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
    const int* bl = T.blist + T.blist_off[node];
    const int nbd = m - ns;
    int loc[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int x = w == 0 ? pr : pc;
      if (x < os + ns) {
        loc[w] = x - os;
      } else {
        int lo = 0, hi = nbd;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (bl[mid] < x)
            lo = mid + 1;
          else
            hi = mid;
        }
        loc[w] = ns + lo;
      }
    }
    S v = alphaA * to_<S, V>(A[k]) + alphaB * to_<S, V>(B[k]);
    F[T.front_off[node] + loc[0] + (long long)loc[1] * m] = v;
  }
}

// ------------------------------------------------------------------------------ extend-add
template <class S>
__global__ void __launch_bounds__(256) k_lu_extend_add(FrontTab T, const int* __restrict__ nodes, int slot,
                                                       S* __restrict__ F) {
  const int p = nodes[blockIdx.x];
  const int c = slot ? T.child1[p] : T.child0[p];
  if (c < 0) return;
  const int mc = T.m[c], nsc = T.ns[c], nbc = mc - nsc;
  const int nt = (nbc + 31) >> 5;
  const int tile = blockIdx.y;
  if (tile >= nt * nt) return;
  const int ti = tile % nt, tj = tile / nt;
  const int mp = T.m[p];
  const int* rel = T.rel + T.blist_off[c];
  const S* Fc = F + T.front_off[c];
  S* Fp = F + T.front_off[p];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = ti * 32 + tx;
  if (i >= nbc) return;
  const int ri = rel[i];
  for (int jj = ty; jj < 32; jj += 8) {
    const int j = tj * 32 + jj;
    if (j >= nbc) break;
    const int rj = rel[j];
    S* dst = Fp + ri + (long long)rj * mp;
    *dst = *dst + Fc[(nsc + i) + (long long)(nsc + j) * mc];
  }
}

// ------------------------------------------------------------------------------ panel (pivot window) kernel
template <class S>
__global__ void __launch_bounds__(HWIN) k_lu_panel(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                   S* __restrict__ F, int* __restrict__ piv, double thr,
                                                   int* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = kstep * PW;
  if (k0 >= ns) return;
  if (ns - k0 > HWIN) return;  // tall panels: k_lu_panel_tall (pivot search over all fully-summed rows)
  const int w = min(PW, ns - k0);
  const int hw = ns - k0;
  constexpr int LD = HWIN + 1;
  S* a = reinterpret_cast<S*>(smem_raw);              // a[c*LD + r]
  __shared__ double red_v[HWIN / 32];
  __shared__ int red_i[HWIN / 32];
  __shared__ int s_piv;
  S* Fn = F + T.front_off[node];
  const int t = threadIdx.x;
  for (int c = 0; c < w; ++c)
    if (t < hw) a[c * LD + t] = Fn[(k0 + t) + (long long)(k0 + c) * m];
  __syncthreads();
  for (int j = 0; j < w; ++j) {
    // ---- pivot search in column j, rows j..hw-1 (one row per thread)
    double v = -1.0;
    int idx = t;
    if (t >= j && t < hw) v = abs1_(a[j * LD + t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_down_sync(0xffffffffu, v, o);
      int oi = __shfl_down_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) v = ov, idx = oi;
    }
    if ((t & 31) == 0) red_v[t >> 5] = v, red_i[t >> 5] = idx;
    __syncthreads();
    if (t == 0) {
      double bv = red_v[0];
      int bi = red_i[0];
      for (int q = 1; q < HWIN / 32; ++q)
        if (red_v[q] > bv || (red_v[q] == bv && red_i[q] < bi)) bv = red_v[q], bi = red_i[q];
      s_piv = bi;
      piv[T.own_start[node] + k0 + j] = k0 + bi;
      if (!(bv >= thr)) {
        // static pivot perturbation (tiny / exactly singular pivot); corrected by iterative refinement
        a[j * LD + bi] = scalar_traits<S>::from(thr, 0.0);
        atomicAdd(&counters[0], 1);
      }
    }
    __syncthreads();
    const int pr = s_piv;
    if (t < w && pr != j) {
      S tmp = a[t * LD + j];
      a[t * LD + j] = a[t * LD + pr];
      a[t * LD + pr] = tmp;
    }
    __syncthreads();
    if (t > j && t < hw) {
      const S l = a[j * LD + t] / a[j * LD + j];
      a[j * LD + t] = l;
      for (int c = j + 1; c < w; ++c) fnma_(a[c * LD + t], l, a[c * LD + j]);
    }
    __syncthreads();
  }
  for (int c = 0; c < w; ++c)
    if (t < hw) Fn[(k0 + t) + (long long)(k0 + c) * m] = a[c * LD + t];
}

// ------------------------------------------------------------------------------ tall panel kernel
// Fronts whose remaining pivot block is taller than the shared-memory window: right-looking LU of the
// (ns-k0) x 32 panel in global memory (L2 resident), partial pivoting over ALL fully-summed rows.
// (A window restricted to 256 rows was measured to let the element growth reach 1e8-1e9 on the
// 500k-triangle pencil; full partial pivoting keeps it at ~1e3-1e4.)
constexpr int TALL_THREADS = 1024;
template <class S>
__global__ void __launch_bounds__(TALL_THREADS) k_lu_panel_tall(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                                S* __restrict__ F, int* __restrict__ piv, double thr,
                                                                int* __restrict__ counters) {
  __shared__ S prow[PW];
  __shared__ double red_v[TALL_THREADS / 32];
  __shared__ int red_i[TALL_THREADS / 32];
  __shared__ int s_piv;
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = kstep * PW;
  if (k0 >= ns) return;
  const int h = ns - k0;
  if (h <= HWIN) return;
  const int w = min(PW, h);
  S* P = F + T.front_off[node] + k0 + (long long)k0 * m;  // P[r + c*m]
  const int t = threadIdx.x;
  for (int j = 0; j < w; ++j) {
    double v = -1.0;
    int idx = 0x7fffffff;
    for (int r = j + t; r < h; r += TALL_THREADS) {
      const double a = abs1_(P[r + (long long)j * m]);
      if (a > v) v = a, idx = r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_down_sync(0xffffffffu, v, o);
      int oi = __shfl_down_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) v = ov, idx = oi;
    }
    if ((t & 31) == 0) red_v[t >> 5] = v, red_i[t >> 5] = idx;
    __syncthreads();
    if (t == 0) {
      double bv = red_v[0];
      int bi = red_i[0];
      for (int q = 1; q < TALL_THREADS / 32; ++q)
        if (red_v[q] > bv || (red_v[q] == bv && red_i[q] < bi)) bv = red_v[q], bi = red_i[q];
      s_piv = bi;
      piv[T.own_start[node] + k0 + j] = k0 + bi;
      if (!(bv >= thr)) {
        P[bi + (long long)j * m] = scalar_traits<S>::from(thr, 0.0);
        atomicAdd(&counters[0], 1);
      }
    }
    __syncthreads();
    const int pr = s_piv;
    if (t < w) {
      S a = P[pr + (long long)t * m];
      if (pr != j) {
        const S b = P[j + (long long)t * m];
        P[pr + (long long)t * m] = b;
        P[j + (long long)t * m] = a;
      }
      prow[t] = a;  // the pivot row after the interchange
    }
    __syncthreads();
    const S pivot = prow[j];
    for (int r = j + 1 + t; r < h; r += TALL_THREADS) {
      const S l = P[r + (long long)j * m] / pivot;
      P[r + (long long)j * m] = l;
      for (int c = j + 1; c < w; ++c) {
        S* dst = P + r + (long long)c * m;
        S x = *dst;
        fnma_(x, l, prow[c]);
        *dst = x;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ swaps + TRSMs
// blockIdx.x < ctiles : thread per column c outside the panel: row interchanges, and for c right of the
//                       panel the unit-lower solve of the U row block.
// blockIdx.x >= ctiles: thread per row below the pivot window: X U_kk = F (L21 / L-panel rows).
template <class S>
__global__ void __launch_bounds__(128) k_lu_swap_trsm(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                      S* __restrict__ F, const int* __restrict__ piv, int ctiles) {
  __shared__ S D[PW][PW + 1];
  __shared__ int sp[PW];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = kstep * PW;
  if (k0 >= ns) return;
  const int w = min(PW, ns - k0);
  const int hw = ns - k0;  // every fully-summed row is handled by the panel kernels
  const bool col_role = (int)blockIdx.y < ctiles;
  const int tile = col_role ? blockIdx.y : blockIdx.y - ctiles;
  if (col_role ? (tile * 128 >= m) : (k0 + hw + tile * 128 >= m)) return;
  S* Fn = F + T.front_off[node];
  for (int idx = threadIdx.x; idx < w * w; idx += 128) {
    int r = idx % w, c = idx / w;
    D[r][c] = Fn[(k0 + r) + (long long)(k0 + c) * m];
  }
  if (threadIdx.x < w) sp[threadIdx.x] = piv[T.own_start[node] + k0 + threadIdx.x];
  __syncthreads();
  if (col_role) {
    const int c = tile * 128 + threadIdx.x;
    if (c >= m || (c >= k0 && c < k0 + w)) return;
    S* col = Fn + (long long)c * m;
    for (int j = 0; j < w; ++j) {
      const int p = sp[j];
      if (p != k0 + j) {
        S tmp = col[k0 + j];
        col[k0 + j] = col[p];
        col[p] = tmp;
      }
    }
    if (c >= k0 + w) {
      S x[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) x[j] = (j < w) ? col[k0 + j] : scalar_traits<S>::zero();
#pragma unroll
      for (int j = 1; j < PW; ++j) {
        if (j < w) {
#pragma unroll
          for (int i = 0; i < j; ++i) fnma_(x[j], D[j][i], x[i]);
        }
      }
#pragma unroll
      for (int j = 0; j < PW; ++j)
        if (j < w) col[k0 + j] = x[j];
    }
  } else {
    const int r = k0 + hw + tile * 128 + threadIdx.x;
    if (r >= m) return;
    S x[PW];
#pragma unroll
    for (int c = 0; c < PW; ++c) x[c] = (c < w) ? Fn[r + (long long)(k0 + c) * m] : scalar_traits<S>::zero();
#pragma unroll
    for (int c = 0; c < PW; ++c) {
      if (c < w) {
#pragma unroll
        for (int i = 0; i < c; ++i) fnma_(x[c], x[i], D[i][c]);
        x[c] = x[c] / D[c][c];
      }
    }
#pragma unroll
    for (int c = 0; c < PW; ++c)
      if (c < w) Fn[r + (long long)(k0 + c) * m] = x[c];
  }
}

// ------------------------------------------------------------------------------ trailing update
template <class S>
struct GemmCfg {
  static constexpr int KC = 32;
};
template <>
struct GemmCfg<cplx> {
  static constexpr int KC = 16;
};

template <class S>
__global__ void __launch_bounds__(256) k_lu_gemm(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                 S* __restrict__ F) {
  constexpr int TM = 64, KC = GemmCfg<S>::KC;
  __shared__ S As[KC][TM];
  __shared__ S Bs[KC][TM + 1];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = kstep * PW;
  if (k0 >= ns) return;
  const int w = min(PW, ns - k0);
  const int r0 = k0 + w;
  const int nrem = m - r0;
  if (nrem <= 0) return;
  const int nt = (nrem + TM - 1) / TM;
  const int tile = blockIdx.y;
  if (tile >= nt * nt) return;
  const int ti = tile % nt, tj = tile / nt;
  S* Fn = F + T.front_off[node];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  S acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = scalar_traits<S>::zero();
  const int rbase = r0 + ti * TM, cbase = r0 + tj * TM;
  for (int kc = 0; kc < w; kc += KC) {
    for (int idx = threadIdx.x; idx < KC * TM; idx += 256) {
      const int r = idx % TM, k = idx / TM;
      const int gr = rbase + r, gk = kc + k;
      As[k][r] = (gr < m && gk < w) ? Fn[gr + (long long)(k0 + gk) * m] : scalar_traits<S>::zero();
    }
    for (int idx = threadIdx.x; idx < KC * TM; idx += 256) {
      const int k = idx % KC, c = idx / KC;
      const int gc = cbase + c, gk = kc + k;
      Bs[k][c] = (gc < m && gk < w) ? Fn[(k0 + gk) + (long long)gc * m] : scalar_traits<S>::zero();
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      S av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fnma_(acc[i][j], av[i], bv[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gc = cbase + ty * 4 + j;
    if (gc >= m) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = rbase + tx * 4 + i;
      if (gr < m) {
        S* dst = Fn + gr + (long long)gc * m;
        *dst = *dst + acc[i][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------ composed row permutation
// LAPACK-style sequential interchanges (row j <-> piv[j], j = 0..ns-1) composed once per front into a
// gather: (P b)[i] = b[rowperm[i]].  The forward substitution needs P b up front because the L rows it
// uses are in their final (fully interchanged) positions.
__global__ void k_lu_compose_perm(FrontTab T, int n_nodes, const int* __restrict__ piv, int* __restrict__ rowperm) {
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n_nodes) return;
  const int ns = T.ns[node], os = T.own_start[node];
  int* rp = rowperm + os;
  const int* pv = piv + os;
  for (int i = 0; i < ns; ++i) rp[i] = i;
  for (int j = 0; j < ns; ++j) {
    const int p = pv[j];
    if (p != j) {
      const int tmp = rp[j];
      rp[j] = rp[p];
      rp[p] = tmp;
    }
  }
}

// ------------------------------------------------------------------------------ factorisation driver
template <class S>
__global__ void k_absmax(const S* __restrict__ v, int64_t n, double* __restrict__ out) {
  double mx = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mx = fmax(mx, abs1_(v[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(mx));
}

template <class S, class V>
static void factor_typed(Handle& h, S alphaA, S alphaB) {
  LU& lu = *h.lu;
  LUDevice& d = *lu.dev;
  const LUSymbolic& Sy = lu.sym;
  cudaStream_t s = h.stream;
  const size_t wsc = sizeof(S) / sizeof(double);
  Trace tr(s);
  d.factors.alloc((size_t)std::max<int64_t>(1, Sy.factor_entries) * wsc);
  tr.mark("lu factor: alloc");
  d.factors.zero(s, (size_t)Sy.factor_entries * wsc);
  tr.mark("lu factor: memset");
  d.counters.zero(s);
  S* F = (S*)d.factors.p;
  FrontTab T = d.tab();
  const Pattern& pat = h.pat;
  k_lu_scatter<S, V><<<cdiv((int64_t)pat.n * 8, 256), 256, 0, s>>>(pat.n, pat.indptr.p, pat.indices.p,
                                                                  (const V*)h.A.v.p, (const V*)h.B.v.p, alphaA, alphaB,
                                                                  T, d.iperm.p, d.node_of.p, F);
  FW_CHECK_CUDA(cudaGetLastError());
  // pivot threshold relative to the largest entry of the pencil
  DevBuf<double> amax;
  amax.alloc(2);
  amax.zero(s);
  k_absmax<V><<<296, 256, 0, s>>>((const V*)h.A.v.p, pat.nnz, amax.p);
  k_absmax<V><<<296, 256, 0, s>>>((const V*)h.B.v.p, pat.nnz, amax.p + 1);
  double am[2] = {0, 0};
  amax.download(am, 2, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  const double scale = absv_(alphaA) * am[0] + absv_(alphaB) * am[1];
  const double thr = 1e-13 * (scale > 0 ? scale : 1.0);
  const size_t panel_smem = (size_t)PW * (HWIN + 1) * sizeof(S);
  FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_panel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem));
  int64_t launches = 3;
  for (int lev = Sy.n_levels - 1; lev >= 0; --lev) {
    const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes = d.level_nodes.p + Sy.level_start[lev];
    if (lev + 1 < Sy.n_levels) {
      const int cnb = Sy.level_max_nb[lev + 1];
      const int nt = (cnb + 31) / 32;
      if (nt > 0) {
        dim3 grid(nfront, nt * nt);
        k_lu_extend_add<S><<<grid, 256, 0, s>>>(T, nodes, 0, F);
        k_lu_extend_add<S><<<grid, 256, 0, s>>>(T, nodes, 1, F);
        launches += 2;
      }
    }
    const int max_ns = Sy.level_max_ns[lev], max_m = Sy.level_max_m[lev];
    const int npan = (max_ns + PW - 1) / PW;
    for (int k = 0; k < npan; ++k) {
      if (max_ns - k * PW > HWIN)
        k_lu_panel_tall<S><<<nfront, TALL_THREADS, 0, s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
      k_lu_panel<S><<<nfront, HWIN, panel_smem, s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
      const int ctiles = (max_m + 127) / 128;
      const int rtiles = (max_m + 127) / 128;
      k_lu_swap_trsm<S><<<dim3(nfront, ctiles + rtiles), 128, 0, s>>>(T, nodes, k, F, d.piv.p, ctiles);
      const int rem = max_m - (k * PW);  // upper bound of the trailing size
      const int nt = (rem + 63) / 64;
      if (nt > 0) k_lu_gemm<S><<<dim3(nfront, nt * nt), 256, 0, s>>>(T, nodes, k, F);
      launches += 3;
    }
  }
  tr.mark("lu factor: levels");
  k_lu_compose_perm<<<cdiv(Sy.n_nodes, 64), 64, 0, s>>>(T, Sy.n_nodes, d.piv.p, d.rowperm.p);
  tr.mark("lu factor: compose perm");
  launches += 1;
  FW_CHECK_CUDA(cudaGetLastError());
  int cnt[4] = {0, 0, 0, 0};
  d.counters.download(cnt, 4, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  lu.perturbed = cnt[0];
  lu.factored = true;
  h.tim.kernel_launches += launches;
}

void lu_factor(Handle& h, double aAr, double aAi, double aBr, double aBi, bool complex_arith) {
  FW_REQUIRE(h.lu && h.lu->dev, "lu_analyse must run first");
  FW_REQUIRE(h.A.valid && h.B.valid, "A and B must be assembled");
  LU& lu = *h.lu;
  const bool vals_complex = h.A.is_complex;
  if (vals_complex || aAi != 0.0 || aBi != 0.0) complex_arith = true;
  lu.is_complex = complex_arith;
  if (!complex_arith)
    factor_typed<double, double>(h, aAr, aBr);
  else if (!vals_complex)
    factor_typed<cplx, double>(h, cplx{aAr, aAi}, cplx{aBr, aBi});
  else
    factor_typed<cplx, cplx>(h, cplx{aAr, aAi}, cplx{aBr, aBi});
}

// ------------------------------------------------------------------------------ triangular solves
template <class S>
__global__ void k_lu_permute_in(int na, const int* __restrict__ perm, const S* __restrict__ b, int n, int nrhs,
                                S* __restrict__ xp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na) return;
  for (int r = 0; r < nrhs; ++r) xp[(size_t)r * na + i] = b[(size_t)r * n + perm[i]];
}
template <class S>
__global__ void k_lu_permute_out(int n, const int* __restrict__ iperm, const S* __restrict__ xp, int na, int nrhs,
                                 S* __restrict__ x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = iperm[i];
  for (int r = 0; r < nrhs; ++r) x[(size_t)r * n + i] = p >= 0 ? xp[(size_t)r * na + p] : scalar_traits<S>::zero();
}

// Level-batched substitution, four kernels per tree level:
//   forward : k_fwd_chain  (one CTA per front: assemble w from the rhs and the children, apply the row
//                           permutation, unit-lower solve of the ns x ns pivot block, panel by panel)
//             k_fwd_update (fronts x 128-row tiles: w_boundary -= L21 x1, the bulk of the L traffic)
//   backward: k_bwd_update (fronts x 128-row tiles: gather x2 from the parent, y1 = w1 - U12 x2)
//             k_bwd_chain  (one CTA per front: upper solve of the pivot block, write x)
constexpr int SOLVE_MAX_THREADS = 512;
constexpr int UPD_ROWS = 128;
constexpr int UPD_CHUNK = 256;

template <class S>
__device__ inline S shfl_bcast_(S v, int src);
template <>
__device__ inline double shfl_bcast_<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <>
__device__ inline cplx shfl_bcast_<cplx>(cplx v, int src) {
  return {__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src)};
}

template <class S, int NRHS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS) k_fwd_chain(FrontTab T, const int* __restrict__ nodes,
                                                                 const S* __restrict__ F,
                                                                 const int* __restrict__ rowperm,
                                                                 S* __restrict__ work, int64_t work_stride,
                                                                 S* __restrict__ xp, int na) {
  __shared__ S D[PW][PW + 1];
  __shared__ S xs[NRHS][PW];
  const int nthr = blockDim.x;
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  const S* Fn = F + T.front_off[node];
  S* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x;
  for (int i = t; i < m; i += nthr)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][i] = i < ns ? xp[(size_t)r * na + os + i] : scalar_traits<S>::zero();
  __syncthreads();
  for (int sl = 0; sl < 2; ++sl) {
    const int c = sl ? T.child1[node] : T.child0[node];
    if (c < 0) continue;
    const int nsc = T.ns[c], nbc = T.m[c] - nsc;
    const int* rel = T.rel + T.blist_off[c];
    for (int i = t; i < nbc; i += nthr) {
      const int ri = rel[i];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        const S* wc = work + (size_t)r * work_stride + T.work_off[c];
        w[r][ri] = w[r][ri] + wc[nsc + i];
      }
    }
    __syncthreads();
  }
  if (ns == 0) return;
  // row interchanges of the whole front first: stage the assembled own part in xp, gather it back
  for (int i = t; i < ns; i += nthr)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) xp[(size_t)r * na + os + i] = w[r][i];
  __syncthreads();
  for (int i = t; i < ns; i += nthr) {
    const int src = rowperm[os + i];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][i] = xp[(size_t)r * na + os + src];
  }
  __syncthreads();
  for (int k0 = 0; k0 < ns; k0 += PW) {
    const int wd = min(PW, ns - k0);
    for (int idx = t; idx < wd * wd; idx += nthr) {
      int r = idx % wd, c = idx / wd;
      D[r][c] = Fn[(k0 + r) + (long long)(k0 + c) * m];
    }
    __syncthreads();
    if (t < 32 * NRHS) {
      const int r = t >> 5, lane = t & 31;
      S x = lane < wd ? w[r][k0 + lane] : scalar_traits<S>::zero();
      for (int j = 0; j < wd; ++j) {
        const S xj = shfl_bcast_<S>(x, j);
        if (lane > j && lane < wd) fnma_(x, D[lane][j], xj);
      }
      if (lane < wd) {
        w[r][k0 + lane] = x;
        xs[r][lane] = x;
      }
    }
    __syncthreads();
    // remaining own rows only; the boundary rows are updated by k_fwd_update
    for (int row = k0 + wd + t; row < ns; row += nthr) {
      S acc[NRHS];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) acc[r] = w[r][row];
#pragma unroll 8
      for (int j = 0; j < wd; ++j) {
        const S l = Fn[row + (long long)(k0 + j) * m];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) fnma_(acc[r], l, xs[r][j]);
      }
#pragma unroll
      for (int r = 0; r < NRHS; ++r) w[r][row] = acc[r];
    }
    __syncthreads();
  }
}

template <class S, int NRHS>
__global__ void __launch_bounds__(UPD_ROWS) k_fwd_update(FrontTab T, const int* __restrict__ nodes,
                                                         const S* __restrict__ F, S* __restrict__ work,
                                                         int64_t work_stride) {
  __shared__ S xs[NRHS][UPD_CHUNK];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int nbd = m - ns;
  const int r0 = blockIdx.y * UPD_ROWS;
  if (r0 >= nbd || ns == 0) return;
  const S* Fn = F + T.front_off[node];
  S* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x;
  const int row = ns + r0 + t;
  const bool valid = row < m;
  S acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = scalar_traits<S>::zero();
  for (int j0 = 0; j0 < ns; j0 += UPD_CHUNK) {
    const int cn = min(UPD_CHUNK, ns - j0);
    for (int j = t; j < cn; j += UPD_ROWS)
#pragma unroll
      for (int r = 0; r < NRHS; ++r) xs[r][j] = w[r][j0 + j];
    __syncthreads();
    if (valid) {
      const S* col = Fn + row + (long long)j0 * m;
#pragma unroll 8
      for (int j = 0; j < cn; ++j) {
        const S l = col[(long long)j * m];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) fnma_(acc[r], l, xs[r][j]);
      }
    }
    __syncthreads();
  }
  if (valid)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][row] = w[r][row] + acc[r];
}

template <class S, int NRHS>
__global__ void __launch_bounds__(UPD_ROWS) k_bwd_update(FrontTab T, const int* __restrict__ nodes,
                                                         const S* __restrict__ F, S* __restrict__ work,
                                                         int64_t work_stride) {
  __shared__ S xs[NRHS][UPD_CHUNK];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int nbd = m - ns;
  if (nbd == 0) return;  // root
  const int r0 = blockIdx.y * UPD_ROWS;
  if (r0 >= ns && blockIdx.y != 0) return;  // tile 0 always runs: it stores x2 for the children
  const int p = T.parent[node];
  const S* Fn = F + T.front_off[node];
  const int* rel = T.rel + T.blist_off[node];
  S* w[NRHS];
  const S* wp[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) {
    w[r] = work + (size_t)r * work_stride + T.work_off[node];
    wp[r] = work + (size_t)r * work_stride + T.work_off[p];
  }
  const int t = threadIdx.x;
  const int row = r0 + t;
  const bool valid = row < ns;
  S acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = scalar_traits<S>::zero();
  for (int c0 = 0; c0 < nbd; c0 += UPD_CHUNK) {
    const int cn = min(UPD_CHUNK, nbd - c0);
    for (int c = t; c < cn; c += UPD_ROWS) {
      const int ri = rel[c0 + c];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        const S v = wp[r][ri];
        xs[r][c] = v;
        if (blockIdx.y == 0) w[r][ns + c0 + c] = v;
      }
    }
    __syncthreads();
    if (valid) {
      const S* col = Fn + row + (long long)(ns + c0) * m;
#pragma unroll 8
      for (int c = 0; c < cn; ++c) {
        const S u = col[(long long)c * m];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) fma_(acc[r], u, xs[r][c]);
      }
    }
    __syncthreads();
  }
  if (valid)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][row] = w[r][row] - acc[r];
}

template <class S, int NRHS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS) k_bwd_chain(FrontTab T, const int* __restrict__ nodes,
                                                                 const S* __restrict__ F, S* __restrict__ work,
                                                                 int64_t work_stride, S* __restrict__ xp, int na) {
  constexpr int NGMAX = SOLVE_MAX_THREADS / 32;
  __shared__ S D[PW][PW + 1];
  __shared__ S part[NRHS][NGMAX][PW];
  const int nthr = blockDim.x, NG = nthr >> 5;
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  if (ns == 0) return;
  const S* Fn = F + T.front_off[node];
  S* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x, lane = t & 31, g = t >> 5;
  const int npan = (ns + PW - 1) / PW;
  for (int kp = npan - 1; kp >= 0; --kp) {
    const int k0 = kp * PW;
    const int wd = min(PW, ns - k0);
    for (int idx = t; idx < wd * wd; idx += nthr) {
      int r = idx % wd, c = idx / wd;
      D[r][c] = Fn[(k0 + r) + (long long)(k0 + c) * m];
    }
    S acc[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) acc[r] = scalar_traits<S>::zero();
    if (lane < wd) {
      // own columns right of the panel (already final); boundary columns were folded in by k_bwd_update
#pragma unroll 4
      for (int c = k0 + wd + g; c < ns; c += NG) {
        const S u = Fn[(k0 + lane) + (long long)c * m];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) fma_(acc[r], u, w[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) part[r][g][lane] = acc[r];
    __syncthreads();
    if (t < 32 * NRHS) {
      const int r = t >> 5;
      S y = scalar_traits<S>::zero();
      if (lane < wd) {
        y = w[r][k0 + lane];
        for (int q = 0; q < NG; ++q) y = y - part[r][q][lane];
      }
      for (int j = wd - 1; j >= 0; --j) {
        S xj = scalar_traits<S>::zero();
        if (lane == j) {
          y = y / D[j][j];
          xj = y;
        }
        xj = shfl_bcast_<S>(xj, j);
        if (lane < j) fnma_(y, D[lane][j], xj);
      }
      if (lane < wd) w[r][k0 + lane] = y;
    }
    __syncthreads();
  }
  for (int i = t; i < ns; i += nthr)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) xp[(size_t)r * na + os + i] = w[r][i];
}

static int chain_threads(int max_ns) {
  if (max_ns <= 64) return 64;
  if (max_ns <= 128) return 128;
  if (max_ns <= 512) return 256;
  return SOLVE_MAX_THREADS;
}

template <class S, int NRHS>
static void solve_typed(Handle& h, const S* b, S* x) {
  LU& lu = *h.lu;
  LUDevice& d = *lu.dev;
  const LUSymbolic& Sy = lu.sym;
  cudaStream_t s = h.stream;
  const size_t wsc = sizeof(S) / sizeof(double);
  const int n = Sy.n, na = Sy.n_active;
  d.work.alloc((size_t)std::max<int64_t>(1, Sy.work_entries) * 2 * wsc);
  d.xperm.alloc((size_t)std::max(1, na) * 2 * wsc);
  S* xp = (S*)d.xperm.p;
  S* work = (S*)d.work.p;
  const S* F = (const S*)d.factors.p;
  FrontTab T = d.tab();
  int64_t launches = 2;
  const char* tenv = getenv("FW_TRACE");
  const bool trace2 = tenv && atoi(tenv) >= 2;
  Trace tr(s);
  if (na > 0) k_lu_permute_in<S><<<cdiv(na, 256), 256, 0, s>>>(na, d.perm.p, b, n, NRHS, xp);
  for (int lev = Sy.n_levels - 1; lev >= 0; --lev) {
    const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes = d.level_nodes.p + Sy.level_start[lev];
    const int thr = chain_threads(Sy.level_max_ns[lev]);
    k_fwd_chain<S, NRHS><<<nfront, thr, 0, s>>>(T, nodes, F, d.rowperm.p, work, Sy.work_entries, xp, na);
    if (trace2) tr.mark(("fwd chain  lev " + std::to_string(lev) + " fronts " + std::to_string(nfront) + " max_ns " +
                         std::to_string(Sy.level_max_ns[lev])).c_str());
    const int tiles = (Sy.level_max_nb[lev] + UPD_ROWS - 1) / UPD_ROWS;
    if (tiles > 0) {
      k_fwd_update<S, NRHS><<<dim3(nfront, tiles), UPD_ROWS, 0, s>>>(T, nodes, F, work, Sy.work_entries);
      ++launches;
      if (trace2) tr.mark(("fwd update lev " + std::to_string(lev) + " max_nb " +
                           std::to_string(Sy.level_max_nb[lev])).c_str());
    }
    ++launches;
  }
  for (int lev = 0; lev < Sy.n_levels; ++lev) {
    const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes = d.level_nodes.p + Sy.level_start[lev];
    const int thr = chain_threads(Sy.level_max_ns[lev]);
    if (Sy.level_max_nb[lev] > 0) {
      const int tiles = std::max(1, (Sy.level_max_ns[lev] + UPD_ROWS - 1) / UPD_ROWS);
      k_bwd_update<S, NRHS><<<dim3(nfront, tiles), UPD_ROWS, 0, s>>>(T, nodes, F, work, Sy.work_entries);
      ++launches;
      if (trace2) tr.mark(("bwd update lev " + std::to_string(lev)).c_str());
    }
    k_bwd_chain<S, NRHS><<<nfront, thr, 0, s>>>(T, nodes, F, work, Sy.work_entries, xp, na);
    ++launches;
    if (trace2) tr.mark(("bwd chain  lev " + std::to_string(lev)).c_str());
  }
  k_lu_permute_out<S><<<cdiv(n, 256), 256, 0, s>>>(n, d.iperm.p, xp, na, NRHS, x);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += launches;
}

void lu_solve_dev(Handle& h, const void* b, void* x, int nrhs) {
  FW_REQUIRE(h.lu && h.lu->factored, "LU has not been factored");
  FW_REQUIRE(nrhs == 1 || nrhs == 2, "nrhs must be 1 or 2");
  if (h.lu->is_complex) {
    if (nrhs == 1)
      solve_typed<cplx, 1>(h, (const cplx*)b, (cplx*)x);
    else
      solve_typed<cplx, 2>(h, (const cplx*)b, (cplx*)x);
  } else {
    if (nrhs == 1)
      solve_typed<double, 1>(h, (const double*)b, (double*)x);
    else
      solve_typed<double, 2>(h, (const double*)b, (double*)x);
  }
}

}  // namespace fw

// ---- debug export (tests only): factor arena and pivot vector
namespace fw {
void lu_debug_download(Handle& h, int which, void* out) {
  FW_REQUIRE(h.lu && h.lu->factored, "LU has not been factored");
  LU& lu = *h.lu;
  LUDevice& d = *lu.dev;
  if (which == 0) {
    size_t cnt = (size_t)lu.sym.factor_entries * (lu.is_complex ? 2 : 1);
    d.factors.download((double*)out, cnt, h.stream);
  } else {
    d.piv.download((int*)out, (size_t)lu.sym.n_active, h.stream);
  }
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
}
}  // namespace fw
