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
#include <cooperative_groups.h>

#include <algorithm>

#include "lu.h"

#include "lu_dev.h"

namespace fw {

LU::LU() {}
LU::~LU() { drop_graphs(); }

template <class S, class V>
__device__ inline S to_(V v);
template <>
__device__ inline double to_<double, double>(double v) { return v; }
template <>
__device__ inline cplx to_<cplx, double>(double v) { return {v, 0.0}; }
template <>
__device__ inline cplx to_<cplx, cplx>(cplx v) { return v; }

// ------------------------------------------------------------------------------ analysis upload
// large tables go through the pinned staging ring (pageable cudaMemcpy runs at a few GB/s)
static thread_local Handle* g_up_handle = nullptr;
struct UpGuard {
  explicit UpGuard(Handle* h) { g_up_handle = h; }
  ~UpGuard() { g_up_handle = nullptr; }
};
template <class T>
static void up(DevBuf<T>& d, const std::vector<T>& v, cudaStream_t s) {
  d.alloc(v.size());
  if (g_up_handle)
    upload_staged(*g_up_handle, d.p, v.data(), v.size() * sizeof(T));
  else
    d.upload(v.data(), v.size(), s);
}
static void up64(DevBuf<long long>& d, const std::vector<int64_t>& v, cudaStream_t s) {
  static_assert(sizeof(long long) == sizeof(int64_t), "");
  d.alloc(v.size());
  if (g_up_handle)
    upload_staged(*g_up_handle, d.p, v.data(), v.size() * sizeof(int64_t));
  else
    d.upload((const long long*)v.data(), v.size(), s);
}

static void lu_host_analysis(const Handle& h, LU& lu) {
  const Space& sp = h.sp;
  if (h.generic.valid) {
    const Connectivity& c = h.generic;
    lu_symbolic_build(lu.sym, c.n, c.nb, c.ne, c.edofs.data(), c.cx.data(), c.cy.data(), lu.active_key.data(),
                      lu.leaf_key);
  } else {
    std::vector<double> cx(sp.ne), cy(sp.ne);
    for (int e = 0; e < sp.ne; ++e) {
      int a = sp.h_t[e], b = sp.h_t[sp.ne + e], c = sp.h_t[2 * (size_t)sp.ne + e];
      cx[e] = (sp.h_p[a] + sp.h_p[b] + sp.h_p[c]) / 3.0;
      cy[e] = (sp.h_p[sp.nv + a] + sp.h_p[sp.nv + b] + sp.h_p[sp.nv + c]) / 3.0;
    }
    lu_symbolic_build(lu.sym, sp.n, sp.nb, sp.ne, sp.h_edofs.data(), cx.data(), cy.data(), lu.active_key.data(),
                      lu.leaf_key);
  }
}

static void lu_analyse_finish(Handle& h);

static bool lu_reusable(const Handle& h, const std::vector<uint8_t>& active, int leaf_elements) {
  return h.lu && h.lu->dev && h.lu->leaf_key == leaf_elements && h.lu->active_key == active;
}

// The analysis runs on the device (lu_symbolic_gpu.cu); FW_HOST_ANALYSIS=1 selects the threaded host version
// (lu_symbolic.cpp, same results bit by bit), which is also what the CPU tests exercise.
static bool host_analysis() {
  static const bool v = getenv("FW_HOST_ANALYSIS") != nullptr;
  return v;
}

void lu_analyse_prefetch(Handle& h, const std::vector<uint8_t>& active, int leaf_elements) {
  if (!host_analysis()) return;  // nothing to overlap: the device analysis takes a few milliseconds
  // an analysis with the same keys may already be running (started by fw_set_space from fw_analysis_hint)
  if (h.lu_prefetch && h.lu_prefetch->lu && h.lu_prefetch->lu->leaf_key == leaf_elements &&
      h.lu_prefetch->lu->active_key == active)
    return;
  h.lu_prefetch.reset();
  if (lu_reusable(h, active, leaf_elements)) return;
  if (getenv("FW_NO_PREFETCH")) return;
  h.lu_prefetch.reset(new LUPrefetch());
  LUPrefetch* pf = h.lu_prefetch.get();
  pf->lu.reset(new LU());
  pf->lu->active_key = active;
  pf->lu->leaf_key = leaf_elements;
  const Handle* hp = &h;
  pf->th = std::thread([pf, hp] {
    try {
      lu_host_analysis(*hp, *pf->lu);
    } catch (const std::exception& e) {
      pf->err = e.what();
      if (pf->err.empty()) pf->err = "analysis failed";
    }
  });
}

void lu_analyse(Handle& h, const std::vector<uint8_t>& active, int leaf_elements) {
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid || h.generic.valid, "space missing");
  std::unique_ptr<LUPrefetch> pf = std::move(h.lu_prefetch);
  if (pf && pf->th.joinable()) pf->th.join();
  if (lu_reusable(h, active, leaf_elements)) {
    // same mesh (fw_set_space / fw_eigs_csr reset h.lu), same Dirichlet set: wavelength / epsilon sweeps
    h.lu->factored = false;
    return;
  }
  Trace tr(h.stream);
  cudaStream_t s = h.stream;
  if (!host_analysis()) {
    h.lu.reset(new LU());
    h.lu->active_key = active;
    h.lu->leaf_key = leaf_elements;
    lu_symbolic_build_gpu(h, *h.lu, active, leaf_elements);
    lu_analyse_finish(h);
    return;
  }
  if (pf && pf->lu && pf->lu->leaf_key == leaf_elements && pf->lu->active_key == active) {
    if (!pf->err.empty()) throw Error(FW_ERR_INTERNAL, pf->err);
    h.lu = std::move(pf->lu);
  } else {
    h.lu.reset(new LU());
    h.lu->active_key = active;
    h.lu->leaf_key = leaf_elements;
    lu_host_analysis(h, *h.lu);
  }
  LU& lu = *h.lu;
  tr.mark("lu analyse: host symbolic");
  lu.dev.reset(new LUDevice(h.lu_arena));
  LUDevice& d = *lu.dev;
  const LUSymbolic& S = lu.sym;
  UpGuard up_guard(&h);
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
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  tr.mark("lu analyse: upload");
  lu_analyse_finish(h);
}

// work arrays of the numeric phase + storage layout of the inverted diagonal blocks (both analysis paths)
static void lu_analyse_finish(Handle& h) {
  LU& lu = *h.lu;
  LUDevice& d = *lu.dev;
  const LUSymbolic& S = lu.sym;
  cudaStream_t s = h.stream;
  d.piv.alloc((size_t)std::max(1, S.n_active));
  d.rowperm.alloc((size_t)std::max(1, S.n_active));
  d.counters.alloc(4);
  {
    // storage for the inverted diagonal blocks of all fronts (cluster chains of the tall fronts, DINV chains of the
    // small ones; lu_solve.cu)
    std::vector<int64_t> doff((size_t)std::max(1, S.n_nodes), -1);
    int64_t tot = 0;
    for (int lev = 0; lev < S.n_levels; ++lev) {
      for (int q = S.level_start[lev]; q < S.level_start[lev + 1]; ++q) {
        const int node = S.level_nodes[q];
        const int npan = (S.ns[node] + PW - 1) / PW;
        if (npan == 0) continue;
        doff[node] = tot;
        tot += (int64_t)2 * npan * PW * PW;
      }
    }
    up64(d.dinv_off, doff, s);
    d.dinv_entries = tot;
  }
  {
    // storage for inv(L11), inv(U11) of the fronts of the upper levels (selective inversion, lu_tinv.cu)
    d.tinv_min_level_ns = getenv("FW_TINV_MIN_NS") ? atoi(getenv("FW_TINV_MIN_NS")) : TINV_DEFAULT_MIN_NS;
    std::vector<int64_t> toff((size_t)std::max(1, S.n_nodes), -1);
    int64_t tot = 0;
    for (int lev = 0; lev < S.n_levels; ++lev) {
      if (!tinv_level_static(S, d.tinv_min_level_ns, lev) || getenv("FW_NO_TINV")) continue;
      for (int q = S.level_start[lev]; q < S.level_start[lev + 1]; ++q) {
        const int node = S.level_nodes[q];
        if (S.ns[node] == 0) continue;
        toff[node] = tot;
        tot += (int64_t)2 * S.ns[node] * S.ns[node];
      }
    }
    up64(d.tinv_off, toff, s);
    d.tinv_entries = tot;
    d.tinv_valid = false;
  }
  {
    // dataflow substitution (lu_solve.cu): nodes deepest level first, one completion flag per node
    std::vector<int> order;
    order.reserve((size_t)S.n_nodes);
    for (int lev = S.n_levels - 1; lev >= 0; --lev)
      for (int q = S.level_start[lev]; q < S.level_start[lev + 1]; ++q) order.push_back(S.level_nodes[q]);
    if (order.empty()) order.push_back(0);
    up(d.order_fwd, order, s);
    d.tree_sync.alloc((size_t)S.n_nodes + 8);
    d.tree_sync.zero(s);
  }
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
}

// ------------------------------------------------------------------------------ scatter of OP into fronts
template <class S, class V>
__global__ void __launch_bounds__(256, 4) k_lu_scatter(int n, const int* __restrict__ indptr,
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
__global__ void __launch_bounds__(256, 4) k_lu_extend_add(FrontTab T, const int* __restrict__ nodes, int slot,
                                                       S* __restrict__ F) {
  const int p = nodes[blockIdx.x];
  const int c = slot ? T.child1[p] : T.child0[p];
  if (c < 0) return;
  const int mc = T.m[c], nsc = T.ns[c], nbc = mc - nsc;
  const int nt = (nbc + 31) >> 5;
  const int ti = blockIdx.y, tj = blockIdx.z;  // 2-D tile index (each <= 65535)
  if (ti >= nt || tj >= nt) return;
  const int mp = T.m[p];
  const int* rel = T.rel + T.blist_off[c];
  const S* Fc = F + T.front_off[c];
  S* Fp = F + T.front_off[p];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = ti * 32 + tx;
  if (i >= nbc) return;
  const int ri = rel[i];
  // the four entries of this thread: all loads first (the targets are distinct, which the compiler cannot see)
  long long dst[4];
  S add[4], old[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tj * 32 + ty + 8 * u;
    dst[u] = j < nbc ? ri + (long long)rel[j] * mp : -1;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tj * 32 + ty + 8 * u;
    if (dst[u] >= 0) {
      add[u] = Fc[(nsc + i) + (long long)(nsc + j) * mc];
      old[u] = Fp[dst[u]];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (dst[u] >= 0) Fp[dst[u]] = old[u] + add[u];
}

// ------------------------------------------------------------------------------ panel (pivot window) kernel
// WIN: rows of the shared-memory window = threads of the CTA (64 / 128 for the levels of small fronts, where the
// 66 KB window of 256 rows would leave three CTAs per SM with most of their threads idle)
template <class S, int WIN>
__global__ void __launch_bounds__(WIN, 1024 / WIN) k_lu_panel(FrontTab T, const int* __restrict__ nodes, int kstep,
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
  constexpr int LD = WIN + 1;
  S* a = reinterpret_cast<S*>(smem_raw);              // a[c*LD + r]
  __shared__ double red_v[WIN / 32];
  __shared__ int red_i[WIN / 32];
  __shared__ int s_piv;
  S* Fn = F + T.front_off[node];
  const int t = threadIdx.x;
  if (t < hw) {  // eight independent loads in flight per thread
    for (int c0 = 0; c0 < w; c0 += 8) {
      S v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = c0 + u < w ? Fn[(k0 + t) + (long long)(k0 + c0 + u) * m] : scalar_traits<S>::zero();
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < w) a[(c0 + u) * LD + t] = v[u];
    }
  }
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
      for (int q = 1; q < WIN / 32; ++q)
        if (red_v[q] > bv || (red_v[q] == bv && red_i[q] < bi)) bv = red_v[q], bi = red_i[q];
      s_piv = bi;
      piv[T.own_start[node] + k0 + j] = k0 + bi;
      if (!(bv >= thr)) {
        // static pivot perturbation (tiny / exactly singular pivot); corrected by iterative refinement
        a[j * LD + bi] = scalar_traits<S>::from(thr, 0.0);
        atomicAdd(&counters[0], 1);
        if (bv == 0.0) atomicAdd(&counters[1], 1);  // whole column exactly zero: SuperLU's "Factor is exactly singular"
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
        if (bv == 0.0) atomicAdd(&counters[1], 1);
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

// ------------------------------------------------------------------------------ tall panel on a thread-block cluster
// f64 panels of 256 < h <= 4096 rows: 8 CTAs x 512 threads, every thread owns ONE row of the panel and keeps its
// 32 entries in registers for the whole panel (the global-memory kernel above re-reads and re-writes the trailing
// columns for every pivot: ~200 us per panel at h = 2000 on one SM).  Per pivot column: block arg-max, candidates
// exchanged through distributed shared memory, cluster barrier, the pivot row (and the row it swaps with) broadcast
// to all CTAs, cluster barrier, rank-1 update in registers.  Same pivot choice as k_lu_panel_tall (largest |.|, lowest
// row on ties), same arithmetic per entry => identical factors.
namespace cgp = cooperative_groups;
constexpr int PCL = 8, PCL_THREADS = 512;

__global__ void __cluster_dims__(PCL, 1, 1) __launch_bounds__(PCL_THREADS)
    k_lu_panel_cluster(FrontTab T, const int* __restrict__ nodes, int kstep, double* __restrict__ F,
                       int* __restrict__ piv, double thr, int* __restrict__ counters) {
  __shared__ double prow[PW], jrow[PW], stage[2][PW];
  __shared__ double cand_v[PCL];
  __shared__ int cand_i[PCL];
  __shared__ double red_v[PCL_THREADS / 32];
  __shared__ int red_i[PCL_THREADS / 32];
  cgp::cluster_group cl = cgp::this_cluster();
  const int rank = (int)cl.block_rank();
  const int node = nodes[blockIdx.x / PCL];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = kstep * PW;
  const int h = ns - k0;
  if (k0 >= ns || h <= HWIN || h > PCL * PCL_THREADS) return;  // uniform over the cluster
  const int w = min(PW, h);
  double* P = F + T.front_off[node] + k0 + (long long)k0 * m;  // P[r + c*m]
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int r = rank * PCL_THREADS + t;  // the row of the panel owned by this thread
  const bool live = r < h;
  double a[PW];
#pragma unroll
  for (int c = 0; c < PW; ++c) a[c] = (live && c < w) ? P[r + (long long)c * m] : 0.0;
  for (int j = 0; j < w; ++j) {
    // ---- arg-max of |a[j]| over the rows j..h-1
    double aj = 0.0;
#pragma unroll
    for (int c = 0; c < PW; ++c)
      if (c == j) aj = a[c];
    double v = (live && r >= j) ? fabs(aj) : -1.0;
    int idx = r;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, v, o);
      const int oi = __shfl_down_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) v = ov, idx = oi;
    }
    if (lane == 0) red_v[wid] = v, red_i[wid] = idx;
    __syncthreads();
    if (t == 0) {
      double bv = red_v[0];
      int bi = red_i[0];
      for (int q = 1; q < PCL_THREADS / 32; ++q)
        if (red_v[q] > bv || (red_v[q] == bv && red_i[q] < bi)) bv = red_v[q], bi = red_i[q];
      for (int c = 0; c < PCL; ++c) {
        cl.map_shared_rank(cand_v, c)[rank] = bv;
        cl.map_shared_rank(cand_i, c)[rank] = bi;
      }
    }
    cl.sync();
    double bv = cand_v[0];
    int pr = cand_i[0];
#pragma unroll
    for (int q = 1; q < PCL; ++q)
      if (cand_v[q] > bv || (cand_v[q] == bv && cand_i[q] < pr)) bv = cand_v[q], pr = cand_i[q];
    const bool tiny = !(bv >= thr);
    if (rank == 0 && t == 0) {
      piv[T.own_start[node] + k0 + j] = k0 + pr;
      if (tiny) atomicAdd(&counters[0], 1);
      if (bv == 0.0) atomicAdd(&counters[1], 1);
    }
    // ---- broadcast the pivot row and row j (their owners stage the registers in shared memory, the owner's
    //      warp copies the 32 values into every CTA)
    const int own_pr_warp = (pr % PCL_THREADS) >> 5, own_pr_rank = pr / PCL_THREADS;
    const int own_j_warp = (j % PCL_THREADS) >> 5, own_j_rank = j / PCL_THREADS;
    if (r == pr) {
      // static pivot perturbation (tiny / exactly singular pivot); corrected by iterative refinement
#pragma unroll
      for (int c = 0; c < PW; ++c) stage[0][c] = (tiny && c == j) ? thr : a[c];
    }
    if (r == j) {
#pragma unroll
      for (int c = 0; c < PW; ++c) stage[1][c] = a[c];
    }
    __syncwarp();
    if (rank == own_pr_rank && wid == own_pr_warp) {
      const double x = stage[0][lane];
#pragma unroll
      for (int c = 0; c < PCL; ++c) cl.map_shared_rank(prow, c)[lane] = x;
    }
    if (rank == own_j_rank && wid == own_j_warp) {
      const double x = stage[1][lane];
#pragma unroll
      for (int c = 0; c < PCL; ++c) cl.map_shared_rank(jrow, c)[lane] = x;
    }
    cl.sync();
    // ---- interchange + rank-1 update in registers
    if (pr != j) {
      if (r == pr) {
#pragma unroll
        for (int c = 0; c < PW; ++c) a[c] = jrow[c];
      } else if (r == j) {
#pragma unroll
        for (int c = 0; c < PW; ++c) a[c] = prow[c];
      }
    } else if (r == j && tiny) {
#pragma unroll
      for (int c = 0; c < PW; ++c) a[c] = prow[c];
    }
    if (live && r > j) {
      double l = 0.0;
#pragma unroll
      for (int c = 0; c < PW; ++c) {
        if (c == j) {
          l = a[c] / prow[c];
          a[c] = l;
        } else if (c > j) {
          a[c] = fma(-l, prow[c], a[c]);
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < PW; ++c)
      if (c < w) P[r + (long long)c * m] = a[c];
  }
}

// ------------------------------------------------------------------------------ tall panel in ONE CTA, registers
// The cluster kernel above pays two hardware cluster barriers per pivot column (~4.7 us per column measured: the
// 2001 pivots of the root front alone cost 9.5 ms).  This kernel keeps the whole panel on one SM: 1024 threads, every
// thread owns R rows, the 32-column panel is processed as 32/SW LEFT-LOOKING sub-panels of SW columns that live in
// registers (R x SW entries per thread).  Per pivot column: warp arg-max, one block barrier, every warp reduces the 32
// warp candidates itself, the owner of the pivot row publishes its SW entries, second block barrier, rank-1 update
// in registers.  Pivoting is IMPLICIT while the panel is factored (a chosen row is frozen where it lies, so no row
// data moves between threads); before a sub-panel starts, the effect of the earlier pivots of this panel on its
// columns is applied at once (small unit-lower solve of the pivot rows in shared memory, then one pass over the
// thread's own L entries, which are L1/L2 hits).  At the end the <= 64 rows whose position changes are moved and
// the implicit order is converted into the LAPACK-style sequential interchanges the other kernels expect.
template <class S, int PR_THREADS, int R, int SW>
__global__ void __launch_bounds__(PR_THREADS, 1) k_lu_panel_regs(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                                 S* __restrict__ F, int* __restrict__ piv, double thr,
                                                                 int* __restrict__ counters) {
  __shared__ S prow[SW];
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  constexpr int NW = PR_THREADS / 32;
  __shared__ int s_orig[PW];      // panel-local row of pivot j
  __shared__ S Ub[PW][SW];        // pivot rows j < c0 in the columns of the current sub-panel
  __shared__ S Lpp[PW][PW + 1];   // Lpp[i][j]: multiplier of pivot row i for pivot j < i
  __shared__ int s_pos[PW], s_occ[PW], s_chosen[PW], s_src[2 * PW], s_dst[2 * PW];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = kstep * PW;
  const int h = ns - k0;
  if (k0 >= ns || h <= HWIN || h > R * PR_THREADS) return;
  S* P = F + T.front_off[node] + k0 + (long long)k0 * m;  // P[r + c*m]
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  int froz[R];  // index of the pivot this row became, 99 while it is still a candidate
#pragma unroll
  for (int q = 0; q < R; ++q) froz[q] = 99;
  for (int c0 = 0; c0 < PW; c0 += SW) {
    S a[R][SW];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = t + q * PR_THREADS;
#pragma unroll
      for (int c = 0; c < SW; ++c) a[q][c] = i < h ? P[i + (long long)(c0 + c) * m] : scalar_traits<S>::zero();
    }
    if (c0 > 0) {
      // ---- left-looking update of these SW columns by the pivots 0 .. c0-1 of this panel
#pragma unroll
      for (int q = 0; q < R; ++q)
        if (froz[q] < c0) {
#pragma unroll
          for (int c = 0; c < SW; ++c) Ub[froz[q]][c] = a[q][c];
        }
      for (int e = t; e < PW * PW; e += PR_THREADS) {
        const int i = e >> 5, j = e & 31;
        if (i < c0 && j < i) Lpp[i][j] = P[s_orig[i] + (long long)j * m];
      }
      __syncthreads();
      if (t < SW) {
        for (int j = 1; j < c0; ++j) {
          S u = Ub[j][t];
          for (int i = 0; i < j; ++i) fnma_(u, Lpp[j][i], Ub[i][t]);
          Ub[j][t] = u;
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int i = t + q * PR_THREADS;
        if (i >= h) continue;
        const int lim = min(froz[q], c0);
        int j = 0;
        constexpr int LU = (R * SW * (int)sizeof(S) >= 128) ? 4 : 8;  // loads in flight (register budget: 64 per thread)
        for (; j + LU <= lim; j += LU) {
          S l[LU];
#pragma unroll
          for (int u = 0; u < LU; ++u) l[u] = P[i + (long long)(j + u) * m];
#pragma unroll
          for (int u = 0; u < LU; ++u)
#pragma unroll
            for (int c = 0; c < SW; ++c) fnma_(a[q][c], l[u], Ub[j + u][c]);
        }
        for (; j < lim; ++j) {
          const S l = P[i + (long long)j * m];
#pragma unroll
          for (int c = 0; c < SW; ++c) fnma_(a[q][c], l, Ub[j][c]);
        }
      }
    }
    // rows frozen by an earlier sub-panel now hold their final U entries of these columns
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = t + q * PR_THREADS;
      if (i < h && froz[q] < c0) {
#pragma unroll
        for (int c = 0; c < SW; ++c) P[i + (long long)(c0 + c) * m] = a[q][c];
      }
    }
    // ---- right-looking elimination of the SW columns in registers.  The columns SHIFT left by one after every
    //      pivot, so the current column is always a[.][0]: the loop body is the same for every column (compact
    //      code: the fully unrolled variant was instruction-fetch bound) and all register indices stay static.
#pragma unroll 1
    for (int jj = 0; jj < SW; ++jj) {
      const int j = c0 + jj;
      double v = -1.0;
      int idx = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int i = t + q * PR_THREADS;
        if (i < h && froz[q] == 99) {
          const double av = abs1_(a[q][0]);
          if (av > v) v = av, idx = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oi = __shfl_down_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) v = ov, idx = oi;
      }
      if (lane == 0) red_v[wid] = v, red_i[wid] = idx;
      __syncthreads();
      double bv = lane < NW ? red_v[lane] : -1.0;
      int pr = lane < NW ? red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, pr, o);
        if (ov > bv || (ov == bv && oi < pr)) bv = ov, pr = oi;
      }
      const bool tiny = !(bv >= thr);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        if (t + q * PR_THREADS == pr) {
          froz[q] = j;
          // static pivot perturbation (tiny / exactly singular pivot); corrected by iterative refinement
          if (tiny) a[q][0] = scalar_traits<S>::from(thr, 0.0);
#pragma unroll
          for (int c = 0; c < SW; ++c) {
            prow[c] = a[q][c];
            if (c < SW - jj) P[pr + (long long)(j + c) * m] = a[q][c];  // the pivot row is final: U entries
          }
          s_orig[j] = pr;
          if (tiny) atomicAdd(&counters[0], 1);
          if (bv == 0.0) atomicAdd(&counters[1], 1);
        }
      }
      __syncthreads();
      const S inv = scalar_traits<S>::one() / prow[0];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int i = t + q * PR_THREADS;
        if (i < h && froz[q] == 99) {
          const S l = a[q][0] * inv;
          P[i + (long long)j * m] = l;
#pragma unroll
          for (int c = 1; c < SW; ++c) {
            S x = a[q][c];
            fnma_(x, l, prow[c]);
            a[q][c - 1] = x;
          }
          a[q][SW - 1] = scalar_traits<S>::zero();
        }
      }
    }
    __syncthreads();  // the sub-panel is visible to the whole CTA (Lpp gather, L reloads, final row moves)
  }
  // ---- implicit pivot order -> sequential interchanges (row j <-> piv[j]) and the final position of every moved row
  if (t < PW) s_pos[t] = t, s_occ[t] = t, s_chosen[t] = 0;
  __syncthreads();
  if (t == 0) {
    int* pv = piv + T.own_start[node] + k0;
    for (int j = 0; j < PW; ++j) {
      const int o = s_orig[j];
      const int cur = o < PW ? s_pos[o] : o;  // rows below the top block stay put until they are chosen
      if (o < PW) s_chosen[o] = 1;
      pv[j] = k0 + cur;
      const int d = s_occ[j];  // the original top row that currently sits at position j
      if (cur != j) {
        s_pos[d] = cur;
        if (cur < PW) s_occ[cur] = d;
      }
      s_occ[j] = -1;
    }
    for (int j = 0; j < PW; ++j) s_src[j] = s_orig[j], s_dst[j] = j;
    for (int d = 0; d < PW; ++d) {
      const bool moved = !s_chosen[d] && s_pos[d] != d;
      s_src[PW + d] = moved ? d : -1;
      s_dst[PW + d] = moved ? s_pos[d] : -1;
    }
  }
  __syncthreads();
  constexpr int MV = 2 * PW * PW / PR_THREADS;
  S mv[MV];
#pragma unroll
  for (int u = 0; u < MV; ++u) {
    const int e = t + u * PR_THREADS, who = e >> 5, c = e & 31;
    const int src = s_src[who];
    mv[u] = (src >= 0 && src != s_dst[who]) ? P[src + (long long)c * m] : scalar_traits<S>::zero();
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < MV; ++u) {
    const int e = t + u * PR_THREADS, who = e >> 5, c = e & 31;
    const int src = s_src[who], dst = s_dst[who];
    if (src >= 0 && src != dst) P[dst + (long long)c * m] = mv[u];
  }
}

// ------------------------------------------------------------------------------ swaps + TRSMs
// blockIdx.x < ctiles : thread per column c outside the panel: row interchanges, and for c right of the
//                       panel the unit-lower solve of the U row block.
// blockIdx.x >= ctiles: thread per row below the pivot window: X U_kk = F (L21 / L-panel rows).
// ob > 0 (two-level blocking, outer blocks of ob panels): the unit-lower solve is applied only to the columns of the
// current outer block; the U strip right of it is solved once per outer block by k_lu_outer_trsm.
// GATHER: the interchanges as one gather (two rounds of independent loads; 150 registers) -- levels of few fronts,
// where the chain of w dependent read-modify-writes through L2 is what the step waits for; the sequential form keeps
// the register count (and the occupancy) of the levels with thousands of small fronts.
template <class S, bool GATHER, int MINB = 0>
__global__ void __launch_bounds__(128, MINB) k_lu_swap_trsm(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                      S* __restrict__ F, const int* __restrict__ piv, int ctiles,
                                                      int ob) {
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
    const int jend = ob > 0 ? min(ns, (kstep / ob + 1) * ob * PW) : m;
    if constexpr (GATHER) {
    // The w sequential interchanges (row k0+j <-> sp[j]) as ONE gather: 64 threads trace the interchanges backwards
      // to find which original row ends at each of the w top positions (s_top) and at every partner position below
      // the top block (s_far; always an original top row).  A column then needs two rounds of independent loads
      // instead of a chain of w dependent read-modify-writes through L2.
      __shared__ int s_top[PW], s_far[PW];
      if (threadIdx.x < 2 * PW) {
        const int j = threadIdx.x & (PW - 1);
        int pos = threadIdx.x < PW ? k0 + j : (j < w ? sp[j] : -1);
        const bool wanted = j < w && (threadIdx.x < PW || pos >= k0 + w);
        if (wanted) {
          for (int i = w - 1; i >= 0; --i) {
            const int a = k0 + i, b = sp[i];
            pos = pos == a ? b : (pos == b ? a : pos);
          }
        }
        if (threadIdx.x < PW)
          s_top[j] = wanted ? pos : k0 + j;
        else
          s_far[j] = wanted ? pos : -1;
      }
      __syncthreads();
      const int c = tile * 128 + threadIdx.x;
      if (c >= m || (c >= k0 && c < k0 + w)) return;
      S* col = Fn + (long long)c * m;
      S x[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) x[j] = (j < w) ? col[s_top[j]] : scalar_traits<S>::zero();
#pragma unroll
      for (int ch = 0; ch < PW; ch += 8) {
        S y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int f = (ch + u < w) ? s_far[ch + u] : -1;
          y[u] = f >= 0 ? col[f] : scalar_traits<S>::zero();
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (ch + u < w && s_far[ch + u] >= 0) col[sp[ch + u]] = y[u];
      }
        if (c >= k0 + w && c < jend) {
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
      } else {
#pragma unroll
        for (int j = 0; j < PW; ++j)
          if (j < w && s_top[j] != k0 + j) col[k0 + j] = x[j];
      }
    } else {
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
      if (c >= k0 + w && c < jend) {
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

// ------------------------------------------------------------------------------ outer-block U strip
// Two-level blocking: after the ob panels of an outer block [j0, j0 + ow) are factored (their row interchanges are
// already applied to every column), the U strip right of the block is solved once with the unit-lower ow x ow block:
// thread per column, 32 rows at a time in registers, the L slab of the current 32 columns in shared memory.
template <class S>
__global__ void __launch_bounds__(128) k_lu_outer_trsm(FrontTab T, const int* __restrict__ nodes, int jblk, int ob,
                                                       S* __restrict__ F) {
  extern __shared__ __align__(16) unsigned char outer_smem[];
  S(*Ls)[PW + 1] = reinterpret_cast<S(*)[PW + 1]>(outer_smem);  // rows of the block from the current 32 columns down
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int j0 = jblk * ob * PW;
  if (j0 >= ns) return;
  const int ow = min(ob * PW, ns - j0);
  const int c = j0 + ow + blockIdx.y * 128 + threadIdx.x;
  if (j0 + ow + (int)blockIdx.y * 128 >= m) return;
  S* Fn = F + T.front_off[node];
  const bool live = c < m;
  S* col = Fn + (long long)(live ? c : 0) * m + j0;
  for (int s0 = 0; s0 < ow; s0 += PW) {
    const int w = min(PW, ow - s0);
    const int rows = ow - s0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * w; idx += 128) {
      const int r = idx % rows, cc = idx / rows;
      Ls[r][cc] = Fn[(j0 + s0 + r) + (long long)(j0 + s0 + cc) * m];
    }
    __syncthreads();
    if (!live) continue;
    S x[PW];
#pragma unroll
    for (int j = 0; j < PW; ++j) x[j] = (j < w) ? col[s0 + j] : scalar_traits<S>::zero();
#pragma unroll
    for (int j = 1; j < PW; ++j) {
      if (j < w) {
#pragma unroll
        for (int i = 0; i < j; ++i) fnma_(x[j], Ls[j][i], x[i]);
      }
    }
#pragma unroll
    for (int j = 0; j < PW; ++j)
      if (j < w) col[s0 + j] = x[j];
    int r = w;
    for (; r + 8 <= rows; r += 8) {  // eight independent read-modify-writes in flight
      S acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = col[s0 + r + u];
#pragma unroll
      for (int j = 0; j < PW; ++j)
        if (j < w) {
#pragma unroll
          for (int u = 0; u < 8; ++u) fnma_(acc[u], Ls[r + u][j], x[j]);
        }
#pragma unroll
      for (int u = 0; u < 8; ++u) col[s0 + r + u] = acc[u];
    }
    for (; r < rows; ++r) {
      S acc = col[s0 + r];
#pragma unroll
      for (int j = 0; j < PW; ++j)
        if (j < w) fnma_(acc, Ls[r][j], x[j]);
      col[s0 + r] = acc;
    }
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

// SCHUR = false: per-panel rank-32 update of everything that feeds later panels (all trailing entries with a
//                 row or a column inside the pivot block); the boundary x boundary block S22 is skipped.
// SCHUR = true : once per front after its last panel, S22 -= L21 U12 with the full inner dimension ns, so the
//                 Schur complement (the bulk of the flops for nb >> ns) is read and written once.
// ob > 0: inner update of two-level blocking -- only the columns of the current outer block of ob panels.
template <class S, bool SCHUR, int MINB = 0>
__global__ void __launch_bounds__(256, MINB) k_lu_gemm(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                       S* __restrict__ F, int ob = 0, int kw = PW) {
  constexpr int TM = 64, KC = GemmCfg<S>::KC;
  __shared__ S As[KC][TM];
  __shared__ S Bs[KC][TM + 1];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = SCHUR ? 0 : kstep * kw;  // kw: one panel, or an outer block of panels
  if (k0 >= ns) return;
  const int w = SCHUR ? ns : min(kw, ns - k0);
  const int r0 = SCHUR ? ns : k0 + w;
  const int nrem = m - r0;
  if (nrem <= 0) return;
  const int nt = (nrem + TM - 1) / TM;
  const int ti = blockIdx.y, tj = blockIdx.z;
  if (ti >= nt || tj >= nt) return;
  if (!SCHUR && r0 + ti * TM >= ns && r0 + tj * TM >= ns) return;  // tile entirely inside S22
  const int cend = (!SCHUR && ob > 0) ? min(ns, (kstep / ob + 1) * ob * PW) : m;  // exclusive column limit (kw = PW)
  if (r0 + tj * TM >= cend) return;
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
      Bs[k][c] = (gc < cend && gk < w) ? Fn[(k0 + gk) + (long long)gc * m] : scalar_traits<S>::zero();
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
  // C += acc: all loads of the thread tile before the first store (one read-modify-write per entry makes every entry
  // wait a full memory round trip -- the compiler must assume that a store aliases the next load)
  S old[4][4];
  bool ok[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gc = cbase + ty * 4 + j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = rbase + tx * 4 + i;
      ok[i][j] = gc < cend && gr < m && (SCHUR || gr < ns || gc < ns);
      old[i][j] = ok[i][j] ? Fn[gr + (long long)gc * m] : scalar_traits<S>::zero();
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gc = cbase + ty * 4 + j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = rbase + tx * 4 + i;
      if (ok[i][j]) Fn[gr + (long long)gc * m] = old[i][j] + acc[i][j];
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

// ------------------------------------------------------------------------------ inverted diagonal blocks
// One warp per (front, panel): X = inv(L_kk) (unit lower) and Y = inv(U_kk) of the PW x PW diagonal block,
// lane c computes column c by substitution (the block is read from shared memory with broadcast loads, the
// column lives in registers).  Written column-major to a side buffer -- the factors themselves stay untouched.
// The cluster substitution kernels (lu_solve.cu) then replace the 32 dependent steps of every diagonal solve
// by one 32 x 32 matrix-vector product.
template <class S>
__global__ void __launch_bounds__(32) k_lu_diag_inverse(FrontTab T, const int* __restrict__ nodes,
                                                        const S* __restrict__ F, S* __restrict__ dinv) {
  using Tr = scalar_traits<S>;
  const int node = nodes[blockIdx.x];
  const int ns = T.ns[node], m = T.m[node];
  const int k0 = blockIdx.y * PW;
  const long long off = T.dinv_off[node];
  if (k0 >= ns || off < 0) return;
  const int npan = (ns + PW - 1) / PW;
  const int wd = min(PW, ns - k0);
  const S* Fn = F + T.front_off[node];
  __shared__ S D[PW][PW + 1];
  __shared__ S R[PW];
  const int lane = threadIdx.x;
  for (int c = 0; c < PW; ++c)
    D[lane][c] = (lane < wd && c < wd) ? Fn[(k0 + lane) + (long long)(k0 + c) * m] : (lane == c ? Tr::one() : Tr::zero());
  __syncwarp();
  R[lane] = Tr::one() / D[lane][lane];
  __syncwarp();
  const int c = lane;
  S x[PW];
  // inv(L): x_c = 1, x_i = -sum_{j<i} L_ij x_j
#pragma unroll
  for (int i = 0; i < PW; ++i) {
    S sacc = Tr::zero();
#pragma unroll
    for (int j = 0; j < i; ++j) fnma_(sacc, D[i][j], x[j]);
    x[i] = i < c ? Tr::zero() : (i == c ? Tr::one() : sacc);
  }
  S* outL = dinv + off + (long long)blockIdx.y * PW * PW + (long long)c * PW;
#pragma unroll
  for (int i = 0; i < PW; ++i) outL[i] = x[i];
  // inv(U): y_c = 1/U_cc, y_i = -(sum_{j>i} U_ij y_j) / U_ii
#pragma unroll
  for (int i = PW - 1; i >= 0; --i) {
    S sacc = Tr::zero();
#pragma unroll
    for (int j = i + 1; j < PW; ++j) fnma_(sacc, D[i][j], x[j]);
    x[i] = i > c ? Tr::zero() : (i == c ? R[i] : sacc * R[i]);
  }
  S* outU = dinv + off + (long long)(npan + blockIdx.y) * PW * PW + (long long)c * PW;
#pragma unroll
  for (int i = 0; i < PW; ++i) outU[i] = x[i];
}

// occupancy hint of the SIMT GEMM selected at run time (FW_GEMM_MINB = 0 | 3 | 4; default 4: 64 registers, 4 blocks per SM,
// -2 ms per config-2 factorisation, profiles/factor_hints_r02.log)
#define LU_GEMM_ARGS_(...) __VA_ARGS__
#define LU_GEMM_LAUNCH(SCH, CFG, ARGS)                                                  \
  do {                                                                                  \
    if (gemm_minb == 4) k_lu_gemm<S, SCH, 4><<<LU_GEMM_ARGS_ CFG>>>(LU_GEMM_ARGS_ ARGS);        \
    else if (gemm_minb == 3) k_lu_gemm<S, SCH, 3><<<LU_GEMM_ARGS_ CFG>>>(LU_GEMM_ARGS_ ARGS);   \
    else k_lu_gemm<S, SCH><<<LU_GEMM_ARGS_ CFG>>>(LU_GEMM_ARGS_ ARGS);                          \
  } while (0)

template <class S, class V>
static void factor_typed(Handle& h, const Values& VA, const Values& VB, S alphaA, S alphaB) {
  LU& lu = *h.lu;
  LUDevice& d = *lu.dev;
  const LUSymbolic& Sy = lu.sym;
  cudaStream_t s = h.stream;
  const size_t wsc = sizeof(S) / sizeof(double);
  lu.drop_graphs();  // captured solves refer to the previous factorisation (scalar type, buffers)
  lu.no_tinv = false;
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
                                                                  (const V*)VA.v.p, (const V*)VB.v.p, alphaA, alphaB,
                                                                  T, d.iperm.p, d.node_of.p, F);
  FW_CHECK_CUDA(cudaGetLastError());
  // pivot threshold relative to the largest entry of the pencil
  DevBuf<double> amax;
  amax.alloc(2);
  amax.zero(s);
  k_absmax<V><<<296, 256, 0, s>>>((const V*)VA.v.p, pat.nnz, amax.p);
  k_absmax<V><<<296, 256, 0, s>>>((const V*)VB.v.p, pat.nnz, amax.p + 1);
  double am[2] = {0, 0};
  amax.download(am, 2, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  const double scale = absv_(alphaA) * am[0] + absv_(alphaB) * am[1];
  const double thr = 1e-13 * (scale > 0 ? scale : 1.0);
  auto panel_smem = [](int win) { return (size_t)PW * (win + 1) * sizeof(S); };
  FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_panel<S, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem(256)));
  FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_panel<S, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem(128)));
  int64_t launches = 3;
  // trailing blocks at least this wide use the 128x128 register-tiled GEMM (FW_BIG_GEMM_MIN to experiment)
  const char* tenv = getenv("FW_TRACE");
  const bool trace2 = tenv && atoi(tenv) >= 2;
  // (the tensor-core GEMM has a 64 x 64 tile variant for trailing blocks below ~900 rows, lu_gemm.cu)
  static const int big_gemm_min = getenv("FW_BIG_GEMM_MIN") ? atoi(getenv("FW_BIG_GEMM_MIN")) : 48;
  static const int gemm_minb_env = getenv("FW_GEMM_MINB") ? atoi(getenv("FW_GEMM_MINB")) : 4;
  const int gemm_minb = scalar_traits<S>::is_complex ? 0 : gemm_minb_env;  // the c128 kernel uses 128 registers
  static const int swap_gather_max_fronts =
      getenv("FW_SWAP_GATHER_MAX_FRONTS") ? atoi(getenv("FW_SWAP_GATHER_MAX_FRONTS")) : 4096;
  static const bool no_cluster_panel = getenv("FW_NO_CLUSTER_PANEL") != nullptr;
  static const bool no_regs_panel = getenv("FW_NO_REGS_PANEL") != nullptr;
  // two-level blocking (outer blocks of OB panels) for levels whose pivot blocks span more than one panel
  constexpr int OB = 4;
  static const int blocked_min_ns = getenv("FW_BLOCKED_MIN_NS") ? atoi(getenv("FW_BLOCKED_MIN_NS")) : 33;
  static const int big_gemm_min_outer = getenv("FW_BIG_GEMM_MIN_OUTER") ? atoi(getenv("FW_BIG_GEMM_MIN_OUTER")) : 96;
  const size_t outer_smem = sizeof(S) * OB * PW * (PW + 1);
  {
    static bool configured[64] = {};
    if (first_use_on_device(configured))
      FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_outer_trsm<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)outer_smem));
  }
  constexpr int RSW2 = scalar_traits<S>::is_complex ? 4 : 8, RSW4 = scalar_traits<S>::is_complex ? 2 : 4;
  // measured at 500k triangles: the level of 16 fronts (ns ~ 1000) took 15.7 ms with the global-memory tall-panel
  // kernel against 3.8 ms for the level of 8 fronts of half the width on the cluster kernel
  static const int cluster_panel_max_fronts =
      getenv("FW_CLUSTER_PANEL_MAX_FRONTS") ? atoi(getenv("FW_CLUSTER_PANEL_MAX_FRONTS")) : 64;
  // per-panel rank-32 updates: the 64x64 tile kernel (3+ CTAs per SM) beats the 128x128 one (244 registers, one CTA
  // per SM, only two K-chunks to pipeline): 69.9 vs 76.5 ms over the nine top levels; the big kernel keeps the Schur
  // updates (K = ns)
  static const int big_gemm_min_panel =
      getenv("FW_BIG_GEMM_MIN_PANEL") ? atoi(getenv("FW_BIG_GEMM_MIN_PANEL")) : (1 << 30);
  // Experiment (FW_FACTOR_L2_MB > 0, off by default): levels of thousands of small fronts processed in CHUNKS of fronts
  // that fit the L2 cache, so that the ~17 kernels of a level find their fronts in L2 instead of streaming the level
  // from HBM every time (the leaf level of config 2 is 5.2 GB).  Measured SLOWER: 127 -> 145 ms (96 MB chunks) .. 189 ms
  // (16 MB): every kernel is bound by the dependent steps inside its CTAs (~30 us), a chunk of 255 fronts is one
  // such wave per launch -- the level needs many independent CTAs in flight, not cache residency.
  static const int64_t l2_budget =
      (int64_t)(getenv("FW_FACTOR_L2_MB") ? atoi(getenv("FW_FACTOR_L2_MB")) : 0) << 20;
  for (int lev = Sy.n_levels - 1; lev >= 0; --lev) {
    const int nfront_all = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes_all = d.level_nodes.p + Sy.level_start[lev];
    const int64_t front_bytes = (int64_t)Sy.level_max_m[lev] * Sy.level_max_m[lev] * (int64_t)sizeof(S);
    int chunk = nfront_all;
    if (l2_budget > 0 && front_bytes * nfront_all > 2 * l2_budget)
      chunk = (int)std::max<int64_t>(148, l2_budget / std::max<int64_t>(1, front_bytes));
    const bool chunked = chunk < nfront_all;
   for (int c0 = 0; c0 < nfront_all; c0 += chunk) {
    const int nfront = std::min(chunk, nfront_all - c0);
    const int* nodes = nodes_all + c0;
    if (lev + 1 < Sy.n_levels) {
      const int cnb = Sy.level_max_nb[lev + 1];
      const int nt = (cnb + 31) / 32;
      if (nt > 0) {
        dim3 grid(nfront, nt, nt);
        k_lu_extend_add<S><<<grid, 256, 0, s>>>(T, nodes, 0, F);
        k_lu_extend_add<S><<<grid, 256, 0, s>>>(T, nodes, 1, F);
        launches += 2;
      }
    }
    const int max_ns = Sy.level_max_ns[lev], max_m = Sy.level_max_m[lev];
    const int npan = (max_ns + PW - 1) / PW;
    // small fronts: one CTA walks through all steps of its front (lu_front_small.cu)
    const bool fused_small = npan > 0 && lu_front_small_applies(max_ns, max_m);
    if (fused_small) {
      lu_front_small_launch<S>(T, nodes, nfront, F, d.piv.p, thr, d.counters.p, s);
      launches += 1;
    }
    for (int k = 0; k < (fused_small ? 0 : npan); ++k) {
      if (max_ns - k * PW > HWIN) {
        bool done = false;
        {
          // one CTA per front, the panel in registers (implicit pivoting): any number of fronts, up to 4096 rows
          const int hmax = max_ns - k * PW;
          // register budget: R * SW entries per thread; 512-thread CTAs (128 registers) for the taller panels
          if (!no_regs_panel && hmax <= 4096) {
            if (hmax <= 1024)
              k_lu_panel_regs<S, 1024, 1, 8><<<nfront, 1024, 0, s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
            else if (hmax <= 2048)
              k_lu_panel_regs<S, 512, 4, RSW2><<<nfront, 512, 0, s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
            else
              k_lu_panel_regs<S, 512, 8, RSW4><<<nfront, 512, 0, s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
            done = true;
          }
        }
        if constexpr (!scalar_traits<S>::is_complex) {
          if (!done)
          // few, tall f64 fronts (top of the tree): one 8-CTA cluster per front, the panel lives in registers
          if (nfront <= cluster_panel_max_fronts && max_ns - k * PW <= PCL * PCL_THREADS && !no_cluster_panel) {
            k_lu_panel_cluster<<<nfront * PCL, PCL_THREADS, 0, s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
            done = true;
          }
        }
        if (!done) k_lu_panel_tall<S><<<nfront, TALL_THREADS, 0, s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
      }
      {
        const int hmax = max_ns - k * PW;  // tallest remaining pivot block of the level
        if (hmax <= 64)
          k_lu_panel<S, 64><<<nfront, 64, panel_smem(64), s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
        else if (hmax <= 128)
          k_lu_panel<S, 128><<<nfront, 128, panel_smem(128), s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
        else
          k_lu_panel<S, 256><<<nfront, 256, panel_smem(256), s>>>(T, nodes, k, F, d.piv.p, thr, d.counters.p);
      }
      const int ctiles = (max_m + 127) / 128;
      const int rtiles = (max_m + 127) / 128;
      // two-level blocking: rank-32 updates stay inside the current outer block of OB panels; the rest of the front
      // gets one rank-(32 OB) update per outer block (the 128 x 128 tensor-core GEMM needs a deep inner dimension)
      const int ob = (max_ns >= blocked_min_ns) ? OB : 0;
      {
        const dim3 grid(nfront, ctiles + rtiles);
        // occupancy hint: 6 blocks of 128 threads (80 registers, a few spills) measured 2 ms faster per factorisation than the
        // unconstrained 153-register build (profiles/factor_hints_r02.log); FW_SWAP_MINB=0 restores it
        // (real arithmetic only: the c128 instantiation needs 234 registers and spills badly under the cap -- config-3 probe:
        // factor 1.93 -> 2.00 s)
        static const int swap_minb_env = getenv("FW_SWAP_MINB") ? atoi(getenv("FW_SWAP_MINB")) : 6;
        const int swap_minb = scalar_traits<S>::is_complex ? 0 : swap_minb_env;
        if (nfront_all <= swap_gather_max_fronts) {
          if (swap_minb == 4) k_lu_swap_trsm<S, true, 4><<<grid, 128, 0, s>>>(T, nodes, k, F, d.piv.p, ctiles, ob);
          else if (swap_minb == 6) k_lu_swap_trsm<S, true, 6><<<grid, 128, 0, s>>>(T, nodes, k, F, d.piv.p, ctiles, ob);
          else k_lu_swap_trsm<S, true><<<grid, 128, 0, s>>>(T, nodes, k, F, d.piv.p, ctiles, ob);
        } else {
          if (swap_minb == 4) k_lu_swap_trsm<S, false, 4><<<grid, 128, 0, s>>>(T, nodes, k, F, d.piv.p, ctiles, ob);
          else if (swap_minb == 6) k_lu_swap_trsm<S, false, 6><<<grid, 128, 0, s>>>(T, nodes, k, F, d.piv.p, ctiles, ob);
          else k_lu_swap_trsm<S, false><<<grid, 128, 0, s>>>(T, nodes, k, F, d.piv.p, ctiles, ob);
        }
      }
      const int rem = max_m - (k * PW);  // upper bound of the trailing size
      const int nt = (rem + 63) / 64;
      launches += 3;
      if (ob == 0) {
        if (rem >= big_gemm_min_panel)
          lu_gemm_big_launch<S>(T, nodes, nfront, k, F, rem, false, s);
        else if (nt > 0)
          LU_GEMM_LAUNCH(false, (dim3(nfront, nt, nt), 256, 0, s), (T, nodes, k, F, 0));
        continue;
      }
      const bool last_inner = (k % ob == ob - 1) || k == npan - 1;
      if (!last_inner) {
        const int ct = (ob * PW - PW + 63) / 64;  // column tiles inside the outer block
        if (nt > 0) LU_GEMM_LAUNCH(false, (dim3(nfront, nt, ct), 256, 0, s), (T, nodes, k, F, ob));
      } else {
        const int jb = k / ob;
        const int orem = max_m - jb * ob * PW;  // columns right of the block start (upper bound)
        const int otiles = (orem + 127) / 128;
        k_lu_outer_trsm<S><<<dim3(nfront, otiles), 128, outer_smem, s>>>(T, nodes, jb, ob, F);
        if (orem >= big_gemm_min_outer)
          lu_gemm_big_launch<S>(T, nodes, nfront, jb, F, orem, false, s, ob * PW);
        else
          LU_GEMM_LAUNCH(false, (dim3(nfront, (orem + 63) / 64, (orem + 63) / 64), 256, 0, s), (T, nodes, jb, F, 0, ob * PW));
        launches += 1;
      }
    }
    if (trace2 && !chunked)
      tr.mark(("  factor lev " + std::to_string(lev) + " panels (" + std::to_string(npan) + " steps, " +
               std::to_string(nfront) + " fronts, max m " + std::to_string(max_m) + ")").c_str());
    {
      const int nt = (Sy.level_max_nb[lev] + 63) / 64;
      if (nt > 0 && npan > 0 && !fused_small) {
        if (Sy.level_max_nb[lev] >= big_gemm_min)
          lu_gemm_big_launch<S>(T, nodes, nfront, 0, F, Sy.level_max_nb[lev], true, s);
        else
          LU_GEMM_LAUNCH(true, (dim3(nfront, nt, nt), 256, 0, s), (T, nodes, 0, F));
        launches += 1;
      }
    }
    if (trace2 && !chunked) tr.mark(("  factor lev " + std::to_string(lev) + " schur gemm").c_str());
   }  // chunks
    if (trace2 && chunked)
      tr.mark(("  factor lev " + std::to_string(lev) + " panels + schur gemm in chunks of " + std::to_string(chunk) +
               " of " + std::to_string(nfront_all) + " fronts").c_str());
  }
  tr.mark("lu factor: levels");
  k_lu_compose_perm<<<cdiv(Sy.n_nodes, 64), 64, 0, s>>>(T, Sy.n_nodes, d.piv.p, d.rowperm.p);
  tr.mark("lu factor: compose perm");
  launches += 1;
  d.dinv_valid = false;
  if (d.dinv_entries > 0 && !scalar_traits<S>::is_complex) {
    d.dinv.alloc((size_t)d.dinv_entries * wsc);
    for (int lev = 0; lev < Sy.n_levels; ++lev) {
      const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
      const int npan = (Sy.level_max_ns[lev] + PW - 1) / PW;
      if (npan == 0) continue;
      k_lu_diag_inverse<S><<<dim3(nfront, npan), 32, 0, s>>>(T, d.level_nodes.p + Sy.level_start[lev], F, (S*)d.dinv.p);
      ++launches;
    }
    d.dinv_valid = true;
    tr.mark("lu factor: diagonal-block inverses");
  }
  d.tinv_valid = false;
  if constexpr (!scalar_traits<S>::is_complex) {
    if (d.tinv_entries > 0) {
      tinv_build(h, d, Sy);
      tr.mark("lu factor: pivot-block inverses (selective inversion)");
    }
  }
  FW_CHECK_CUDA(cudaGetLastError());
  int cnt[4] = {0, 0, 0, 0};
  d.counters.download(cnt, 4, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  lu.perturbed = cnt[0];
  if (cnt[1] > 0) {
    lu.factored = false;
    throw Error(FW_ERR_SINGULAR, "Factor is exactly singular");  // the message of scipy's splu (SuperLU) for this case
  }
  lu.factored = true;
  h.tim.kernel_launches += launches;
}

void lu_factor_values(Handle& h, const Values& VA, const Values& VB, double aAr, double aAi, double aBr, double aBi,
                      bool complex_arith) {
  FW_REQUIRE(h.lu && h.lu->dev, "lu_analyse must run first");
  FW_REQUIRE(VA.valid && VB.valid && VA.is_complex == VB.is_complex, "operators must be assembled (same scalar type)");
  LU& lu = *h.lu;
  const bool vals_complex = VA.is_complex;
  if (vals_complex || aAi != 0.0 || aBi != 0.0) complex_arith = true;
  lu.is_complex = complex_arith;
  if (!complex_arith)
    factor_typed<double, double>(h, VA, VB, aAr, aBr);
  else if (!vals_complex)
    factor_typed<cplx, double>(h, VA, VB, cplx{aAr, aAi}, cplx{aBr, aBi});
  else
    factor_typed<cplx, cplx>(h, VA, VB, cplx{aAr, aAi}, cplx{aBr, aBi});
}

void lu_factor(Handle& h, double aAr, double aAi, double aBr, double aBi, bool complex_arith) {
  FW_REQUIRE(h.A.valid && h.B.valid, "A and B must be assembled");
  lu_factor_values(h, h.A, h.B, aAr, aAi, aBr, aBi, complex_arith);
}

// ---- debug export (tests only): factor arena and pivot vector
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
