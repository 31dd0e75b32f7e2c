// Level-batched forward/backward substitution with the multifrontal factors (device).  See lu.h.
// Replaces the two triangular solves per ARPACK iteration of scipy's SuperLU object
// (scipy arpack.py:1011-1019, ido==1 branch :817-823).
//
// Four kernels per tree level:
//   forward : chain  (assemble w from the rhs and the children, apply the composed row permutation,
//                     unit-lower solve of the ns x ns pivot block, 32 columns at a time)
//             update (fronts x 128-row tiles: w_boundary -= L21 x1 -- the bulk of the L traffic)
//   backward: update (fronts x 128-row tiles: gather x2 from the parent, y1 = w1 - U12 x2)
//             chain  (right-looking upper solve of the pivot block, write x)
// The chain is the only sequential part (ns/32 dependent steps).  Fronts with a tall pivot block
// (top of the tree: ns up to a few thousand, only a handful of fronts per level) run it on a
// thread-block CLUSTER of 8 CTAs -- every CTA owns a slice of the rows, the 32x32 diagonal solve is
// replicated per CTA, one hardware cluster barrier per step -- instead of a single CTA.
// HBM-bound work (each factor entry is read once per solve); no tensor cores.
#include <cooperative_groups.h>

#include <algorithm>

#include "lu_dev.h"

namespace cg = cooperative_groups;

namespace fw {

constexpr int SOLVE_MAX_THREADS = 512;
constexpr int UPD_ROWS = 128;
constexpr int UPD_CHUNK = 256;
constexpr int CL = 8;             // CTAs per cluster in the tall-front chain kernels
constexpr int CL_THREADS = 512;
constexpr int CL_MIN_NS = 384;    // levels whose largest pivot block exceeds this use the cluster kernels

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

template <class S>
__device__ inline S shfl_bcast_(S v, int src);
template <>
__device__ inline double shfl_bcast_<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <>
__device__ inline cplx shfl_bcast_<cplx>(cplx v, int src) {
  return {__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src)};
}

// acc[r] -= sum_j col[j*m] * xs[r][j]: the strided factor entries of one row are fetched 16 at a time so
// that 16 independent HBM requests are in flight per thread (the chain is latency bound, not bandwidth bound)
template <class S, int NRHS, int LD>
__device__ __forceinline__ void row_dot_sub(S (&acc)[NRHS], const S* __restrict__ col, long long m, int cnt,
                                            S (&xs)[NRHS][LD]) {
  int j = 0;
  for (; j + 16 <= cnt; j += 16) {
    S l[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) l[q] = col[(long long)(j + q) * m];
#pragma unroll
    for (int q = 0; q < 16; ++q)
#pragma unroll
      for (int r = 0; r < NRHS; ++r) fnma_(acc[r], l[q], xs[r][j + q]);
  }
  for (; j < cnt; ++j) {
    const S l = col[(long long)j * m];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) fnma_(acc[r], l, xs[r][j]);
  }
}

// Synchronisation policy of the chain kernels: one CTA (block barrier) or one cluster (cluster barrier).
struct SyncCta {
  __device__ static int rank() { return 0; }
  __device__ static int nctas() { return 1; }
  __device__ static int front(int bx) { return bx; }
  __device__ static void sync() { __syncthreads(); }
};
struct SyncCluster {
  __device__ static int rank() { return (int)cg::this_cluster().block_rank(); }
  __device__ static int nctas() { return CL; }
  __device__ static int front(int bx) { return bx / CL; }
  // barrier.cluster arrive.release / wait.acquire: orders the global-memory writes of all CTAs of the
  // cluster (and invalidates L1) before anyone proceeds
  __device__ static void sync() { cg::this_cluster().sync(); }
};

// ------------------------------------------------------------------------------ forward chain
template <class S, int NRHS, class SY>
__device__ __forceinline__ void fwd_chain_body(FrontTab T, const int* __restrict__ nodes, const S* __restrict__ F,
                                               const int* __restrict__ rowperm, S* work, int64_t work_stride, S* xp,
                                               int na) {
  __shared__ S D[PW][PW + 1];
  __shared__ S xs[NRHS][PW];
  const int node = nodes[SY::front(blockIdx.x)];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  const S* Fn = F + T.front_off[node];
  S* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x;
  const int nth = blockDim.x * SY::nctas();
  const int gt = SY::rank() * blockDim.x + t;
  for (int i = gt; i < m; i += nth)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][i] = i < ns ? xp[(size_t)r * na + os + i] : scalar_traits<S>::zero();
  SY::sync();
  for (int sl = 0; sl < 2; ++sl) {
    const int c = sl ? T.child1[node] : T.child0[node];
    if (c < 0) continue;
    const int nsc = T.ns[c], nbc = T.m[c] - nsc;
    const int* rel = T.rel + T.blist_off[c];
    for (int i = gt; i < nbc; i += nth) {
      const int ri = rel[i];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        const S* wc = work + (size_t)r * work_stride + T.work_off[c];
        w[r][ri] = w[r][ri] + wc[nsc + i];
      }
    }
    SY::sync();
  }
  if (ns == 0) return;
  // row interchanges of the whole front first: stage the assembled own part in xp, gather it back
  for (int i = gt; i < ns; i += nth)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) xp[(size_t)r * na + os + i] = w[r][i];
  SY::sync();
  for (int i = gt; i < ns; i += nth) {
    const int src = rowperm[os + i];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][i] = xp[(size_t)r * na + os + src];
  }
  SY::sync();
  for (int k0 = 0; k0 < ns; k0 += PW) {
    const int wd = min(PW, ns - k0);
    for (int idx = t; idx < wd * wd; idx += blockDim.x) {
      int r = idx % wd, c = idx / wd;
      D[r][c] = Fn[(k0 + r) + (long long)(k0 + c) * m];
    }
    __syncthreads();
    if (t < 32 * NRHS) {  // replicated in every CTA of a cluster; rank 0 stores the result
      const int r = t >> 5, lane = t & 31;
      S x = lane < wd ? w[r][k0 + lane] : scalar_traits<S>::zero();
      for (int j = 0; j < wd; ++j) {
        const S xj = shfl_bcast_<S>(x, j);
        if (lane > j && lane < wd) fnma_(x, D[lane][j], xj);
      }
      if (lane < wd) xs[r][lane] = x;
    }
    __syncthreads();
    // remaining own rows only; the boundary rows are updated by k_fwd_update
    for (int row = k0 + wd + gt; row < ns; row += nth) {
      S acc[NRHS];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) acc[r] = w[r][row];
      row_dot_sub<S, NRHS, PW>(acc, Fn + row + (long long)k0 * m, m, wd, xs);
#pragma unroll
      for (int r = 0; r < NRHS; ++r) w[r][row] = acc[r];
    }
    SY::sync();
    // x_k is stored only now: the other CTAs of a cluster have read the old w[k0:k0+wd) by this point
    if (SY::rank() == 0 && t < 32 * NRHS && (t & 31) < wd) w[t >> 5][k0 + (t & 31)] = xs[t >> 5][t & 31];
  }
}

template <class S, int NRHS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS) k_fwd_chain(FrontTab T, const int* __restrict__ nodes,
                                                                 const S* __restrict__ F,
                                                                 const int* __restrict__ rowperm, S* work,
                                                                 int64_t work_stride, S* xp, int na) {
  fwd_chain_body<S, NRHS, SyncCta>(T, nodes, F, rowperm, work, work_stride, xp, na);
}
template <class S, int NRHS>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CL_THREADS)
    k_fwd_chain_cluster(FrontTab T, const int* __restrict__ nodes, const S* __restrict__ F,
                        const int* __restrict__ rowperm, S* work, int64_t work_stride, S* xp, int na) {
  fwd_chain_body<S, NRHS, SyncCluster>(T, nodes, F, rowperm, work, work_stride, xp, na);
}

// ------------------------------------------------------------------------------ backward chain (right-looking)
template <class S, int NRHS, class SY>
__device__ __forceinline__ void bwd_chain_body(FrontTab T, const int* __restrict__ nodes, const S* __restrict__ F,
                                               S* work, int64_t work_stride, S* xp, int na) {
  __shared__ S D[PW][PW + 1];
  __shared__ S xs[NRHS][PW];
  const int node = nodes[SY::front(blockIdx.x)];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  if (ns == 0) return;
  const S* Fn = F + T.front_off[node];
  S* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x;
  const int nth = blockDim.x * SY::nctas();
  const int gt = SY::rank() * blockDim.x + t;
  const int npan = (ns + PW - 1) / PW;
  for (int kp = npan - 1; kp >= 0; --kp) {
    const int k0 = kp * PW;
    const int wd = min(PW, ns - k0);
    for (int idx = t; idx < wd * wd; idx += blockDim.x) {
      int r = idx % wd, c = idx / wd;
      const S v = Fn[(k0 + r) + (long long)(k0 + c) * m];
      // reciprocal of the diagonal here (32 threads in parallel) keeps the divisions out of the serial chain
      D[r][c] = (r == c) ? (scalar_traits<S>::one() / v) : v;
    }
    __syncthreads();
    if (t < 32 * NRHS) {
      const int r = t >> 5, lane = t & 31;
      S y = lane < wd ? w[r][k0 + lane] : scalar_traits<S>::zero();
      for (int j = wd - 1; j >= 0; --j) {
        S xj = scalar_traits<S>::zero();
        if (lane == j) {
          y = y * D[j][j];
          xj = y;
        }
        xj = shfl_bcast_<S>(xj, j);
        if (lane < j) fnma_(y, D[lane][j], xj);
      }
      if (lane < wd) xs[r][lane] = y;
    }
    __syncthreads();
    // fold x_k into the rows above: w[0:k0) -= U[0:k0, panel] x_k  (thread owns rows: no reduction)
    for (int row = gt; row < k0; row += nth) {
      S acc[NRHS];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) acc[r] = w[r][row];
      row_dot_sub<S, NRHS, PW>(acc, Fn + row + (long long)k0 * m, m, wd, xs);
#pragma unroll
      for (int r = 0; r < NRHS; ++r) w[r][row] = acc[r];
    }
    SY::sync();
    if (SY::rank() == 0 && t < 32 * NRHS && (t & 31) < wd) w[t >> 5][k0 + (t & 31)] = xs[t >> 5][t & 31];
  }
  SY::sync();
  for (int i = gt; i < ns; i += nth)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) xp[(size_t)r * na + os + i] = w[r][i];
}

template <class S, int NRHS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS) k_bwd_chain(FrontTab T, const int* __restrict__ nodes,
                                                                 const S* __restrict__ F, S* work,
                                                                 int64_t work_stride, S* xp, int na) {
  bwd_chain_body<S, NRHS, SyncCta>(T, nodes, F, work, work_stride, xp, na);
}
template <class S, int NRHS>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CL_THREADS)
    k_bwd_chain_cluster(FrontTab T, const int* __restrict__ nodes, const S* __restrict__ F, S* work,
                        int64_t work_stride, S* xp, int na) {
  bwd_chain_body<S, NRHS, SyncCluster>(T, nodes, F, work, work_stride, xp, na);
}

// ------------------------------------------------------------------------------ off-diagonal updates
template <class S, int NRHS>
__global__ void __launch_bounds__(UPD_ROWS) k_fwd_update(FrontTab T, const int* __restrict__ nodes,
                                                         const S* __restrict__ F, S* work, int64_t work_stride) {
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
    if (valid) row_dot_sub<S, NRHS, UPD_CHUNK>(acc, Fn + row + (long long)j0 * m, m, cn, xs);
    __syncthreads();
  }
  if (valid)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][row] = w[r][row] + acc[r];
}

template <class S, int NRHS>
__global__ void __launch_bounds__(UPD_ROWS) k_bwd_update(FrontTab T, const int* __restrict__ nodes,
                                                         const S* __restrict__ F, S* work, int64_t work_stride) {
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
    if (valid) row_dot_sub<S, NRHS, UPD_CHUNK>(acc, Fn + row + (long long)(ns + c0) * m, m, cn, xs);
    __syncthreads();
  }
  if (valid)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][row] = w[r][row] + acc[r];
}

// ------------------------------------------------------------------------------ driver
static int chain_threads(int max_ns) {
  if (max_ns <= 64) return 64;
  if (max_ns <= 128) return 128;
  if (max_ns <= 256) return 256;
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
  const bool no_cluster = getenv("FW_NO_CLUSTER") != nullptr;
  Trace tr(s);
  if (na > 0) k_lu_permute_in<S><<<cdiv(na, 256), 256, 0, s>>>(na, d.perm.p, b, n, NRHS, xp);
  for (int lev = Sy.n_levels - 1; lev >= 0; --lev) {
    const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes = d.level_nodes.p + Sy.level_start[lev];
    const int max_ns = Sy.level_max_ns[lev];
    if (max_ns > CL_MIN_NS && !no_cluster)
      k_fwd_chain_cluster<S, NRHS><<<nfront * CL, CL_THREADS, 0, s>>>(T, nodes, F, d.rowperm.p, work,
                                                                      Sy.work_entries, xp, na);
    else
      k_fwd_chain<S, NRHS><<<nfront, chain_threads(max_ns), 0, s>>>(T, nodes, F, d.rowperm.p, work,
                                                                    Sy.work_entries, xp, na);
    ++launches;
    if (trace2)
      tr.mark(("fwd chain  lev " + std::to_string(lev) + " fronts " + std::to_string(nfront) + " max_ns " +
               std::to_string(max_ns)).c_str());
    const int tiles = (Sy.level_max_nb[lev] + UPD_ROWS - 1) / UPD_ROWS;
    if (tiles > 0) {
      k_fwd_update<S, NRHS><<<dim3(nfront, tiles), UPD_ROWS, 0, s>>>(T, nodes, F, work, Sy.work_entries);
      ++launches;
      if (trace2)
        tr.mark(("fwd update lev " + std::to_string(lev) + " max_nb " + std::to_string(Sy.level_max_nb[lev])).c_str());
    }
  }
  for (int lev = 0; lev < Sy.n_levels; ++lev) {
    const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes = d.level_nodes.p + Sy.level_start[lev];
    const int max_ns = Sy.level_max_ns[lev];
    if (Sy.level_max_nb[lev] > 0) {
      const int tiles = std::max(1, (max_ns + UPD_ROWS - 1) / UPD_ROWS);
      k_bwd_update<S, NRHS><<<dim3(nfront, tiles), UPD_ROWS, 0, s>>>(T, nodes, F, work, Sy.work_entries);
      ++launches;
      if (trace2) tr.mark(("bwd update lev " + std::to_string(lev)).c_str());
    }
    if (max_ns > CL_MIN_NS && !no_cluster)
      k_bwd_chain_cluster<S, NRHS><<<nfront * CL, CL_THREADS, 0, s>>>(T, nodes, F, work, Sy.work_entries, xp, na);
    else
      k_bwd_chain<S, NRHS><<<nfront, chain_threads(max_ns), 0, s>>>(T, nodes, F, work, Sy.work_entries, xp, na);
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
