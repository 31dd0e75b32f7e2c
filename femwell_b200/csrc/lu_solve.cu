// Level-batched forward/backward substitution with the multifrontal factors (device).  See lu.h.
// Replaces the two triangular solves per ARPACK iteration of scipy's SuperLU object
// (scipy arpack.py:1011-1019, ido==1 branch :817-823).
//
// Four kernels per tree level:
//   forward : chain  (assemble w from the rhs and the children, apply the composed row permutation,
//                     unit-lower solve of the ns x ns pivot block, 32 columns at a time)
//             update (fronts x 128-row tiles x 4 column groups: w_boundary -= L21 x1 -- the bulk of
//                     the L traffic)
//   backward: update (same tiling: gather x2 from the parent, y1 = w1 - U12 x2)
//             chain  (right-looking upper solve of the pivot block, write x)
// The chain is the only sequential part (ns/32 dependent steps).  Every thread owns one fixed row of the
// pivot block; the 32 factor entries it needs for the NEXT step and the next 32x32 diagonal block are
// prefetched into registers while the current step runs (software pipelining: the chain is bound by HBM
// latency, not bandwidth).  Fronts with a tall pivot block (top of the tree: ns up to a few thousand,
// a handful of fronts per level) run the chain on a thread-block CLUSTER of 8 CTAs -- every CTA owns a
// slice of the rows, the 32x32 diagonal solve is replicated per CTA, one hardware cluster barrier per
// step.  HBM-bound work (each factor entry is read once per solve); no tensor cores.
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
constexpr int DPT = 4;            // prefetched diagonal-block entries per thread (blockDim >= 256)

// TMA bulk prefetch of a contiguous global range into L2 (no destination, no completion tracking)
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, long long bytes) {
  const unsigned long long a = (unsigned long long)p;
  const unsigned long long a0 = a & ~15ull;
  const long long nb = (bytes + (long long)(a - a0)) & ~15ll;
  if (nb > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(a0), "r"((unsigned)nb) : "memory");
}

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

__device__ __forceinline__ double ldcg_(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cplx ldcg_(const cplx* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
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
// that 16 independent HBM requests are in flight per thread
template <class S, int NRHS, int LD>
__device__ __forceinline__ void row_dot_sub(S (&acc)[NRHS], const S* __restrict__ col, long long m, int cnt,
                                            S (&xs)[NRHS][LD], int x0 = 0, int stride = 1) {
  int j = 0;
  for (; j + 16 <= cnt; j += 16) {
    S l[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) l[q] = col[(long long)(j + q) * stride * m];
#pragma unroll
    for (int q = 0; q < 16; ++q)
#pragma unroll
      for (int r = 0; r < NRHS; ++r) fnma_(acc[r], l[q], xs[r][x0 + (j + q) * stride]);
  }
  for (; j < cnt; ++j) {
    const S l = col[(long long)j * stride * m];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) fnma_(acc[r], l, xs[r][x0 + j * stride]);
  }
}

// Occupancy hints.  A kernel with __launch_bounds__(N) and NO minimum block count makes ptxas schedule for the smallest
// register footprint (32 registers here): it then sinks every load of an unrolled batch down to the FMA that consumes
// it, and ONE factor load is in flight per thread (SASS: LDG, DFMA, LDG, DFMA ...; ncu: every DFMA waits on the long
// scoreboard).  With an explicit minimum the same source keeps the 16 loads of row_dot_sub back to back (58-64
// registers).  Measured on the config-2 factors: substitution pair 4.30 -> 3.28 ms (profiles/occupancy_hint_r02.log).
__host__ __device__ constexpr int upd_min_blocks(int G) { return G == 4 ? 2 : 8; }

// Synchronisation policy of the chain kernels: one CTA (block barrier) or one cluster (cluster barrier).
struct SyncCta {
  static constexpr bool pipe = false;  // small fronts: occupancy matters more than the register prefetch
  __device__ static int rank() { return 0; }
  __device__ static int nctas() { return 1; }
  __device__ static int front(int bx) { return bx; }
  __device__ static void sync() { __syncthreads(); }
};
struct SyncCluster {
  static constexpr bool pipe = true;
  __device__ static int rank() { return (int)cg::this_cluster().block_rank(); }
  __device__ static int nctas() { return CL; }
  __device__ static int front(int bx) { return bx / CL; }
  // barrier.cluster arrive.release / wait.acquire: orders the global-memory writes of all CTAs of the
  // cluster (and invalidates L1) before anyone proceeds
  __device__ static void sync() { cg::this_cluster().sync(); }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- DINV chains (real arithmetic, blocks of up to DINV_MAX_THREADS = 256 threads): the 32 x 32 diagonal blocks come pre-inverted from
// LUDevice::dinv (k_lu_diag_inverse, all fronts), so the 32 dependent shuffle/FMA steps of a diagonal solve become one
// 32 x 32 matrix-vector product, and the block itself is copied with cp.async into one of two shared-memory buffers
// one panel ahead (ncu on the leaf level of config 2: 30 % of the backward kernel sat in the eight serialised
// load + divide iterations per thread that filled D, another 35 % at the barriers around the serial solve).
// Only the non-zero triangle of the first wd columns is fetched (16-byte chunks = two rows of a column); the other chunks
// of the buffer are zeroed, so the full 32 x 32 product of dinv_apply stays valid.  (Fetching the whole 8 KB block put
// 1 GB per pass on top of the 2 GB of factor entries of the leaf level: ncu DRAM traffic 2.93 GB for the forward kernel.)
template <bool LOWER>
__device__ __forceinline__ void dinv_fetch(double* buf, const double* __restrict__ src, int wd, int t, int nthr) {
  for (int c = t; c < PW * PW / 2; c += nthr) {
    const int col = c >> 4, r2 = (c & 15) * 2;
    const bool need = col < wd && r2 < wd && (LOWER ? r2 + 1 >= col : r2 <= col);
    if (need)
      cp_async16(buf + 2 * c, src + 2 * c);
    else
      *reinterpret_cast<double2*>(buf + 2 * c) = make_double2(0.0, 0.0);
  }
}
// x = inv(D_kk) w_k by warp r (right-hand side r); Dv is column-major, ragged blocks are padded with the identity
template <int NRHS>
__device__ __forceinline__ void dinv_apply(const double* Dv, const double* wk, int wd, int t, double (*xs)[PW]) {
  const int r = t >> 5, lane = t & 31;
  const double wv = lane < wd ? wk[lane] : 0.0;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
  for (int j = 0; j < PW; j += 4) {
    a0 = fma(Dv[(j + 0) * PW + lane], __shfl_sync(0xffffffffu, wv, j + 0), a0);
    a1 = fma(Dv[(j + 1) * PW + lane], __shfl_sync(0xffffffffu, wv, j + 1), a1);
    a2 = fma(Dv[(j + 2) * PW + lane], __shfl_sync(0xffffffffu, wv, j + 2), a2);
    a3 = fma(Dv[(j + 3) * PW + lane], __shfl_sync(0xffffffffu, wv, j + 3), a3);
  }
  if (lane < wd) xs[r][lane] = (a0 + a1) + (a2 + a3);
}

// prefetch / commit of a 32x32 diagonal block (column-major panel of the front) through registers
template <class S>
__device__ __forceinline__ void diag_fetch(S (&dn)[DPT], const S* __restrict__ Fn, int m, int k0, int wd, int t,
                                           int nthr) {
#pragma unroll
  for (int q = 0; q < DPT; ++q) {
    const int idx = t + q * nthr;
    dn[q] = scalar_traits<S>::zero();
    if (idx < wd * wd) {
      const int r = idx % wd, c = idx / wd;
      dn[q] = Fn[(k0 + r) + (long long)(k0 + c) * m];
    }
  }
}
template <class S, bool RECIP_DIAG>
__device__ __forceinline__ void diag_commit(S (*D)[PW + 1], const S (&dn)[DPT], int wd, int t, int nthr) {
#pragma unroll
  for (int q = 0; q < DPT; ++q) {
    const int idx = t + q * nthr;
    if (idx < wd * wd) {
      const int r = idx % wd, c = idx / wd;
      // reciprocal of the diagonal here (32 threads in parallel) keeps the divisions out of the serial chain
      D[r][c] = (RECIP_DIAG && r == c) ? (scalar_traits<S>::one() / dn[q]) : dn[q];
    }
  }
}

// ------------------------------------------------------------------------------ forward chain
// FUSE: the CTA also updates the boundary rows (w2 -= L21 x1) panel by panel, so the level needs no k_fwd_update
// launch (levels of many small fronts: one contiguous pass over the L panel, no second kernel, no tail).
template <class S, int NRHS, class SY, bool FUSE = false, bool DINV = false>
__device__ __forceinline__ void fwd_chain_body(FrontTab T, const int node, const S* __restrict__ F,
                                               const int* __restrict__ rowperm, S* work, int64_t work_stride, S* xp,
                                               int na, const double* __restrict__ dinv = nullptr) {
  __shared__ S D[DINV ? 1 : PW][PW + 1];
  __shared__ __align__(16) double Dv[DINV ? 2 : 1][DINV ? PW * PW : 2];
  __shared__ S xs[NRHS][PW];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  const S* Fn = F + T.front_off[node];
  S* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x;
  const int nthr = blockDim.x;
  const int nth = nthr * SY::nctas();
  const int gt = SY::rank() * nthr + t;
  if (FUSE && T.pf > 0) {
    // small fronts: the L panel (m x ns, contiguous) is pulled into L2 by TMA bulk prefetches before the dependent
    // chain of loads starts, so that the chain runs at L2 latency instead of loaded-HBM latency
    const long long bytes = min((long long)m * ns * (long long)sizeof(S), (long long)T.pf * 1024);
    const long long chunk = 4096;
    for (long long o = (long long)t * chunk; o < bytes; o += (long long)nthr * chunk)
      prefetch_l2_bulk((const char*)Fn + o, min(chunk, bytes - o));
  }
  // register prefetch for real arithmetic and block sizes >= 256 (32 c128 values would not fit the budget)
  const bool pipelined = !DINV && SY::pipe && !scalar_traits<S>::is_complex && nthr * DPT >= PW * PW;
  const double* dv = nullptr;
  if constexpr (DINV) {
    dv = dinv + T.dinv_off[node];
    if (ns > 0) dinv_fetch<true>(Dv[0], dv, min(PW, ns), t, nthr);  // independent of the right-hand side: in flight during the assembly
  }
  // the first diagonal block and this thread's first row segment do not depend on the rhs: fetch them now
  S dn[DPT], ln[PW];
  const int row0 = gt;  // the fixed row of the pivot block owned by this thread (further rows: gt + c*nth)
  if (ns > 0 && pipelined) {
    diag_fetch<S>(dn, Fn, m, 0, min(PW, ns), t, nthr);
    if (row0 >= PW && row0 < ns) {
#pragma unroll
      for (int j = 0; j < PW; ++j) ln[j] = Fn[row0 + (long long)j * m];
    }
  }
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
        w[r][ri] = w[r][ri] + ldcg_(wc + nsc + i);  // written by another CTA, possibly of this launch: not through L1
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
  const int rlim = FUSE ? m : ns;  // rows folded by this kernel
  for (int k0 = 0; k0 < ns; k0 += PW) {
    const int wd = min(PW, ns - k0);
    if constexpr (DINV) {
      cp_async_wait_all();
    } else if (pipelined) {
      diag_commit<S, false>(D, dn, wd, t, nthr);
    } else {
      for (int idx = t; idx < wd * wd; idx += nthr) {
        int r = idx % wd, c = idx / wd;
        D[r][c] = Fn[(k0 + r) + (long long)(k0 + c) * m];
      }
    }
    __syncthreads();
    // prefetch the next diagonal block while warp 0 runs the serial 32-step solve
    const int k1 = k0 + PW;
    if (pipelined && k1 < ns) diag_fetch<S>(dn, Fn, m, k1, min(PW, ns - k1), t, nthr);
    if constexpr (DINV) {
      const int kp = k0 / PW;
      if (k1 < ns) dinv_fetch<true>(Dv[(kp + 1) & 1], dv + (long long)(kp + 1) * PW * PW, min(PW, ns - k1), t, nthr);
      if (t < 32 * NRHS) dinv_apply<NRHS>(Dv[kp & 1], work + (size_t)(t >> 5) * work_stride + T.work_off[node] + k0, wd, t, xs);
    } else if (t < 32 * NRHS) {  // replicated in every CTA of a cluster; rank 0 stores the result
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
    if (pipelined && wd == PW && row0 >= k0 + PW && row0 < ns) {
      S acc[NRHS];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) acc[r] = w[r][row0];
#pragma unroll
      for (int j = 0; j < PW; ++j)
#pragma unroll
        for (int r = 0; r < NRHS; ++r) fnma_(acc[r], ln[j], xs[r][j]);
#pragma unroll
      for (int r = 0; r < NRHS; ++r) w[r][row0] = acc[r];
      // this row's entries of the next panel (used one step later)
      if (row0 >= k1 + PW && k1 + PW <= ns) {
#pragma unroll
        for (int j = 0; j < PW; ++j) ln[j] = Fn[row0 + (long long)(k1 + j) * m];
      } else if (row0 >= k1 + PW) {
        // the next panel is the ragged last one: it is handled by the generic path below
      }
      for (int row = row0 + nth; row < ns; row += nth) {
        S a2[NRHS];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) a2[r] = w[r][row];
        row_dot_sub<S, NRHS, PW>(a2, Fn + row + (long long)k0 * m, m, wd, xs);
#pragma unroll
        for (int r = 0; r < NRHS; ++r) w[r][row] = a2[r];
      }
    } else {
      const int first = (pipelined && wd == PW) ? row0 + nth : k0 + wd + gt;
      for (int row = first; row < rlim; row += nth) {
        if (row < k0 + wd) continue;
        S acc[NRHS];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) acc[r] = w[r][row];
        row_dot_sub<S, NRHS, PW>(acc, Fn + row + (long long)k0 * m, m, wd, xs);
#pragma unroll
        for (int r = 0; r < NRHS; ++r) w[r][row] = acc[r];
      }
      // keep the pipeline primed for the next full panel
      if (pipelined && wd == PW && row0 >= k1 + PW && row0 < ns && k1 + PW <= ns) {
#pragma unroll
        for (int j = 0; j < PW; ++j) ln[j] = Fn[row0 + (long long)(k1 + j) * m];
      }
    }
    SY::sync();
    // x_k is stored only now: the other CTAs of a cluster have read the old w[k0:k0+wd) by this point
    if (SY::rank() == 0 && t < 32 * NRHS && (t & 31) < wd) w[t >> 5][k0 + (t & 31)] = xs[t >> 5][t & 31];
  }
}

template <class S, int NRHS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS, 2) k_fwd_chain(FrontTab T, const int* __restrict__ nodes,
                                                                 const S* __restrict__ F,
                                                                 const int* __restrict__ rowperm, S* work,
                                                                 int64_t work_stride, S* xp, int na) {
  fwd_chain_body<S, NRHS, SyncCta>(T, nodes[blockIdx.x], F, rowperm, work, work_stride, xp, na);
}
template <class S, int NRHS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS, 2) k_fwd_chain_fused(FrontTab T, const int* __restrict__ nodes,
                                                                       const S* __restrict__ F,
                                                                       const int* __restrict__ rowperm, S* work,
                                                                       int64_t work_stride, S* xp, int na) {
  fwd_chain_body<S, NRHS, SyncCta, true>(T, nodes[blockIdx.x], F, rowperm, work, work_stride, xp, na);
}
// DINV variants (f64, <= DINV_MAX_THREADS threads): see dinv_fetch / dinv_apply
constexpr int DINV_MAX_THREADS = 256;
template <int NRHS, bool FUSE>
__global__ void __launch_bounds__(DINV_MAX_THREADS, 4) k_fwd_chain_dinv(FrontTab T, const int* __restrict__ nodes,
                                                                     const double* __restrict__ F,
                                                                     const double* __restrict__ dinv,
                                                                     const int* __restrict__ rowperm, double* work,
                                                                     int64_t work_stride, double* xp, int na) {
  fwd_chain_body<double, NRHS, SyncCta, FUSE, true>(T, nodes[blockIdx.x], F, rowperm, work, work_stride, xp, na, dinv);
}
template <class S, int NRHS>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CL_THREADS)
    k_fwd_chain_cluster(FrontTab T, const int* __restrict__ nodes, const S* __restrict__ F,
                        const int* __restrict__ rowperm, S* work, int64_t work_stride, S* xp, int na) {
  fwd_chain_body<S, NRHS, SyncCluster>(T, nodes[blockIdx.x / CL], F, rowperm, work, work_stride, xp, na);
}

// ------------------------------------------------------------------------------ backward chain (right-looking)
constexpr int FUSE_CHUNK = 128;  // boundary columns staged per step of the fused backward update

// FUSE: the CTA first gathers x2 from the parent and folds it into the own rows (w1 -= U12 x2), so the level needs
// no k_bwd_update launch.
template <class S, int NRHS, class SY, bool FUSE = false, bool DINV = false>
__device__ __forceinline__ void bwd_chain_body(FrontTab T, const int node, const S* __restrict__ F,
                                               S* work, int64_t work_stride, S* xp, int na,
                                               const double* __restrict__ dinv = nullptr) {
  __shared__ S D[DINV ? 1 : PW][PW + 1];
  __shared__ __align__(16) double Dv[DINV ? 2 : 1][DINV ? PW * PW : 2];
  __shared__ S xs[NRHS][PW];
  __shared__ S x2[FUSE ? NRHS : 1][FUSE ? FUSE_CHUNK : 1];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  const S* Fn = F + T.front_off[node];
  S* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x;
  const int nthr = blockDim.x;
  if (FUSE && T.pf > 0) {
    // the U rows of the front (rows 0..ns of every column right of the diagonal) into L2, column by column
    const int cmax = min(m, (int)((long long)T.pf * 1024 / max(1, ns * (int)sizeof(S))));
    for (int c = t; c < cmax; c += nthr) {  // in the order of use: U12 left to right, then U11 right to left
      const int col = c < m - ns ? ns + c : m - 1 - c;
      const int rows = c < m - ns ? ns : col + 1;
      prefetch_l2_bulk(Fn + (long long)col * m, (long long)rows * (long long)sizeof(S));
    }
  }
  if constexpr (DINV) {
    // the first inv(U_kk) block does not depend on the right-hand side: in flight during the U12 pass
    const int np = (ns + PW - 1) / PW;
    if (np > 0)
      dinv_fetch<false>(Dv[(np - 1) & 1], dinv + T.dinv_off[node] + (long long)(2 * np - 1) * PW * PW,
                        ns - (np - 1) * PW, t, nthr);
  }
  if constexpr (FUSE) if (m > ns) {
    const int p = T.parent[node];
    const int nbd = m - ns;
    const int* rel = T.rel + T.blist_off[node];
    for (int c0 = 0; c0 < nbd; c0 += FUSE_CHUNK) {
      const int cn = min(FUSE_CHUNK, nbd - c0);
      for (int c = t; c < cn; c += nthr) {
        const int ri = rel[c0 + c];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
          const S v = ldcg_(work + (size_t)r * work_stride + T.work_off[p] + ri);  // the parent's CTA wrote it
          x2[r][c] = v;
          w[r][ns + c0 + c] = v;  // the children of this front gather from here
        }
      }
      __syncthreads();
      for (int row = t; row < ns; row += nthr) {
        S acc[NRHS];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) acc[r] = w[r][row];
        row_dot_sub<S, NRHS, FUSE_CHUNK>(acc, Fn + row + (long long)(ns + c0) * m, m, cn, x2);
#pragma unroll
        for (int r = 0; r < NRHS; ++r) w[r][row] = acc[r];
      }
      __syncthreads();
    }
  }
  if (ns == 0) return;
  const int nth = nthr * SY::nctas();
  const int gt = SY::rank() * nthr + t;
  const int npan = (ns + PW - 1) / PW;
  const bool pipelined = !DINV && SY::pipe && !scalar_traits<S>::is_complex && nthr * DPT >= PW * PW;
  const int row0 = gt;
  S dn[DPT], un[PW];
  {
    const int k0 = (npan - 1) * PW;
    if (pipelined) diag_fetch<S>(dn, Fn, m, k0, min(PW, ns - k0), t, nthr);
  }
  const double* dv = nullptr;  // inv(U_kk) blocks follow the npan inv(L_kk) blocks
  if constexpr (DINV) dv = dinv + T.dinv_off[node] + (long long)npan * PW * PW;
  bool un_valid = false;  // un holds U[row0, panel kp] for the panel about to be processed
  for (int kp = npan - 1; kp >= 0; --kp) {
    const int k0 = kp * PW;
    const int wd = min(PW, ns - k0);
    if constexpr (DINV) {
      cp_async_wait_all();
    } else if (pipelined) {
      diag_commit<S, true>(D, dn, wd, t, nthr);
    } else {
      for (int idx = t; idx < wd * wd; idx += nthr) {
        int r = idx % wd, c = idx / wd;
        const S v = Fn[(k0 + r) + (long long)(k0 + c) * m];
        D[r][c] = (r == c) ? (scalar_traits<S>::one() / v) : v;
      }
    }
    __syncthreads();
    if (pipelined && kp > 0) diag_fetch<S>(dn, Fn, m, k0 - PW, PW, t, nthr);
    if constexpr (DINV) {
      if (kp > 0) dinv_fetch<false>(Dv[(kp - 1) & 1], dv + (long long)(kp - 1) * PW * PW, PW, t, nthr);
      if (t < 32 * NRHS) dinv_apply<NRHS>(Dv[kp & 1], work + (size_t)(t >> 5) * work_stride + T.work_off[node] + k0, wd, t, xs);
    } else if (t < 32 * NRHS) {
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
    if (row0 < k0) {
      S acc[NRHS];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) acc[r] = w[r][row0];
      if (un_valid && wd == PW) {
#pragma unroll
        for (int j = 0; j < PW; ++j)
#pragma unroll
          for (int r = 0; r < NRHS; ++r) fnma_(acc[r], un[j], xs[r][j]);
      } else {
        row_dot_sub<S, NRHS, PW>(acc, Fn + row0 + (long long)k0 * m, m, wd, xs);
      }
#pragma unroll
      for (int r = 0; r < NRHS; ++r) w[r][row0] = acc[r];
    }
    // this row's entries of the next (lower-index, always full) panel
    un_valid = false;
    if (pipelined && kp > 0 && row0 < k0 - PW) {
#pragma unroll
      for (int j = 0; j < PW; ++j) un[j] = Fn[row0 + (long long)(k0 - PW + j) * m];
      un_valid = true;
    }
    for (int row = row0 + nth; row < k0; row += nth) {
      S a2[NRHS];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) a2[r] = w[r][row];
      row_dot_sub<S, NRHS, PW>(a2, Fn + row + (long long)k0 * m, m, wd, xs);
#pragma unroll
      for (int r = 0; r < NRHS; ++r) w[r][row] = a2[r];
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
__global__ void __launch_bounds__(SOLVE_MAX_THREADS, 2) k_bwd_chain(FrontTab T, const int* __restrict__ nodes,
                                                                 const S* __restrict__ F, S* work,
                                                                 int64_t work_stride, S* xp, int na) {
  bwd_chain_body<S, NRHS, SyncCta>(T, nodes[blockIdx.x], F, work, work_stride, xp, na);
}
template <class S, int NRHS>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS, 2) k_bwd_chain_fused(FrontTab T, const int* __restrict__ nodes,
                                                                       const S* __restrict__ F, S* work,
                                                                       int64_t work_stride, S* xp, int na) {
  bwd_chain_body<S, NRHS, SyncCta, true>(T, nodes[blockIdx.x], F, work, work_stride, xp, na);
}
template <int NRHS, bool FUSE>
__global__ void __launch_bounds__(DINV_MAX_THREADS, 4) k_bwd_chain_dinv(FrontTab T, const int* __restrict__ nodes,
                                                                     const double* __restrict__ F,
                                                                     const double* __restrict__ dinv, double* work,
                                                                     int64_t work_stride, double* xp, int na) {
  bwd_chain_body<double, NRHS, SyncCta, FUSE, true>(T, nodes[blockIdx.x], F, work, work_stride, xp, na, dinv);
}
template <class S, int NRHS>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CL_THREADS)
    k_bwd_chain_cluster(FrontTab T, const int* __restrict__ nodes, const S* __restrict__ F, S* work,
                        int64_t work_stride, S* xp, int na) {
  bwd_chain_body<S, NRHS, SyncCluster>(T, nodes[blockIdx.x / CL], F, work, work_stride, xp, na);
}

// ------------------------------------------------------------------------------ cluster chains, register resident
// Chains of the tall fronts (f64, ns <= CL * CL_THREADS) with the inverted diagonal blocks of the factorisation
// (LUDevice::dinv).  Every thread owns ONE row of the pivot block and keeps its entry of the right-hand side in
// a register for the whole chain; warp g of the cluster therefore owns panel g.  Per 32-column step:
//   owner warp : x_k = inv(D_kk) w_k as a 32 x 32 matrix-vector product (w_k gathered with shuffles, the
//                inverse block was copied into shared memory with cp.async at kernel start), then stores x_k
//                into the shared memory of all CTAs of the cluster (distributed shared memory);
//   one hardware cluster barrier;
//   all threads: acc -= L[row, panel k] . x_k with the 32 factor entries prefetched into registers one step
//                earlier.
// No global-memory access sits on the dependent path any more (the first version re-read w from L2 and ran a
// 32-step shuffle substitution per panel: 4.8 us per step; see DESIGN.md for the measured effect).


template <int NRHS>
__device__ __forceinline__ void inv_block_matvec(const double* __restrict__ Dw, int lane, double (&acc)[NRHS]) {
  // x_lane = sum_j Dinv[lane][j] acc_j  (column-major block: conflict-free)
#pragma unroll
  for (int r = 0; r < NRHS; ++r) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int j = 0; j < PW; j += 4) {
      s0 = fma(Dw[(j + 0) * PW + lane], __shfl_sync(0xffffffffu, acc[r], j + 0), s0);
      s1 = fma(Dw[(j + 1) * PW + lane], __shfl_sync(0xffffffffu, acc[r], j + 1), s1);
      s2 = fma(Dw[(j + 2) * PW + lane], __shfl_sync(0xffffffffu, acc[r], j + 2), s2);
      s3 = fma(Dw[(j + 3) * PW + lane], __shfl_sync(0xffffffffu, acc[r], j + 3), s3);
    }
    acc[r] = (s0 + s1) + (s2 + s3);
  }
}

template <int NRHS>
__device__ __forceinline__ void panel_dot_sub(double (&acc)[NRHS], const double (&ln)[PW], const double* xsb) {
#pragma unroll
  for (int r = 0; r < NRHS; ++r) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int j = 0; j < PW; j += 4) {
      s0 = fma(ln[j + 0], xsb[r * PW + j + 0], s0);
      s1 = fma(ln[j + 1], xsb[r * PW + j + 1], s1);
      s2 = fma(ln[j + 2], xsb[r * PW + j + 2], s2);
      s3 = fma(ln[j + 3], xsb[r * PW + j + 3], s3);
    }
    acc[r] -= (s0 + s1) + (s2 + s3);
  }
}

template <int NRHS>
__device__ __forceinline__ void bcast_to_cluster(double* xs_local, int buf, int lane, const double (&acc)[NRHS]) {
  cg::cluster_group cl = cg::this_cluster();
#pragma unroll
  for (int c = 0; c < CL; ++c) {
    double* remote = cl.map_shared_rank(xs_local, c);
#pragma unroll
    for (int r = 0; r < NRHS; ++r) remote[(buf * NRHS + r) * PW + lane] = acc[r];
  }
}

// ---- step signalling without a cluster barrier: the owner warp stores x_k into every CTA with st.async, which
// completes transaction bytes on an mbarrier of the destination CTA; the consumers wait on their local mbarrier.
// (A barrier.cluster with release/acquire semantics makes every thread wait for its outstanding prefetch loads:
// 2.1 us per step; the mbarrier hand-off leaves the loads in flight.)  XRING slots, one cluster barrier every
// XRING steps bounds how far the CTAs can drift apart, so a slot is never overwritten while it is still read.
constexpr int XRING = 8;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* b, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_async_remote(double* local_dst, unsigned long long* local_bar, unsigned cta, double v) {
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(smem_u32(local_dst)), "r"(cta));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(cta));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(ra),
               "l"(__double_as_longlong(v)), "r"(rb)
               : "memory");
}
// one chain step: publish x_k (owner warp) and wait for it (everybody); returns the slot of the ring that holds x_k
template <int NRHS>
__device__ __forceinline__ const double* chain_step_exchange(int step, bool owner, int lane, int t, const double* Dw,
                                                             double (&acc)[NRHS], double* xs, unsigned long long* mbar,
                                                             int use_mbar) {
  cg::cluster_group cl = cg::this_cluster();
  if (!use_mbar) {
    const int buf = step & 1;
    if (owner) {
      inv_block_matvec<NRHS>(Dw, lane, acc);
      bcast_to_cluster<NRHS>(xs, buf, lane, acc);
    }
    cl.sync();
    return xs + buf * NRHS * PW;
  }
  const int slot = step & (XRING - 1);
  if (slot == 0 && step > 0) cl.sync();
  if (t == 0) mbar_expect_tx(&mbar[slot], 8u * PW * NRHS);
  if (owner) {
    inv_block_matvec<NRHS>(Dw, lane, acc);
#pragma unroll
    for (int c = 0; c < CL; ++c)
#pragma unroll
      for (int r = 0; r < NRHS; ++r) st_async_remote(xs + (slot * NRHS + r) * PW + lane, &mbar[slot], (unsigned)c, acc[r]);
  }
  const unsigned par = (unsigned)(step / XRING) & 1u;
  while (!mbar_try_wait(&mbar[slot], par)) {
  }
  return xs + slot * NRHS * PW;
}

// THREADS x CL rows per cluster; STAGES panels of factor entries are kept in flight per thread (registers): with one
// stage the load issued at the end of step k is needed one step (~1 us) later and its L2 latency bounds the step;
// 256 threads per CTA leave room for two or three stages.
template <int NRHS, int THREADS, int STAGES>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS)
    k_fwd_chain_inv(FrontTab T, const int* __restrict__ nodes, const double* __restrict__ F,
                    const double* __restrict__ dinv, const int* __restrict__ rowperm, double* work,
                    int64_t work_stride, double* xp, int na, int use_mbar) {
  extern __shared__ __align__(16) double inv_smem[];
  __shared__ __align__(8) unsigned long long mbar[XRING];
  double* Dsm = inv_smem;                               // [warps][PW*PW]
  double* xs = inv_smem + (THREADS / 32) * PW * PW;     // [XRING][NRHS][PW]
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0) {
    for (int b = 0; b < XRING; ++b) mbar_init(&mbar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  const int node = nodes[blockIdx.x / CL];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  const double* Fn = F + T.front_off[node];
  double* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int nth = THREADS * CL;
  const int gt = (int)cl.block_rank() * THREADS + t;
  // panels are dealt round-robin to the CTAs of the cluster: the rows still active at any step are spread
  // evenly over the 8 SMs (the chain is bound by the L2 -> SM fill bandwidth of the SMs that hold rows)
  const int gw = wid * CL + (int)cl.block_rank();  // the panel owned by this warp
  const int row0 = gw * PW + lane;
  const int npan = (ns + PW - 1) / PW;
  double* Dw = Dsm + wid * PW * PW;
  if (gw < npan) {
    const double* src = dinv + T.dinv_off[node] + (long long)gw * PW * PW;
#pragma unroll
    for (int q = 0; q < PW * PW / 2 / 32; ++q) cp_async16(Dw + 2 * (lane + 32 * q), src + 2 * (lane + 32 * q));
  }
  // The chain walks L11 panel by panel with one dependent step per panel: pull the whole lower triangle into
  // L2 now (one bulk prefetch per column, issued by the thread that owns the row of the same index), so that
  // the one-step-ahead register prefetch below is served by L2 instead of HBM.
  if (row0 < ns) prefetch_l2_bulk(Fn + (long long)row0 * m + row0, (long long)(ns - row0) * 8);
  double ln[STAGES][PW];
#pragma unroll
  for (int st = 0; st < STAGES; ++st) {
    if (st < npan && row0 >= (st + 1) * PW && row0 < ns) {
#pragma unroll
      for (int j = 0; j < PW; ++j) ln[st][j] = Fn[row0 + (long long)(st * PW + j) * m];
    }
  }
  // assemble w: own rhs entries + contributions of the children (same as the generic chain)
  for (int i = gt; i < m; i += nth)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][i] = i < ns ? xp[(size_t)r * na + os + i] : 0.0;
  cl.sync();
  for (int sl = 0; sl < 2; ++sl) {
    const int c = sl ? T.child1[node] : T.child0[node];
    if (c < 0) continue;
    const int nsc = T.ns[c], nbc = T.m[c] - nsc;
    const int* rel = T.rel + T.blist_off[c];
    for (int i = gt; i < nbc; i += nth) {
      const int ri = rel[i];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        const double* wc = work + (size_t)r * work_stride + T.work_off[c];
        w[r][ri] = w[r][ri] + wc[nsc + i];
      }
    }
    cl.sync();
  }
  if (ns == 0) return;
  for (int i = gt; i < ns; i += nth)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) xp[(size_t)r * na + os + i] = w[r][i];
  cl.sync();
  double acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = row0 < ns ? xp[(size_t)r * na + os + rowperm[os + row0]] : 0.0;
  cp_async_wait_all();
  __syncwarp();
  for (int kb = 0; kb < npan; kb += STAGES) {
#pragma unroll
    for (int st = 0; st < STAGES; ++st) {
      const int k = kb + st;
      if (k < npan) {
        const double* xk = chain_step_exchange<NRHS>(k, gw == k, lane, t, Dw, acc, xs, mbar, use_mbar);
        if (row0 >= (k + 1) * PW && row0 < ns) panel_dot_sub<NRHS>(acc, ln[st], xk);
        // this stage is free again: entries of panel k + STAGES (a row below a panel means the panel is full)
        const int kn = k + STAGES;
        if (kn < npan && row0 >= (kn + 1) * PW && row0 < ns) {
#pragma unroll
          for (int j = 0; j < PW; ++j) ln[st][j] = Fn[row0 + (long long)(kn * PW + j) * m];
        }
      }
    }
  }
  if (row0 < ns) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][row0] = acc[r];
  }
}

template <int NRHS, int THREADS, int STAGES>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS)
    k_bwd_chain_inv(FrontTab T, const int* __restrict__ nodes, const double* __restrict__ F,
                    const double* __restrict__ dinv, double* work, int64_t work_stride, double* xp, int na,
                    int use_mbar) {
  extern __shared__ __align__(16) double inv_smem[];
  __shared__ __align__(8) unsigned long long mbar[XRING];
  double* Dsm = inv_smem;
  double* xs = inv_smem + (THREADS / 32) * PW * PW;
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0) {
    for (int b = 0; b < XRING; ++b) mbar_init(&mbar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  cl.sync();  // the barriers of every CTA are initialised before anyone signals them
  const int node = nodes[blockIdx.x / CL];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  if (ns == 0) return;
  const double* Fn = F + T.front_off[node];
  double* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int gw = wid * CL + (int)cl.block_rank();  // round-robin panel ownership (see k_fwd_chain_inv)
  const int row0 = gw * PW + lane;
  const int npan = (ns + PW - 1) / PW;
  double* Dw = Dsm + wid * PW * PW;
  if (gw < npan) {
    const double* src = dinv + T.dinv_off[node] + (long long)(npan + gw) * PW * PW;
#pragma unroll
    for (int q = 0; q < PW * PW / 2 / 32; ++q) cp_async16(Dw + 2 * (lane + 32 * q), src + 2 * (lane + 32 * q));
  }
  // upper triangle of U11 into L2 (see k_fwd_chain_inv)
  if (row0 < ns) prefetch_l2_bulk(Fn + (long long)row0 * m, (long long)(row0 + 1) * 8);
  // stage st holds the entries of panel npan - 1 - (step) for the steps = st (mod STAGES); the last panel may be ragged
  double un[STAGES][PW];
#pragma unroll
  for (int st = 0; st < STAGES; ++st) {
    const int kp = npan - 1 - st;
    if (kp >= 0 && row0 < kp * PW) {
      const int wd = min(PW, ns - kp * PW);
#pragma unroll
      for (int j = 0; j < PW; ++j) un[st][j] = j < wd ? Fn[row0 + (long long)(kp * PW + j) * m] : 0.0;
    }
  }
  double acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = row0 < ns ? w[r][row0] : 0.0;
  cp_async_wait_all();
  __syncwarp();
  for (int sb = 0; sb < npan; sb += STAGES) {
#pragma unroll
    for (int st = 0; st < STAGES; ++st) {
      const int step = sb + st;
      if (step < npan) {
        const int k = npan - 1 - step;
        const double* xk = chain_step_exchange<NRHS>(step, gw == k, lane, t, Dw, acc, xs, mbar, use_mbar);
        if (row0 < k * PW) panel_dot_sub<NRHS>(acc, un[st], xk);
        const int kn = k - STAGES;  // always a full panel
        if (kn >= 0 && row0 < kn * PW) {
#pragma unroll
          for (int j = 0; j < PW; ++j) un[st][j] = Fn[row0 + (long long)(kn * PW + j) * m];
        }
      }
    }
  }
  if (row0 < ns) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
      w[r][row0] = acc[r];
      xp[(size_t)r * na + os + row0] = acc[r];
    }
  }
}

// ------------------------------------------------------------------------------ off-diagonal updates
// block = UPD_ROWS rows x UPD_GROUPS column groups; group g handles the columns j = g (mod UPD_GROUPS) of a
// chunk, the partial sums are combined through shared memory in a fixed order (deterministic).
// gridDim.z > 1: the column chunks are dealt round-robin to the z-slices; every slice writes its partial
// sums to scratch[z][work_off + row] and k_upd_reduce adds them to w in a fixed order (deterministic).
template <class S, int NRHS, int G>
__global__ void __launch_bounds__(UPD_ROWS* G, upd_min_blocks(G)) k_fwd_update(FrontTab T, const int* __restrict__ nodes,
                                                            const S* __restrict__ F, S* work, int64_t work_stride,
                                                            S* scratch) {
  constexpr int UPD_GROUPS = G;
  __shared__ S xs[NRHS][UPD_CHUNK];
  __shared__ S part[NRHS][UPD_GROUPS][UPD_ROWS];
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
  const int lr = t % UPD_ROWS, g = t / UPD_ROWS;
  const int row = ns + r0 + lr;
  const bool valid = row < m;
  S acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = scalar_traits<S>::zero();
  for (int j0 = blockIdx.z * UPD_CHUNK; j0 < ns; j0 += UPD_CHUNK * gridDim.z) {
    const int cn = min(UPD_CHUNK, ns - j0);
    for (int j = t; j < cn; j += UPD_ROWS * UPD_GROUPS)
#pragma unroll
      for (int r = 0; r < NRHS; ++r) xs[r][j] = w[r][j0 + j];
    __syncthreads();
    if (valid && g < cn) {
      const int cnt = (cn - g + UPD_GROUPS - 1) / UPD_GROUPS;
      row_dot_sub<S, NRHS, UPD_CHUNK>(acc, Fn + row + (long long)(j0 + g) * m, m, cnt, xs, g, UPD_GROUPS);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < NRHS; ++r) part[r][g][lr] = acc[r];
  __syncthreads();
  if (g == 0 && valid) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
      S s = part[r][0][lr];
#pragma unroll
      for (int q = 1; q < UPD_GROUPS; ++q) s = s + part[r][q][lr];
      if (gridDim.z == 1)
        w[r][row] = w[r][row] + s;
      else
        scratch[((size_t)blockIdx.z * NRHS + r) * work_stride + T.work_off[node] + row] = s;
    }
  }
}

template <class S, int NRHS, int G>
__global__ void __launch_bounds__(UPD_ROWS* G, upd_min_blocks(G)) k_bwd_update(FrontTab T, const int* __restrict__ nodes,
                                                            const S* __restrict__ F, S* work, int64_t work_stride,
                                                            S* scratch) {
  constexpr int UPD_GROUPS = G;
  __shared__ S xs[NRHS][UPD_CHUNK];
  __shared__ S part[NRHS][UPD_GROUPS][UPD_ROWS];
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
  const int lr = t % UPD_ROWS, g = t / UPD_ROWS;
  const int row = r0 + lr;
  const bool valid = row < ns;
  S acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = scalar_traits<S>::zero();
  for (int c0 = blockIdx.z * UPD_CHUNK; c0 < nbd; c0 += UPD_CHUNK * gridDim.z) {
    const int cn = min(UPD_CHUNK, nbd - c0);
    for (int c = t; c < cn; c += UPD_ROWS * UPD_GROUPS) {
      const int ri = rel[c0 + c];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        const S v = wp[r][ri];
        xs[r][c] = v;
        if (blockIdx.y == 0) w[r][ns + c0 + c] = v;
      }
    }
    __syncthreads();
    if (valid && g < cn) {
      const int cnt = (cn - g + UPD_GROUPS - 1) / UPD_GROUPS;
      row_dot_sub<S, NRHS, UPD_CHUNK>(acc, Fn + row + (long long)(ns + c0 + g) * m, m, cnt, xs, g, UPD_GROUPS);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < NRHS; ++r) part[r][g][lr] = acc[r];
  __syncthreads();
  if (g == 0 && valid) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
      S s = part[r][0][lr];
#pragma unroll
      for (int q = 1; q < UPD_GROUPS; ++q) s = s + part[r][q][lr];
      if (gridDim.z == 1)
        w[r][row] = w[r][row] + s;
      else
        scratch[((size_t)blockIdx.z * NRHS + r) * work_stride + T.work_off[node] + row] = s;
    }
  }
}

// w[row] += sum_z scratch[z][row] for the rows [lo_own ? 0 : ns, ...) of every front of a level
template <class S, int NRHS>
__global__ void k_upd_reduce(FrontTab T, const int* __restrict__ nodes, S* work, int64_t work_stride,
                             const S* __restrict__ scratch, int nz, int boundary_rows) {
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int lo = boundary_rows ? ns : 0, hi = boundary_rows ? m : ns;
  if (boundary_rows ? ns == 0 : m == ns) return;  // the update kernel did not run for this front
  for (int row = lo + blockIdx.y * blockDim.x + threadIdx.x; row < hi; row += gridDim.y * blockDim.x) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
      S* w = work + (size_t)r * work_stride + T.work_off[node];
      S s = w[row];
      for (int z = 0; z < nz; ++z) s = s + scratch[((size_t)z * NRHS + r) * work_stride + T.work_off[node] + row];
      w[row] = s;
    }
  }
}

// ------------------------------------------------------------------------------ driver
// ------------------------------------------------------------------------------ dataflow bands
// The lower levels of the tree (many fronts per level) do not run level by level: one persistent launch covers a BAND
// of levels.  CTAs draw fronts from a ticket counter in dependency order (forward: deepest level first, backward: top
// level first) and wait on the completion flags of the fronts they depend on (forward: the two children, backward: the
// parent), so a front starts as soon as its own inputs exist and the tail of one level overlaps the head of the next.
// A ticket is only ever drawn by a running CTA and every dependency has a smaller ticket, so the wait cannot deadlock
// whatever the number of resident CTAs.  Fronts of the first level of a band depend on an earlier launch: no wait.
// Flags carry an epoch (tree_sync[0], bumped by k_tree_reset at the start of every solve) instead of being cleared.
constexpr int TS_EPOCH = 0, TS_TICKET = 1, TS_FLAGS = 8;

__global__ void k_tree_reset(int* __restrict__ sync) {
  sync[TS_EPOCH] += 2;
  sync[TS_TICKET + 0] = sync[TS_TICKET + 1] = sync[TS_TICKET + 2] = sync[TS_TICKET + 3] = 0;
}

__device__ __forceinline__ void wait_flag(const int* flag, int epoch) {
  while (*(const volatile int*)flag != epoch) __nanosleep(40);
}

template <class S, int NRHS, bool FWD>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS, 2) k_tree_band(FrontTab T, const int* __restrict__ order, int n_items,
                                                                 int n_first, int* sync, int ticket_slot,
                                                                 const S* __restrict__ F,
                                                                 const int* __restrict__ rowperm, S* work,
                                                                 int64_t work_stride, S* xp, int na) {
  __shared__ int s_item;
  int* flags = sync + TS_FLAGS;
  const int epoch = sync[TS_EPOCH] + (FWD ? 0 : 1);
  for (;;) {
    __syncthreads();  // the shared memory of the previous front is free
    if (threadIdx.x == 0) s_item = atomicAdd(&sync[TS_TICKET + ticket_slot], 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) return;
    const int node = order[item];
    if (item >= n_first) {
      if (threadIdx.x == 0) {
        if (FWD) {
          const int c0 = T.child0[node], c1 = T.child1[node];
          if (c0 >= 0) wait_flag(flags + c0, epoch);
          if (c1 >= 0) wait_flag(flags + c1, epoch);
        } else {
          wait_flag(flags + T.parent[node], epoch);
        }
        __threadfence();
      }
      __syncthreads();
    }
    if constexpr (FWD)
      fwd_chain_body<S, NRHS, SyncCta, true>(T, node, F, rowperm, work, work_stride, xp, na);
    else
      bwd_chain_body<S, NRHS, SyncCta, true>(T, node, F, work, work_stride, xp, na);
    __syncthreads();  // every thread of the CTA has issued its stores
    if (threadIdx.x == 0) {
      __threadfence();
      *(volatile int*)(flags + node) = epoch;
    }
  }
}

// z-slices so that a level with few, huge fronts still fills the 148 SMs
static int upd_splits(int nfront, int tiles, int cols) {
  const int chunks = (cols + UPD_CHUNK - 1) / UPD_CHUNK;
  const int want = (2 * 148 + nfront * tiles - 1) / std::max(1, nfront * tiles);
  return std::max(1, std::min(std::min(want, chunks), 16));
}
static int chain_threads(int max_ns) {
  if (max_ns <= 64) return 64;
  if (max_ns <= 128) return 128;
  if (max_ns <= 256) return 256;
  return SOLVE_MAX_THREADS;
}

template <class S, int NRHS>
struct InvChain {
  static void launch(bool, int, int, FrontTab, const int*, const S*, LUDevice&, S*, int64_t, S*, int, cudaStream_t) {}
};
template <int NRHS>
struct InvChain<double, NRHS> {
  template <int THREADS, int STAGES>
  static void run(bool fwd, int nfront, FrontTab T, const int* nodes, const double* F, LUDevice& d, double* work,
                  int64_t work_stride, double* xp, int na, cudaStream_t s, int use_mbar) {
    constexpr int smem = (THREADS / 32) * PW * PW * 8 + XRING * NRHS * PW * 8;
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
      FW_CHECK_CUDA(cudaFuncSetAttribute(k_fwd_chain_inv<NRHS, THREADS, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      FW_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_chain_inv<NRHS, THREADS, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    if (fwd)
      k_fwd_chain_inv<NRHS, THREADS, STAGES><<<nfront * CL, THREADS, smem, s>>>(T, nodes, F, d.dinv.p, d.rowperm.p, work,
                                                                              work_stride, xp, na, use_mbar);
    else
      k_bwd_chain_inv<NRHS, THREADS, STAGES><<<nfront * CL, THREADS, smem, s>>>(T, nodes, F, d.dinv.p, work, work_stride,
                                                                              xp, na, use_mbar);
  }
  static void launch(bool fwd, int nfront, int max_ns, FrontTab T, const int* nodes, const double* F, LUDevice& d,
                     double* work, int64_t work_stride, double* xp, int na, cudaStream_t s) {
    static const int use_mbar = getenv("FW_CHAIN_MBAR") ? atoi(getenv("FW_CHAIN_MBAR")) : 1;
    // measured on the 2001-row root front (forward chain): 512 threads / 1 stage 0.125 ms, 256 / 2 stages 0.102 ms,
    // 256 / 3 stages 0.125 ms (255 registers)
    static const int stages = getenv("FW_CHAIN_STAGES") ? atoi(getenv("FW_CHAIN_STAGES")) : 2;
    // pivot blocks of up to 8 x 256 rows: one row per thread with 256-thread CTAs and a deeper register pipeline
    if (max_ns <= CL * 256 && stages >= 3 && NRHS == 1)
      run<256, 3>(fwd, nfront, T, nodes, F, d, work, work_stride, xp, na, s, use_mbar);
    else if (max_ns <= CL * 256 && stages == 2)
      run<256, 2>(fwd, nfront, T, nodes, F, d, work, work_stride, xp, na, s, use_mbar);
    else
      run<CL_THREADS, 1>(fwd, nfront, T, nodes, F, d, work, work_stride, xp, na, s, use_mbar);
  }
};
template <class S, int NRHS>
static void launch_inv_chain(bool fwd, int nfront, int max_ns, FrontTab T, const int* nodes, const S* F, LUDevice& d,
                             S* work, int64_t work_stride, S* xp, int na, cudaStream_t s) {
  InvChain<S, NRHS>::launch(fwd, nfront, max_ns, T, nodes, F, d, work, work_stride, xp, na, s);
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
  T.pf = getenv("FW_SOLVE_PF_KB") ? atoi(getenv("FW_SOLVE_PF_KB")) : 0;  // L2 prefetch budget per small front
  int64_t launches = 2;
  const char* tenv = getenv("FW_TRACE");
  const bool trace2 = tenv && atoi(tenv) >= 2;
  const bool no_cluster = getenv("FW_NO_CLUSTER") != nullptr;
  const bool no_inv = getenv("FW_NO_DINV") != nullptr;
  const bool no_tinv = getenv("FW_NO_TINV_SOLVE") != nullptr || lu.no_tinv;
  // levels of many small fronts: chain and off-diagonal update in one kernel (FW_FUSE_MAX_M=0 disables)
  static const int fuse_max_m = getenv("FW_FUSE_MAX_M") ? atoi(getenv("FW_FUSE_MAX_M")) : 512;
  // forward fused kernel: two rows per thread (blockDim = m / 2) doubles the fronts in flight per SM -- the kernel is
  // bound by the latency of its dependent steps, not by issue slots: leaf level 0.98 -> 0.70 ms (FW_FUSE_TDIV)
  static const int fuse_min_ns = getenv("FW_FUSE_MIN_NS") ? atoi(getenv("FW_FUSE_MIN_NS")) : 0;
  static const int fuse_tdiv = getenv("FW_FUSE_TDIV") ? std::max(1, atoi(getenv("FW_FUSE_TDIV"))) : 2;
  auto use_inv = [&](int max_ns) {
    return !scalar_traits<S>::is_complex && d.dinv_valid && !no_cluster && !no_inv && max_ns > CL_MIN_NS &&
           max_ns <= CL * CL_THREADS;
  };
  // small fronts in real arithmetic: pre-inverted diagonal blocks + cp.async double buffer (FW_NO_DINV_CHAIN=1 disables)
  const bool no_dinv_chain = getenv("FW_NO_DINV_CHAIN") != nullptr || lu.no_tinv;
  static const int dinv_max_threads =
      getenv("FW_DINV_MAX_THREADS") ? std::min(DINV_MAX_THREADS, atoi(getenv("FW_DINV_MAX_THREADS"))) : DINV_MAX_THREADS;
  // FW_DINV_FWD_ONLY=1 / FW_DINV_BWD_ONLY=1: inverse diagonal blocks in one pass only (inv(L_kk) of the unit lower blocks is
  // benign, the accuracy price of the DINV chains comes from inv(U_kk))
  const bool dinv_fwd_only = getenv("FW_DINV_FWD_ONLY") != nullptr, dinv_bwd_only = getenv("FW_DINV_BWD_ONLY") != nullptr;
  auto use_dinv = [&](int threads, bool fwd) {
    if ((fwd && dinv_bwd_only) || (!fwd && dinv_fwd_only)) return false;
    return !scalar_traits<S>::is_complex && d.dinv_valid && !no_dinv_chain && threads <= dinv_max_threads;
  };
  // ---- dataflow bands over the lower levels (see k_tree_band): A = small fronts (128 threads), B = the rest
  // (the knobs are read per call so that the tests can switch the path; a captured graph keeps its own choice)
  const bool no_tree = !(getenv("FW_TREE") && atoi(getenv("FW_TREE")) > 0);  // measured slower than the levels: opt-in
  const int tree_min_fronts = getenv("FW_TREE_MIN_FRONTS") ? atoi(getenv("FW_TREE_MIN_FRONTS")) : 64;
  const int tree_small_m = getenv("FW_TREE_SMALL_M") ? atoi(getenv("FW_TREE_SMALL_M")) : 320;
  const int tree_max_m = getenv("FW_TREE_MAX_M") ? atoi(getenv("FW_TREE_MAX_M")) : 2048;
  const int tree_threads_a = getenv("FW_TREE_THREADS_A") ? atoi(getenv("FW_TREE_THREADS_A")) : 128;
  const int tree_threads_b = getenv("FW_TREE_THREADS_B") ? atoi(getenv("FW_TREE_THREADS_B")) : 512;
  int band_lo = Sy.n_levels, split = Sy.n_levels;  // A = levels [split, n_levels), B = [band_lo, split)
  if (!no_tree) {
    for (int lev = Sy.n_levels - 1; lev >= 0; --lev) {
      if (Sy.level_start[lev + 1] - Sy.level_start[lev] < tree_min_fronts || Sy.level_max_m[lev] > tree_max_m) break;
      band_lo = lev;
    }
    for (int lev = Sy.n_levels - 1; lev >= band_lo && Sy.level_max_m[lev] <= tree_small_m; --lev) split = lev;
  }
  const int n_all = Sy.level_start[Sy.n_levels];
  const int items_a = n_all - Sy.level_start[split], items_b = Sy.level_start[split] - Sy.level_start[band_lo];
  auto band = [&](bool fwd, bool is_a) {
    const int items = is_a ? items_a : items_b;
    if (items <= 0) return;
    const int threads = is_a ? tree_threads_a : tree_threads_b;
    const int grid = std::min(items, 148 * std::max(1, 2048 / threads));
    // forward: d.order_fwd lists the deepest level first (A then B); backward: d.level_nodes lists the top level first
    const int* order = fwd ? d.order_fwd.p + (is_a ? 0 : items_a)
                           : d.level_nodes.p + Sy.level_start[is_a ? split : band_lo];
    const int first_lev = fwd ? (is_a ? Sy.n_levels - 1 : split - 1) : (is_a ? split : band_lo);
    const int n_first = Sy.level_start[first_lev + 1] - Sy.level_start[first_lev];
    const int slot = (fwd ? 0 : 2) + (is_a ? 0 : 1);
    if (fwd)
      k_tree_band<S, NRHS, true><<<grid, threads, 0, s>>>(T, order, items, n_first, d.tree_sync.p, slot, F, d.rowperm.p,
                                                        work, Sy.work_entries, xp, na);
    else
      k_tree_band<S, NRHS, false><<<grid, threads, 0, s>>>(T, order, items, n_first, d.tree_sync.p, slot, F,
                                                         d.rowperm.p, work, Sy.work_entries, xp, na);
    ++launches;
  };
  Trace tr(s);
  if (na > 0) k_lu_permute_in<S><<<cdiv(na, 256), 256, 0, s>>>(na, d.perm.p, b, n, NRHS, xp);
  if (band_lo < Sy.n_levels) {
    k_tree_reset<<<1, 1, 0, s>>>(d.tree_sync.p);
    ++launches;
    band(true, true);
    if (trace2) tr.mark("fwd band A (dataflow)");
    band(true, false);
    if (trace2) tr.mark("fwd band B (dataflow)");
  }
  for (int lev = band_lo - 1; lev >= 0; --lev) {
    const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes = d.level_nodes.p + Sy.level_start[lev];
    const int max_ns = Sy.level_max_ns[lev];
    // measured (500k triangles): forward fusion pays only when the pivot block spans several panels
    bool tinv = false;  // pivot-block solves as triangular matrix-vector products (selective inversion, lu_tinv.cu)
    if constexpr (!scalar_traits<S>::is_complex) tinv = !no_tinv && tinv_level(Sy, d, lev);
    const bool fuse = !tinv && Sy.level_max_m[lev] <= fuse_max_m && nfront >= 296 && max_ns >= fuse_min_ns;
    if (tinv) {
      if constexpr (!scalar_traits<S>::is_complex) {
        tinv_forward_level<NRHS>(h, d, Sy, lev, work, xp, na);
      }
    } else if (fuse) {
      const int threads = chain_threads(Sy.level_max_m[lev] / fuse_tdiv);
      if (use_dinv(threads, true)) {
        if constexpr (!scalar_traits<S>::is_complex)
          k_fwd_chain_dinv<NRHS, true><<<nfront, threads, 0, s>>>(T, nodes, F, d.dinv.p, d.rowperm.p, work, Sy.work_entries, xp, na);
      } else
        k_fwd_chain_fused<S, NRHS><<<nfront, threads, 0, s>>>(T, nodes, F, d.rowperm.p, work, Sy.work_entries, xp, na);
    } else if (use_dinv(chain_threads(max_ns), true) && !use_inv(max_ns) && !(max_ns > CL_MIN_NS && !no_cluster)) {
      if constexpr (!scalar_traits<S>::is_complex)
        k_fwd_chain_dinv<NRHS, false><<<nfront, chain_threads(max_ns), 0, s>>>(T, nodes, F, d.dinv.p, d.rowperm.p, work, Sy.work_entries, xp, na);
    } else if (use_inv(max_ns))
      launch_inv_chain<S, NRHS>(true, nfront, max_ns, T, nodes, F, d, work, Sy.work_entries, xp, na, s);
    else if (max_ns > CL_MIN_NS && !no_cluster)
      k_fwd_chain_cluster<S, NRHS><<<nfront * CL, CL_THREADS, 0, s>>>(T, nodes, F, d.rowperm.p, work,
                                                                      Sy.work_entries, xp, na);
    else
      k_fwd_chain<S, NRHS><<<nfront, chain_threads(max_ns), 0, s>>>(T, nodes, F, d.rowperm.p, work,
                                                                    Sy.work_entries, xp, na);
    ++launches;
    if (trace2)
      tr.mark(("fwd chain  lev " + std::to_string(lev) + " fronts " + std::to_string(nfront) + " max_ns " +
               std::to_string(max_ns)).c_str());
    const int tiles = fuse ? 0 : (Sy.level_max_nb[lev] + UPD_ROWS - 1) / UPD_ROWS;
    if (tiles > 0) {
      const int nz = upd_splits(nfront, tiles, max_ns);
      if (nz > 1) d.scratch.alloc((size_t)std::max<int64_t>(1, Sy.work_entries) * 16 * 2 * wsc);
      S* scratch = (S*)d.scratch.p;
      if (max_ns > 512)
        k_fwd_update<S, NRHS, 4><<<dim3(nfront, tiles, nz), UPD_ROWS * 4, 0, s>>>(T, nodes, F, work, Sy.work_entries, scratch);
      else
        k_fwd_update<S, NRHS, 1><<<dim3(nfront, tiles, nz), UPD_ROWS, 0, s>>>(T, nodes, F, work, Sy.work_entries, scratch);
      ++launches;
      if (nz > 1) {
        k_upd_reduce<S, NRHS><<<dim3(nfront, tiles), UPD_ROWS, 0, s>>>(T, nodes, work, Sy.work_entries, scratch, nz, 1);
        ++launches;
      }
      if (trace2)
        tr.mark(("fwd update lev " + std::to_string(lev) + " max_nb " + std::to_string(Sy.level_max_nb[lev])).c_str());
    }
  }
  for (int lev = 0; lev < band_lo; ++lev) {
    const int nfront = Sy.level_start[lev + 1] - Sy.level_start[lev];
    const int* nodes = d.level_nodes.p + Sy.level_start[lev];
    const int max_ns = Sy.level_max_ns[lev];
    bool tinv = false;
    if constexpr (!scalar_traits<S>::is_complex) tinv = !no_tinv && tinv_level(Sy, d, lev);
    const bool fuse = !tinv && Sy.level_max_m[lev] <= fuse_max_m && nfront >= 296;
    if (Sy.level_max_nb[lev] > 0 && !fuse) {
      const int tiles = std::max(1, (max_ns + UPD_ROWS - 1) / UPD_ROWS);
      const int max_nb = Sy.level_max_nb[lev];
      const int nz = upd_splits(nfront, tiles, max_nb);
      if (nz > 1) d.scratch.alloc((size_t)std::max<int64_t>(1, Sy.work_entries) * 16 * 2 * wsc);
      S* scratch = (S*)d.scratch.p;
      if (max_nb > 512)
        k_bwd_update<S, NRHS, 4><<<dim3(nfront, tiles, nz), UPD_ROWS * 4, 0, s>>>(T, nodes, F, work, Sy.work_entries, scratch);
      else
        k_bwd_update<S, NRHS, 1><<<dim3(nfront, tiles, nz), UPD_ROWS, 0, s>>>(T, nodes, F, work, Sy.work_entries, scratch);
      ++launches;
      if (nz > 1) {
        k_upd_reduce<S, NRHS><<<dim3(nfront, tiles), UPD_ROWS, 0, s>>>(T, nodes, work, Sy.work_entries, scratch, nz, 0);
        ++launches;
      }
      if (trace2) tr.mark(("bwd update lev " + std::to_string(lev)).c_str());
    }
    if (tinv) {
      if constexpr (!scalar_traits<S>::is_complex) {
        tinv_backward_level<NRHS>(h, d, Sy, lev, work, xp, na);
      }
    } else if (fuse) {
      const int threads = chain_threads(std::max(max_ns, 64));
      if (use_dinv(threads, false)) {
        if constexpr (!scalar_traits<S>::is_complex)
          k_bwd_chain_dinv<NRHS, true><<<nfront, threads, 0, s>>>(T, nodes, F, d.dinv.p, work, Sy.work_entries, xp, na);
      } else
        k_bwd_chain_fused<S, NRHS><<<nfront, threads, 0, s>>>(T, nodes, F, work, Sy.work_entries, xp, na);
    } else if (use_dinv(chain_threads(max_ns), false) && !use_inv(max_ns) && !(max_ns > CL_MIN_NS && !no_cluster)) {
      if constexpr (!scalar_traits<S>::is_complex)
        k_bwd_chain_dinv<NRHS, false><<<nfront, chain_threads(max_ns), 0, s>>>(T, nodes, F, d.dinv.p, work, Sy.work_entries, xp, na);
    } else if (use_inv(max_ns))
      launch_inv_chain<S, NRHS>(false, nfront, max_ns, T, nodes, F, d, work, Sy.work_entries, xp, na, s);
    else if (max_ns > CL_MIN_NS && !no_cluster)
      k_bwd_chain_cluster<S, NRHS><<<nfront * CL, CL_THREADS, 0, s>>>(T, nodes, F, work, Sy.work_entries, xp, na);
    else
      k_bwd_chain<S, NRHS><<<nfront, chain_threads(max_ns), 0, s>>>(T, nodes, F, work, Sy.work_entries, xp, na);
    ++launches;
    if (trace2) tr.mark(("bwd chain  lev " + std::to_string(lev)).c_str());
  }
  if (band_lo < Sy.n_levels) {
    band(false, false);
    if (trace2) tr.mark("bwd band B (dataflow)");
    band(false, true);
    if (trace2) tr.mark("bwd band A (dataflow)");
  }
  k_lu_permute_out<S><<<cdiv(n, 256), 256, 0, s>>>(n, d.iperm.p, xp, na, NRHS, x);
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += launches;
}

static void lu_solve_launch(Handle& h, const void* b, void* x, int nrhs);

void lu_solve_dev(Handle& h, const void* b, void* x, int nrhs) {
  FW_REQUIRE(h.lu && h.lu->factored, "LU has not been factored");
  FW_REQUIRE(nrhs == 1 || nrhs == 2, "nrhs must be 1 or 2");
  const bool no_graph = getenv("FW_NO_SOLVE_GRAPH") != nullptr;
  if (no_graph || getenv("FW_TRACE")) {  // (the trace synchronises the stream between the levels)
    lu_solve_launch(h, b, x, nrhs);
    return;
  }
  LU& lu = *h.lu;
  SolveGraph* g = nullptr;
  for (auto& c : lu.graphs)
    if (c.b == b && c.x == x && c.nrhs == nrhs) g = &c;
  if (!g) {
    if (lu.graphs.size() >= 8) lu.drop_graphs();
    lu.graphs.push_back(SolveGraph{b, x, nrhs, 0, 0, nullptr});
    g = &lu.graphs.back();
  }
  ++g->uses;
  if (g->exec) {
    FW_CHECK_CUDA(cudaGraphLaunch(g->exec, h.stream));
    h.tim.kernel_launches += g->kernels;  // kernel nodes replayed by this graph launch
    return;
  }
  if (g->uses < 2) {  // first use: eager (allocates the work buffers, sets the function attributes)
    lu_solve_launch(h, b, x, nrhs);
    return;
  }
  cudaGraph_t graph = nullptr;
  FW_CHECK_CUDA(cudaStreamBeginCapture(h.stream, cudaStreamCaptureModeRelaxed));
  const int64_t launches_before = h.tim.kernel_launches;
  try {
    lu_solve_launch(h, b, x, nrhs);
  } catch (...) {
    cudaStreamEndCapture(h.stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  g->kernels = (int)(h.tim.kernel_launches - launches_before);
  FW_CHECK_CUDA(cudaStreamEndCapture(h.stream, &graph));
  cudaGraphExec_t exec = nullptr;
  FW_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  g->exec = exec;
  FW_CHECK_CUDA(cudaGraphLaunch(exec, h.stream));
}

static void lu_solve_launch(Handle& h, const void* b, void* x, int nrhs) {
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
