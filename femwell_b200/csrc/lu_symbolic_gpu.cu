// Analysis phase of the multifrontal LU ON THE DEVICE.  Same algorithm and bit-identical results as the host analysis
// (lu_symbolic.cpp): element-based nested dissection by a kd-tree over the element centroids, a dof is owned by the
// lowest tree node that holds all its elements, postorder numbering with the dofs of a node in ascending original
// order, sorted boundary lists and extend-add maps.  The host version costs ~120 ms with 16 threads at 500k triangles
// (of which 65-95 ms were exposed per cold solve, and 8 ranks x 16 threads oversubscribed the host of an 8-GPU node);
// here everything that touches per-dof / per-element data is integer work on the GPU built from the radix sort of
// sort.cu, and the host only handles the geometry-independent tree topology (O(number of fronts)).
//
//   1. tree topology: the kd-tree splits [lo, hi) at (lo + hi) / 2, so parents / children / depths / ranges depend on
//      (n_elements, leaf size) only                                                                         (host)
//   2. ranks of the centroid coordinates along x and y: two stable 64-bit sorts (ties by element index)
//   3. level by level: bounding box of every node (warp-aggregated 64-bit atomics), then ONE stable sort of the keys
//      (lo of the node, rank along the longer side) -- elements of a node stay inside its range and end up ordered
//      along the split axis, so the lower half is exactly the set std::nth_element selects on the host
//   4. owner of a dof = deepest node whose range holds [min, max] of the positions of its elements (atomicMin/Max +
//      a descent), ns = histogram of the owners
//   5. numbering: one sort of (postorder rank of the owner, dof)
//   6. boundary lists: every (dof, element) pair emits (node, new index) for the nodes between the leaf of the element
//      and the owner of the dof; sort + unique; extend-add maps by binary search in the parent's list
#include <algorithm>
#include <numeric>

#include "lu_dev.h"

namespace fw {

namespace {

struct TreeDev {
  const int *parent, *child0, *child1, *depth, *lo, *hi;
};

__device__ __forceinline__ unsigned long long ordered_u64(double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double from_ordered_u64(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
#ifdef __CUDA_ARCH__
  d = __longlong_as_double((long long)u);
#else
  memcpy(&d, &u, 8);
#endif
  return d;
}

__global__ void k_centroid_keys(int ne, int nv, const double* __restrict__ p, const int* __restrict__ t,
                                double* __restrict__ cx, double* __restrict__ cy, uint64_t* __restrict__ kx,
                                uint64_t* __restrict__ ky) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int a = t[e], b = t[ne + e], c = t[2 * (size_t)ne + e];
  // the same expression as the host analysis (lu_numeric.cu lu_host_analysis): identical doubles
  const double x = (p[a] + p[b] + p[c]) / 3.0;
  const double y = (p[nv + a] + p[nv + b] + p[nv + c]) / 3.0;
  cx[e] = x, cy[e] = y;
  kx[e] = ordered_u64(x), ky[e] = ordered_u64(y);
}
__global__ void k_generic_keys(int ne, const double* __restrict__ cx, const double* __restrict__ cy,
                               uint64_t* __restrict__ kx, uint64_t* __restrict__ ky) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  kx[e] = ordered_u64(cx[e]), ky[e] = ordered_u64(cy[e]);
}
__global__ void k_rank_of(int n, const uint32_t* __restrict__ sorted_payload, uint32_t* __restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rank[sorted_payload[i]] = (uint32_t)i;
}
__global__ void k_iota_u32(int n, uint32_t* __restrict__ a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}
__global__ void k_fill_u64(int n, uint64_t v, uint64_t* __restrict__ a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_fill_i32(int n, int v, int* __restrict__ a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// bounding boxes of the nodes that split at this level: bbox[0..3][node] = min x, max x, min y, max y (ordered keys)
__global__ void __launch_bounds__(256) k_level_bbox(int ne, int level, TreeDev T, const int* __restrict__ pnode,
                                                    const uint32_t* __restrict__ eperm, const uint64_t* __restrict__ kx,
                                                    const uint64_t* __restrict__ ky, int nn,
                                                    unsigned long long* __restrict__ bbox) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < ne;
  int v = -1;
  unsigned long long x = 0, y = 0;
  if (live) {
    v = pnode[i];
    if (T.child0[v] < 0 || T.depth[v] != level) v = -1;  // does not split at this level
    if (v >= 0) {
      const uint32_t e = eperm[i];
      x = kx[e], y = ky[e];
    }
  }
  // one atomic per warp when the whole warp sits in one node (the common case on the upper levels)
  const int v0 = __shfl_sync(0xffffffffu, v, 0);
  if (__all_sync(0xffffffffu, v == v0)) {
    if (v0 < 0) return;
    unsigned long long xmin = x, xmax = x, ymin = y, ymax = y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, xmin, o), b = __shfl_xor_sync(0xffffffffu, xmax, o);
      const unsigned long long c = __shfl_xor_sync(0xffffffffu, ymin, o), d = __shfl_xor_sync(0xffffffffu, ymax, o);
      xmin = a < xmin ? a : xmin, xmax = b > xmax ? b : xmax;
      ymin = c < ymin ? c : ymin, ymax = d > ymax ? d : ymax;
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&bbox[v0], xmin);
      atomicMax(&bbox[(size_t)nn + v0], xmax);
      atomicMin(&bbox[2 * (size_t)nn + v0], ymin);
      atomicMax(&bbox[3 * (size_t)nn + v0], ymax);
    }
    return;
  }
  if (v >= 0) {
    atomicMin(&bbox[v], x);
    atomicMax(&bbox[(size_t)nn + v], x);
    atomicMin(&bbox[2 * (size_t)nn + v], y);
    atomicMax(&bbox[3 * (size_t)nn + v], y);
  }
}

// key = (lo of the node << sb) | (rank along the longer side of its bounding box), or (position - lo) where nothing splits
__global__ void __launch_bounds__(256) k_level_keys(int ne, int level, TreeDev T, const int* __restrict__ pnode,
                                                    const uint32_t* __restrict__ eperm, const uint32_t* __restrict__ rx,
                                                    const uint32_t* __restrict__ ry, int nn,
                                                    const unsigned long long* __restrict__ bbox, int sb,
                                                    uint64_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  const int v = pnode[i];
  const int lo = T.lo[v];
  uint64_t sub;
  if (T.child0[v] >= 0 && T.depth[v] == level) {
    const double x0 = from_ordered_u64(bbox[v]), x1 = from_ordered_u64(bbox[(size_t)nn + v]);
    const double y0 = from_ordered_u64(bbox[2 * (size_t)nn + v]), y1 = from_ordered_u64(bbox[3 * (size_t)nn + v]);
    const uint32_t e = eperm[i];
    sub = (x1 - x0 >= y1 - y0) ? rx[e] : ry[e];
  } else {
    sub = (uint64_t)(i - lo);
  }
  keys[i] = ((uint64_t)lo << sb) | sub;
}
// after the sort of a level: positions of a node that split move on to its children
__global__ void k_level_advance(int ne, int level, TreeDev T, int* __restrict__ pnode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  const int v = pnode[i];
  if (T.child0[v] >= 0 && T.depth[v] == level) {
    const int mid = (T.lo[v] + T.hi[v]) / 2;
    pnode[i] = i < mid ? T.child0[v] : T.child1[v];
  }
}
// leaf_of[e] and epos[e] from the final order
__global__ void k_leaf_of(int ne, const uint32_t* __restrict__ eperm, const int* __restrict__ pnode,
                          int* __restrict__ epos, int* __restrict__ leaf_of) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  const uint32_t e = eperm[i];
  epos[e] = i;
  leaf_of[e] = pnode[i];
}
__global__ void k_minmax_pos(int64_t total, int ne, const int* __restrict__ edofs, const int* __restrict__ epos,
                             const uint8_t* __restrict__ active, int* __restrict__ minpos, int* __restrict__ maxpos) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx % ne);
  const int d = edofs[idx];
  if (!active[d]) return;
  const int pos = epos[e];
  atomicMin(&minpos[d], pos);
  atomicMax(&maxpos[d], pos);
}
__global__ void k_owner(int n, int max_depth, TreeDev T, const int* __restrict__ minpos, const int* __restrict__ maxpos,
                        const uint8_t* __restrict__ active, int* __restrict__ owner, int* __restrict__ ns) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const int a = minpos[d], b = maxpos[d];
  if (!active[d] || b < 0) {
    owner[d] = -1;
    return;
  }
  int v = 0;
  for (int it = 0; it < max_depth; ++it) {
    const int c0 = T.child0[v];
    if (c0 < 0) break;
    const int mid = (T.lo[v] + T.hi[v]) / 2;
    if (b < mid)
      v = c0;
    else if (a >= mid)
      v = T.child1[v];
    else
      break;
  }
  owner[d] = v;
  atomicAdd(&ns[v], 1);
}
__global__ void k_number_keys(int n, int nn, int nbits, const int* __restrict__ owner, const int* __restrict__ prank,
                              uint64_t* __restrict__ keys) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const int o = owner[d];
  keys[d] = ((uint64_t)(o >= 0 ? prank[o] : nn) << nbits) | (uint64_t)d;
}
__global__ void k_numbering(int n, int na, const uint32_t* __restrict__ sorted_dof, const int* __restrict__ owner,
                            int* __restrict__ perm, int* __restrict__ iperm, int* __restrict__ node_of) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = (int)sorted_dof[i];
  if (i < na) {
    perm[i] = d;
    iperm[d] = i;
    node_of[i] = owner[d];
  } else {
    iperm[d] = -1;
  }
}
// number of (node, index) pairs a (local dof, element) pair emits = depth(leaf) - depth(owner)
__global__ void k_emit_count(int64_t total, int ne, const int* __restrict__ edofs, const int* __restrict__ owner,
                             const int* __restrict__ leaf_of, const int* __restrict__ depth,
                             uint32_t* __restrict__ cnt) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx % ne);
  const int o = owner[edofs[idx]];
  cnt[idx] = o < 0 ? 0u : (uint32_t)(depth[leaf_of[e]] - depth[o]);
}
__global__ void k_emit(int64_t total, int ne, const int* __restrict__ edofs, const int* __restrict__ owner,
                       const int* __restrict__ leaf_of, const int* __restrict__ parent, const int* __restrict__ iperm,
                       const uint32_t* __restrict__ offs, int ibits, uint64_t* __restrict__ keys) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx % ne);
  const int d = edofs[idx];
  const int o = owner[d];
  if (o < 0) return;
  const uint64_t ni = (uint64_t)iperm[d];
  uint32_t w = offs[idx];
  for (int v = leaf_of[e]; v != o; v = parent[v]) keys[w++] = ((uint64_t)v << ibits) | ni;
}
__global__ void k_unique_heads(int64_t n, const uint64_t* __restrict__ keys, uint32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void k_unique_compact(int64_t n, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pos,
                                 int ibits, int* __restrict__ blist, int* __restrict__ bnode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || keys[i] != keys[i - 1]) {
    const uint32_t w = pos[i];
    blist[w] = (int)(keys[i] & ((1ull << ibits) - 1ull));
    bnode[w] = (int)(keys[i] >> ibits);
  }
}
// first entry of every node in the (node-sorted) unique list: boff[v] = lower_bound(bnode, v), boff[nn] = nu
__global__ void k_node_ranges(int nn, int64_t nu, const int* __restrict__ bnode, long long* __restrict__ boff) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > nn) return;
  int64_t lo = 0, hi = nu;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (bnode[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  boff[v] = lo;
}
__global__ void k_front_sizes(int nn, const int* __restrict__ ns, const long long* __restrict__ boff,
                              int* __restrict__ m) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nn) m[v] = ns[v] + (int)(boff[v + 1] - boff[v]);
}
__global__ void k_rel(int64_t nu, const int* __restrict__ blist, const int* __restrict__ bnode,
                      const int* __restrict__ parent, const int* __restrict__ own_start, const int* __restrict__ ns,
                      const long long* __restrict__ boff, int* __restrict__ rel, int* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nu) return;
  const int v = bnode[i];
  const int p = parent[v];
  const int x = blist[i];
  if (p < 0) {  // the root has no boundary
    atomicExch(err, 1);
    return;
  }
  const int pos = own_start[p], pns = ns[p];
  if (x < pos) {
    atomicExch(err, 2);  // a front never holds a descendant index
  } else if (x < pos + pns) {
    rel[i] = x - pos;
  } else {
    long long lo = boff[p], hi = boff[p + 1];
    const long long base = lo;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (blist[mid] < x)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo >= boff[p + 1] || blist[lo] != x) atomicExch(err, 3);
    rel[i] = pns + (int)(lo - base);
  }
}

int bit_len(int64_t v) {
  int b = 0;
  while (v > 0) ++b, v >>= 1;
  return b < 1 ? 1 : b;
}

template <class T>
void up_small(DevBuf<T>& d, const std::vector<T>& v, cudaStream_t s) {
  d.alloc(std::max<size_t>(1, v.size()));
  if (!v.empty()) FW_CHECK_CUDA(cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

}  // namespace

// Fills lu.sym (small host arrays: topology, per-front sizes and offsets, level statistics) and lu.dev (all
// device tables).  Elements: the FE mesh of the space, or the clique cover of the generic CSR path.
void lu_symbolic_build_gpu(Handle& h, LU& lu, const std::vector<uint8_t>& active_host, int leaf_elements) {
  if (leaf_elements <= 0) leaf_elements = 16;
  cudaStream_t s = h.stream;
  Trace tr(s);
  const bool generic = h.generic.valid;
  const Space& sp = h.sp;
  const int n = generic ? h.generic.n : sp.n;
  const int nb = generic ? h.generic.nb : sp.nb;
  const int ne = generic ? h.generic.ne : sp.ne;
  FW_REQUIRE(ne > 0 && n > 0, "empty problem");
  LUSymbolic& S = lu.sym;
  S = LUSymbolic();
  S.n = n;
  // ---- 1. topology (host; breadth first, node ids as in lu_symbolic.cpp)
  std::vector<int> lo{0}, hi{ne};
  S.parent = {-1}, S.child0 = {-1}, S.child1 = {-1}, S.depth = {0};
  {
    std::vector<int> level{0}, next;
    while (!level.empty()) {
      next.clear();
      for (int v : level) {
        if (hi[v] - lo[v] <= leaf_elements) continue;
        const int mid = (lo[v] + hi[v]) / 2, d = S.depth[v];
        for (int c = 0; c < 2; ++c) {
          const int id = (int)S.parent.size();
          S.parent.push_back(v), S.child0.push_back(-1), S.child1.push_back(-1), S.depth.push_back(d + 1);
          lo.push_back(c ? mid : lo[v]), hi.push_back(c ? hi[v] : mid);
          (c ? S.child1[v] : S.child0[v]) = id;
          next.push_back(id);
        }
      }
      level.swap(next);
    }
  }
  const int nn = (int)S.parent.size();
  S.n_nodes = nn;
  int maxd = 0;
  for (int v = 0; v < nn; ++v) maxd = std::max(maxd, S.depth[v]);
  S.n_levels = maxd + 1;
  S.level_start.assign(S.n_levels + 1, 0);
  for (int v = 0; v < nn; ++v) S.level_start[S.depth[v] + 1]++;
  for (int d = 0; d < S.n_levels; ++d) S.level_start[d + 1] += S.level_start[d];
  S.level_nodes.assign(nn, 0);
  {
    std::vector<int> fill(S.level_start.begin(), S.level_start.end() - 1);
    for (int v = 0; v < nn; ++v) S.level_nodes[fill[S.depth[v]]++] = v;
  }
  std::vector<int> post;
  post.reserve(nn);
  {
    std::vector<std::pair<int, int>> st{{0, 0}};
    while (!st.empty()) {
      auto& top = st.back();
      const int v = top.first;
      if (S.child0[v] < 0) {
        post.push_back(v);
        st.pop_back();
      } else if (top.second == 0) {
        top.second = 1;
        st.push_back({S.child0[v], 0});
      } else if (top.second == 1) {
        top.second = 2;
        st.push_back({S.child1[v], 0});
      } else {
        post.push_back(v);
        st.pop_back();
      }
    }
  }
  std::vector<int> prank(nn);
  for (int i = 0; i < nn; ++i) prank[post[i]] = i;

  lu.dev.reset(new LUDevice(h.lu_arena));
  LUDevice& d = *lu.dev;
  DevBuf<int> dlo, dhi, ddepth, dprank;
  up_small(d.parent, S.parent, s), up_small(d.child0, S.child0, s), up_small(d.child1, S.child1, s);
  up_small(ddepth, S.depth, s), up_small(dlo, lo, s), up_small(dhi, hi, s), up_small(dprank, prank, s);
  up_small(d.level_nodes, S.level_nodes, s);
  d.active.upload(active_host.data(), active_host.size(), s);
  const TreeDev T{d.parent.p, d.child0.p, d.child1.p, ddepth.p, dlo.p, dhi.p};
  tr.mark("lu analyse (device): topology");

  // ---- 2. centroid ranks
  const int ge = cdiv(ne, 256);
  DevBuf<double> cx, cy;
  DevBuf<uint64_t> kx, ky, ka, kb;
  DevBuf<uint32_t> pa, pb, rx, ry;
  cx.alloc(ne), cy.alloc(ne), kx.alloc(ne), ky.alloc(ne), rx.alloc(ne), ry.alloc(ne);
  DevBuf<int> gedofs;  // generic path: the clique cover lives on the host
  const int* edofs_dev = nullptr;
  if (generic) {
    cx.upload(h.generic.cx.data(), ne, s), cy.upload(h.generic.cy.data(), ne, s);
    gedofs.upload(h.generic.edofs.data(), (size_t)nb * ne, s);
    edofs_dev = gedofs.p;
    k_generic_keys<<<ge, 256, 0, s>>>(ne, cx.p, cy.p, kx.p, ky.p);
  } else {
    edofs_dev = sp.edofs.p;
    k_centroid_keys<<<ge, 256, 0, s>>>(ne, sp.nv, sp.p.p, sp.t.p, cx.p, cy.p, kx.p, ky.p);
  }
  for (int axis = 0; axis < 2; ++axis) {
    ka.alloc(ne);
    FW_CHECK_CUDA(cudaMemcpyAsync(ka.p, axis ? ky.p : kx.p, (size_t)ne * 8, cudaMemcpyDeviceToDevice, s));
    radix_sort_pairs_dev(h, ka, kb, pa, pb, ne, 64, true);
    k_rank_of<<<ge, 256, 0, s>>>(ne, pa.p, axis ? ry.p : rx.p);
  }
  tr.mark("lu analyse (device): centroid ranks");

  // ---- 3. kd-tree order of the elements, level by level
  DevBuf<int> pnode, epos, leaf_of;
  DevBuf<unsigned long long> bbox;
  pnode.alloc(ne), epos.alloc(ne), leaf_of.alloc(ne), bbox.alloc((size_t)4 * nn);
  pnode.zero(s);
  pa.alloc(ne);
  k_iota_u32<<<ge, 256, 0, s>>>(ne, pa.p);  // eperm
  const int sb = bit_len(ne);
  for (int level = 0; level < maxd; ++level) {
    k_fill_u64<<<cdiv(2 * nn, 256), 256, 0, s>>>(nn, ~0ull, (uint64_t*)bbox.p);                         // min x
    k_fill_u64<<<cdiv(2 * nn, 256), 256, 0, s>>>(nn, 0ull, (uint64_t*)bbox.p + nn);                     // max x
    k_fill_u64<<<cdiv(2 * nn, 256), 256, 0, s>>>(nn, ~0ull, (uint64_t*)bbox.p + 2 * (size_t)nn);        // min y
    k_fill_u64<<<cdiv(2 * nn, 256), 256, 0, s>>>(nn, 0ull, (uint64_t*)bbox.p + 3 * (size_t)nn);         // max y
    k_level_bbox<<<ge, 256, 0, s>>>(ne, level, T, pnode.p, pa.p, kx.p, ky.p, nn, bbox.p);
    ka.alloc(ne);
    k_level_keys<<<ge, 256, 0, s>>>(ne, level, T, pnode.p, pa.p, rx.p, ry.p, nn, bbox.p, sb, ka.p);
    radix_sort_pairs_dev(h, ka, kb, pa, pb, ne, 2 * sb, false);
    k_level_advance<<<ge, 256, 0, s>>>(ne, level, T, pnode.p);
  }
  k_leaf_of<<<ge, 256, 0, s>>>(ne, pa.p, pnode.p, epos.p, leaf_of.p);
  tr.mark("lu analyse (device): kd-tree order");

  // ---- 4. owners
  const int64_t total = (int64_t)nb * ne;
  DevBuf<int> minpos, maxpos, owner, dns;
  minpos.alloc(n), maxpos.alloc(n), owner.alloc(n), dns.alloc(nn);
  const int gn = cdiv(n, 256);
  k_fill_i32<<<gn, 256, 0, s>>>(n, ne, minpos.p);
  k_fill_i32<<<gn, 256, 0, s>>>(n, -1, maxpos.p);
  dns.zero(s);
  k_minmax_pos<<<cdiv(total, 256), 256, 0, s>>>(total, ne, edofs_dev, epos.p, d.active.p, minpos.p, maxpos.p);
  k_owner<<<gn, 256, 0, s>>>(n, maxd, T, minpos.p, maxpos.p, d.active.p, owner.p, dns.p);
  S.ns.assign(nn, 0);
  dns.download(S.ns.data(), nn, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  S.own_start.assign(nn, 0);
  int64_t na64 = 0;
  for (int v : post) {
    S.own_start[v] = (int)na64;
    na64 += S.ns[v];
  }
  const int na = (int)na64;
  S.n_active = na;
  up_small(d.own_start, S.own_start, s);
  d.ns = std::move(dns);
  tr.mark("lu analyse (device): owners");

  // ---- 5. numbering
  const int nbits = bit_len(n);
  ka.alloc(n);
  k_number_keys<<<gn, 256, 0, s>>>(n, nn, nbits, owner.p, dprank.p, ka.p);
  radix_sort_pairs_dev(h, ka, kb, pa, pb, n, nbits + bit_len(nn + 1), true);
  d.perm.alloc(std::max(1, na)), d.iperm.alloc(n), d.node_of.alloc(std::max(1, na));
  k_numbering<<<gn, 256, 0, s>>>(n, na, pa.p, owner.p, d.perm.p, d.iperm.p, d.node_of.p);
  tr.mark("lu analyse (device): numbering");

  // ---- 6. boundary lists and extend-add maps
  DevBuf<uint32_t> cnt;
  cnt.alloc((size_t)total + 1);
  k_emit_count<<<cdiv(total, 256), 256, 0, s>>>(total, ne, edofs_dev, owner.p, leaf_of.p, ddepth.p, cnt.p);
  uint32_t last_cnt = 0, last_off = 0;
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_cnt, cnt.p + (total - 1), 4, cudaMemcpyDeviceToHost, s));
  exclusive_scan_u32(cnt.p, cnt.p, total, h.scan_tmp, s);
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_off, cnt.p + (total - 1), 4, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  const int64_t n_emit = (int64_t)last_off + last_cnt;
  FW_REQUIRE(n_emit < (int64_t)0xffffffffll, "too many boundary pairs for 32-bit offsets");
  const int ibits = bit_len(std::max(1, na));
  int64_t nu = 0;
  DevBuf<int> bnode;
  if (n_emit > 0) {
    ka.alloc((size_t)n_emit);
    k_emit<<<cdiv(total, 256), 256, 0, s>>>(total, ne, edofs_dev, owner.p, leaf_of.p, d.parent.p, d.iperm.p, cnt.p, ibits,
                                           ka.p);
    radix_sort_pairs_dev(h, ka, kb, pa, pb, n_emit, ibits + bit_len(nn), true);
    DevBuf<uint32_t> flag;
    flag.alloc((size_t)n_emit);
    k_unique_heads<<<cdiv(n_emit, 256), 256, 0, s>>>(n_emit, ka.p, flag.p);
    uint32_t lf = 0, ls = 0;
    FW_CHECK_CUDA(cudaMemcpyAsync(&lf, flag.p + (n_emit - 1), 4, cudaMemcpyDeviceToHost, s));
    exclusive_scan_u32(flag.p, flag.p, n_emit, h.scan_tmp, s);
    FW_CHECK_CUDA(cudaMemcpyAsync(&ls, flag.p + (n_emit - 1), 4, cudaMemcpyDeviceToHost, s));
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    nu = (int64_t)ls + lf;
    d.blist.alloc((size_t)nu), d.rel.alloc((size_t)nu), bnode.alloc((size_t)nu);
    k_unique_compact<<<cdiv(n_emit, 256), 256, 0, s>>>(n_emit, ka.p, flag.p, ibits, d.blist.p, bnode.p);
  } else {
    d.blist.alloc(1), d.rel.alloc(1), bnode.alloc(1);
  }
  DevBuf<long long> boff;
  boff.alloc((size_t)nn + 1);
  k_node_ranges<<<cdiv(nn + 1, 256), 256, 0, s>>>(nn, nu, bnode.p, boff.p);
  d.m.alloc(nn);
  k_front_sizes<<<cdiv(nn, 256), 256, 0, s>>>(nn, d.ns.p, boff.p, d.m.p);
  DevBuf<int> err;
  err.alloc(1);
  err.zero(s);
  if (nu > 0)
    k_rel<<<cdiv(nu, 256), 256, 0, s>>>(nu, d.blist.p, bnode.p, d.parent.p, d.own_start.p, d.ns.p, boff.p, d.rel.p, err.p);
  S.m.assign(nn, 0);
  std::vector<long long> hboff((size_t)nn + 1);
  int herr = 0;
  d.m.download(S.m.data(), nn, s);
  boff.download(hboff.data(), (size_t)nn + 1, s);
  err.download(&herr, 1, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  FW_CHECK_CUDA(cudaGetLastError());
  if (herr) throw Error(FW_ERR_INTERNAL, "lu symbolic (device): extend-add map broken (code " + std::to_string(herr) + ")");
  S.blist_off.assign(hboff.begin(), hboff.begin() + nn);
  d.blist_off = std::move(boff);
  tr.mark("lu analyse (device): boundary lists + extend-add maps");

  // ---- arena offsets and level statistics (host, O(fronts))
  S.front_off.assign(nn, 0);
  S.work_off.assign(nn, 0);
  S.level_max_ns.assign(S.n_levels, 0);
  S.level_max_m.assign(S.n_levels, 0);
  S.level_max_nb.assign(S.n_levels, 0);
  int64_t fo = 0, wo = 0;
  for (int dd = S.n_levels - 1; dd >= 0; --dd)
    for (int i = S.level_start[dd]; i < S.level_start[dd + 1]; ++i) {
      const int v = S.level_nodes[i];
      S.front_off[v] = fo;
      S.work_off[v] = wo;
      fo += (int64_t)S.m[v] * S.m[v];
      wo += S.m[v];
      S.level_max_ns[dd] = std::max(S.level_max_ns[dd], S.ns[v]);
      S.level_max_m[dd] = std::max(S.level_max_m[dd], S.m[v]);
      S.level_max_nb[dd] = std::max(S.level_max_nb[dd], S.m[v] - S.ns[v]);
      S.max_m = std::max(S.max_m, S.m[v]);
      S.max_ns = std::max(S.max_ns, S.ns[v]);
      const double m = S.m[v], ns = S.ns[v], nbd = m - ns;
      S.solve_entries += (int64_t)S.m[v] * S.m[v] - (int64_t)(S.m[v] - S.ns[v]) * (S.m[v] - S.ns[v]);
      auto sq = [](double x) { return x * (x + 1) * (2 * x + 1) / 6; };
      S.factor_flops += 2.0 * (sq(m - 1) - sq(nbd - 1 < 0 ? 0 : nbd - 1)) + 0.5 * (m - 1 + nbd) * ns;
    }
  S.factor_entries = fo;
  S.work_entries = wo;
  {
    std::vector<long long> f64(S.front_off.begin(), S.front_off.end()), w64(S.work_off.begin(), S.work_off.end());
    up_small(d.front_off, f64, s);
    up_small(d.work_off, w64, s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));  // the staging vectors go out of scope
  }
  tr.mark("lu analyse (device): offsets");
}

// tests: download the device tables of the current analysis.  which: 0 perm, 1 iperm, 2 node_of, 7 m, 8 ns,
// 9 own_start, 10 blist, 11 rel (int32); 21 blist_off (int64)
void lu_debug_analysis_download(Handle& h, int which, void* out) {
  FW_REQUIRE(h.lu && h.lu->dev, "no analysis");
  const LUSymbolic& S = h.lu->sym;
  LUDevice& d = *h.lu->dev;
  cudaStream_t s = h.stream;
  const size_t nbl = (size_t)(S.work_entries - S.n_active);
  switch (which) {
    case 0: d.perm.download((int*)out, (size_t)S.n_active, s); break;
    case 1: d.iperm.download((int*)out, (size_t)S.n, s); break;
    case 2: d.node_of.download((int*)out, (size_t)S.n_active, s); break;
    case 7: d.m.download((int*)out, (size_t)S.n_nodes, s); break;
    case 8: d.ns.download((int*)out, (size_t)S.n_nodes, s); break;
    case 9: d.own_start.download((int*)out, (size_t)S.n_nodes, s); break;
    case 10: d.blist.download((int*)out, nbl, s); break;
    case 11: d.rel.download((int*)out, nbl, s); break;
    case 21: d.blist_off.download((long long*)out, (size_t)S.n_nodes, s); break;
    default: throw Error(FW_ERR_INVALID, "unknown array id");
  }
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
}

}  // namespace fw
