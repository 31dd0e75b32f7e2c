// Analysis phase of the multifrontal LU (host).  See lu.h.
//
// Element-based nested dissection: the elements are bisected recursively by the median of their
// centroids along the longer side of the bounding box (a kd-tree).  A dof is owned by the lowest
// tree node whose element set contains every element touching the dof; dofs shared between the
// two halves of a node therefore form its separator.  Because FE couplings only exist inside
// elements, the front of a node is exactly {own dofs} + {ancestor dofs touched by its elements}.
#include <algorithm>
#include <numeric>

#include "lu.h"

namespace fw {

void lu_symbolic_build(LUSymbolic& S, int n, int nb, int ne, const int* edofs, const double* cx, const double* cy,
                       const uint8_t* active, int leaf_elements) {
  if (leaf_elements <= 0) leaf_elements = 32;
  S = LUSymbolic();
  S.n = n;
  // ---- 1. kd-tree over elements
  std::vector<int> eperm(ne);
  std::iota(eperm.begin(), eperm.end(), 0);
  std::vector<int> node_lo, node_hi;
  struct Item {
    int lo, hi, parent, depth;
  };
  std::vector<Item> stack;
  stack.push_back({0, ne, -1, 0});
  while (!stack.empty()) {
    Item it = stack.back();
    stack.pop_back();
    int id = (int)S.parent.size();
    S.parent.push_back(it.parent);
    S.child0.push_back(-1);
    S.child1.push_back(-1);
    S.depth.push_back(it.depth);
    node_lo.push_back(it.lo);
    node_hi.push_back(it.hi);
    if (it.parent >= 0) {
      if (S.child0[it.parent] < 0)
        S.child0[it.parent] = id;
      else
        S.child1[it.parent] = id;
    }
    if (it.hi - it.lo <= leaf_elements) continue;
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int i = it.lo; i < it.hi; ++i) {
      int e = eperm[i];
      x0 = std::min(x0, cx[e]), x1 = std::max(x1, cx[e]);
      y0 = std::min(y0, cy[e]), y1 = std::max(y1, cy[e]);
    }
    const double* c = (x1 - x0 >= y1 - y0) ? cx : cy;
    int mid = (it.lo + it.hi) / 2;
    std::nth_element(eperm.begin() + it.lo, eperm.begin() + mid, eperm.begin() + it.hi,
                     [c](int a, int b) { return c[a] < c[b] || (c[a] == c[b] && a < b); });
    stack.push_back({mid, it.hi, id, it.depth + 1});
    stack.push_back({it.lo, mid, id, it.depth + 1});
  }
  const int nn = (int)S.parent.size();
  S.n_nodes = nn;
  std::vector<int> leaf_of(ne, -1);
  for (int v = 0; v < nn; ++v)
    if (S.child0[v] < 0)
      for (int i = node_lo[v]; i < node_hi[v]; ++i) leaf_of[eperm[i]] = v;

  // ---- 2. owner of every dof = LCA of the leaves of its elements
  std::vector<int> owner(n, -1);
  const std::vector<int>& par = S.parent;
  const std::vector<int>& dep = S.depth;
  for (int e = 0; e < ne; ++e) {
    const int lf = leaf_of[e];
    for (int l = 0; l < nb; ++l) {
      const int d = edofs[(size_t)l * ne + e];
      if (!active[d]) continue;
      int a = owner[d];
      if (a < 0) {
        owner[d] = lf;
        continue;
      }
      int b = lf;
      while (a != b) {
        if (dep[a] > dep[b])
          a = par[a];
        else if (dep[b] > dep[a])
          b = par[b];
        else
          a = par[a], b = par[b];
      }
      owner[d] = a;
    }
  }

  // ---- 3. postorder numbering
  S.ns.assign(nn, 0);
  int n_active = 0;
  for (int d = 0; d < n; ++d)
    if (owner[d] >= 0) {
      S.ns[owner[d]]++;
      n_active++;
    }
  S.n_active = n_active;
  std::vector<int> post;
  post.reserve(nn);
  {
    // iterative postorder
    std::vector<std::pair<int, int>> st;
    st.push_back({0, 0});
    while (!st.empty()) {
      auto& top = st.back();
      int v = top.first;
      if (S.child0[v] < 0) {
        post.push_back(v);
        st.pop_back();
        continue;
      }
      if (top.second == 0) {
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
  S.own_start.assign(nn, 0);
  {
    int run = 0;
    for (int v : post) {
      S.own_start[v] = run;
      run += S.ns[v];
    }
  }
  S.perm.assign(n_active, -1);
  S.iperm.assign(n, -1);
  S.node_of.assign(n_active, -1);
  {
    std::vector<int> fill(nn, 0);
    for (int d = 0; d < n; ++d) {
      int o = owner[d];
      if (o < 0) continue;
      int ni = S.own_start[o] + fill[o]++;
      S.perm[ni] = d;
      S.iperm[d] = ni;
      S.node_of[ni] = o;
    }
  }

  // ---- 4. boundary lists (children before parents) and extend-add maps
  S.m.assign(nn, 0);
  S.blist_off.assign(nn, 0);
  std::vector<std::vector<int>> bl(nn);
  std::vector<int> tmp;
  for (int v : post) {
    const int os = S.own_start[v], nsv = S.ns[v];
    std::vector<int>& out = bl[v];
    int own_seen = 0;
    if (S.child0[v] < 0) {
      tmp.clear();
      for (int i = node_lo[v]; i < node_hi[v]; ++i) {
        int e = eperm[i];
        for (int l = 0; l < nb; ++l) {
          int ni = S.iperm[edofs[(size_t)l * ne + e]];
          if (ni >= 0) tmp.push_back(ni);
        }
      }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      for (int x : tmp) {
        if (x >= os && x < os + nsv)
          own_seen++;
        else if (x >= os + nsv)
          out.push_back(x);
        else
          throw Error(FW_ERR_INTERNAL, "lu symbolic: leaf sees a descendant index");
      }
    } else {
      const std::vector<int>& a = bl[S.child0[v]];
      const std::vector<int>& b = bl[S.child1[v]];
      tmp.resize(a.size() + b.size());
      auto end = std::set_union(a.begin(), a.end(), b.begin(), b.end(), tmp.begin());
      tmp.resize(end - tmp.begin());
      for (int x : tmp) {
        if (x >= os && x < os + nsv)
          own_seen++;
        else if (x >= os + nsv)
          out.push_back(x);
        else
          throw Error(FW_ERR_INTERNAL, "lu symbolic: child boundary holds a descendant index");
      }
    }
    if (own_seen != nsv) throw Error(FW_ERR_INTERNAL, "lu symbolic: separator dofs missing from the front");
    S.m[v] = nsv + (int)out.size();
  }
  {
    int64_t run = 0;
    for (int v = 0; v < nn; ++v) {
      S.blist_off[v] = run;
      run += (int64_t)bl[v].size();
    }
    S.blist.resize((size_t)run);
    S.rel.assign((size_t)run, -1);
    for (int v = 0; v < nn; ++v) std::copy(bl[v].begin(), bl[v].end(), S.blist.begin() + S.blist_off[v]);
    for (int v = 0; v < nn; ++v) {
      int p = S.parent[v];
      if (p < 0) continue;
      const std::vector<int>& mine = bl[v];
      const std::vector<int>& pb = bl[p];
      const int pos = S.own_start[p], pns = S.ns[p];
      size_t k = 0;
      for (size_t i = 0; i < mine.size(); ++i) {
        int x = mine[i];
        int r;
        if (x < pos + pns) {
          r = x - pos;
        } else {
          while (k < pb.size() && pb[k] < x) ++k;
          if (k >= pb.size() || pb[k] != x) throw Error(FW_ERR_INTERNAL, "lu symbolic: extend-add map broken");
          r = pns + (int)k;
        }
        S.rel[(size_t)S.blist_off[v] + i] = r;
      }
    }
  }

  // ---- 5. levels and arena offsets (deepest level first in memory)
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
  S.front_off.assign(nn, 0);
  S.work_off.assign(nn, 0);
  S.level_max_ns.assign(S.n_levels, 0);
  S.level_max_m.assign(S.n_levels, 0);
  S.level_max_nb.assign(S.n_levels, 0);
  int64_t fo = 0, wo = 0;
  for (int d = S.n_levels - 1; d >= 0; --d)
    for (int i = S.level_start[d]; i < S.level_start[d + 1]; ++i) {
      int v = S.level_nodes[i];
      S.front_off[v] = fo;
      S.work_off[v] = wo;
      fo += (int64_t)S.m[v] * S.m[v];
      wo += S.m[v];
      S.level_max_ns[d] = std::max(S.level_max_ns[d], S.ns[v]);
      S.level_max_m[d] = std::max(S.level_max_m[d], S.m[v]);
      S.level_max_nb[d] = std::max(S.level_max_nb[d], S.m[v] - S.ns[v]);
      S.max_m = std::max(S.max_m, S.m[v]);
      S.max_ns = std::max(S.max_ns, S.ns[v]);
    }
  S.factor_entries = fo;
  S.work_entries = wo;
}

}  // namespace fw
