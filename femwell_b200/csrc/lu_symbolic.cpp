// Analysis phase of the multifrontal LU (host).  See lu.h.
//
// Element-based nested dissection: the elements are bisected recursively by the median of their
// centroids along the longer side of the bounding box (a kd-tree).  A dof is owned by the lowest
// tree node whose element set contains every element touching the dof; dofs shared between the
// two halves of a node therefore form its separator.  Because FE couplings only exist inside
// elements, the front of a node is exactly {own dofs} + {ancestor dofs touched by its elements}.
//
// The tree levels are processed with a small std::thread pool (nodes of one level are independent).
#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>

#include <climits>
#include <mutex>
#include <functional>
#include <condition_variable>

#include "lu.h"

namespace fw {

namespace {
int worker_count() {
  static int n = [] {
    const char* e = getenv("FW_HOST_THREADS");
    int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(v, 16));
  }();
  return n;
}

// Persistent workers: one analysis issues ~40 parallel loops (one per tree level and phase); creating and joining
// 16 threads for each of them costs more than many of the loops themselves.
class WorkerPool {
 public:
  static WorkerPool& instance() {
    static WorkerPool p(worker_count() - 1);
    return p;
  }
  // f(i) for i in [0, n); the caller takes part.  Returns false when the pool is busy (another analysis).
  bool run(int n, const std::function<void(int)>& f) {
    std::unique_lock<std::mutex> busy(busy_, std::try_to_lock);
    if (!busy.owns_lock()) return false;
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &f;
      n_ = n;
      next_ = 0;
      pending_ = n;
      ++generation_;
    }
    cv_start_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
    return true;
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_start_.notify_all();
    for (auto& t : th_) t.join();
  }

 private:
  explicit WorkerPool(int nthreads) {
    for (int i = 0; i < nthreads; ++i) th_.emplace_back([this] { loop(); });
  }
  void work() {
    for (;;) {
      int i;
      const std::function<void(int)>* f;
      {
        std::lock_guard<std::mutex> lk(m_);
        if (job_ == nullptr || next_ >= n_) return;
        i = next_++;
        f = job_;
      }
      (*f)(i);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) cv_done_.notify_all();
      }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_start_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
      }
      work();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_, busy_;
  std::condition_variable cv_start_, cv_done_;
  const std::function<void(int)>* job_ = nullptr;
  int n_ = 0, next_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

// f(begin, end, thread_id) over [0, n) in contiguous chunks
template <class F>
void parallel_chunks(int n, int min_per_thread, F&& f) {
  int nt = std::min(worker_count(), std::max(1, n / std::max(1, min_per_thread)));
  if (nt <= 1) {
    f(0, n, 0);
    return;
  }
  std::atomic<bool> failed{false};
  std::string msg;
  std::mutex msg_m;
  auto body = [&](int t) {
    const int a = (int)((int64_t)n * t / nt), b = (int)((int64_t)n * (t + 1) / nt);
    try {
      f(a, b, t);
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> lk(msg_m);
      if (!failed.exchange(true)) msg = e.what();
    }
  };
  const std::function<void(int)> job = body;
  if (!WorkerPool::instance().run(nt, job)) {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { body(t); });
    for (auto& x : th) x.join();
  }
  if (failed) throw Error(FW_ERR_INTERNAL, msg);
}
}  // namespace

void lu_symbolic_build(LUSymbolic& S, int n, int nb, int ne, const int* edofs, const double* cx, const double* cy,
                       const uint8_t* active, int leaf_elements) {
  if (leaf_elements <= 0) leaf_elements = 16;  // measured at 500k triangles: 8: 6.0, 16: 5.75, 32: 6.4 ms per solve
  S = LUSymbolic();
  S.n = n;
  const bool trace = getenv("FW_TRACE") != nullptr;
  double t_last = Trace::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    double t = Trace::now();
    fprintf(stderr, "[fw-trace]   symbolic/%-20s %9.3f ms\n", what, t - t_last);
    t_last = t;
  };
  // ---- 1. kd-tree over elements, breadth first (the nodes of a level own disjoint ranges of eperm)
  std::vector<int> eperm(ne);
  std::iota(eperm.begin(), eperm.end(), 0);
  std::vector<int> node_lo, node_hi;
  auto new_node = [&](int lo, int hi, int parent, int depth) {
    int id = (int)S.parent.size();
    S.parent.push_back(parent);
    S.child0.push_back(-1);
    S.child1.push_back(-1);
    S.depth.push_back(depth);
    node_lo.push_back(lo);
    node_hi.push_back(hi);
    return id;
  };
  {
    std::vector<int> level{new_node(0, ne, -1, 0)};
    std::vector<int> mids;
    while (!level.empty()) {
      mids.assign(level.size(), -1);
      parallel_chunks((int)level.size(), 1, [&](int a, int b, int) {
        for (int q = a; q < b; ++q) {
          const int v = level[q], lo = node_lo[v], hi = node_hi[v];
          if (hi - lo <= leaf_elements) continue;
          double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
          for (int i = lo; i < hi; ++i) {
            int e = eperm[i];
            x0 = std::min(x0, cx[e]), x1 = std::max(x1, cx[e]);
            y0 = std::min(y0, cy[e]), y1 = std::max(y1, cy[e]);
          }
          const double* c = (x1 - x0 >= y1 - y0) ? cx : cy;
          int mid = (lo + hi) / 2;
          std::nth_element(eperm.begin() + lo, eperm.begin() + mid, eperm.begin() + hi,
                           [c](int p, int r) { return c[p] < c[r] || (c[p] == c[r] && p < r); });
          mids[q] = mid;
        }
      });
      std::vector<int> next;
      for (size_t q = 0; q < level.size(); ++q) {
        if (mids[q] < 0) continue;
        const int v = level[q];
        const int lo = node_lo[v], hi = node_hi[v], d = S.depth[v];
        const int c0 = new_node(lo, mids[q], v, d + 1);
        const int c1 = new_node(mids[q], hi, v, d + 1);
        S.child0[v] = c0;
        S.child1[v] = c1;
        next.push_back(c0);
        next.push_back(c1);
      }
      level.swap(next);
    }
  }
  const int nn = (int)S.parent.size();
  S.n_nodes = nn;
  std::vector<int> leaf_of(ne, -1);
  for (int v = 0; v < nn; ++v)
    if (S.child0[v] < 0)
      for (int i = node_lo[v]; i < node_hi[v]; ++i) leaf_of[eperm[i]] = v;
  mark("kd-tree");

  // ---- 2. owner of every dof = LCA of the leaves of its elements.  Threads own contiguous dof ranges: every
  // thread streams the whole connectivity and folds only its own dofs (no atomics, deterministic).
  std::vector<int> owner(n, -1);
  const std::vector<int>& par = S.parent;
  const std::vector<int>& dep = S.depth;
  parallel_chunks(n, 1 << 18, [&](int da, int db, int) {
    for (int l = 0; l < nb; ++l) {
      const int* row = edofs + (size_t)l * ne;
      for (int e = 0; e < ne; ++e) {
        const int d = row[e];
        if (d < da || d >= db || !active[d]) continue;
        const int lf = leaf_of[e];
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
  });
  mark("owners (LCA)");

  // ---- 3. postorder numbering (dofs of a node in ascending original order)
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
  {
    // per-thread histograms over contiguous dof chunks -> deterministic positions without a serial sweep
    const int min_per = 1 << 16;
    const int nt = std::min(worker_count(), std::max(1, n / min_per));
    std::vector<std::vector<int>> cnt(nt, std::vector<int>(nn, 0));
    parallel_chunks(n, min_per, [&](int da, int db, int tid) {
      std::vector<int>& c = cnt[tid];
      for (int d = da; d < db; ++d)
        if (owner[d] >= 0) c[owner[d]]++;
    });
    S.ns.assign(nn, 0);
    int n_active = 0;
    for (int v = 0; v < nn; ++v) {
      int run = 0;
      for (int t = 0; t < nt; ++t) {
        const int c = cnt[t][v];
        cnt[t][v] = run;  // offset of thread t inside node v
        run += c;
      }
      S.ns[v] = run;
      n_active += run;
    }
    S.n_active = n_active;
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
    parallel_chunks(n, min_per, [&](int da, int db, int tid) {
      std::vector<int>& c = cnt[tid];
      for (int d = da; d < db; ++d) {
        const int o = owner[d];
        if (o < 0) continue;
        const int ni = S.own_start[o] + c[o]++;
        S.perm[ni] = d;
        S.iperm[d] = ni;
        S.node_of[ni] = o;
      }
    });
  }
  mark("numbering");

  // ---- levels (by depth)
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

  // ---- 4. boundary lists, deepest level first (children before parents), nodes of a level in parallel
  S.m.assign(nn, 0);
  S.blist_off.assign(nn, 0);
  std::vector<std::vector<int>> bl(nn);
  for (int lev = S.n_levels - 1; lev >= 0; --lev) {
    const int l0 = S.level_start[lev], cnt = S.level_start[lev + 1] - l0;
    double tl0 = Trace::now();
    parallel_chunks(cnt, 8, [&](int qa, int qb, int) {
      std::vector<int> tmp;
      for (int q = qa; q < qb; ++q) {
        const int v = S.level_nodes[l0 + q];
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
        } else {
          const std::vector<int>& a = bl[S.child0[v]];
          const std::vector<int>& b = bl[S.child1[v]];
          tmp.resize(a.size() + b.size());
          auto end = std::set_union(a.begin(), a.end(), b.begin(), b.end(), tmp.begin());
          tmp.resize(end - tmp.begin());
        }
        out.reserve(tmp.size());
        for (int x : tmp) {
          if (x >= os && x < os + nsv)
            own_seen++;
          else if (x >= os + nsv)
            out.push_back(x);
          else
            throw Error(FW_ERR_INTERNAL, "lu symbolic: front holds a descendant index");
        }
        if (own_seen != nsv) throw Error(FW_ERR_INTERNAL, "lu symbolic: separator dofs missing from the front");
        S.m[v] = nsv + (int)out.size();
      }
    });
    if (trace && getenv("FW_TRACE_LEVELS")) fprintf(stderr, "[fw-trace]     blist level %d (%d nodes) %.3f ms\n", lev, cnt, Trace::now() - tl0);
  }
  mark("boundary lists");
  {
    int64_t run = 0;
    for (int v = 0; v < nn; ++v) {
      S.blist_off[v] = run;
      run += (int64_t)bl[v].size();
    }
    S.blist.resize((size_t)run);
    S.rel.assign((size_t)run, -1);
    parallel_chunks(nn, 64, [&](int va, int vb, int) {
      for (int v = va; v < vb; ++v) {
        std::copy(bl[v].begin(), bl[v].end(), S.blist.begin() + S.blist_off[v]);
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
    });
  }
  mark("extend-add maps");

  // ---- 5. arena offsets (deepest level first in memory)
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
      const double m = S.m[v], ns = S.ns[v], nbd = m - ns;
      S.solve_entries += (int64_t)S.m[v] * S.m[v] - (int64_t)(S.m[v] - S.ns[v]) * (S.m[v] - S.ns[v]);
      // sum_{j<ns} [ (m-j-1) divisions + 2 (m-j-1)^2 ] = 2/3 (m^3 - nbd^3) - ... ; written as the exact sum of squares
      auto sq = [](double x) { return x * (x + 1) * (2 * x + 1) / 6; };  // sum_{i<=x} i^2
      S.factor_flops += 2.0 * (sq(m - 1) - sq(nbd - 1 < 0 ? 0 : nbd - 1)) + 0.5 * (m - 1 + nbd) * ns;
    }
  S.factor_entries = fo;
  S.work_entries = wo;
  mark("levels/offsets");
}

}  // namespace fw
