// Generic-CSR entry of the eigen-solver: the scikit-fem EigenSolver protocol solver(K, M) -> (lams, xs)
// (reference femwell/solver.py:87-142, solver_eigen_slepc reads K.indptr/K.indices/K.data exactly like this).
// No mesh is available here, so the nested-dissection analysis runs on a clique cover of the matrix graph
// (each row with its lower neighbours, cut into cliques of <= 16 unknowns) and on coordinates that are
// either supplied by the caller or derived from two breadth-first sweeps (pseudo-peripheral distances).
#include <algorithm>
#include <numeric>
#include <queue>

#include "lu.h"
#include "solver.h"

namespace fw {

namespace {
struct Row {
  std::vector<int> idx;
  std::vector<double> val;  // 1 or 2 doubles per entry
};

void bfs(int n, const std::vector<int>& ap, const std::vector<int>& ai, int src, std::vector<int>& dist) {
  dist.assign(n, -1);
  std::vector<int> q;
  q.reserve(n);
  q.push_back(src);
  dist[src] = 0;
  for (size_t h = 0; h < q.size(); ++h) {
    int v = q[h];
    for (int e = ap[v]; e < ap[v + 1]; ++e) {
      int u = ai[e];
      if (dist[u] < 0) {
        dist[u] = dist[v] + 1;
        q.push_back(u);
      }
    }
  }
  int mx = 0;
  for (int v = 0; v < n; ++v) mx = std::max(mx, dist[v]);
  for (int v = 0; v < n; ++v)
    if (dist[v] < 0) dist[v] = mx + 1;  // other connected components
}
}  // namespace

void setup_generic_csr(Handle& h, int n, const int* Kp, const int* Ki, const void* Kx, const int* Mp, const int* Mi,
                       const void* Mx, bool is_complex, const double* coords) {
  FW_REQUIRE(n > 0 && Kp && Ki && Kx && Mp && Mi && Mx, "bad CSR arguments");
  h.lu_prefetch.reset();  // joins a pending analysis of the previous space
  h.Bc.valid = h.MassC.valid = false;
  const int w = is_complex ? 2 : 1;
  cudaStream_t s = h.stream;
  // ---- union pattern, A = -K, B = -M on it
  std::vector<int> up(n + 1, 0);
  std::vector<int> ui;
  std::vector<double> av, bv;
  ui.reserve((size_t)Kp[n] + Mp[n] / 4);
  av.reserve(((size_t)Kp[n] + Mp[n] / 4) * w);
  bv.reserve(((size_t)Kp[n] + Mp[n] / 4) * w);
  std::vector<std::pair<int, int>> ka, ma;  // (col, source position), sorted by col
  const double* kx = (const double*)Kx;
  const double* mx = (const double*)Mx;
  for (int r = 0; r < n; ++r) {
    ka.clear();
    ma.clear();
    for (int e = Kp[r]; e < Kp[r + 1]; ++e) {
      FW_REQUIRE(Ki[e] >= 0 && Ki[e] < n, "K column index out of range");
      ka.push_back({Ki[e], e});
    }
    for (int e = Mp[r]; e < Mp[r + 1]; ++e) {
      FW_REQUIRE(Mi[e] >= 0 && Mi[e] < n, "M column index out of range");
      ma.push_back({Mi[e], e});
    }
    if (!std::is_sorted(ka.begin(), ka.end())) std::sort(ka.begin(), ka.end());
    if (!std::is_sorted(ma.begin(), ma.end())) std::sort(ma.begin(), ma.end());
    size_t a = 0, b = 0;
    while (a < ka.size() || b < ma.size()) {
      int c;
      if (b >= ma.size() || (a < ka.size() && ka[a].first <= ma[b].first))
        c = ka[a].first;
      else
        c = ma[b].first;
      double ar = 0, ai_ = 0, br = 0, bi = 0;
      while (a < ka.size() && ka[a].first == c) {  // duplicates are summed
        ar += kx[(size_t)ka[a].second * w];
        if (is_complex) ai_ += kx[(size_t)ka[a].second * w + 1];
        ++a;
      }
      while (b < ma.size() && ma[b].first == c) {
        br += mx[(size_t)ma[b].second * w];
        if (is_complex) bi += mx[(size_t)ma[b].second * w + 1];
        ++b;
      }
      ui.push_back(c);
      av.push_back(-ar);
      bv.push_back(-br);
      if (is_complex) {
        av.push_back(-ai_);
        bv.push_back(-bi);
      }
    }
    up[r + 1] = (int)ui.size();
  }
  const int64_t nnz = (int64_t)ui.size();
  Pattern& pat = h.pat;
  pat.n = n;
  pat.nnz = nnz;
  pat.n_coo = 0;
  pat.indptr.upload(up.data(), up.size(), s);
  pat.indices.upload(ui.data(), ui.size(), s);
  for (Values* v : {&h.A, &h.B}) {
    v->is_complex = is_complex;
    v->nnz_kept = -1;
    v->valid = true;
  }
  h.A.v.upload(av.data(), av.size(), s);
  h.B.v.upload(bv.data(), bv.size(), s);
  h.has_symbolic = true;
  h.sp.valid = false;
  h.sp.n = n;
  h.Mass.valid = h.C0.valid = h.C1.valid = false;
  h.lu.reset();

  // ---- symmetrised adjacency (for the clique cover and the pseudo coordinates)
  std::vector<int> deg(n + 1, 0);
  for (int r = 0; r < n; ++r)
    for (int e = up[r]; e < up[r + 1]; ++e) {
      int c = ui[e];
      if (c == r) continue;
      deg[r + 1]++;
      deg[c + 1]++;
    }
  std::vector<int> ap(n + 1, 0);
  for (int r = 0; r < n; ++r) ap[r + 1] = ap[r] + deg[r + 1];
  std::vector<int> ai(ap[n]);
  {
    std::vector<int> fill(ap.begin(), ap.end() - 1);
    for (int r = 0; r < n; ++r)
      for (int e = up[r]; e < up[r + 1]; ++e) {
        int c = ui[e];
        if (c == r) continue;
        ai[fill[r]++] = c;
        ai[fill[c]++] = r;
      }
    // sort + unique per vertex (pairs present in both triangles appear twice)
    std::vector<int> np(n + 1, 0);
    size_t outp = 0;
    for (int r = 0; r < n; ++r) {
      std::sort(ai.begin() + ap[r], ai.begin() + ap[r + 1]);
      int last = -1;
      size_t start = outp;
      for (int e = ap[r]; e < ap[r + 1]; ++e)
        if (ai[e] != last) {
          last = ai[e];
          ai[outp++] = last;
        }
      np[r] = (int)start;
    }
    np[n] = (int)outp;
    ap = np;
    ai.resize(outp);
  }
  // ---- coordinates
  std::vector<double> px(n), py(n);
  if (coords) {
    for (int v = 0; v < n; ++v) px[v] = coords[v], py[v] = coords[(size_t)n + v];
  } else {
    std::vector<int> d0, da, db;
    bfs(n, ap, ai, 0, d0);
    int a = (int)(std::max_element(d0.begin(), d0.end()) - d0.begin());
    bfs(n, ap, ai, a, da);
    int b = (int)(std::max_element(da.begin(), da.end()) - da.begin());
    bfs(n, ap, ai, b, db);
    for (int v = 0; v < n; ++v) px[v] = da[v], py[v] = db[v];
  }
  // ---- clique cover: row r with its lower neighbours, in chunks of 15 (+ r itself)
  Connectivity& cn = h.generic;
  cn = Connectivity();
  cn.n = n;
  cn.nb = 16;
  int64_t ne = 0;
  for (int r = 0; r < n; ++r) {
    int low = (int)(std::lower_bound(ai.begin() + ap[r], ai.begin() + ap[r + 1], r) - (ai.begin() + ap[r]));
    ne += std::max(1, (low + 14) / 15);
  }
  FW_REQUIRE(ne < (int64_t)0x7fffffff, "too many cliques");
  cn.ne = (int)ne;
  cn.edofs.assign((size_t)cn.nb * cn.ne, 0);
  cn.cx.resize(cn.ne);
  cn.cy.resize(cn.ne);
  int e = 0;
  for (int r = 0; r < n; ++r) {
    int low = (int)(std::lower_bound(ai.begin() + ap[r], ai.begin() + ap[r + 1], r) - (ai.begin() + ap[r]));
    int done = 0;
    do {
      int take = std::min(15, low - done);
      cn.edofs[(size_t)0 * cn.ne + e] = r;
      for (int q = 0; q < 15; ++q)
        cn.edofs[(size_t)(q + 1) * cn.ne + e] = q < take ? ai[ap[r] + done + q] : r;  // padded with r
      cn.cx[e] = px[r];
      cn.cy[e] = py[r];
      ++e;
      done += take;
    } while (done < low);
  }
  cn.valid = true;
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
}

}  // namespace fw
