// Device-side mirrors of the multifrontal analysis shared by lu_numeric.cu (factorisation) and
// lu_solve.cu (triangular solves).
#pragma once
#include "lu.h"

namespace fw {

constexpr int PW = 32;    // panel width
constexpr int HWIN = 256; // pivot window height (rows eligible as pivots per panel)
// tree levels whose largest pivot block exceeds CL_MIN_NS run their substitution chains on thread-block
// clusters (lu_solve.cu); their fronts also get explicit inverses of the PW x PW diagonal blocks (dinv)
constexpr int CL_MIN_NS = 384;

struct FrontTab {
  const int *m, *ns, *own_start, *parent, *child0, *child1;
  const long long *front_off, *work_off, *blist_off;
  const int *blist, *rel;
  const long long* dinv_off;  // per front: offset of its inverted diagonal blocks in LUDevice::dinv, or -1
  int pf = 0;                 // substitution kernels: L2 prefetch budget per small front in KB (0 = off)
  const long long* tinv_off = nullptr;  // per front: offset of inv(L11), inv(U11) (2 ns^2 entries) in LUDevice::tinv, or -1
};

struct LUDevice {
  DevBuf<int> iperm, perm, node_of;
  DevBuf<int> parent, child0, child1, m, ns, own_start;
  DevBuf<long long> front_off, work_off, blist_off;
  DevBuf<int> blist, rel, level_nodes;
  DevBuf<int> piv, rowperm;
  DevBuf<double>&factors, &work, &xperm, &scratch;  // Handle::lu_arena
  // inverses of the PW x PW diagonal blocks of L11 (unit lower) and U11 of the cluster-level fronts: per front
  // npan column-major blocks of inv(L_kk) followed by npan blocks of inv(U_kk); ragged blocks identity-padded
  DevBuf<long long> dinv_off;
  DevBuf<double>& dinv;
  int64_t dinv_entries = 0;
  bool dinv_valid = false;
  // selective inversion (lu_tinv.cu): explicit inverses of the whole pivot blocks L11 (unit lower) and U11 of the fronts
  // of the upper levels, so that their substitution "chains" become dense triangular matrix-vector products
  DevBuf<long long> tinv_off;
  DevBuf<double>&tinv, &tinv_tmp;
  int64_t tinv_entries = 0;
  int tinv_min_level_ns = 0;  // levels whose largest pivot block reaches this size carry the inverses
  bool tinv_valid = false;
  DevBuf<int> counters;
  DevBuf<uint8_t> active;
  // dataflow substitution: nodes in deepest-level-first order; tree_sync = [epoch, 4 tickets, pad, one flag per node]
  DevBuf<int> order_fwd, tree_sync;
  explicit LUDevice(Handle::LUArena& a)
      : factors(a.factors), work(a.work), xperm(a.xperm), scratch(a.scratch), dinv(a.dinv), tinv(a.tinv),
        tinv_tmp(a.tinv_tmp) {}
  FrontTab tab() const {
    FrontTab t{m.p, ns.p, own_start.p, parent.p, child0.p, child1.p, front_off.p, work_off.p, blist_off.p,
               blist.p, rel.p, dinv_off.p};
    t.tinv_off = tinv_off.p;
    return t;
  }
};

// selective inversion (lu_tinv.cu)
constexpr int TINV_DEFAULT_MIN_NS = 200;  // re-tuned after the occupancy hints of the chain / update kernels (110 before)
bool tinv_level_static(const LUSymbolic& S, int min_ns, int lev);
bool tinv_level(const LUSymbolic& S, const LUDevice& d, int lev);
void tinv_build(Handle& h, LUDevice& d, const LUSymbolic& S);  // after the numeric factorisation (f64 only)
template <int NRHS>
void tinv_forward_level(Handle& h, LUDevice& d, const LUSymbolic& S, int lev, double* work, double* xp, int na);
template <int NRHS>
void tinv_backward_level(Handle& h, LUDevice& d, const LUSymbolic& S, int lev, double* work, double* xp, int na);

// fused partial factorisation of small fronts (lu_front_small.cu)
bool lu_front_small_applies(int max_ns, int max_m);
template <class S>
void lu_front_small_launch(FrontTab T, const int* nodes, int nfront, S* F, int* piv, double thr, int* counters,
                           cudaStream_t s);

// large-front GEMM (lu_gemm.cu)
template <class S>
void lu_gemm_big_launch(FrontTab T, const int* nodes, int nfront, int kstep, S* F, int max_rem, bool schur,
                        cudaStream_t s, int kw = PW);

}  // namespace fw
