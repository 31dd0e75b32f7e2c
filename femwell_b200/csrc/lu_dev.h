// Device-side mirrors of the multifrontal analysis shared by lu_numeric.cu (factorisation) and
// lu_solve.cu (triangular solves).
#pragma once
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
  DevBuf<double> factors, work, xperm, scratch;
  DevBuf<int> counters;
  DevBuf<uint8_t> active;
  FrontTab tab() const {
    return FrontTab{m.p, ns.p, own_start.p, parent.p, child0.p, child1.p, front_off.p, work_off.p, blist_off.p,
                    blist.p, rel.p};
  }
};

// large-front GEMM (lu_gemm.cu)
template <class S>
void lu_gemm_big_launch(FrontTab T, const int* nodes, int nfront, int kstep, S* F, int max_rem, bool schur,
                        cudaStream_t s);

}  // namespace fw
