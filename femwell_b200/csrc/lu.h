// Sparse multifrontal LU of the shifted pencil (K - sigma M) -- stands in for the cuDSS call
// named by the north star (cuDSS is not present in this image).  Replaces
// scipy's splu(OP) + OP.solve inside eigs mode 3 (scipy arpack.py:1006-1019, 1179-1181).
//
// Analysis (host, lu_symbolic.cpp): element-based geometric nested dissection (kd-tree over
// element centroids; a dof belongs to the lowest tree node that contains all its elements),
// postorder numbering, front index sets, extend-add maps.
// Numeric (device, lu_numeric.cu): level-by-level batched dense partial LU of the fronts with
// windowed partial pivoting, hand-written FP64/C128 TRSM + GEMM kernels, and level-batched
// forward/backward substitution.
#pragma once
#include <thread>

#include "handle.h"

namespace fw {

struct LUSymbolic {
  int n = 0;         // total dofs
  int n_active = 0;  // dofs that take part in the factorisation (not Dirichlet)
  int n_nodes = 0, n_levels = 0;
  int64_t factor_entries = 0;  // sum m^2
  int64_t work_entries = 0;    // sum m
  int64_t solve_entries = 0;   // sum m^2 - (m - ns)^2: entries of L and U read by one forward + backward substitution
  double factor_flops = 0;     // sum over fronts of the partial-LU flops (ns pivots of an m x m front)
  int max_m = 0, max_ns = 0;
  std::vector<int> perm;      // [n_active] new -> old dof
  std::vector<int> iperm;     // [n] old -> new (or -1)
  std::vector<int> node_of;   // [n_active] owning node of a new index
  // per node
  std::vector<int> parent, child0, child1, depth, m, ns, own_start;
  std::vector<int64_t> front_off;  // offset of the m x m front in the factor arena
  std::vector<int64_t> work_off;   // offset of the length-m work vector
  std::vector<int64_t> blist_off;  // offset into blist/rel (length m - ns)
  std::vector<int> blist;          // boundary indices (new numbering, ascending)
  std::vector<int> rel;            // position of each boundary index in the parent's front
  // levels (by depth, deepest first when iterated backwards)
  std::vector<int> level_start;    // [n_levels + 1]
  std::vector<int> level_nodes;    // nodes sorted by depth
  std::vector<int> level_max_ns, level_max_m, level_max_nb;
};

// centroids: [2, ne]; edofs: [nb, ne]; active: [n] (1 = in the system)
void lu_symbolic_build(LUSymbolic& S, int n, int nb, int ne, const int* edofs, const double* cx, const double* cy,
                       const uint8_t* active, int leaf_elements);

struct LUDevice;  // device-side mirrors (lu_numeric.cu)
struct LU;
// the same analysis on the device (lu_symbolic_gpu.cu): fills lu.sym (small host arrays) and lu.dev (device tables)
void lu_symbolic_build_gpu(Handle& h, LU& lu, const std::vector<uint8_t>& active, int leaf_elements);
void lu_debug_analysis_download(Handle& h, int which, void* out);

// One forward + backward substitution is ~60 dependent kernel launches with fixed arguments for a fixed (rhs, solution)
// pointer pair; the Arnoldi loop repeats it ~50 times per factorisation.  The launch sequence is captured into a CUDA
// graph the second time a pointer pair is seen and replayed afterwards (invalidated by every new factorisation).
struct SolveGraph {
  const void* b = nullptr;
  void* x = nullptr;
  int nrhs = 0;
  int uses = 0;
  int kernels = 0;  // kernel nodes of the captured sequence
  cudaGraphExec_t exec = nullptr;
};

struct LU {
  LUSymbolic sym;
  bool is_complex = false;
  bool factored = false;
  std::vector<SolveGraph> graphs;
  void drop_graphs() {
    for (auto& g : graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }
  int perturbed = 0;
  // set by the eigen-solver when its residual check failed: the substitutions fall back from the inverse pivot blocks of
  // the upper levels to substitution chains (backward stable) for the retry; cleared by the next factorisation
  bool no_tinv = false;
  std::vector<uint8_t> active_key;  // the analysis is reusable while mesh, Dirichlet set and leaf size stay
  int leaf_key = 0;
  std::unique_ptr<LUDevice> dev;
  LU();
  ~LU();
};

// host analysis started ahead of time on a worker thread (fw_compute_modes: it overlaps the COO sort, the
// element kernels and the reduce on the GPU); lu_analyse joins it
struct LUPrefetch {
  std::unique_ptr<LU> lu;
  std::thread th;
  std::string err;
  ~LUPrefetch() {
    if (th.joinable()) th.join();
  }
};
void lu_analyse_prefetch(Handle& h, const std::vector<uint8_t>& active, int leaf_elements);

// factor OP = alphaA * A + alphaB * B (values on the structural pattern h.pat)
void lu_analyse(Handle& h, const std::vector<uint8_t>& active, int leaf_elements);
void lu_factor(Handle& h, double alphaA_re, double alphaA_im, double alphaB_re, double alphaB_im, bool complex_arith);
// the same for any two value arrays on the structural pattern (the H-field fallback factors the mass matrix)
void lu_factor_values(Handle& h, const Values& VA, const Values& VB, double alphaA_re, double alphaA_im,
                      double alphaB_re, double alphaB_im, bool complex_arith);
// solve OP x = b for nrhs right-hand sides given in ORIGINAL dof order on the device
// (b, x: [nrhs][n] of the LU scalar type; x[inactive] = 0).  b and x may alias.
void lu_solve_dev(Handle& h, const void* b, void* x, int nrhs);
void lu_debug_download(Handle& h, int which, void* out);

}  // namespace fw
