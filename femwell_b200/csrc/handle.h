// The library handle: one device, one stream, all resident state of the waveguide path.
#pragma once
#include "common.h"

namespace fw {

struct LU;          // sparse multifrontal factorisation (lu.h)
struct LUPrefetch;  // host analysis running on a worker thread (lu.h)

// structural CSR pattern + the sorted-COO bookkeeping that produced it
struct Pattern {
  int n = 0;              // rows
  int64_t n_coo = 0;      // COO entries that were sorted
  int64_t nnz = 0;        // distinct (row, col)
  DevBuf<int> indptr;     // [n+1]
  DevBuf<int> indices;    // [nnz]
  DevBuf<uint32_t> perm;  // [n_coo]  source COO index of sorted position
  DevBuf<uint32_t> seg;   // [nnz+1]  first sorted position of every CSR slot
  DevBuf<int> tt_slots;   // slots of the Nedelec x Nedelec sub-pattern (built on first use)
  int64_t n_tt = -1;
};

// values on a Pattern (+ "any non-zero contribution" flags for eliminate_zeros parity)
struct Values {
  bool is_complex = false;
  bool valid = false;
  DevBuf<double> v;       // nnz (f64) or 2*nnz (c128)
  DevBuf<uint8_t> flag;   // nnz
  int64_t nnz_kept = -1;  // after eliminate_zeros
};

struct Space {
  int order = 0, nv = 0, ne = 0, n = 0, nb = 0, nq = 0, nt = 0, nz = 0;
  int kind[MAXNB];
  std::vector<double> vec, sc, X, W;
  std::vector<double> h_p;
  std::vector<int> h_t, h_edofs;
  std::vector<uint8_t> h_is_z;  // [n] 1 for Lagrange (E_z) dofs (host copy of is_z)
  DevBuf<double> p;
  DevBuf<int> t, edofs;
  DevBuf<uint8_t> is_z;  // [n] 1 for Lagrange (E_z) dofs
  DevBuf<uint8_t> emask; // [ne] 1 for the elements of the basis (CellBasis.tind); empty: all elements
  bool has_emask = false;
  bool valid = false;
};

// Clique cover of the matrix graph used by the nested-dissection analysis when no FE mesh is available
// (generic CSR eigen-solver, fw_eigs_csr): every coupled pair of unknowns shares at least one clique.
struct Connectivity {
  int n = 0, nb = 0, ne = 0;
  std::vector<int> edofs;      // [nb, ne]
  std::vector<double> cx, cy;  // clique "centroids" (real or pseudo coordinates)
  bool valid = false;
};

struct Handle {
  DevPool pool;  // first member: destroyed last, after every buffer has gone back to it
  int device = 0;
  cudaStream_t stream = nullptr;
  Space sp;
  Pattern pat;
  bool has_symbolic = false;
  bool eliminate_zeros = true;  // export semantics of scipy's coo.eliminate_zeros()
  int mass_solver = 0;          // FW_OPT_MASS_SOLVER: 0 PCG with direct fallback, 1 PCG only, 2 direct (sparse LU)
  double pcg_tol = 1e-10;       // relative residual of the H-field mass solves (FW_OPT_PCG_TOL_EXP)
  bool smooth_v0 = true;        // FW_OPT_START_VECTOR: contrast-weighted smooth default start vector (start_vector.cu)
  DevBuf<double> v0_dev;        // start vector prepared on the device for the next eigs call (fw_compute_modes)
  bool v0_dev_valid = false;
  Values A, B, Mass, C0, C1;
  // mass matrix without the structurally-zero cross blocks (own compact CSR; used by the PCG of the H field)
  struct CompactCsr {
    DevBuf<int> indptr, indices;
    DevBuf<double> vals;
    DevBuf<int> pair;  // order 2: partner of every Nedelec dof (same edge / same triangle), -1 otherwise
    int64_t nnz = 0;
    bool valid = false;
  } MassC, Bc;  // Bc: bform without the zero slots of the shared pattern (M x of the Arnoldi operator)
  // the mass matrix is block diagonal (Nedelec | Lagrange): the two blocks are solved as separate PCG systems -- the
  // Lagrange block converges in ~1/3 of the iterations of the Nedelec block
  struct MassBlock {
    DevBuf<int> idx, indptr, indices, pair;  // idx: block dof -> dof of the space
    DevBuf<double> vals;
    int n = 0;
    int64_t nnz = 0;
  } MassN, MassP;
  bool mass_split = false;
  Connectivity generic;  // set by fw_eigs_csr instead of a space
  DevBuf<double> blocks;  // scratch: local element blocks (SoA [entry][Ne])
  // generic COO->CSR result
  Pattern coo_pat;
  Values coo_val;
  // compacted export scratch
  DevBuf<uint32_t> scan_tmp, scan_tmp2;
  std::unique_ptr<LU> lu;
  std::unique_ptr<LUPrefetch> lu_prefetch;
  // fw_analysis_hint: Dirichlet set / leaf size of the solve that will follow the next fw_set_space (one shot)
  std::vector<int> hint_dirichlet;
  int hint_leaf = 0;
  bool hint_valid = false;
  // big LU buffers live here so that they survive a new space / analysis (no 25 GB cudaFree + cudaMalloc per solve)
  struct LUArena {
    DevBuf<double> factors, work, xperm, scratch, dinv, tinv, tinv_tmp;
  } lu_arena;
  // pinned staging ring for large device -> pageable-host copies (results E, H)
  void* pinned[2] = {nullptr, nullptr};
  cudaEvent_t pinned_ev[2] = {nullptr, nullptr};
  double lu_sigma_re = 0, lu_sigma_im = 0;
  fw_timings tim{};
  Handle();
  ~Handle();
};

// device -> pageable host through the pinned ring, chunk copies overlapped with a threaded host memcpy (api.cu)
void download_staged(Handle& h, void* host_dst, const void* dev_src, size_t bytes);
void upload_staged(Handle& h, void* dev_dst, const void* host_src, size_t bytes);

// ---- scan / sort / COO->CSR (sort.cu)
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, DevBuf<uint32_t>& tmp, cudaStream_t s);
// stable LSD radix sort of (key, payload) pairs by the low `bits` bits; the result ends up in keys_a / pay_a
void radix_sort_pairs_dev(Handle& h, DevBuf<uint64_t>& keys_a, DevBuf<uint64_t>& keys_b, DevBuf<uint32_t>& pay_a,
                          DevBuf<uint32_t>& pay_b, int64_t n, int bits, bool payload_iota);
void build_pattern_from_keys(Pattern& pat, DevBuf<uint64_t>& keys, DevBuf<uint32_t>& payload, int64_t n_coo,
                             int n_rows, int col_bits, Handle& h);
void symbolic_space(Handle& h);
template <class S>
void reduce_onto_pattern(const Pattern& pat, const S* coo_vals, const int* lut, int lut_n, int64_t stride,
                         S* out, uint8_t* flag, cudaStream_t s, const int* slots = nullptr, int64_t nslots = 0);
void build_slot_list(Handle& h, Pattern& pat, const int* lut_dev, int64_t stride);
int64_t count_kept(Handle& h, const Pattern& pat, Values& val);
void export_compacted(Handle& h, const Pattern& pat, Values& val, int* indptr, int* indices, void* data);
void coo_to_csr_dev(Handle& h, int n_rows, int64_t n_entries, DevBuf<int>& dr, DevBuf<int>& dc, DevBuf<double>& dv,
                    bool is_complex);
// device-side compaction: keeps the slots whose flag has a bit of `mask` set; returns the number kept
int64_t compact_csr_dev(Handle& h, const Pattern& pat, Values& val, uint8_t mask, DevBuf<int>& oip, DevBuf<int>& oind,
                        DevBuf<double>& oval);
void coo_to_csr(Handle& h, int n_rows, int64_t n_entries, const int* rows, const int* cols, const void* data,
                bool is_complex);

// ---- element kernels (assemble.cu)
void upload_tables(Handle& h);
void assemble_waveguide(Handle& h, int eps_kind, bool eps_complex, const void* eps_host, double k0, double mu_r,
                        double radius);
void assemble_aux(Handle& h);  // Mass, C0, C1 (geometry only, f64)
// start_vector.cu: smooth default start vector of the Arnoldi iteration; false if none applies
bool smooth_start_vector(Handle& h, int eps_kind, bool eps_complex, const void* eps_host, DevBuf<double>& v0);

// ---- CSR kernels (csr.cu)
template <class MT, class VT>
void spmv(int n, const int* indptr, const int* indices, const MT* vals, const VT* x, VT* y, cudaStream_t s);

}  // namespace fw
