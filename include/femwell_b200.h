/*
 * femwell_b200 -- C ABI of the B200-native waveguide eigenmode hot path.
 *
 * Drop-in boundary for femwell.maxwell.waveguide.compute_modes and the Mode/Modes
 * post-processing (reference: /root/reference/femwell/maxwell/waveguide.py:528-677,
 * 699-736, 112-258) and for the EigenSolver protocol of femwell/solver.py:87-142.
 * The reference is pure Python over scikit-fem + scipy; it has no FFI of its own, so
 * each entry point cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (never throws across the ABI);
 *    fw_last_error() returns a thread-local message.
 *  - all pointers are HOST pointers unless the name ends in _dev; the library owns its
 *    device memory (one handle <-> one device <-> one CUDA stream).
 *  - complex numbers are interleaved (re, im) float64 pairs ("c128").
 *  - integer index arrays are int32; sizes that can exceed 2^31 are int64.
 *  - matrices are CSR with sorted, unique column indices (scipy "canonical format").
 */
#ifndef FEMWELL_B200_H
#define FEMWELL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fw_handle fw_handle;

enum {
  FW_OK = 0,
  FW_ERR_INVALID = -1,   /* bad argument / call order                      */
  FW_ERR_CUDA = -2,      /* CUDA runtime error                             */
  FW_ERR_NOCONV = -3,    /* Arnoldi did not converge (ArpackNoConvergence) */
  FW_ERR_SINGULAR = -4,  /* factorisation hit an exactly singular pivot    */
  FW_ERR_INTERNAL = -5
};

/* epsilon representation: basis_epsilon_r.elem of the reference call sites
 * (waveguide_modes.py:69 P0, bent_waveguide.py:73 DG-P1, waveguide_anisotropic.py:81 Vector-P0) */
enum { FW_EPS_P0 = 0, FW_EPS_DG_P1 = 1, FW_EPS_VEC_P0 = 2 };

/* which assembled operator to export */
enum { FW_MAT_A = 0, FW_MAT_B = 1, FW_MAT_MASS = 2, FW_MAT_C0 = 3, FW_MAT_C1 = 4 };

/* functionals of waveguide.py:112-258 evaluated by fw_functional */
enum {
  FW_FUN_OVERLAP = 0,      /* 0.5*int conj(E_i)xH_j + E_j x conj(H_i)      :723-736 */
  FW_FUN_EX2_EY2 = 1,      /* out[0]=int|Ex|^2, out[1]=int|Ey|^2           :116-127 */
  FW_FUN_EX2_EY2_EALL = 2, /* + out[2]=int(|Ex|^2+|Ey|^2+|Ez|^2)           :139-159 */
  FW_FUN_COUPLING = 3,     /* int d_eps (conj(E_i).E_j incl. z)            :167-179 */
  FW_FUN_E2_E4 = 4,        /* out[0]=int e2, out[1]=int e2^2 (field sel.)  :181-213 */
  FW_FUN_CONFINEMENT = 5   /* int sqrt(eps) conj(E).E (t and z)            :225-249 */
};

const char* fw_last_error(void);
int fw_version(void);

/* ---- lifetime ------------------------------------------------------------------ */
int fw_create(fw_handle** out, int device);
int fw_destroy(fw_handle* h);
int fw_synchronize(fw_handle* h);
/* FW_OPT_ELIMINATE_ZEROS (default 1): matrix export drops local contributions that are exactly 0.0
 * like scikit-fem's COOData (coo_matrix.eliminate_zeros()); 0 exports the structural
 * pattern (every (test dof, trial dof) pair of every element).                                   */
enum {
  FW_OPT_ELIMINATE_ZEROS = 0,
  /* relative residual 10^-value of the H-field mass solves (block PCG replacing the spsolve of waveguide.py:671);
   * default 10: the reference's H carries complex64 rounding (1e-7, :658,:666)                                  */
  FW_OPT_PCG_TOL_EXP = 1,
  /* H-field mass solves: 0 (default) block PCG, falling back to the sparse LU of the mass block when the PCG stagnates
   * (elongated triangles: the Nedelec mass matrix is ill conditioned); 1 PCG only (FW_ERR_NOCONV on stagnation);
   * 2 always the direct solve (what the reference does: spsolve, waveguide.py:671)                              */
  FW_OPT_MASS_SOLVER = 2,
  /* default start vector of the Arnoldi iteration inside fw_compute_modes when fw_eigs_params.v0 is NULL: 1 (default)
   * a smooth field weighted by the index contrast (fewer operator applications; csrc/start_vector.cu), 0 the
   * uniform(-1, 1) stream that mirrors scipy's default (arpack.py:361-364)                                        */
  FW_OPT_START_VECTOR = 3
};
int fw_set_option(fw_handle* h, int32_t option, int32_t value);
/* Optional page-locked host buffers for the large results (E, H of fw_compute_modes: 2 * k * n_dofs * 16 bytes).  A
 * destination obtained here is filled by one DMA at bus speed; any other host pointer goes through the library's pinned
 * staging ring plus a threaded memcpy (about twice as slow).  NULL on failure (no CUDA device, out of lockable memory). */
void* fw_host_alloc(size_t bytes);
int fw_host_free(void* p);

/* ---- a1: mesh + FE space (waveguide.py:567-575; scikit-fem CellBasis) ------------ */
typedef struct {
  int32_t order;            /* 1: N1xP1, 2: N2xP2                                     */
  int32_t n_vertices, n_elements, n_dofs, nbfun, nqp;
  const double* p;          /* [2, Nv]   vertex coordinates                          */
  const int32_t* t;         /* [3, Ne]   sorted per column                            */
  const int32_t* element_dofs; /* [nbfun, Ne]                                       */
  const int32_t* kind;      /* [nbfun]   0 = Nedelec (E_t), 1 = Lagrange (E_z)        */
  const double* vec;        /* [nbfun, 2, nqp] reference phi_hat | grad_hat psi      */
  const double* sc;         /* [nbfun, nqp]    reference curl_hat | psi              */
  const double* X;          /* [2, nqp]  quadrature points                           */
  const double* W;          /* [nqp]     quadrature weights                          */
} fw_space;
int fw_set_space(fw_handle* h, const fw_space* s);
/* Element subset of the basis (scikit-fem CellBasis(elements=...); basis_epsilon_r.with_element keeps it,
 * waveguide.py:574): the forms are assembled over `elements` only.  Call after fw_set_space; n_sel < 0 or NULL
 * restores "all elements" (the state after fw_set_space).  Dofs no selected element touches get empty rows, as in
 * the reference (whose factorisation then fails: FW_ERR_SINGULAR here).                                     */
int fw_set_element_subset(fw_handle* h, int32_t n_sel, const int32_t* elements);
/* Optional, before fw_set_space: announces the Dirichlet set (condense, waveguide.py:610-619) and the leaf size of
 * the solve that will follow, so that the host analysis of its sparse LU starts inside fw_set_space and overlaps
 * the uploads.  One shot; a solve with other parameters simply redoes the analysis. */
int fw_analysis_hint(fw_handle* h, int32_t n_dirichlet, const int32_t* dirichlet, int32_t leaf_elements);

/* ---- a5 (structural half): sort-by-key of the (row, col) pairs of all local blocks;
 *      replaces scikit-fem COOData -> scipy coo_matrix.tocsr() index work              */
int fw_symbolic(fw_handle* h, int64_t* nnz_structural);

/* ---- a2+a3+a4: fused element kernel for aform/bform (waveguide.py:577-603) followed
 *      by the segmented reduce onto the structural pattern.  eps is [Ne] (P0) or [3*Ne];
 *      eps_is_complex selects c128 (dtype of A,B follows epsilon_r.dtype, :577)        */
int fw_assemble_waveguide(fw_handle* h, int32_t eps_kind, int32_t eps_is_complex, const void* eps,
                          double k0, double mu_r, double radius);

/* scipy-canonical CSR after eliminate_zeros.  Two-step export:
 * fw_matrix_nnz then fw_matrix_export into caller buffers (data: f64 or c128 as assembled;
 * FW_MAT_MASS/C0/C1 are always f64).                                                   */
int fw_matrix_nnz(fw_handle* h, int32_t which, int64_t* nnz, int32_t* is_complex);
int fw_matrix_export(fw_handle* h, int32_t which, int32_t* indptr, int32_t* indices, void* data);

/* ---- generic COO -> CSR (scipy coo_matrix((data,(rows,cols))).eliminate_zeros().tocsr())
 *      radix sort by key + segmented reduce; returns nnz, then export.                 */
int fw_coo_to_csr(fw_handle* h, int32_t n_rows, int64_t n_entries, const int32_t* rows,
                  const int32_t* cols, const void* data, int32_t is_complex, int64_t* nnz);
int fw_coo_to_csr_export(fw_handle* h, int32_t* indptr, int32_t* indices, void* data);

/* ---- K4: CSR SpMV y = M x on the assembled operators (x,y host, length n_dofs)       */
int fw_matrix_spmv(fw_handle* h, int32_t which, int32_t x_is_complex, const void* x, void* y);

/* Device-resident timing of K4 (measurement only): `reps` applications y = M x of an assembled operator on an
 * internal vector, CUDA events on the handle's stream; *ms_per_apply = average.  x_is_complex selects the
 * f64 or c128 vector kernel.                                                                              */
int fw_bench_spmv(fw_handle* h, int32_t which, int32_t x_is_complex, int32_t reps, double* ms_per_apply);

/* ---- a6-a9: shift-invert Arnoldi on (-A) x = lam (-B) x  (waveguide.py:605-625,
 *      scipy eigs mode 3).  dirichlet: dofs condensed out (waveguide.py:610-619).
 *      v0: optional start vector [n_dofs] f64 (NULL: internal fixed-seed vector).
 *      lams [k] c128 ; X [k, n_dofs] c128 (row i = eigenvector i), raw (not un-scaled). */
typedef struct {
  int32_t k;            /* num_modes                                              */
  int32_t ncv;          /* 0: max(2k+1, 20) like scipy                             */
  int32_t maxiter;      /* max restarts, 0: default                               */
  double tol;           /* 0: machine epsilon like ARPACK                          */
  double sigma_re, sigma_im;
  int32_t n_dirichlet;
  const int32_t* dirichlet;
  const double* v0;
  int32_t refine;       /* iterative-refinement steps per LU solve: 0 = none, retried with 1 if
                           the returned pairs miss 1e-10; >0 fixed; <0 never            */
  int32_t leaf_elements;/* nested-dissection leaf size, 0: default                  */
} fw_eigs_params;
typedef struct {
  int32_t nconv, iterations, restarts, n_op, pivots_perturbed, lu_levels, lu_fronts;
  int64_t lu_factor_entries;
  double t_symbolic_ms, t_factor_ms, t_arnoldi_ms, t_solve_ms_total, max_residual;
} fw_eigs_stats;
int fw_eigs(fw_handle* h, const fw_eigs_params* prm, double* lams, double* X, fw_eigs_stats* st);

/* ---- the scikit-fem EigenSolver protocol on caller-assembled matrices (femwell/solver.py:87-142:
 *      solver(K, M) -> (lams, xs) for K x = lam M x, shift-invert around sigma).  K, M: n x n CSR (f64 or both
 *      c128); coords: optional [2, n] coordinates of the unknowns for the nested dissection (NULL: pseudo
 *      coordinates from two breadth-first sweeps of the matrix graph).  X [k, n] c128, unit 2-norm.        */
int fw_eigs_csr(fw_handle* h, int32_t n, const int32_t* K_indptr, const int32_t* K_indices, const void* K_data,
                const int32_t* M_indptr, const int32_t* M_indices, const void* M_data, int32_t is_complex,
                const double* coords, const fw_eigs_params* prm, double* lams, double* X, fw_eigs_stats* st);

/* ---- sparse LU exposed for tests: factor (A_which*alpha + B_which*beta) and solve      */
int fw_lu_factor_shifted(fw_handle* h, double sigma_re, double sigma_im, int32_t n_dirichlet,
                         const int32_t* dirichlet, int32_t leaf_elements);
int fw_lu_solve(fw_handle* h, int32_t rhs_is_complex, const void* b, void* x, int32_t refine);

/* ---- a10-a13: E_z un-scaling, H-field, power normalisation (waveguide.py:627-639,
 *      657-677).  In: lams [k], X [k, n] raw from fw_eigs.  Out: E,H [k, n] c128 normalised,
 *      powers [k] c128 (the pre-normalisation self overlap).                           */
int fw_postprocess_modes(fw_handle* h, int32_t k, const double* lams, const double* X, double k0,
                         double omega, double mu0, double* E, double* H, double* powers);
int fw_hfield(fw_handle* h, const double* E, double beta_re, double beta_im, double omega, double mu0,
              double* H);

/* ---- a12/a15: interpolate + pointwise + reduce functionals.  fields: up to 4 c128 dof
 *      vectors [n_dofs] (unused = NULL); aux: per-element (or 3/elem) coefficient, kind as
 *      FW_EPS_*, may be NULL; elements: optional subset [n_sel] (NULL = all);
 *      out: up to 3 c128 values.                                                        */
int fw_functional(fw_handle* h, int32_t kind, const double* f0, const double* f1, const double* f2,
                  const double* f3, int32_t aux_kind, int32_t aux_is_complex, const void* aux,
                  int32_t n_sel, const int32_t* elements, int32_t option, double* out);

/* ---- a16: Mode.calculate_intensity (waveguide.py:260-280): 0.5 Re(Ex Hy* - Ey Hx*) at the quadrature points,
 *      L2-projected onto discontinuous P1.  E, H [n_dofs] c128; out [3*Ne] f64, dof = 3*e + a.          */
int fw_intensity(fw_handle* h, const double* E, const double* H, double* out);

/* Cross-mesh branch of calculate_overlap (femwell/maxwell/waveguide.py:737-759) and calculate_scalar_product
 * (:762-782).  basis_i is the current space of the handle, basis_j is passed as a second space description; its mesh may
 * differ.  The transverse parts of E_j, H_j are L2-projected onto continuous P1 on mesh j and evaluated at the
 * quadrature points of mesh i.  mode 0: 0.5 int [conj(E_i) x H_j + E_j x conj(H_i)];
 * mode 1: int conj(E_i) x H_j (H_i, E_j may be NULL).  elements: optional element subset of basis_i.
 * Fields complex128 [n_dofs of their basis]; out: complex128 [1].  A quadrature point of mesh i outside mesh j
 * is an error ("Point is outside of the mesh.").                                                          */
int fw_overlap_cross(fw_handle* h, const fw_space* space_j, const double* E_i, const double* H_i, const double* E_j,
                     const double* H_j, int32_t mode, int32_t n_sel, const int32_t* elements, double* out);

/* Mode.poynting (femwell/maxwell/waveguide.py:74-95): P = E x H at the quadrature points projected onto
 * ElementVector(ElementDG(P1 | P2), 3) (per-element mass solve).  E, H: complex128 [n_dofs];
 * out: complex128 [n_elements * 3 * nl] with nl = 3 (order 1) or 6 (order 2), dof = e*3*nl + a*3 + c. */
int fw_poynting(fw_handle* h, const double* E, const double* H, double* out);

/* calculate_energy_current_density (femwell/maxwell/waveguide.py:680-696): P0 projection of
 * |e_x|^2 + |e_y|^2 + |e_z|.  xs: complex128 [n_dofs]; out: complex128 [n_elements]. */
int fw_energy_current_density(fw_handle* h, const double* xs, double* out);

/* Mode.eval_error_estimator (femwell/maxwell/waveguide.py:473-502): eta[e] = 1/2 sum_{f in e} int_f h (|[grad E_z].n|^2 +
 * |[eps E_t].n|^2) ds over the interior facets (h = |f|), the indicator of the adaptive refinement loop of
 * docs/photonics/examples/refinement.py:58-105.  E: complex128 [n_dofs]; eps as in fw_assemble_waveguide (FW_EPS_P0 or
 * FW_EPS_DG_P1); t2f int32 [3, Ne] facet of local edge (0,1) (1,2) (0,2); f2t int32 [2, n_facets] the two elements
 * of a facet (-1: boundary); s, w [nqf] Gauss-Legendre rule on [0, 1] (the reference: degree 2 * maxdeg);
 * facet_vec float64 [nbfun, 3, 2, nqf]: reference vector of every local basis function (Nedelec value / Lagrange
 * gradient) at the points of the three local edges.  out: float64 [Ne].                                          */
int fw_error_estimator(fw_handle* h, const double* E, int32_t eps_kind, int32_t eps_is_complex, const void* eps,
                       int32_t n_facets, const int32_t* t2f, const int32_t* f2t, int32_t nqf, const double* s,
                       const double* w, const double* facet_vec, double* out);

/* ---- whole path with HOST buffers (the call the Python shim makes for compute_modes) */
typedef struct {
  double wavelength, mu_r, radius;
  int32_t eps_kind, eps_is_complex;
  const void* eps;
  fw_eigs_params eigs;       /* sigma must be filled by the caller (waveguide.py:605-608) */
  int32_t reuse_symbolic;    /* 1: mesh unchanged since the last call (sweeps)          */
} fw_modes_params;
int fw_compute_modes(fw_handle* h, const fw_modes_params* prm, double* lams, double* E, double* H,
                     double* powers, fw_eigs_stats* st);

/* ---- timing of the most recent stages (ms, CUDA events on the handle's stream)        */
typedef struct {
  double upload_ms, symbolic_ms, element_kernel_ms, reduce_ms, lu_symbolic_ms, lu_factor_ms,
      arnoldi_ms, postprocess_ms, total_ms, download_ms;
  int64_t element_kernel_launches, kernel_launches, nnz_structural;
  /* roofline inputs of the sparse LU (from the analysis tables): entries of L and U that one pair of triangular
   * solves reads (sum over fronts of m^2 - (m - ns)^2) and the flops of one numeric factorisation              */
  int64_t lu_solve_entries;
  double lu_factor_flops;
  /* H-field mass solves (block PCG): iterations of the slowest system and its final ||r|| / ||b||             */
  int64_t pcg_iterations;
  double pcg_rel_residual;
  double hfield_ms;      /* calculate_hfield part of postprocess_ms                                             */
  int64_t mass_direct_blocks; /* mass blocks that went through the direct (sparse LU) solve in the last H-field     */
} fw_timings;
int fw_get_timings(fw_handle* h, fw_timings* t);

/* ---- host-only helpers exposed for the CPU tests                                     */
int fw_host_dense_eig(int32_t n, const double* a_c128_colmajor, double* w_c128, double* v_c128_colmajor);
/* analysis phase (nested dissection, fronts, extend-add maps) of the sparse LU, host only */
int fw_debug_lu_symbolic(int32_t n, int32_t nb, int32_t ne, const int32_t* edofs, const double* cx,
                         const double* cy, const uint8_t* active, int32_t leaf_elements, int64_t* sizes8);
int fw_debug_lu_symbolic_get(int32_t which, void* out);
/* tests only: device tables of the current sparse-LU analysis; which as fw_debug_lu_symbolic_get (0 perm, 1 iperm,
 * 2 node_of, 7 m, 8 ns, 9 own_start, 10 blist, 11 rel: int32; 21 blist_off: int64); sizes8 as fw_debug_lu_symbolic */
int fw_debug_lu_analysis(fw_handle* h, int32_t n_dirichlet, const int32_t* dirichlet, int32_t leaf_elements,
                         int64_t* sizes8);
int fw_debug_lu_analysis_get(fw_handle* h, int32_t which, void* out);
/* tests only: which = 0 factor arena (f64/c128, sum m^2 entries), 1 pivot vector (int32, n_active) */
int fw_debug_lu_get(fw_handle* h, int32_t which, void* out);

#ifdef __cplusplus
}
#endif
#endif
