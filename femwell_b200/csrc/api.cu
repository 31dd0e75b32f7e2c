// extern "C" boundary of libfemwell_b200 (declarations + reference citations: include/femwell_b200.h).
#include <thread>

#include "handle.h"
#include "lu.h"
#include "solver.h"

using namespace fw;

static thread_local std::string g_err;

#define FW_API_BEGIN try {
#define FW_API_END                         \
  }                                        \
  catch (const fw::Error& e) {             \
    g_err = e.what();                      \
    return e.code;                         \
  }                                        \
  catch (const std::exception& e) {        \
    g_err = e.what();                      \
    return FW_ERR_INTERNAL;                \
  }                                        \
  return FW_OK;

struct fw_handle {
  Handle h;
};

namespace fw {
Handle::Handle() {}
Handle::~Handle() {
  lu_prefetch.reset();
  if (stream) cudaStreamSynchronize(stream);
  lu.reset();
  for (int i = 0; i < 2; ++i) {
    if (pinned[i]) cudaFreeHost(pinned[i]);
    if (pinned_ev[i]) cudaEventDestroy(pinned_ev[i]);
  }
  if (stream) cudaStreamDestroy(stream);
}

// A cudaMemcpy into pageable memory runs at ~4 GB/s (driver-internal staging, single threaded).  Here the
// device data is copied in 32 MB chunks into a ring of two pinned buffers and a small thread team moves each
// chunk to its pageable destination while the next chunk is already on the bus.
void download_staged(Handle& h, void* host_dst, const void* dev_src, size_t bytes) {
  constexpr size_t CHUNK = (size_t)32 << 20;
  bool dst_pinned = false;
  if (bytes >= 4 * CHUNK) {  // page-locked destination (fw_host_alloc): one DMA at bus speed, no staging copy
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, host_dst) == cudaSuccess) dst_pinned = attr.type == cudaMemoryTypeHost;
    cudaGetLastError();  // an unregistered pointer is not an error here
  }
  if (bytes < 4 * CHUNK || dst_pinned) {
    FW_CHECK_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, h.stream));
    FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
    return;
  }
  for (int i = 0; i < 2; ++i) {
    if (!h.pinned[i]) FW_CHECK_CUDA(cudaHostAlloc(&h.pinned[i], CHUNK, cudaHostAllocDefault));
    if (!h.pinned_ev[i]) FW_CHECK_CUDA(cudaEventCreateWithFlags(&h.pinned_ev[i], cudaEventDisableTiming));
  }
  const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
  auto issue = [&](size_t c) {
    const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
    FW_CHECK_CUDA(cudaMemcpyAsync(h.pinned[c & 1], (const char*)dev_src + off, len, cudaMemcpyDeviceToHost, h.stream));
    FW_CHECK_CUDA(cudaEventRecord(h.pinned_ev[c & 1], h.stream));
  };
  const int nthreads = 8;
  issue(0);
  for (size_t c = 0; c < nchunk; ++c) {
    FW_CHECK_CUDA(cudaEventSynchronize(h.pinned_ev[c & 1]));
    if (c + 1 < nchunk) issue(c + 1);  // the other buffer was drained in the previous iteration
    const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
    char* dst = (char*)host_dst + off;
    const char* src = (const char*)h.pinned[c & 1];
    std::vector<std::thread> team;
    for (int t = 1; t < nthreads; ++t) {
      const size_t a = len * t / nthreads, b = len * (t + 1) / nthreads;
      team.emplace_back([=] { memcpy(dst + a, src + a, b - a); });
    }
    memcpy(dst, src, len / nthreads);
    for (auto& th : team) th.join();
  }
}
// The mirror image for large pageable-host -> device copies (mesh arrays, LU analysis tables): a thread team fills
// one pinned buffer while the other one is on the bus.
void upload_staged(Handle& h, void* dev_dst, const void* host_src, size_t bytes) {
  constexpr size_t CHUNK = (size_t)32 << 20;
  if (bytes == 0) return;
  if (bytes < ((size_t)4 << 20)) {
    FW_CHECK_CUDA(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, h.stream));
    return;
  }
  for (int i = 0; i < 2; ++i) {
    if (!h.pinned[i]) FW_CHECK_CUDA(cudaHostAlloc(&h.pinned[i], CHUNK, cudaHostAllocDefault));
    if (!h.pinned_ev[i]) FW_CHECK_CUDA(cudaEventCreateWithFlags(&h.pinned_ev[i], cudaEventDisableTiming));
  }
  // pending device->host use of the ring (download_staged) has been synchronised by its caller
  const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
  const int nthreads = 4;
  for (size_t c = 0; c < nchunk; ++c) {
    const int b = (int)(c & 1);
    if (c >= 2) FW_CHECK_CUDA(cudaEventSynchronize(h.pinned_ev[b]));  // the DMA out of this buffer has finished
    const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
    const char* src = (const char*)host_src + off;
    char* dst = (char*)h.pinned[b];
    std::vector<std::thread> team;
    for (int t = 1; t < nthreads; ++t) {
      const size_t a = len * t / nthreads, e = len * (t + 1) / nthreads;
      team.emplace_back([=] { memcpy(dst + a, src + a, e - a); });
    }
    memcpy(dst, src, len / nthreads);
    for (auto& th : team) th.join();
    FW_CHECK_CUDA(cudaMemcpyAsync((char*)dev_dst + off, h.pinned[b], len, cudaMemcpyHostToDevice, h.stream));
    FW_CHECK_CUDA(cudaEventRecord(h.pinned_ev[b], h.stream));
  }
  // the ring may be reused by the next staged copy only after these DMAs: wait here (the host data may also go away)
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
}
}  // namespace fw

static Values& pick(Handle& h, int which) {
  switch (which) {
    case FW_MAT_A:
      return h.A;
    case FW_MAT_B:
      return h.B;
    case FW_MAT_MASS:
      return h.Mass;
    case FW_MAT_C0:
      return h.C0;
    case FW_MAT_C1:
      return h.C1;
  }
  throw Error(FW_ERR_INVALID, "unknown matrix id");
}

extern "C" {

const char* fw_last_error(void) { return g_err.c_str(); }
int fw_version(void) { return 100; }

int fw_create(fw_handle** out, int device) {
  FW_API_BEGIN
  FW_REQUIRE(out != nullptr, "out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(FW_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                                 " -- libfemwell_b200 has no CPU fallback");
  FW_REQUIRE(device >= 0 && device < count, "bad device index");
  FW_CHECK_CUDA(cudaSetDevice(device));
  std::unique_ptr<fw_handle> hh(new fw_handle());
  hh->h.device = device;
  FW_CHECK_CUDA(cudaStreamCreateWithFlags(&hh->h.stream, cudaStreamNonBlocking));
  *out = hh.release();
  FW_API_END
}

int fw_destroy(fw_handle* h) {
  FW_API_BEGIN
  if (h) {
    cudaSetDevice(h->h.device);
    delete h;
  }
  FW_API_END
}

int fw_synchronize(fw_handle* h) {
  FW_API_BEGIN
  FW_REQUIRE(h, "handle is NULL");
  FW_CHECK_CUDA(cudaSetDevice(h->h.device));
  FW_CHECK_CUDA(cudaStreamSynchronize(h->h.stream));
  FW_API_END
}

int fw_set_space(fw_handle* hh, const fw_space* s) {
  FW_API_BEGIN
  FW_REQUIRE(hh && s, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  Space& sp = h.sp;
  h.lu_prefetch.reset();  // joins a pending analysis before the host arrays change
  // whatever happens below, the handle no longer holds the previous space: a failed call must leave it in the
  // "no space" state (host sizes, device buffers and the CSR pattern would otherwise disagree), and a one-shot
  // analysis hint must not leak to an unrelated mesh
  const bool hinted = h.hint_valid;
  h.hint_valid = false;
  sp.valid = false;
  sp.has_emask = false;
  h.generic.valid = false;
  h.has_symbolic = false;
  h.A.valid = h.B.valid = h.Mass.valid = h.C0.valid = h.C1.valid = h.MassC.valid = h.Bc.valid = false;
  h.lu.reset();
  FW_REQUIRE(s->order == 1 || s->order == 2, "Only order 1 and 2 implemented by now.");
  const int nb = s->order == 1 ? 6 : 14;
  FW_REQUIRE(s->nbfun == nb, "nbfun does not match order");
  FW_REQUIRE(s->nqp >= 1 && s->nqp <= MAXQ, "nqp out of range (1..12)");
  FW_REQUIRE(s->n_vertices > 0 && s->n_elements > 0 && s->n_dofs > 0, "empty mesh");
  FW_REQUIRE(s->p && s->t && s->element_dofs && s->kind && s->vec && s->sc && s->X && s->W, "NULL array");
  Timer tu(h.stream, "fw:set_space upload");
  sp.order = s->order;
  sp.nv = s->n_vertices;
  sp.ne = s->n_elements;
  sp.n = s->n_dofs;
  sp.nb = nb;
  sp.nq = s->nqp;
  sp.nt = 0;
  for (int l = 0; l < nb; ++l) {
    sp.kind[l] = s->kind[l];
    const int expect = nb == 6 ? (l < 3 ? 1 : 0) : (l < 3 ? 1 : (l >= 12 ? 0 : (((l - 3) % 3) == 2 ? 1 : 0)));
    FW_REQUIRE(sp.kind[l] == expect, "local dof layout must be entity-major (vertex, facet, interior)");
    sp.nt += sp.kind[l] == 0;
  }
  sp.nz = nb - sp.nt;
  sp.vec.assign(s->vec, s->vec + (size_t)nb * 2 * sp.nq);
  sp.sc.assign(s->sc, s->sc + (size_t)nb * sp.nq);
  sp.X.assign(s->X, s->X + (size_t)2 * sp.nq);
  sp.W.assign(s->W, s->W + (size_t)sp.nq);
  // host copies of the mesh are only needed by the host version of the LU analysis (FW_HOST_ANALYSIS=1)
  const bool keep_host = getenv("FW_HOST_ANALYSIS") != nullptr;
  sp.h_p.clear(), sp.h_t.clear(), sp.h_edofs.clear();
  if (keep_host) {
    sp.h_p.assign(s->p, s->p + (size_t)2 * sp.nv);
    sp.h_t.assign(s->t, s->t + (size_t)3 * sp.ne);
    sp.h_edofs.assign(s->element_dofs, s->element_dofs + (size_t)nb * sp.ne);
  }
  const int32_t* in_t = s->t;
  const int32_t* in_ed = s->element_dofs;
  // validate indices on the host (avoids device faults on bad input) and mark the Lagrange dofs; a few threads,
  // the connectivity has 14 x Ne entries
  std::vector<uint8_t> isz((size_t)sp.n, 0);
  {
    const int nthr = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    std::vector<int> bad(nthr, 0);
    std::vector<std::thread> team;
    uint8_t* iszp = isz.data();
    for (int w = 0; w < nthr; ++w) {
      team.emplace_back([&, w] {
        const int ea = (int)((int64_t)sp.ne * w / nthr), eb = (int)((int64_t)sp.ne * (w + 1) / nthr);
        int b = 0;
        for (int a = 0; a < 3; ++a)
          for (int e = ea; e < eb; ++e) {
            const int v = in_t[(size_t)a * sp.ne + e];
            b |= (v < 0 || v >= sp.nv) ? 1 : 0;
          }
        for (int l = 0; l < nb; ++l) {
          const bool z = sp.kind[l] == 1;
          for (int e = ea; e < eb; ++e) {
            const int d = in_ed[(size_t)l * sp.ne + e];
            if (d < 0 || d >= sp.n)
              b |= 2;
            else if (z)
              iszp[d] = 1;  // benign race: every writer stores the same value
          }
        }
        bad[w] = b;
      });
    }
    for (auto& th : team) th.join();
    int b = 0;
    for (int v : bad) b |= v;
    FW_REQUIRE(!(b & 1), "t out of range");
    FW_REQUIRE(!(b & 2), "element_dofs out of range");
  }
  if (hinted) {
    // the caller announced the solve that follows (fw_analysis_hint): the host analysis of its sparse LU starts now and
    // overlaps the uploads below and everything up to the factorisation
    std::vector<uint8_t> active((size_t)sp.n, 1);
    for (int d : h.hint_dirichlet) {
      FW_REQUIRE(d >= 0 && d < sp.n, "dirichlet dof out of range");
      active[d] = 0;
    }
    h.generic.valid = false;
    lu_analyse_prefetch(h, active, h.hint_leaf);
  }
  sp.p.alloc((size_t)2 * sp.nv);
  sp.t.alloc((size_t)3 * sp.ne);
  sp.edofs.alloc((size_t)nb * sp.ne);
  sp.is_z.alloc(isz.size());
  upload_staged(h, sp.p.p, s->p, (size_t)2 * sp.nv * sizeof(double));
  upload_staged(h, sp.t.p, s->t, (size_t)3 * sp.ne * sizeof(int));
  upload_staged(h, sp.edofs.p, s->element_dofs, (size_t)nb * sp.ne * sizeof(int));
  upload_staged(h, sp.is_z.p, isz.data(), isz.size());
  sp.h_is_z = isz;
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  sp.valid = true;
  h.generic.valid = false;
  h.has_symbolic = false;
  h.A.valid = h.B.valid = h.Mass.valid = h.C0.valid = h.C1.valid = h.MassC.valid = h.Bc.valid = false;
  h.lu.reset();
  h.tim = fw_timings{};
  h.tim.upload_ms = tu.stop();
  FW_API_END
}

int fw_set_element_subset(fw_handle* hh, int32_t n_sel, const int32_t* elements) {
  FW_API_BEGIN
  FW_REQUIRE(hh, "handle is NULL");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "fw_set_space must be called first");
  // the assembled operators no longer match the basis
  h.A.valid = h.B.valid = h.Mass.valid = h.C0.valid = h.C1.valid = h.MassC.valid = h.Bc.valid = false;
  if (n_sel < 0 || !elements) {
    sp.has_emask = false;
    sp.emask.release();
    return FW_OK;
  }
  std::vector<uint8_t> m((size_t)sp.ne, 0);
  for (int i = 0; i < n_sel; ++i) {
    FW_REQUIRE(elements[i] >= 0 && elements[i] < sp.ne, "element index out of range");
    m[elements[i]] = 1;
  }
  sp.emask.upload(m.data(), m.size(), h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  sp.has_emask = true;
  FW_API_END
}

int fw_analysis_hint(fw_handle* hh, int32_t n_dirichlet, const int32_t* dirichlet, int32_t leaf_elements) {
  FW_API_BEGIN
  FW_REQUIRE(hh && (n_dirichlet == 0 || dirichlet), "NULL argument");
  Handle& h = hh->h;
  h.hint_dirichlet.assign(dirichlet, dirichlet + (n_dirichlet > 0 ? n_dirichlet : 0));
  h.hint_leaf = leaf_elements;
  h.hint_valid = true;
  FW_API_END
}

int fw_set_option(fw_handle* hh, int32_t option, int32_t value) {
  FW_API_BEGIN
  FW_REQUIRE(hh, "handle is NULL");
  switch (option) {
    case FW_OPT_ELIMINATE_ZEROS:
      hh->h.eliminate_zeros = value != 0;
      break;
    case FW_OPT_MASS_SOLVER:
      FW_REQUIRE(value >= 0 && value <= 2, "mass solver: 0 auto, 1 PCG, 2 direct");
      hh->h.mass_solver = value;
      break;
    case FW_OPT_PCG_TOL_EXP:
      FW_REQUIRE(value >= 1 && value <= 15, "PCG tolerance exponent out of range (1..15)");
      hh->h.pcg_tol = std::pow(10.0, -(double)value);
      break;
    case FW_OPT_START_VECTOR:
      hh->h.smooth_v0 = value != 0;
      break;

    default:
      throw Error(FW_ERR_INVALID, "unknown option");
  }
  FW_API_END
}

int fw_symbolic(fw_handle* hh, int64_t* nnz) {
  FW_API_BEGIN
  FW_REQUIRE(hh, "handle is NULL");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  Timer t(h.stream, "fw:symbolic (COO sort + CSR pattern)");
  symbolic_space(h);
  h.tim.symbolic_ms = t.stop();
  if (nnz) *nnz = h.pat.nnz;
  FW_API_END
}

int fw_assemble_waveguide(fw_handle* hh, int32_t eps_kind, int32_t eps_is_complex, const void* eps, double k0,
                          double mu_r, double radius) {
  FW_API_BEGIN
  FW_REQUIRE(hh, "handle is NULL");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  assemble_waveguide(h, eps_kind, eps_is_complex != 0, eps, k0, mu_r, radius);
  FW_API_END
}

int fw_matrix_nnz(fw_handle* hh, int32_t which, int64_t* nnz, int32_t* is_complex) {
  FW_API_BEGIN
  FW_REQUIRE(hh && nnz, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  if (which >= FW_MAT_MASS) assemble_aux(h);
  Values& v = pick(h, which);
  FW_REQUIRE(v.valid, "matrix has not been assembled");
  *nnz = count_kept(h, h.pat, v);
  if (is_complex) *is_complex = v.is_complex;
  FW_API_END
}

int fw_matrix_export(fw_handle* hh, int32_t which, int32_t* indptr, int32_t* indices, void* data) {
  FW_API_BEGIN
  FW_REQUIRE(hh && indptr && indices && data, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  if (which >= FW_MAT_MASS) assemble_aux(h);
  Values& v = pick(h, which);
  FW_REQUIRE(v.valid, "matrix has not been assembled");
  export_compacted(h, h.pat, v, indptr, indices, data);
  FW_API_END
}

int fw_coo_to_csr(fw_handle* hh, int32_t n_rows, int64_t n_entries, const int32_t* rows, const int32_t* cols,
                  const void* data, int32_t is_complex, int64_t* nnz) {
  FW_API_BEGIN
  FW_REQUIRE(hh && nnz, "NULL argument");
  FW_REQUIRE(n_rows > 0 && n_entries >= 0, "bad sizes");
  FW_REQUIRE(n_entries == 0 || (rows && cols && data), "NULL array");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  for (int64_t i = 0; i < n_entries; ++i)
    FW_REQUIRE(rows[i] >= 0 && rows[i] < n_rows && cols[i] >= 0 && cols[i] < n_rows, "COO index out of range");
  coo_to_csr(h, n_rows, n_entries, rows, cols, data, is_complex != 0);
  *nnz = h.coo_pat.nnz;
  FW_API_END
}

int fw_coo_to_csr_export(fw_handle* hh, int32_t* indptr, int32_t* indices, void* data) {
  FW_API_BEGIN
  FW_REQUIRE(hh && indptr && indices && data, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.coo_val.valid, "fw_coo_to_csr has not been called");
  export_compacted(h, h.coo_pat, h.coo_val, indptr, indices, data);
  FW_API_END
}

int fw_matrix_spmv(fw_handle* hh, int32_t which, int32_t x_is_complex, const void* x, void* y) {
  FW_API_BEGIN
  FW_REQUIRE(hh && x && y, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  if (which >= FW_MAT_MASS) assemble_aux(h);
  Values& v = pick(h, which);
  FW_REQUIRE(v.valid, "matrix has not been assembled");
  const int n = h.sp.n;
  cudaStream_t s = h.stream;
  const bool out_complex = v.is_complex || x_is_complex;
  DevBuf<double> dx, dy;
  if (out_complex && !x_is_complex) {
    // promote x to complex on the host side of the copy
    std::vector<double> xc((size_t)2 * n, 0.0);
    for (int i = 0; i < n; ++i) xc[2 * (size_t)i] = ((const double*)x)[i];
    dx.upload(xc.data(), xc.size(), s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
  } else {
    dx.upload((const double*)x, (size_t)n * (x_is_complex ? 2 : 1), s);
  }
  dy.alloc((size_t)n * (out_complex ? 2 : 1));
  if (!out_complex)
    spmv<double, double>(n, h.pat.indptr.p, h.pat.indices.p, v.v.p, dx.p, dy.p, s);
  else if (!v.is_complex)
    spmv<double, cplx>(n, h.pat.indptr.p, h.pat.indices.p, v.v.p, (const cplx*)dx.p, (cplx*)dy.p, s);
  else
    spmv<cplx, cplx>(n, h.pat.indptr.p, h.pat.indices.p, (const cplx*)v.v.p, (const cplx*)dx.p, (cplx*)dy.p, s);
  dy.download((double*)y, (size_t)n * (out_complex ? 2 : 1), s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  FW_API_END
}

int fw_bench_spmv(fw_handle* hh, int32_t which, int32_t x_is_complex, int32_t reps, double* ms_per_apply) {
  FW_API_BEGIN
  FW_REQUIRE(hh && ms_per_apply && reps >= 1, "bad argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  if (which >= FW_MAT_MASS) assemble_aux(h);
  Values& v = pick(h, which);
  FW_REQUIRE(v.valid, "matrix has not been assembled");
  const int n = h.sp.n;
  cudaStream_t s = h.stream;
  const bool cx = v.is_complex || x_is_complex;
  DevBuf<double> dx, dy;
  dx.alloc((size_t)n * (cx ? 2 : 1));
  dy.alloc((size_t)n * (cx ? 2 : 1));
  std::vector<double> x0((size_t)n * (cx ? 2 : 1));
  for (size_t i = 0; i < x0.size(); ++i) x0[i] = 1.0 / (1.0 + (double)(i % 97));
  dx.upload(x0.data(), x0.size(), s);
  auto once = [&] {
    if (!cx)
      spmv<double, double>(n, h.pat.indptr.p, h.pat.indices.p, v.v.p, dx.p, dy.p, s);
    else if (!v.is_complex)
      spmv<double, cplx>(n, h.pat.indptr.p, h.pat.indices.p, v.v.p, (const cplx*)dx.p, (cplx*)dy.p, s);
    else
      spmv<cplx, cplx>(n, h.pat.indptr.p, h.pat.indices.p, (const cplx*)v.v.p, (const cplx*)dx.p, (cplx*)dy.p, s);
  };
  for (int i = 0; i < 3; ++i) once();
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  Timer t(s);
  for (int i = 0; i < reps; ++i) once();
  *ms_per_apply = t.stop() / reps;
  FW_CHECK_CUDA(cudaGetLastError());
  FW_API_END
}

int fw_eigs(fw_handle* hh, const fw_eigs_params* prm, double* lams, double* X, fw_eigs_stats* st) {
  FW_API_BEGIN
  FW_REQUIRE(hh && prm && lams && X, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  DevBuf<double> Xd;
  std::string why;
  const bool ok = eigs_shift_invert(h, *prm, (std::complex<double>*)lams, Xd, st, &why);
  Xd.download(X, (size_t)prm->k * h.sp.n * 2, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  if (!ok) throw Error(FW_ERR_NOCONV, why);  // the converged pairs (st->nconv) have been handed over
  FW_API_END
}

int fw_eigs_csr(fw_handle* hh, int32_t n, const int32_t* K_indptr, const int32_t* K_indices, const void* K_data,
                const int32_t* M_indptr, const int32_t* M_indices, const void* M_data, int32_t is_complex,
                const double* coords, const fw_eigs_params* prm, double* lams, double* X, fw_eigs_stats* st) {
  FW_API_BEGIN
  FW_REQUIRE(hh && prm && lams && X, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  h.tim = fw_timings{};
  {
    Timer t(h.stream, "fw:generic CSR setup");
    setup_generic_csr(h, n, K_indptr, K_indices, K_data, M_indptr, M_indices, M_data, is_complex != 0, coords);
    h.tim.upload_ms = t.stop();
  }
  DevBuf<double> Xd;
  std::string why;
  const bool ok = eigs_shift_invert(h, *prm, (std::complex<double>*)lams, Xd, st, &why);
  Xd.download(X, (size_t)prm->k * n * 2, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  if (!ok) throw Error(FW_ERR_NOCONV, why);
  FW_API_END
}

int fw_lu_factor_shifted(fw_handle* hh, double sigma_re, double sigma_im, int32_t n_dirichlet,
                         const int32_t* dirichlet, int32_t leaf_elements) {
  FW_API_BEGIN
  FW_REQUIRE(hh, "handle is NULL");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.A.valid && h.B.valid, "assemble A and B first");
  std::vector<uint8_t> active((size_t)h.sp.n, 1);
  for (int i = 0; i < n_dirichlet; ++i) {
    FW_REQUIRE(dirichlet[i] >= 0 && dirichlet[i] < h.sp.n, "dirichlet dof out of range");
    active[dirichlet[i]] = 0;
  }
  {
    Timer t(h.stream, "fw:LU analysis");
    lu_analyse(h, active, leaf_elements);
    h.tim.lu_symbolic_ms = t.stop();
  }
  {
    Timer t(h.stream, "fw:LU factor");
    lu_factor(h, -1.0, 0.0, sigma_re, sigma_im, h.A.is_complex || sigma_im != 0.0);
    h.tim.lu_factor_ms = t.stop();
  }
  h.lu_sigma_re = sigma_re;
  h.lu_sigma_im = sigma_im;
  FW_API_END
}

int fw_lu_solve(fw_handle* hh, int32_t rhs_is_complex, const void* b, void* x, int32_t refine) {
  FW_API_BEGIN
  FW_REQUIRE(hh && b && x, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.lu && h.lu->factored, "fw_lu_factor_shifted must be called first");
  (void)refine;
  const int n = h.sp.n;
  cudaStream_t s = h.stream;
  const bool lu_c = h.lu->is_complex;
  DevBuf<double> db, dx;
  // FW_SOLVE_TIME=1: device time of the substitution pair alone (scripts/nrhs_probe.py)
  auto timed_solve = [&](const void* bp, void* xp, int nrhs) {
    if (!getenv("FW_SOLVE_TIME")) return lu_solve_dev(h, bp, xp, nrhs);
    Timer t(s);
    lu_solve_dev(h, bp, xp, nrhs);
    fprintf(stderr, "[fw-solve] nrhs %d: %.3f ms\n", nrhs, t.stop());
  };
  if (lu_c) {
    std::vector<double> tmp((size_t)2 * n);
    if (rhs_is_complex)
      memcpy(tmp.data(), b, sizeof(double) * 2 * n);
    else
      for (int i = 0; i < n; ++i) tmp[2 * (size_t)i] = ((const double*)b)[i], tmp[2 * (size_t)i + 1] = 0.0;
    db.upload(tmp.data(), tmp.size(), s);
    dx.alloc((size_t)2 * n);
    timed_solve(db.p, dx.p, 1);
    std::vector<double> out((size_t)2 * n);
    dx.download(out.data(), out.size(), s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    if (rhs_is_complex)
      memcpy(x, out.data(), sizeof(double) * 2 * n);
    else
      for (int i = 0; i < n; ++i) ((double*)x)[i] = out[2 * (size_t)i];
  } else if (!rhs_is_complex) {
    db.upload((const double*)b, (size_t)n, s);
    dx.alloc((size_t)n);
    timed_solve(db.p, dx.p, 1);
    dx.download((double*)x, (size_t)n, s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
  } else {
    // real factors, complex rhs: solve real and imaginary parts as two right-hand sides
    std::vector<double> tmp((size_t)2 * n);
    for (int i = 0; i < n; ++i) {
      tmp[i] = ((const double*)b)[2 * (size_t)i];
      tmp[(size_t)n + i] = ((const double*)b)[2 * (size_t)i + 1];
    }
    db.upload(tmp.data(), tmp.size(), s);
    dx.alloc((size_t)2 * n);
    timed_solve(db.p, dx.p, 2);
    dx.download(tmp.data(), tmp.size(), s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) {
      ((double*)x)[2 * (size_t)i] = tmp[i];
      ((double*)x)[2 * (size_t)i + 1] = tmp[(size_t)n + i];
    }
  }
  FW_API_END
}

int fw_postprocess_modes(fw_handle* hh, int32_t k, const double* lams, const double* X, double k0, double omega,
                         double mu0, double* E, double* H, double* powers) {
  FW_API_BEGIN
  FW_REQUIRE(hh && lams && X && E && H && powers && k >= 1, "bad argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid && h.has_symbolic, "space/symbolic missing");
  const size_t cnt = (size_t)k * h.sp.n * 2;
  DevBuf<double> dE, dH;
  dE.upload(X, cnt, h.stream);
  dH.alloc(cnt);
  Timer t(h.stream, "fw:postprocess (H field, power)");
  postprocess_modes_dev(h, k, (const std::complex<double>*)lams, dE.p, k0, omega, mu0, dE.p, dH.p,
                        (std::complex<double>*)powers);
  h.tim.postprocess_ms = t.stop();
  dE.download(E, cnt, h.stream);
  dH.download(H, cnt, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_API_END
}

int fw_hfield(fw_handle* hh, const double* E, double beta_re, double beta_im, double omega, double mu0, double* H) {
  FW_API_BEGIN
  FW_REQUIRE(hh && E && H, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid && h.has_symbolic, "space/symbolic missing");
  const size_t cnt = (size_t)h.sp.n * 2;
  DevBuf<double> dE, dH;
  dE.upload(E, cnt, h.stream);
  dH.alloc(cnt);
  hfield_dev(h, (const cplx*)dE.p, cplx{beta_re, beta_im}, omega, mu0, (cplx*)dH.p);
  dH.download(H, cnt, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_API_END
}

int fw_intensity(fw_handle* hh, const double* E, const double* H, double* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && E && H && out, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid, "space missing");
  const size_t cnt = (size_t)h.sp.n * 2;
  DevBuf<double> dE, dH, dout;
  dE.upload(E, cnt, h.stream);
  dH.upload(H, cnt, h.stream);
  dout.alloc((size_t)3 * h.sp.ne);
  intensity_dev(h, (const cplx*)dE.p, (const cplx*)dH.p, dout.p);
  dout.download(out, (size_t)3 * h.sp.ne, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_API_END
}

int fw_overlap_cross(fw_handle* hh, const fw_space* space_j, const double* E_i, const double* H_i, const double* E_j,
                     const double* H_j, int32_t mode, int32_t n_sel, const int32_t* elements, double* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && space_j && E_i && H_j && out, "NULL argument");
  FW_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (overlap) or 1 (scalar product)");
  FW_REQUIRE(mode == 1 || (H_i && E_j), "the overlap needs all four fields");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "fw_set_space must be called first");
  cudaStream_t s = h.stream;
  DevBuf<double> dEi, dHi;
  dEi.upload(E_i, (size_t)sp.n * 2, s);
  if (H_i) dHi.upload(H_i, (size_t)sp.n * 2, s);
  DevBuf<int> dsel;
  if (elements) {
    for (int i = 0; i < n_sel; ++i) FW_REQUIRE(elements[i] >= 0 && elements[i] < sp.ne, "element index out of range");
    dsel.upload(elements, (size_t)n_sel, s);
  }
  const std::complex<double> r = overlap_cross_dev(h, *space_j, (const cplx*)dEi.p, H_i ? (const cplx*)dHi.p : nullptr,
                                                   E_j, H_j, mode, n_sel, elements ? dsel.p : nullptr);
  out[0] = r.real(), out[1] = r.imag();
  FW_API_END
}

int fw_poynting(fw_handle* hh, const double* E, const double* H, double* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && E && H && out, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid, "space missing");
  const size_t cnt = (size_t)h.sp.n * 2;
  const size_t nout = (size_t)h.sp.ne * 3 * (h.sp.nb == 6 ? 3 : 6) * 2;
  DevBuf<double> dE, dH, dout;
  dE.upload(E, cnt, h.stream);
  dH.upload(H, cnt, h.stream);
  dout.alloc(nout);
  poynting_dev(h, (const cplx*)dE.p, (const cplx*)dH.p, (cplx*)dout.p);
  dout.download(out, nout, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_API_END
}

int fw_energy_current_density(fw_handle* hh, const double* xs, double* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && xs && out, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid, "space missing");
  DevBuf<double> dE, dout;
  dE.upload(xs, (size_t)h.sp.n * 2, h.stream);
  dout.alloc((size_t)h.sp.ne * 2);
  energy_density_dev(h, (const cplx*)dE.p, (cplx*)dout.p);
  dout.download(out, (size_t)h.sp.ne * 2, h.stream);
  FW_CHECK_CUDA(cudaStreamSynchronize(h.stream));
  FW_API_END
}

int fw_error_estimator(fw_handle* hh, const double* E, int32_t eps_kind, int32_t eps_is_complex, const void* eps,
                       int32_t n_facets, const int32_t* t2f, const int32_t* f2t, int32_t nqf, const double* sq,
                       const double* wq, const double* facet_vec, double* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && E && eps && t2f && f2t && sq && wq && facet_vec && out, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  const Space& sp = h.sp;
  FW_REQUIRE(sp.valid, "space missing");
  FW_REQUIRE(n_facets > 0, "no facets");
  for (size_t i = 0; i < (size_t)3 * sp.ne; ++i) FW_REQUIRE(t2f[i] >= 0 && t2f[i] < n_facets, "t2f out of range");
  for (size_t i = 0; i < (size_t)2 * n_facets; ++i) FW_REQUIRE(f2t[i] >= -1 && f2t[i] < sp.ne, "f2t out of range");
  cudaStream_t s = h.stream;
  DevBuf<double> dE, deps, dout;
  DevBuf<int> dt2f, df2t;
  dE.upload(E, (size_t)sp.n * 2, s);
  const size_t neps = (eps_kind == FW_EPS_P0 ? (size_t)sp.ne : (size_t)3 * sp.ne) * (eps_is_complex ? 2 : 1);
  deps.upload((const double*)eps, neps, s);
  dt2f.upload(t2f, (size_t)3 * sp.ne, s);
  df2t.upload(f2t, (size_t)2 * n_facets, s);
  dout.alloc((size_t)sp.ne);
  error_estimator_dev(h, (const cplx*)dE.p, eps_kind, eps_is_complex != 0, deps.p, n_facets, dt2f.p, df2t.p, nqf, sq, wq,
                      facet_vec, dout.p);
  dout.download(out, (size_t)sp.ne, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  FW_API_END
}

int fw_functional(fw_handle* hh, int32_t kind, const double* f0, const double* f1, const double* f2,
                  const double* f3, int32_t aux_kind, int32_t aux_is_complex, const void* aux, int32_t n_sel,
                  const int32_t* elements, int32_t option, double* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && f0 && out, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid, "space missing");
  const Space& sp = h.sp;
  cudaStream_t s = h.stream;
  const size_t cnt = (size_t)sp.n * 2;
  const double* hf[4] = {f0, f1, f2, f3};
  DevBuf<double> df[4];
  const cplx* dp[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < 4; ++i)
    if (hf[i]) {
      // identical host pointers (e.g. the self overlap) share one device copy
      int same = -1;
      for (int j = 0; j < i; ++j)
        if (hf[j] == hf[i]) same = j;
      if (same >= 0) {
        dp[i] = dp[same];
      } else {
        df[i].upload(hf[i], cnt, s);
        dp[i] = (const cplx*)df[i].p;
      }
    }
  DevBuf<double> daux;
  if (aux) {
    const size_t na = (aux_kind == FW_EPS_P0 ? (size_t)sp.ne : (size_t)3 * sp.ne) * (aux_is_complex ? 2 : 1);
    daux.upload((const double*)aux, na, s);
  }
  DevBuf<int> dsel;
  if (elements) {
    for (int i = 0; i < n_sel; ++i) FW_REQUIRE(elements[i] >= 0 && elements[i] < sp.ne, "element index out of range");
    dsel.upload(elements, (size_t)n_sel, s);
  }
  std::complex<double> o3[3];
  functional_dev(h, kind, dp[0], dp[1], dp[2], dp[3], aux_kind, aux_is_complex != 0, aux ? daux.p : nullptr, n_sel,
                 elements ? dsel.p : nullptr, option, o3);
  for (int k = 0; k < 3; ++k) out[2 * k] = o3[k].real(), out[2 * k + 1] = o3[k].imag();
  FW_API_END
}

int fw_compute_modes(fw_handle* hh, const fw_modes_params* prm, double* lams, double* E, double* H, double* powers,
                     fw_eigs_stats* st) {
  FW_API_BEGIN
  FW_REQUIRE(hh && prm && lams && E && H && powers, "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid, "fw_set_space must be called first");
  cudaStream_t s = h.stream;
  const double keep_upload = h.tim.upload_ms;
  const double keep_symbolic = h.tim.symbolic_ms;
  h.tim = fw_timings{};
  h.tim.upload_ms = keep_upload;
  h.tim.symbolic_ms = keep_symbolic;
  Timer total(s, "fw:compute_modes");
  {
    // the host analysis of the sparse LU only needs the mesh and the Dirichlet set: start it now, it runs on worker
    // threads while the GPU sorts the COO keys and assembles
    std::vector<uint8_t> active((size_t)h.sp.n, 1);
    for (int i = 0; i < prm->eigs.n_dirichlet; ++i) {
      const int d = prm->eigs.dirichlet[i];
      FW_REQUIRE(d >= 0 && d < h.sp.n, "dirichlet dof out of range");
      active[d] = 0;
    }
    lu_analyse_prefetch(h, active, prm->eigs.leaf_elements);
  }
  if (!(prm->reuse_symbolic && h.has_symbolic)) {
    Timer t(s, "fw:symbolic (COO sort + CSR pattern)");
    symbolic_space(h);
    h.tim.symbolic_ms = t.stop();
  }
  const double k0 = 2.0 * M_PI / prm->wavelength;
  assemble_waveguide(h, prm->eps_kind, prm->eps_is_complex != 0, prm->eps, k0, prm->mu_r, prm->radius);
  // geometry-only operators of the H-field recovery: assembled now, while the host analysis of the LU is still running
  assemble_aux(h);
  DevBuf<double> Xd, Hd;
  std::string why;
  h.v0_dev_valid = !prm->eigs.v0 && h.smooth_v0 &&
                   smooth_start_vector(h, prm->eps_kind, prm->eps_is_complex != 0, prm->eps, h.v0_dev);
  if (getenv("FW_TRACE"))
    fprintf(stderr, "[fw-trace] start vector: user v0 %d, option %d, smooth vector ready %d\n", prm->eigs.v0 != nullptr,
            (int)h.smooth_v0, (int)h.v0_dev_valid);
  bool ok = false;
  try {
    ok = eigs_shift_invert(h, prm->eigs, (std::complex<double>*)lams, Xd, st, &why);
  } catch (...) {
    h.v0_dev_valid = false;
    throw;
  }
  h.v0_dev_valid = false;  // one shot: fw_eigs / fw_eigs_csr keep the uniform(-1, 1) default
  const int k = prm->eigs.k;
  const size_t cnt = (size_t)k * h.sp.n * 2;
  if (!ok) {
    // no convergence: the raw (not un-scaled, unit 2-norm) converged eigenvectors go to E[0:nconv], like the
    // eigenvectors attribute of scipy's ArpackNoConvergence
    Xd.download(E, cnt, s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    throw Error(FW_ERR_NOCONV, why);
  }
  Hd.alloc(cnt);
  {
    Timer t(s, "fw:postprocess (H field, power)");
    const double c0 = 299792458.0, mu0 = 1.25663706212e-06;
    postprocess_modes_dev(h, k, (const std::complex<double>*)lams, Xd.p, k0, k0 * c0, mu0, Xd.p, Hd.p,
                          (std::complex<double>*)powers);
    h.tim.postprocess_ms = t.stop();
  }
  // total_ms: device work complete (results resident in HBM); download_ms: D2H of lams/E/H
  h.tim.total_ms = total.stop();
  h.tim.nnz_structural = h.pat.nnz;
  {
    Timer t(s, "fw:download");
    download_staged(h, E, Xd.p, cnt * sizeof(double));
    download_staged(h, H, Hd.p, cnt * sizeof(double));
    h.tim.download_ms = t.stop();
  }
  FW_API_END
}

void* fw_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

int fw_host_free(void* p) {
  FW_API_BEGIN
  if (p) FW_CHECK_CUDA(cudaFreeHost(p));
  FW_API_END
}

int fw_host_dense_eig(int32_t n, const double* a, double* w, double* v) {
  FW_API_BEGIN
  FW_REQUIRE(n >= 1 && a && w && v, "bad argument");
  if (!dense_eig(n, (const std::complex<double>*)a, (std::complex<double>*)w, (std::complex<double>*)v))
    throw Error(FW_ERR_NOCONV, "QR iteration failed");
  FW_API_END
}

// host-only: analysis phase of the sparse LU on caller-provided element connectivity (CPU tests)
static thread_local std::unique_ptr<LUSymbolic> g_dbg_sym;
int fw_debug_lu_symbolic(int32_t n, int32_t nb, int32_t ne, const int32_t* edofs, const double* cx, const double* cy,
                         const uint8_t* active, int32_t leaf_elements, int64_t* sizes /*[8]*/) {
  FW_API_BEGIN
  FW_REQUIRE(edofs && cx && cy && active && sizes, "NULL argument");
  g_dbg_sym.reset(new LUSymbolic());
  lu_symbolic_build(*g_dbg_sym, n, nb, ne, edofs, cx, cy, active, leaf_elements);
  const LUSymbolic& S = *g_dbg_sym;
  sizes[0] = S.n_active, sizes[1] = S.n_nodes, sizes[2] = S.n_levels, sizes[3] = (int64_t)S.blist.size();
  sizes[4] = S.factor_entries, sizes[5] = S.work_entries, sizes[6] = S.max_m, sizes[7] = S.max_ns;
  FW_API_END
}
// which: 0 perm, 1 iperm, 2 node_of, 3 parent, 4 child0, 5 child1, 6 depth, 7 m, 8 ns, 9 own_start,
//        10 blist, 11 rel, 12 level_start, 13 level_nodes (int32) ; 20 front_off, 21 blist_off (int64)
int fw_debug_lu_symbolic_get(int32_t which, void* out) {
  FW_API_BEGIN
  FW_REQUIRE(g_dbg_sym && out, "fw_debug_lu_symbolic must be called first");
  const LUSymbolic& S = *g_dbg_sym;
  auto cp = [&](const std::vector<int>& v) { memcpy(out, v.data(), v.size() * sizeof(int)); };
  auto cp64 = [&](const std::vector<int64_t>& v) { memcpy(out, v.data(), v.size() * sizeof(int64_t)); };
  switch (which) {
    case 0: cp(S.perm); break;
    case 1: cp(S.iperm); break;
    case 2: cp(S.node_of); break;
    case 3: cp(S.parent); break;
    case 4: cp(S.child0); break;
    case 5: cp(S.child1); break;
    case 6: cp(S.depth); break;
    case 7: cp(S.m); break;
    case 8: cp(S.ns); break;
    case 9: cp(S.own_start); break;
    case 10: cp(S.blist); break;
    case 11: cp(S.rel); break;
    case 12: cp(S.level_start); break;
    case 13: cp(S.level_nodes); break;
    case 20: cp64(S.front_off); break;
    case 21: cp64(S.blist_off); break;
    default: throw Error(FW_ERR_INVALID, "unknown array id");
  }
  FW_API_END
}

int fw_debug_lu_analysis(fw_handle* hh, int32_t n_dirichlet, const int32_t* dirichlet, int32_t leaf_elements,
                         int64_t* sizes) {
  FW_API_BEGIN
  FW_REQUIRE(hh && sizes && (n_dirichlet == 0 || dirichlet), "NULL argument");
  Handle& h = hh->h;
  FW_CHECK_CUDA(cudaSetDevice(h.device));
  current_pool() = &h.pool;
  FW_REQUIRE(h.sp.valid, "fw_set_space must be called first");
  std::vector<uint8_t> active((size_t)h.sp.n, 1);
  for (int i = 0; i < n_dirichlet; ++i) {
    FW_REQUIRE(dirichlet[i] >= 0 && dirichlet[i] < h.sp.n, "dirichlet dof out of range");
    active[dirichlet[i]] = 0;
  }
  h.lu.reset();  // never reuse: this call exists to look at a fresh analysis
  lu_analyse(h, active, leaf_elements);
  const LUSymbolic& S = h.lu->sym;
  sizes[0] = S.n_active, sizes[1] = S.n_nodes, sizes[2] = S.n_levels, sizes[3] = S.work_entries - S.n_active;
  sizes[4] = S.factor_entries, sizes[5] = S.work_entries, sizes[6] = S.max_m, sizes[7] = S.max_ns;
  FW_API_END
}
int fw_debug_lu_analysis_get(fw_handle* hh, int32_t which, void* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && out, "NULL argument");
  FW_CHECK_CUDA(cudaSetDevice(hh->h.device));
  current_pool() = &hh->h.pool;
  lu_debug_analysis_download(hh->h, which, out);
  FW_API_END
}

int fw_debug_lu_get(fw_handle* hh, int32_t which, void* out) {
  FW_API_BEGIN
  FW_REQUIRE(hh && out, "NULL argument");
  FW_CHECK_CUDA(cudaSetDevice(hh->h.device));
  current_pool() = &hh->h.pool;
  lu_debug_download(hh->h, which, out);
  FW_API_END
}

int fw_get_timings(fw_handle* hh, fw_timings* t) {
  FW_API_BEGIN
  FW_REQUIRE(hh && t, "NULL argument");
  *t = hh->h.tim;
  FW_API_END
}

}  // extern "C"
