// Multi-right-hand-side Jacobi-preconditioned CG for the L2 projection of calculate_hfield
// (femwell/maxwell/waveguide.py:666-677: scipy.sparse.linalg.spsolve(bform.assemble(basis), ...)).
//
// The mass matrix (block diagonal: Nedelec mass + Lagrange mass) is real SPD and mode independent, the
// right-hand sides are the k complex vectors C(beta_i) e_i.  All k systems advance together so that the
// matrix is streamed once per iteration for all modes, and every CG scalar (alpha, beta, residual norms)
// lives in device memory -- the host only looks at the residuals every few iterations.
//
// Layout: the CG vectors are INTERLEAVED, v[i*k + r] (dof i, system r): the SpMV gather of one matrix entry
// then reads k consecutive complex values (64 B for k = 4, two full sectors) instead of k scattered 16 B
// pieces, which halves the L2 -> SM traffic of the dominant kernel.  The matrix is the compact copy without the
// structurally-zero Nedelec x Lagrange blocks of the shared pattern (Handle::MassC).  The <p, M p> partial
// sums are produced by the SpMV itself.  HBM-bound: nnz*12 + ~9 k N 16 bytes per iteration, 5 launches.
#include <algorithm>

#include "lu.h"
#include "solver.h"

namespace fw {

constexpr int CG_MAXK = 8;

__device__ inline double block_sum_256(double s, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double q = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) q += red[i];
  __syncthreads();
  return q;  // valid in thread 0
}

// q[row] = sum_c M[row,c] p[c] for all K systems; part_pq[r][block] = sum_rows Re(conj(p_r[row]) q_r[row]).
// The matrix is real, so the K complex systems are 2K independent real columns: 2K lanes per row, lane c owns
// real component c of the interleaved vectors and walks all entries of the row.  The lanes of a row read the
// same (value, column) pair (one broadcast sector each) and 16 K contiguous bytes of p -- no shuffle reduction.
template <int K>
__global__ void __launch_bounds__(256) k_spmv_multi(int n, const int* __restrict__ indptr,
                                                    const int* __restrict__ indices, const double* __restrict__ vals,
                                                    const double* __restrict__ x, double* __restrict__ y,
                                                    double* __restrict__ part_pq, int nblk) {
  constexpr int C = 2 * K;            // real components per dof
  constexpr int RPB = 256 / C;        // rows per block
  __shared__ double red[8][C];
  const int g = threadIdx.x / C, c = threadIdx.x % C;
  const int row = blockIdx.x * RPB + g;
  double acc = 0.0, dot = 0.0;
  if (row < n) {
    const int a = indptr[row], b = indptr[row + 1];
    int e = a;
    for (; e + 4 <= b; e += 4) {
      const double v0 = vals[e], v1 = vals[e + 1], v2 = vals[e + 2], v3 = vals[e + 3];
      const int c0 = indices[e], c1 = indices[e + 1], c2 = indices[e + 2], c3 = indices[e + 3];
      const double x0 = x[(size_t)c0 * C + c], x1 = x[(size_t)c1 * C + c], x2 = x[(size_t)c2 * C + c],
                   x3 = x[(size_t)c3 * C + c];
      acc = fma(v0, x0, acc);
      acc = fma(v1, x1, acc);
      acc = fma(v2, x2, acc);
      acc = fma(v3, x3, acc);
    }
    for (; e < b; ++e) acc = fma(vals[e], x[(size_t)indices[e] * C + c], acc);
    y[(size_t)row * C + c] = acc;
    dot = x[(size_t)row * C + c] * acc;
  }
  // per-component block sums: lanes with equal c sit C apart
#pragma unroll
  for (int o = 16; o >= C; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < C) red[wid][lane] = dot;
  __syncthreads();
  if (threadIdx.x < K) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w][2 * threadIdx.x] + red[w][2 * threadIdx.x + 1];
    part_pq[(size_t)threadIdx.x * nblk + blockIdx.x] = t;
  }
}

__global__ void k_cg_diag(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                          const double* __restrict__ vals, double* __restrict__ idiag) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0;
  for (int e = indptr[r]; e < indptr[r + 1]; ++e)
    if (indices[e] == r) d = vals[e];
  idiag[r] = d != 0.0 ? 1.0 / d : 0.0;
}

// Block-Jacobi weights: z_d = ws[d] r_d + wp[d] r_pair(d).  The two Nedelec functions of an edge (and the two
// interior ones of a triangle) of the second-order element are strongly coupled in the mass matrix; inverting these
// 2 x 2 blocks exactly instead of only their diagonal saves ~20 % of the CG iterations (128 -> 101 on a 22k-dof test).
__global__ void k_cg_block_weights(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                   const double* __restrict__ vals, const double* __restrict__ idiag,
                                   const int* __restrict__ pair, double* __restrict__ ws, double* __restrict__ wp) {
  int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const int p = pair ? pair[d] : -1;
  double b = 0.0;
  if (p >= 0)
    for (int e = indptr[d]; e < indptr[d + 1]; ++e)
      if (indices[e] == p) b = vals[e];
  const double ia = idiag[d];
  if (p < 0 || ia == 0.0 || idiag[p] == 0.0) {
    ws[d] = ia, wp[d] = 0.0;
    return;
  }
  const double a = 1.0 / ia, c = 1.0 / idiag[p];
  const double det = a * c - b * b;
  if (!(det > 1e-14 * a * c)) {
    ws[d] = ia, wp[d] = 0.0;
    return;
  }
  ws[d] = c / det;
  wp[d] = -b / det;
}

// out[r] = sum_b partial[r*nblk + b]  (one block per system)
__global__ void __launch_bounds__(256) k_cg_reduce(const double* __restrict__ partial, int nblk,
                                                   double* __restrict__ out) {
  __shared__ double sh[256];
  const int r = blockIdx.x, t = threadIdx.x;
  double s = 0;
  for (int b = t; b < nblk; b += 256) s += partial[(size_t)r * nblk + b];
  sh[t] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  if (t == 0) out[r] = sh[0];
}
// the two reductions that follow k_cg_step in one launch: blockIdx.y selects (rr, rz)
__global__ void __launch_bounds__(256) k_cg_reduce2(const double* __restrict__ pa, double* __restrict__ oa,
                                                    const double* __restrict__ pb, double* __restrict__ ob, int nblk) {
  __shared__ double sh[256];
  const double* partial = blockIdx.y ? pb : pa;
  double* out = blockIdx.y ? ob : oa;
  const int r = blockIdx.x, t = threadIdx.x;
  double s = 0;
  for (int b = t; b < nblk; b += 256) s += partial[(size_t)r * nblk + b];
  sh[t] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) sh[t] += sh[t + o];
    __syncthreads();
  }
  if (t == 0) out[r] = sh[0];
}

// x += alpha p ; r_new = r - alpha q ; z = Minv_blk r_new ; partials of <r,r> and <r,z>   (alpha_r = rz_r / pq_r)
// INIT: only z and the partials (x = 0, r = b).  Flat, fully coalesced 16-byte accesses over the interleaved
// vectors: element o = i*K + r, and because 256 % K == 0 a thread always meets the same system r.  The residual is
// ping-ponged (rv -> rn) so that the partner entry of the 2 x 2 block can be recomputed from the old vectors.
template <int K, bool INIT>
__global__ void __launch_bounds__(256) k_cg_step(int n, double2* __restrict__ x, const double2* __restrict__ rv,
                                                 double2* __restrict__ rn, const double2* __restrict__ p,
                                                 const double2* __restrict__ q, double2* __restrict__ z,
                                                 const double* __restrict__ ws, const double* __restrict__ wp,
                                                 const int* __restrict__ pair, const double* __restrict__ rz,
                                                 const double* __restrict__ pq, double* __restrict__ part_rr,
                                                 double* __restrict__ part_rz, int nblk) {
  __shared__ double red[2][8][K];
  const int64_t total = (int64_t)n * K;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  const int r = threadIdx.x & (K - 1);
  double alpha = 0.0;
  if (!INIT) {
    const double den = pq[r];
    alpha = (den != 0.0 && isfinite(den)) ? rz[r] / den : 0.0;
  }
  double s_rr = 0.0, s_rz = 0.0;
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    const int64_t o = base + qq * 256 + threadIdx.x;
    if (o < total) {
      const int64_t i = o / K;
      double2 res = rv[o];
      if (!INIT) {
        const double2 xo = x[o], po = p[o], qo = q[o];
        x[o] = make_double2(fma(alpha, po.x, xo.x), fma(alpha, po.y, xo.y));
        res = make_double2(fma(-alpha, qo.x, res.x), fma(-alpha, qo.y, res.y));
        rn[o] = res;
      }
      const double w_s = ws[i];
      double2 zz = make_double2(res.x * w_s, res.y * w_s);
      const int pi = pair ? pair[i] : -1;
      if (pi >= 0) {
        const int64_t op = (int64_t)pi * K + r;
        double2 rp = rv[op];
        if (!INIT) {
          const double2 qp = q[op];
          rp = make_double2(fma(-alpha, qp.x, rp.x), fma(-alpha, qp.y, rp.y));
        }
        const double w_p = wp[i];
        zz = make_double2(fma(w_p, rp.x, zz.x), fma(w_p, rp.y, zz.y));
      }
      z[o] = zz;
      s_rr += res.x * res.x + res.y * res.y;
      s_rz += res.x * zz.x + res.y * zz.y;
    }
  }
#pragma unroll
  for (int o = 16; o >= K; o >>= 1) {
    s_rr += __shfl_xor_sync(0xffffffffu, s_rr, o);
    s_rz += __shfl_xor_sync(0xffffffffu, s_rz, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < K) red[0][wid][lane] = s_rr, red[1][wid][lane] = s_rz;
  __syncthreads();
  if (threadIdx.x < K) {
    double t1 = 0, t2 = 0;
    for (int w = 0; w < 8; ++w) t1 += red[0][w][threadIdx.x], t2 += red[1][w][threadIdx.x];
    part_rr[(size_t)threadIdx.x * nblk + blockIdx.x] = t1;
    part_rz[(size_t)threadIdx.x * nblk + blockIdx.x] = t2;
  }
}

// p = z + (rz_new / rz_old) p
template <int K>
__global__ void __launch_bounds__(256) k_cg_direction(int n, double2* __restrict__ p, const double2* __restrict__ z,
                                                      const double* __restrict__ rz_new,
                                                      const double* __restrict__ rz_old) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (int64_t)n * K) return;
  const int r = threadIdx.x & (K - 1);
  const double den = rz_old[r];
  const double beta = (den != 0.0 && isfinite(den)) ? rz_new[r] / den : 0.0;
  const double2 zo = z[o], po = p[o];
  p[o] = make_double2(fma(beta, po.x, zo.x), fma(beta, po.y, zo.y));
}

// [k][n] <-> interleaved [n][K]
template <int K>
__global__ void k_cg_interleave(int n, const cplx* __restrict__ in, cplx* __restrict__ out, bool to_interleaved,
                                const int* __restrict__ idx, int n_full) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t d = idx ? idx[i] : i;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    if (to_interleaved)
      out[(size_t)i * K + r] = in[(size_t)r * n_full + d];
    else
      out[(size_t)r * n_full + d] = in[(size_t)i * K + r];
  }
}

struct CsrView {
  const int *indptr, *indices;
  const double* vals;
  int64_t nnz;
  const int* pair;  // optional: partner dof of the 2 x 2 blocks of the block-Jacobi preconditioner (-1: none)
  const int* idx;   // optional: row i of this system is dof idx[i] of the right-hand side / solution arrays
  int n_full;       // leading dimension of those arrays
};

// returns false when the iteration does not reach the tolerance (stagnation on badly shaped meshes: the condition
// number of the diagonally scaled Nedelec mass matrix grows with the square of the element aspect ratio) -- the
// caller then falls back to a direct solve; X is left untouched in that case
template <int K>
static bool mass_solve_k(Handle& h, int n, const CsrView& M, const cplx* B, cplx* X) {
  cudaStream_t s = h.stream;
  const int nblk = cdiv((int64_t)n * K, 1024);
  const int nblk_mv = cdiv(n, 256 / (2 * K));
  const size_t vec = (size_t)K * n * 2;
  DevBuf<double> idiag, wself, wpart, part, scal, xb, r, r2, z, p, q;
  idiag.alloc(n), wself.alloc(n), wpart.alloc(n);
  part.alloc((size_t)2 * K * nblk + (size_t)K * nblk_mv);
  scal.alloc((size_t)5 * CG_MAXK);
  xb.alloc(vec), r.alloc(vec), r2.alloc(vec), z.alloc(vec), p.alloc(vec), q.alloc(vec);
  cplx *xv = (cplx*)xb.p, *zv = (cplx*)z.p, *pv = (cplx*)p.p, *qv = (cplx*)q.p;
  cplx* rbuf[2] = {(cplx*)r.p, (cplx*)r2.p};  // residual, ping-ponged
  cplx* rv = rbuf[0];
  double* part_rr = part.p;
  double* part_rz = part.p + (size_t)K * nblk;
  double* part_pq = part.p + (size_t)2 * K * nblk;
  double* d_rr = scal.p;  // [K]
  double* d_rz[2] = {scal.p + CG_MAXK, scal.p + 2 * CG_MAXK};
  double* d_pq = scal.p + 3 * CG_MAXK;
  const int g256 = cdiv(n, 256);
  Trace tr(s);
  tr.mark("    mass PCG: workspace");
  k_cg_diag<<<g256, 256, 0, s>>>(n, M.indptr, M.indices, M.vals, idiag.p);
  k_cg_block_weights<<<g256, 256, 0, s>>>(n, M.indptr, M.indices, M.vals, idiag.p, M.pair, wself.p, wpart.p);
  FW_CHECK_CUDA(cudaMemsetAsync(xv, 0, vec * sizeof(double), s));
  k_cg_interleave<K><<<g256, 256, 0, s>>>(n, B, rv, true, M.idx, M.n_full);
  k_cg_step<K, true><<<nblk, 256, 0, s>>>(n, (double2*)xv, (const double2*)rv, nullptr, (const double2*)pv,
                                          (const double2*)qv, (double2*)zv, wself.p, wpart.p, M.pair, nullptr, nullptr,
                                          part_rr, part_rz, nblk);
  k_cg_reduce2<<<dim3(K, 2), 256, 0, s>>>(part_rr, d_rr, part_rz, d_rz[0], nblk);
  FW_CHECK_CUDA(cudaMemcpyAsync(pv, zv, vec * sizeof(double), cudaMemcpyDeviceToDevice, s));
  double bnorm2[CG_MAXK], rr[CG_MAXK];
  FW_CHECK_CUDA(cudaMemcpyAsync(bnorm2, d_rr, sizeof(double) * K, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < K; ++i)
    if (!std::isfinite(bnorm2[i])) throw Error(FW_ERR_INVALID, "mass PCG (H field): non-finite right-hand side");
  tr.mark("    mass PCG: init");
  int64_t launches = 6;
  int cur = 0, it = 0;
  const int check_every = 4;
  const int max_it = 600;
  // ||r|| <= tol ||b||.  The reference's H carries complex64 rounding (1e-7, waveguide.py:658,666) and the north star
  // asks 1e-6 on overlaps; the default 1e-10 keeps H four digits beyond that (FW_OPT_PCG_TOL_EXP).
  const double tol2 = h.pcg_tol * h.pcg_tol;
  bool converged = false;
  double worst = 0.0;
  auto check = [&]() {
    FW_CHECK_CUDA(cudaMemcpyAsync(rr, d_rr, sizeof(double) * K, cudaMemcpyDeviceToHost, s));
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    bool done = true;
    worst = 0.0;
    for (int i = 0; i < K; ++i) {
      if (!(rr[i] <= tol2 * bnorm2[i])) done = false;
      const double rel = bnorm2[i] > 0 ? std::sqrt(rr[i] / bnorm2[i]) : (rr[i] > 0 ? INFINITY : 0.0);
      if (!(rel <= worst)) worst = rel;  // NaN propagates
    }
    return done;
  };
  for (; it < max_it; ++it) {
    k_spmv_multi<K><<<nblk_mv, 256, 0, s>>>(n, M.indptr, M.indices, M.vals, (const double*)pv, (double*)qv, part_pq,
                                            nblk_mv);
    k_cg_reduce<<<K, 256, 0, s>>>(part_pq, nblk_mv, d_pq);
    k_cg_step<K, false><<<nblk, 256, 0, s>>>(n, (double2*)xv, (const double2*)rbuf[cur], (double2*)rbuf[cur ^ 1],
                                             (const double2*)pv, (const double2*)qv, (double2*)zv, wself.p, wpart.p,
                                             M.pair, d_rz[cur], d_pq, part_rr, part_rz, nblk);
    k_cg_reduce2<<<dim3(K, 2), 256, 0, s>>>(part_rr, d_rr, part_rz, d_rz[cur ^ 1], nblk);
    k_cg_direction<K><<<cdiv((int64_t)n * K, 256), 256, 0, s>>>(n, (double2*)pv, (const double2*)zv, d_rz[cur ^ 1], d_rz[cur]);
    cur ^= 1;
    launches += 5;
    if ((it + 1) % check_every == 0) {
      if (check()) {
        converged = true;
        ++it;
        break;
      }
      // give up early when the observed rate cannot reach the tolerance within max_it iterations
      if (it + 1 >= 48 && worst > 0 && worst < 1.0) {
        const double need = (it + 1) * std::log(h.pcg_tol) / std::log(worst);
        if (need > max_it) {
          ++it;
          break;
        }
      } else if (it + 1 >= 48 && !(worst < 1.0)) {
        ++it;
        break;
      }
    }
  }
  if (!converged) converged = check();  // the loop may end between two checks (or without any iteration)
  tr.mark("    mass PCG: iterations");
  h.tim.pcg_iterations = std::max<int64_t>(h.tim.pcg_iterations, it);
  if (converged && !(worst <= h.tim.pcg_rel_residual)) h.tim.pcg_rel_residual = worst;
  if (!converged) {
    if (getenv("FW_TRACE"))
      fprintf(stderr, "[fw-trace]   mass PCG: gave up after %d iterations at rel. residual %.2e\n", it, worst);
    h.tim.kernel_launches += launches;
    return false;
  }
  k_cg_interleave<K><<<g256, 256, 0, s>>>(n, xv, X, false, M.idx, M.n_full);
  ++launches;
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += launches;
  if (getenv("FW_TRACE"))
    fprintf(stderr, "[fw-trace]   mass PCG: %d rhs, %d iterations, rel. residual %.2e, %lld kernel launches, mass nnz %lld\n",
            K, it, worst, (long long)launches, (long long)M.nnz);
  return true;
}

// M X = B for a real SPD CSR matrix and k <= CG_MAXK complex right-hand sides ([k][n] each); X overwritten.
// Returns false when the PCG does not converge (X then holds nothing useful for the rows of this system).
static bool pcg_solve_view(Handle& h, int n_rows, const CsrView& V, int k, const cplx* B, cplx* X);

[[noreturn]] static void pcg_failed(const Handle& h, const char* what) {
  throw Error(FW_ERR_NOCONV, std::string(what) + ": the mass PCG did not converge (tolerance " + std::to_string(h.pcg_tol) + ")");
}

// ---- direct fallback for one diagonal block of the mass matrix (Nedelec or Lagrange dofs): the multifrontal LU of
// the eigen-solve factors the block (the dofs of the other block are "inactive") -- what the reference does for every
// mode (scipy spsolve, waveguide.py:671).  Used when the PCG stagnates: on graded meshes with elongated triangles the
// diagonally scaled Nedelec mass matrix is ill conditioned (kappa ~ aspect ratio squared).
__global__ void __launch_bounds__(256) k_split_re_im(int n, const cplx* __restrict__ b, double* __restrict__ out2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cplx v = b[i];
  out2[i] = v.re;
  out2[(size_t)n + i] = v.im;
}
// X[d] (+)= (x2[0][d], x2[1][d]) for the dofs of the block
__global__ void __launch_bounds__(256) k_merge_re_im(int n, const double* __restrict__ x2, const uint8_t* __restrict__ is_z,
                                                     int block_z, bool add, cplx* __restrict__ X) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (is_z[i] != 0) != (block_z != 0)) return;
  cplx v{x2[i], x2[(size_t)n + i]};
  if (add) v = v + X[i];
  X[i] = v;
}
// r = b - M x on the rows of the block (M = full mass matrix on the structural pattern: the cross blocks are zeros)
__global__ void __launch_bounds__(256) k_mass_residual(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                                       const double* __restrict__ vals, const cplx* __restrict__ x,
                                                       const cplx* __restrict__ b, const uint8_t* __restrict__ is_z,
                                                       int block_z, cplx* __restrict__ r, double* __restrict__ norms) {
  __shared__ double red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double rr = 0.0, bb = 0.0;
  if (i < n) {
    cplx res{0.0, 0.0};
    if ((is_z[i] != 0) == (block_z != 0)) {
      cplx acc{0.0, 0.0};
      for (int e = indptr[i]; e < indptr[i + 1]; ++e) fma_(acc, x[indices[e]], vals[e]);
      res = b[i] - acc;
      rr = abs2_(res), bb = abs2_(b[i]);
    }
    r[i] = res;
  }
  rr = block_sum_256(rr, red);
  bb = block_sum_256(bb, red);
  if (threadIdx.x == 0) {
    atomicAdd(&norms[0], rr);
    atomicAdd(&norms[1], bb);
  }
}

static void mass_solve_direct(Handle& h, bool block_z, int k, const cplx* B, cplx* X) {
  const Space& sp = h.sp;
  const int n = sp.n;
  cudaStream_t s = h.stream;
  FW_REQUIRE((int)sp.h_is_z.size() == n, "space missing");
  Trace tr(s);
  std::vector<uint8_t> active((size_t)n);
  for (int d = 0; d < n; ++d) active[d] = (sp.h_is_z[d] != 0) == block_z ? 1 : 0;
  lu_analyse(h, active, 0);
  lu_factor_values(h, h.Mass, h.Mass, 1.0, 0.0, 0.0, 0.0, false);
  tr.mark("    mass direct: analysis + factorisation of the block");
  const Pattern& pat = h.pat;
  DevBuf<double> b2, x2, rbuf, nrm;
  b2.alloc((size_t)2 * n), x2.alloc((size_t)2 * n), rbuf.alloc((size_t)2 * n), nrm.alloc(2);
  const int g = cdiv(n, 256);
  for (int r = 0; r < k; ++r) {
    const cplx* b = B + (size_t)r * n;
    cplx* x = X + (size_t)r * n;
    k_split_re_im<<<g, 256, 0, s>>>(n, b, b2.p);
    lu_solve_dev(h, b2.p, x2.p, 2);
    k_merge_re_im<<<g, 256, 0, s>>>(n, x2.p, sp.is_z.p, block_z, false, x);
    double rel = 0.0;
    for (int it = 0; it < 3; ++it) {  // residual check + iterative refinement
      nrm.zero(s);
      k_mass_residual<<<g, 256, 0, s>>>(n, pat.indptr.p, pat.indices.p, h.Mass.v.p, x, b, sp.is_z.p, block_z,
                                        (cplx*)rbuf.p, nrm.p);
      double nn[2];
      nrm.download(nn, 2, s);
      FW_CHECK_CUDA(cudaStreamSynchronize(s));
      rel = nn[1] > 0 ? std::sqrt(nn[0] / nn[1]) : std::sqrt(nn[0]);
      if (!std::isfinite(rel)) throw Error(FW_ERR_INVALID, "mass solve (H field): non-finite right-hand side");
      if (rel <= h.pcg_tol || it == 2) break;
      k_split_re_im<<<g, 256, 0, s>>>(n, (const cplx*)rbuf.p, b2.p);
      lu_solve_dev(h, b2.p, x2.p, 2);
      k_merge_re_im<<<g, 256, 0, s>>>(n, x2.p, sp.is_z.p, block_z, true, x);
    }
    if (!(rel <= std::max(h.pcg_tol, 1e-9)))
      throw Error(FW_ERR_NOCONV, "mass solve (H field): the direct fallback misses the tolerance, ||r||/||b|| = " +
                                     std::to_string(rel));
    if (!(rel <= h.tim.pcg_rel_residual)) h.tim.pcg_rel_residual = rel;
  }
  h.lu->factored = false;  // the arena no longer holds the factors of the shifted pencil
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += 6 * k;
  tr.mark("    mass direct: solves");
}

void mass_solve_multi(Handle& h, int k, const cplx* B, cplx* X) {
  FW_REQUIRE(h.Mass.valid && h.MassC.valid, "mass matrix missing");
  const bool force_direct = h.mass_solver == 2, pcg_only = h.mass_solver == 1;
  if (h.mass_split) {
    // Nedelec and Lagrange blocks as two independent systems (every dof belongs to exactly one of them)
    int blk = 0;
    for (const Handle::MassBlock* b : {&h.MassN, &h.MassP}) {
      const CsrView V{b->indptr.p, b->indices.p, b->vals.p, b->nnz, b->pair.n ? b->pair.p : nullptr, b->idx.p, h.sp.n};
      if (force_direct || !pcg_solve_view(h, b->n, V, k, B, X)) {
        if (pcg_only) pcg_failed(h, "H field");
        mass_solve_direct(h, blk == 1, k, B, X);
        h.tim.mass_direct_blocks += 1;
      }
      ++blk;
    }
    return;
  }
  const CsrView V{h.MassC.indptr.p, h.MassC.indices.p, h.MassC.vals.p, h.MassC.nnz,
                  h.MassC.pair.n ? h.MassC.pair.p : nullptr, nullptr, h.sp.n};
  if (!pcg_solve_view(h, h.sp.n, V, k, B, X)) pcg_failed(h, "H field");
}

void pcg_solve_multi(Handle& h, int n_rows, const int* indptr, const int* indices, const double* vals, int64_t nnz,
                     int k, const cplx* B, cplx* X, const int* pair) {
  const CsrView V{indptr, indices, vals, nnz, pair, nullptr, n_rows};
  if (!pcg_solve_view(h, n_rows, V, k, B, X)) pcg_failed(h, "P1 projection");
}

// ---- real packing.  With a real permittivity the modes come out of the real-arithmetic Arnoldi as real vectors, so
// every right-hand side C(beta) e restricted to one block of the mass matrix is purely real (Lagrange rows) or purely
// imaginary (Nedelec rows).  Two such real systems are then solved as ONE complex system M (x_a + i x_b) = b_a + i b_b
// (the matrix is real): half the vector traffic of the PCG.  The structure is detected on the device, never assumed.
struct CompSel {
  int comp[CG_MAXK];  // 0: the real part carries the data, 1: the imaginary part
};
// out[2r] = max |Re b_r|, out[2r+1] = max |Im b_r| over the rows of the block (non-negative doubles order like their
// bit patterns)
__global__ void __launch_bounds__(256) k_cg_component_max(int n, int k, const cplx* __restrict__ B,
                                                          const int* __restrict__ idx, int n_full,
                                                          unsigned long long* __restrict__ out) {
  for (int r = 0; r < k; ++r) {
    double mre = 0.0, mim = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const cplx v = B[(size_t)r * n_full + (idx ? idx[i] : i)];
      const double a = fabs(v.re), b = fabs(v.im);
      mre = a > mre || a != a ? a : mre;  // a NaN sticks (and counts as "non-zero")
      mim = b > mim || b != b ? b : mim;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double a = __shfl_xor_sync(0xffffffffu, mre, o), b = __shfl_xor_sync(0xffffffffu, mim, o);
      mre = a > mre || a != a ? a : mre;
      mim = b > mim || b != b ? b : mim;
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMax(&out[2 * r], (unsigned long long)__double_as_longlong(mre));
      atomicMax(&out[2 * r + 1], (unsigned long long)__double_as_longlong(mim));
    }
  }
}
// Bp[j][d] = (b'_{2j}[d], b'_{2j+1}[d]) with b'_r the data-carrying component of b_r (0 beyond k)
__global__ void __launch_bounds__(256) k_cg_pack_real(int n, int k, const cplx* __restrict__ B,
                                                      const int* __restrict__ idx, int n_full, CompSel sel,
                                                      cplx* __restrict__ Bp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t d = idx ? idx[i] : i;
  for (int j = 0; 2 * j < k; ++j) {
    const cplx a = B[(size_t)(2 * j) * n_full + d];
    cplx o{sel.comp[2 * j] ? a.im : a.re, 0.0};
    if (2 * j + 1 < k) {
      const cplx b = B[(size_t)(2 * j + 1) * n_full + d];
      o.im = sel.comp[2 * j + 1] ? b.im : b.re;
    }
    Bp[(size_t)j * n_full + d] = o;
  }
}
__global__ void __launch_bounds__(256) k_cg_unpack_real(int n, int k, const cplx* __restrict__ Xp,
                                                        const int* __restrict__ idx, int n_full, CompSel sel,
                                                        cplx* __restrict__ X) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t d = idx ? idx[i] : i;
  for (int r = 0; r < k; ++r) {
    const cplx v = Xp[(size_t)(r >> 1) * n_full + d];
    const double x = (r & 1) ? v.im : v.re;
    X[(size_t)r * n_full + d] = sel.comp[r] ? cplx{0.0, x} : cplx{x, 0.0};
  }
}

static bool pcg_solve_groups(Handle& h, int n_rows, const CsrView& V, int k, const cplx* B, cplx* X);

static bool pcg_solve_view(Handle& h, int n_rows, const CsrView& V, int k, const cplx* B, cplx* X) {
  FW_REQUIRE(k >= 1 && k <= CG_MAXK, "pcg_solve_multi: 1 <= k <= 8");
  static const bool no_pack = getenv("FW_PCG_NO_PACK") != nullptr;
  if (k >= 2 && !no_pack) {
    cudaStream_t s = h.stream;
    DevBuf<double> cm;
    cm.alloc(2 * CG_MAXK);
    cm.zero(s);
    k_cg_component_max<<<296, 256, 0, s>>>(n_rows, k, B, V.idx, V.n_full, (unsigned long long*)cm.p);
    double mx[2 * CG_MAXK];
    cm.download(mx, 2 * CG_MAXK, s);
    FW_CHECK_CUDA(cudaStreamSynchronize(s));
    h.tim.kernel_launches += 1;
    CompSel sel{};
    bool packable = true;
    for (int r = 0; r < k; ++r) {
      if (mx[2 * r + 1] == 0.0)
        sel.comp[r] = 0;
      else if (mx[2 * r] == 0.0)
        sel.comp[r] = 1;
      else
        packable = false;  // genuinely complex (or non-finite) right-hand side
    }
    if (packable) {
      const int k2 = (k + 1) / 2;
      DevBuf<double> bp, xp;
      bp.alloc((size_t)2 * V.n_full * k2);
      xp.alloc((size_t)2 * V.n_full * k2);
      const int g = cdiv(n_rows, 256);
      k_cg_pack_real<<<g, 256, 0, s>>>(n_rows, k, B, V.idx, V.n_full, sel, (cplx*)bp.p);
      if (!pcg_solve_groups(h, n_rows, V, k2, (const cplx*)bp.p, (cplx*)xp.p)) return false;
      k_cg_unpack_real<<<g, 256, 0, s>>>(n_rows, k, (const cplx*)xp.p, V.idx, V.n_full, sel, X);
      FW_CHECK_CUDA(cudaGetLastError());
      h.tim.kernel_launches += 2;
      return true;
    }
  }
  return pcg_solve_groups(h, n_rows, V, k, B, X);
}

static bool pcg_solve_groups(Handle& h, int n_rows, const CsrView& V, int k, const cplx* B, cplx* X) {
  const size_t n = (size_t)V.n_full;  // stride between the systems in B and X
  // systems are solved in groups of 4 / 2 / 1 (compile-time interleave factor)
  int done = 0;
  while (done < k) {
    const int left = k - done;
    bool ok;
    if (left >= 4) {
      ok = mass_solve_k<4>(h, n_rows, V, B + done * n, X + done * n);
      done += 4;
    } else if (left >= 2) {
      ok = mass_solve_k<2>(h, n_rows, V, B + done * n, X + done * n);
      done += 2;
    } else {
      ok = mass_solve_k<1>(h, n_rows, V, B + done * n, X + done * n);
      done += 1;
    }
    if (!ok) return false;
  }
  return true;
}

}  // namespace fw
