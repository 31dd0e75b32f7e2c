// Selective inversion for the upper levels of the multifrontal tree (f64).
//
// A triangular solve with the ns x ns pivot block of a front is a chain of ns/32 dependent steps; on the levels with
// few, large fronts these chains -- not the bytes -- set the time of a substitution pass (config 2: 1.05 of 5.1 ms per
// forward/backward pair in 14 chain kernels that read 3 % of the factor).  Here the pivot blocks of those fronts are
// inverted once per factorisation, inv(L11) (unit lower) and inv(U11), and the chain becomes a dense triangular
// matrix-vector product: every row is independent, the launch is bound by bandwidth instead of latency.  The
// off-diagonal blocks L21 / U12 are NOT multiplied by the inverses (selective inversion in the sense of Raghavan,
// "Efficient parallel sparse triangular solution using selective inversion", 1998): their updates keep using the
// factors, only the pivot-block solves change.
//
// Inversion by recursive doubling on the 32 x 32 diagonal blocks:
//   inv [A 0; C B] = [inv A, 0; -inv(B) C inv(A), inv B]      (lower; the upper case is the mirror image)
// two batched GEMMs per triangle and doubling level, all fronts of a level and all block pairs in one launch.
// Replaces nothing in the reference: scipy's SuperLU solves by substitution (arpack.py:1011-1019); this is an
// implementation detail of the hand-written sparse LU (lu.h).
#include "lu_dev.h"

namespace fw {

// Levels that carry inverse pivot blocks: pivot blocks of at least tinv_min_level_ns rows, and never the three deepest
// levels of the tree -- the pivot blocks of the leaf fronts (interior dofs of a few elements of the shifted, indefinite
// pencil) are too ill conditioned for explicit inverses: with them the solve residual rose from 9e-13 to 5e-10.
bool tinv_level_static(const LUSymbolic& S, int min_ns, int lev) {
  if (getenv("FW_TINV_ALL_LEVELS")) return S.level_max_ns[lev] >= min_ns;  // tests: exercise the accuracy fallback
  return S.level_max_ns[lev] >= min_ns && lev < S.n_levels - 3;
}
bool tinv_level(const LUSymbolic& S, const LUDevice& d, int lev) {
  return d.tinv_valid && tinv_level_static(S, d.tinv_min_level_ns, lev);
}

// ------------------------------------------------------------------------------ 32 x 32 diagonal blocks
__global__ void __launch_bounds__(32) k_tinv_diag(FrontTab T, const int* __restrict__ nodes,
                                                  const double* __restrict__ F, double* __restrict__ tinv) {
  const int node = nodes[blockIdx.x];
  const int ns = T.ns[node], m = T.m[node];
  const int k0 = blockIdx.y * PW;
  const long long off = T.tinv_off[node];
  if (k0 >= ns || off < 0) return;
  const int wd = min(PW, ns - k0);
  const double* Fn = F + T.front_off[node];
  __shared__ double D[PW][PW + 1];
  __shared__ double R[PW];
  const int lane = threadIdx.x;
  for (int c = 0; c < PW; ++c)
    D[lane][c] = (lane < wd && c < wd) ? Fn[(k0 + lane) + (long long)(k0 + c) * m] : (lane == c ? 1.0 : 0.0);
  __syncwarp();
  R[lane] = 1.0 / D[lane][lane];
  __syncwarp();
  const int c = lane;
  double x[PW];
  // column c of inv(L): x_c = 1, x_i = -sum_{j<i} L_ij x_j
#pragma unroll
  for (int i = 0; i < PW; ++i) {
    double sacc = 0.0;
#pragma unroll
    for (int j = 0; j < i; ++j) sacc = fma(-D[i][j], x[j], sacc);
    x[i] = i < c ? 0.0 : (i == c ? 1.0 : sacc);
  }
  double* Li = tinv + off;
  double* Ui = Li + (long long)ns * ns;
  if (c < wd) {
#pragma unroll
    for (int i = 0; i < PW; ++i)
      if (i < wd) Li[(k0 + i) + (long long)(k0 + c) * ns] = x[i];
  }
  // column c of inv(U): y_c = 1/U_cc, y_i = -(sum_{j>i} U_ij y_j) / U_ii
#pragma unroll
  for (int i = PW - 1; i >= 0; --i) {
    double sacc = 0.0;
#pragma unroll
    for (int j = i + 1; j < PW; ++j) sacc = fma(-D[i][j], x[j], sacc);
    x[i] = i > c ? 0.0 : (i == c ? R[i] : sacc * R[i]);
  }
  if (c < wd) {
#pragma unroll
    for (int i = 0; i < PW; ++i)
      if (i < wd) Ui[(k0 + i) + (long long)(k0 + c) * ns] = x[i];
  }
}

// ------------------------------------------------------------------------------ doubling step
// block pair q of a front: A = [2 q s, 2 q s + s), B = [2 q s + s, min(ns, 2 q s + 2 s))
//   MODE 0: Tmp[B,A]  =  L11[B,A]  inv(L)[A,A]        MODE 1: inv(L)[B,A] = - inv(L)[B,B] Tmp[B,A]
//   MODE 2: Tmp[A,B]  =  U11[A,B]  inv(U)[B,B]        MODE 3: inv(U)[A,B] = - inv(U)[A,A] Tmp[A,B]
// FP64 tensor-core GEMM (mma.sync m8n8k4 = DMMA, the mainloop of lu_gemm.cu on generic operands): 64 x 64 tiles, 128
// threads = 2 x 2 warps of 32 x 32, k-chunks of 16 through a 3-stage cp.async pipeline.  One operand of every product
// is triangular: the k-range of a tile is cut to the chunks where that operand is non-zero (half of the flops).
constexpr int TG_TM = 64, TG_KC = 16, TG_LDA = TG_TM + 4, TG_LDB = TG_KC + 4, TG_STAGES = 3, TG_THREADS = 128;
constexpr int TG_STAGE_DOUBLES = TG_KC * TG_LDA + TG_TM * TG_LDB;

__device__ __forceinline__ void tg_cp_async8_zfill(double* smem, const double* gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void tg_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(TG_THREADS, 3) k_tinv_gemm(FrontTab T, const int* __restrict__ nodes, int s,
                                                             int npairs, const double* __restrict__ F,
                                                             double* __restrict__ tinv, double* __restrict__ tmp) {
  extern __shared__ __align__(16) double tg_smem[];
  const int node = nodes[blockIdx.x / npairs];
  const int q = blockIdx.x % npairs;
  const int ns = T.ns[node], m = T.m[node];
  const long long off = T.tinv_off[node];
  if (off < 0) return;
  const int a0 = 2 * q * s, b0 = a0 + s;
  if (b0 >= ns) return;
  const int na = s, nbk = min(ns, b0 + s) - b0;
  const double* Fn = F + T.front_off[node];
  double* Li = tinv + off;
  double* Ui = Li + (long long)ns * ns;
  double* Tp = tmp + off / 2;
  int M, N, K;
  const double *X, *Y;
  double* C;
  long long ldx, ldy;
  const long long ldc = ns;
  if (MODE == 0) {
    M = nbk, N = na, K = na;
    X = Fn + b0 + (long long)a0 * m, ldx = m;
    Y = Li + a0 + (long long)a0 * ns, ldy = ns;
    C = Tp + b0 + (long long)a0 * ns;
  } else if (MODE == 1) {
    M = nbk, N = na, K = nbk;
    X = Li + b0 + (long long)b0 * ns, ldx = ns;
    Y = Tp + b0 + (long long)a0 * ns, ldy = ns;
    C = Li + b0 + (long long)a0 * ns;
  } else if (MODE == 2) {
    M = na, N = nbk, K = nbk;
    X = Fn + a0 + (long long)b0 * m, ldx = m;
    Y = Ui + b0 + (long long)b0 * ns, ldy = ns;
    C = Tp + a0 + (long long)b0 * ns;
  } else {
    M = na, N = nbk, K = na;
    X = Ui + a0 + (long long)a0 * ns, ldx = ns;
    Y = Tp + a0 + (long long)b0 * ns, ldy = ns;
    C = Ui + a0 + (long long)b0 * ns;
  }
  const int rbase = blockIdx.y * TG_TM, cbase = blockIdx.z * TG_TM;
  if (rbase >= M || cbase >= N) return;
  // non-zero k-range of the triangular operand for this tile
  int kbeg = 0, kend = K;
  if (MODE == 0) kbeg = cbase;                     // Y lower triangular: Y[k, c] = 0 for k < c
  if (MODE == 1) kend = min(K, rbase + TG_TM);     // X lower triangular: X[r, k] = 0 for k > r
  if (MODE == 2) kend = min(K, cbase + TG_TM);     // Y upper triangular: Y[k, c] = 0 for k > c
  if (MODE == 3) kbeg = rbase;                     // X upper triangular: X[r, k] = 0 for k < r
  kbeg = (kbeg / TG_KC) * TG_KC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, qq = lane & 3;

  auto issue = [&](int chunk, int stage) {
    double* As = tg_smem + (size_t)stage * TG_STAGE_DOUBLES;
    double* Bs = As + TG_KC * TG_LDA;
    const int kc = kbeg + chunk * TG_KC;
#pragma unroll
    for (int it = 0; it < TG_KC * TG_TM / TG_THREADS; ++it) {
      const int idx = tid + it * TG_THREADS;
      {
        const int r = idx % TG_TM, k = idx / TG_TM;
        const bool ok = rbase + r < M && kc + k < kend;
        tg_cp_async8_zfill(As + k * TG_LDA + r, ok ? X + (rbase + r) + (long long)(kc + k) * ldx : X, ok);
      }
      {
        const int k = idx % TG_KC, c = idx / TG_KC;
        const bool ok = cbase + c < N && kc + k < kend;
        tg_cp_async8_zfill(Bs + c * TG_LDB + k, ok ? Y + (kc + k) + (long long)(cbase + c) * ldy : Y, ok);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nchunk = max(0, (kend - kbeg + TG_KC - 1) / TG_KC);
  if (nchunk > 0) issue(0, 0);
  if (nchunk > 1)
    issue(1, 1);
  else
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  for (int c = 0; c < nchunk; ++c) {
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    if (c + 2 < nchunk)
      issue(c + 2, (c + 2) % TG_STAGES);
    else
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    const double* As = tg_smem + (size_t)(c % TG_STAGES) * TG_STAGE_DOUBLES;
    const double* Bs = As + TG_KC * TG_LDA;
#pragma unroll
    for (int kk = 0; kk < TG_KC / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(kk * 4 + qq) * TG_LDA + wm * 32 + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[(wn * 32 + j * 8 + g) * TG_LDB + kk * 4 + qq];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) tg_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  constexpr double sign = (MODE == 1 || MODE == 3) ? -1.0 : 1.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gc = cbase + wn * 32 + j * 8 + 2 * qq + h;
      if (gc >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gr = rbase + wm * 32 + i * 8 + g;
        if (gr < M) C[gr + (long long)gc * ldc] = sign * acc[i][j][h];
      }
    }
  }
}

void tinv_build(Handle& h, LUDevice& d, const LUSymbolic& S) {
  cudaStream_t s = h.stream;
  d.tinv.alloc((size_t)d.tinv_entries);
  d.tinv_tmp.alloc((size_t)d.tinv_entries / 2 + 1);
  d.tinv.zero(s, (size_t)d.tinv_entries);
  const double* F = d.factors.p;
  FrontTab T = d.tab();
  int64_t launches = 0;
  constexpr size_t gemm_smem = sizeof(double) * TG_STAGES * TG_STAGE_DOUBLES;
  {
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
      FW_CHECK_CUDA(cudaFuncSetAttribute(k_tinv_gemm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
      FW_CHECK_CUDA(cudaFuncSetAttribute(k_tinv_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
      FW_CHECK_CUDA(cudaFuncSetAttribute(k_tinv_gemm<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
      FW_CHECK_CUDA(cudaFuncSetAttribute(k_tinv_gemm<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
    }
  }
  for (int lev = 0; lev < S.n_levels; ++lev) {
    const int max_ns = S.level_max_ns[lev];
    if (!tinv_level_static(S, d.tinv_min_level_ns, lev)) continue;
    const int nfront = S.level_start[lev + 1] - S.level_start[lev];
    const int* nodes = d.level_nodes.p + S.level_start[lev];
    const int npan = (max_ns + PW - 1) / PW;
    k_tinv_diag<<<dim3(nfront, npan), 32, 0, s>>>(T, nodes, F, d.tinv.p);
    ++launches;
    for (int bs = PW; bs < max_ns; bs *= 2) {
      const int npairs = (max_ns + 2 * bs - 1) / (2 * bs);
      const int nt = (bs + 63) / 64;
      const dim3 grid((unsigned)(nfront * npairs), nt, nt);
      k_tinv_gemm<0><<<grid, TG_THREADS, gemm_smem, s>>>(T, nodes, bs, npairs, F, d.tinv.p, d.tinv_tmp.p);
      k_tinv_gemm<1><<<grid, TG_THREADS, gemm_smem, s>>>(T, nodes, bs, npairs, F, d.tinv.p, d.tinv_tmp.p);
      k_tinv_gemm<2><<<grid, TG_THREADS, gemm_smem, s>>>(T, nodes, bs, npairs, F, d.tinv.p, d.tinv_tmp.p);
      k_tinv_gemm<3><<<grid, TG_THREADS, gemm_smem, s>>>(T, nodes, bs, npairs, F, d.tinv.p, d.tinv_tmp.p);
      launches += 4;
    }
  }
  FW_CHECK_CUDA(cudaGetLastError());
  h.tim.kernel_launches += launches;
  d.tinv_valid = true;
}

// ------------------------------------------------------------------------------ substitution with the inverses
// forward, step 1: assemble the front's right-hand side (own entries from xp, boundary entries zero, plus the
// children's boundary contributions) and stage the own part in xp; the row interchanges are applied by the
// matrix-vector kernel when it gathers its input
template <int NRHS>
__global__ void __launch_bounds__(1024, 1) k_tinv_fwd_prep(FrontTab T, const int* __restrict__ nodes, double* work,
                                                           int64_t work_stride, double* xp, int na) {
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node], os = T.own_start[node];
  double* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x, nth = blockDim.x;
  for (int i = t; i < m; i += nth)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) w[r][i] = i < ns ? xp[(size_t)r * na + os + i] : 0.0;
  __syncthreads();
  for (int sl = 0; sl < 2; ++sl) {
    const int c = sl ? T.child1[node] : T.child0[node];
    if (c < 0) continue;
    const int nsc = T.ns[c], nbc = T.m[c] - nsc;
    const int* rel = T.rel + T.blist_off[c];
    // four entries per thread and pass, all loads before the first store (the targets of one child are distinct, but
    // the compiler cannot know that: written as a plain loop every read-modify-write waits for the previous one)
    for (int i0 = t; i0 < nbc; i0 += 4 * nth) {
      int ri[4];
      double v[NRHS][4], o[NRHS][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nth;
        ri[u] = i < nbc ? rel[i] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nth;
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
          v[r][u] = ri[u] >= 0 ? work[(size_t)r * work_stride + T.work_off[c] + nsc + i] : 0.0;
          o[r][u] = ri[u] >= 0 ? w[r][ri[u]] : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
          if (ri[u] >= 0) w[r][ri[u]] = o[r][u] + v[r][u];
    }
    __syncthreads();
  }
  for (int i = t; i < ns; i += nth)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) xp[(size_t)r * na + os + i] = w[r][i];
}

// x = inv(L11) P b (FWD: input gathered from xp through rowperm, output w[0:ns)) or x = inv(U11) y (input w[0:ns),
// output xp).  CTA = 32 rows x 16 threads per row on interleaved columns, 8 independent loads in flight per thread
// (64 rows x 4 threads left the 2001-row root front on 32 SMs with 16 KB in flight each: 160 us).
constexpr int TV_ROWS = 32, TV_G = 16;
template <int NRHS, bool FWD>
__global__ void __launch_bounds__(TV_ROWS* TV_G, 2) k_tinv_matvec(FrontTab T, const int* __restrict__ nodes,
                                                               const double* __restrict__ tinv,
                                                               const int* __restrict__ rowperm, double* work,
                                                               int64_t work_stride, double* xp, int na, int ld_b) {
  extern __shared__ double tv_b[];  // [NRHS][ld_b]
  __shared__ double part[NRHS][TV_G][TV_ROWS];
  const int node = nodes[blockIdx.x];
  const int ns = T.ns[node], os = T.own_start[node];
  const long long off = T.tinv_off[node];
  const int R0 = blockIdx.y * TV_ROWS;
  if (R0 >= ns || off < 0) return;
  const double* M = tinv + off + (FWD ? 0 : (long long)ns * ns);
  double* w[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) w[r] = work + (size_t)r * work_stride + T.work_off[node];
  const int t = threadIdx.x;
  const int jlo = FWD ? 0 : R0, jhi = FWD ? min(ns, R0 + TV_ROWS) : ns;  // columns this tile touches
  for (int j = jlo + t; j < jhi; j += TV_ROWS * TV_G) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r)
      tv_b[r * ld_b + j] = FWD ? xp[(size_t)r * na + os + rowperm[os + j]] : w[r][j];
  }
  __syncthreads();
  const int lr = t % TV_ROWS, g = t / TV_ROWS;
  const int row = R0 + lr;
  const bool valid = row < ns;
  double acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
  if (valid) {
    // this row's non-zero columns: [0, row] (lower) or [row, ns) (upper)
    const int c0 = FWD ? 0 : row, c1 = FWD ? row + 1 : ns;
    const double* Mr = M + row;
    int j = c0 + g;
    for (; j + 7 * TV_G < c1; j += 8 * TV_G) {
      double l[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) l[u] = Mr[(long long)(j + u * TV_G) * ns];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int r = 0; r < NRHS; ++r) acc[r] = fma(l[u], tv_b[r * ld_b + j + u * TV_G], acc[r]);
    }
    for (; j < c1; j += TV_G) {
      const double l = Mr[(long long)j * ns];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) acc[r] = fma(l, tv_b[r * ld_b + j], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < NRHS; ++r) part[r][g][lr] = acc[r];
  __syncthreads();
  if (g == 0 && valid) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
      double x = part[r][0][lr];
#pragma unroll
      for (int q = 1; q < TV_G; ++q) x += part[r][q][lr];
      if (FWD)
        w[r][row] = x;
      else
        xp[(size_t)r * na + os + row] = x;
    }
  }
}

// backward, last step: the children gather x from the front's work vector
template <int NRHS>
__global__ void __launch_bounds__(256) k_tinv_bwd_copy(FrontTab T, const int* __restrict__ nodes, double* work,
                                                       int64_t work_stride, const double* __restrict__ xp, int na) {
  const int node = nodes[blockIdx.x];
  const int ns = T.ns[node], os = T.own_start[node];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < ns; i += gridDim.y * blockDim.x)
#pragma unroll
    for (int r = 0; r < NRHS; ++r) work[(size_t)r * work_stride + T.work_off[node] + i] = xp[(size_t)r * na + os + i];
}

template <int NRHS, bool FWD>
static void launch_matvec(LUDevice& d, const LUSymbolic& S, int lev, double* work, double* xp, int na, cudaStream_t s) {
  const int nfront = S.level_start[lev + 1] - S.level_start[lev];
  const int* nodes = d.level_nodes.p + S.level_start[lev];
  const int max_ns = S.level_max_ns[lev];
  const int ld_b = max_ns + 1;
  const size_t smem = sizeof(double) * NRHS * ld_b;
  static bool configured[64] = {};
  if (first_use_on_device(configured))
    FW_CHECK_CUDA(cudaFuncSetAttribute(k_tinv_matvec<NRHS, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  FW_REQUIRE(smem <= 160 * 1024, "pivot block too large for the inverse-based substitution");
  FrontTab T = d.tab();
  k_tinv_matvec<NRHS, FWD><<<dim3(nfront, (max_ns + TV_ROWS - 1) / TV_ROWS), TV_ROWS * TV_G, smem, s>>>(
      T, nodes, d.tinv.p, d.rowperm.p, work, S.work_entries, xp, na, ld_b);
}

template <int NRHS>
void tinv_forward_level(Handle& h, LUDevice& d, const LUSymbolic& S, int lev, double* work, double* xp, int na) {
  cudaStream_t s = h.stream;
  const int nfront = S.level_start[lev + 1] - S.level_start[lev];
  const int* nodes = d.level_nodes.p + S.level_start[lev];
  const int prep_threads = S.level_max_m[lev] > 2048 ? 1024 : S.level_max_m[lev] > 768 ? 512 : 256;
  k_tinv_fwd_prep<NRHS><<<nfront, prep_threads, 0, s>>>(d.tab(), nodes, work, S.work_entries, xp, na);
  launch_matvec<NRHS, true>(d, S, lev, work, xp, na, s);
  h.tim.kernel_launches += 1;  // (the caller counts one launch per level)
}

template <int NRHS>
void tinv_backward_level(Handle& h, LUDevice& d, const LUSymbolic& S, int lev, double* work, double* xp, int na) {
  cudaStream_t s = h.stream;
  const int nfront = S.level_start[lev + 1] - S.level_start[lev];
  const int* nodes = d.level_nodes.p + S.level_start[lev];
  launch_matvec<NRHS, false>(d, S, lev, work, xp, na, s);
  const int ny = std::max(1, std::min(8, (S.level_max_ns[lev] + 255) / 256));
  k_tinv_bwd_copy<NRHS><<<dim3(nfront, ny), 256, 0, s>>>(d.tab(), nodes, work, S.work_entries, xp, na);
  h.tim.kernel_launches += 1;
}

template void tinv_forward_level<1>(Handle&, LUDevice&, const LUSymbolic&, int, double*, double*, int);
template void tinv_forward_level<2>(Handle&, LUDevice&, const LUSymbolic&, int, double*, double*, int);
template void tinv_backward_level<1>(Handle&, LUDevice&, const LUSymbolic&, int, double*, double*, int);
template void tinv_backward_level<2>(Handle&, LUDevice&, const LUSymbolic&, int, double*, double*, int);

}  // namespace fw
