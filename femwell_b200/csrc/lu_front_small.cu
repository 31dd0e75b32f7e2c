// Fused partial factorisation of SMALL fronts (m <= a few hundred rows, pivot block <= 128 rows): one CTA walks through
// every step of its front -- panel LU with partial pivoting in a shared-memory window, row interchanges, unit-lower
// solve of the U rows inside the pivot block, L rows of the boundary, rank-32 update of the pivot-block columns, then
// the U strip of the boundary columns and the Schur complement -- with block barriers between the steps instead of
// kernel boundaries.  The level-by-level kernels of lu_numeric.cu launch 3 kernels per 32-column panel over ALL fronts
// of a level: on the levels with thousands of small fronts every one of those launches streams the level from HBM and
// waits for the slowest CTA of its last wave (config 2: the leaf level streams its 5.2 GB about twelve times, 26 ms
// for 54 GFLOP).  Here a front is read and written once (it stays in L1/L2 while its CTA works on it) and the SMs
// stay busy with independent fronts at different steps.  Same arithmetic and the same pivot choice as the
// level-by-level kernels (largest modulus, lowest row on ties).
#include "lu_dev.h"

namespace fw {

namespace {

constexpr int FS_THREADS = 128;  // small CTAs: more fronts in flight per SM hide the dependent steps of each
constexpr int FS_WIN = 128;  // rows of the panel window = largest pivot block handled here
constexpr int FS_TM = 64, FS_TN = 32, FS_KC = 32;  // GEMM tile: 64 rows x 32 columns, 4 x 4 entries per thread

template <class S>
struct FsSmem {
  // the panel window and the GEMM tiles are never live at the same time
  union {
    S a[PW][FS_WIN + 1];
    struct {
      S As[FS_KC][FS_TM];
      S Bs[FS_KC][FS_TN + 1];
    } g;
    S Ls[FS_WIN][PW + 1];  // L slab of the U-strip solve
  } u;
  S D[PW][PW + 1];
  int sp[PW], s_top[PW], s_far[PW];
  double red_v[FS_THREADS / 32];
  int red_i[FS_THREADS / 32];
  int s_piv;
};

// C[rows r_lo.., cols c_lo..c_hi) -= L[rows, k0:k0+klen) * U[k0:k0+klen, cols), one 64 x 32 tile per call;
// skip_s22: entries with row >= ns and col >= ns are left alone (they belong to the Schur complement)
template <class S>
__device__ __forceinline__ void fs_gemm_tile(FsSmem<S>& sm, S* Fn, int m, int ns, int k0, int klen, int rbase, int cbase,
                                             int c_hi, bool skip_s22) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  S acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = scalar_traits<S>::zero();
  for (int kc = 0; kc < klen; kc += FS_KC) {
    for (int idx = tid; idx < FS_KC * FS_TM; idx += FS_THREADS) {
      const int r = idx % FS_TM, k = idx / FS_TM;
      const int gr = rbase + r, gk = kc + k;
      sm.u.g.As[k][r] = (gr < m && gk < klen) ? Fn[gr + (long long)(k0 + gk) * m] : scalar_traits<S>::zero();
    }
    for (int idx = tid; idx < FS_KC * FS_TN; idx += FS_THREADS) {
      const int k = idx % FS_KC, c = idx / FS_KC;
      const int gc = cbase + c, gk = kc + k;
      sm.u.g.Bs[k][c] = (gc < c_hi && gk < klen) ? Fn[(k0 + gk) + (long long)gc * m] : scalar_traits<S>::zero();
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FS_KC; ++k) {
      S av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sm.u.g.As[k][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sm.u.g.Bs[k][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fnma_(acc[i][j], av[i], bv[j]);
    }
    __syncthreads();
  }
  // all loads of the thread tile before the first store (see k_lu_gemm)
  S old[4][4];
  bool ok[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gc = cbase + ty * 4 + j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = rbase + tx * 4 + i;
      ok[i][j] = gc < c_hi && gr < m && !(skip_s22 && gr >= ns && gc >= ns);
      old[i][j] = ok[i][j] ? Fn[gr + (long long)gc * m] : scalar_traits<S>::zero();
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gc = cbase + ty * 4 + j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = rbase + tx * 4 + i;
      if (ok[i][j]) Fn[gr + (long long)gc * m] = old[i][j] + acc[i][j];
    }
  }
}

}  // namespace

template <class S>
__global__ void __launch_bounds__(FS_THREADS, sizeof(S) == 8 ? 5 : 2) k_lu_front_small(FrontTab T, const int* __restrict__ nodes,
                                                               S* __restrict__ F, int* __restrict__ piv, double thr,
                                                               int* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char fs_raw[];
  FsSmem<S>& sm = *reinterpret_cast<FsSmem<S>*>(fs_raw);
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  if (ns == 0) return;
  S* Fn = F + T.front_off[node];
  const int t = threadIdx.x;
  constexpr int LD = FS_WIN + 1;
  S* a = &sm.u.a[0][0];  // a[c * LD + r]
  const int npan = (ns + PW - 1) / PW;
  for (int kp = 0; kp < npan; ++kp) {
    const int k0 = kp * PW;
    const int w = min(PW, ns - k0);
    const int hw = ns - k0;  // fully-summed rows still in play (<= FS_WIN)
    // ---- (1) panel LU with partial pivoting in the shared-memory window (one row per thread)
    for (int c = 0; c < w; ++c)
      if (t < hw) a[c * LD + t] = Fn[(k0 + t) + (long long)(k0 + c) * m];
    __syncthreads();
    for (int j = 0; j < w; ++j) {
      double v = -1.0;
      int idx = t;
      if (t >= j && t < hw) v = abs1_(a[j * LD + t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oi = __shfl_down_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) v = ov, idx = oi;
      }
      if ((t & 31) == 0) sm.red_v[t >> 5] = v, sm.red_i[t >> 5] = idx;
      __syncthreads();
      if (t == 0) {
        double bv = sm.red_v[0];
        int bi = sm.red_i[0];
        for (int q = 1; q < FS_THREADS / 32; ++q)
          if (sm.red_v[q] > bv || (sm.red_v[q] == bv && sm.red_i[q] < bi)) bv = sm.red_v[q], bi = sm.red_i[q];
        sm.s_piv = bi;
        sm.sp[j] = k0 + bi;
        piv[T.own_start[node] + k0 + j] = k0 + bi;
        if (!(bv >= thr)) {
          // static pivot perturbation (tiny / exactly singular pivot); corrected by iterative refinement
          a[j * LD + bi] = scalar_traits<S>::from(thr, 0.0);
          atomicAdd(&counters[0], 1);
          if (bv == 0.0) atomicAdd(&counters[1], 1);
        }
      }
      __syncthreads();
      const int pr = sm.s_piv;
      if (t < w && pr != j) {
        const S tmp = a[t * LD + j];
        a[t * LD + j] = a[t * LD + pr];
        a[t * LD + pr] = tmp;
      }
      __syncthreads();
      if (t > j && t < hw) {
        const S l = a[j * LD + t] / a[j * LD + j];
        a[j * LD + t] = l;
        for (int c = j + 1; c < w; ++c) fnma_(a[c * LD + t], l, a[c * LD + j]);
      }
      __syncthreads();
    }
    for (int c = 0; c < w; ++c)
      if (t < hw) Fn[(k0 + t) + (long long)(k0 + c) * m] = a[c * LD + t];
    for (int idx = t; idx < w * w; idx += FS_THREADS) {
      const int r = idx % w, c = idx / w;
      sm.D[r][c] = a[c * LD + r];
    }
    __syncthreads();
    // ---- (2) the other columns: row interchanges; unit-lower solve of the U rows for the columns of the pivot block
    // (the w sequential interchanges as ONE gather: threads 0..63 trace them backwards to find which original row ends at
    //  each top position and at each partner position below the top block; see k_lu_swap_trsm<GATHER>)
    if (t < 2 * PW) {
      const int j = t & (PW - 1);
      int pos = t < PW ? k0 + j : (j < w ? sm.sp[j] : -1);
      const bool wanted = j < w && (t < PW || pos >= k0 + w);
      if (wanted) {
        for (int i = w - 1; i >= 0; --i) {
          const int a0 = k0 + i, b0 = sm.sp[i];
          pos = pos == a0 ? b0 : (pos == b0 ? a0 : pos);
        }
      }
      if (t < PW)
        sm.s_top[j] = wanted ? pos : k0 + j;
      else
        sm.s_far[j] = wanted ? pos : -1;
    }
    __syncthreads();
    for (int c = t; c < m; c += FS_THREADS) {
      if (c >= k0 && c < k0 + w) continue;
      S* col = Fn + (long long)c * m;
      S x[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) x[j] = (j < w) ? col[sm.s_top[j]] : scalar_traits<S>::zero();
#pragma unroll
      for (int ch = 0; ch < PW; ch += 8) {
        S y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int f = (ch + u < w) ? sm.s_far[ch + u] : -1;
          y[u] = f >= 0 ? col[f] : scalar_traits<S>::zero();
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (ch + u < w && sm.s_far[ch + u] >= 0) col[sm.sp[ch + u]] = y[u];
      }
      if (c >= k0 + w && c < ns) {
#pragma unroll
        for (int j = 1; j < PW; ++j) {
          if (j < w) {
#pragma unroll
            for (int i = 0; i < j; ++i) fnma_(x[j], sm.D[j][i], x[i]);
          }
        }
#pragma unroll
        for (int j = 0; j < PW; ++j)
          if (j < w) col[k0 + j] = x[j];
      } else {
#pragma unroll
        for (int j = 0; j < PW; ++j)
          if (j < w && sm.s_top[j] != k0 + j) col[k0 + j] = x[j];
      }
    }
    // ---- (3) boundary rows: L21 block of this panel, X U_kk = F
    for (int r = ns + t; r < m; r += FS_THREADS) {
      S x[PW];
#pragma unroll
      for (int c = 0; c < PW; ++c) x[c] = (c < w) ? Fn[r + (long long)(k0 + c) * m] : scalar_traits<S>::zero();
#pragma unroll
      for (int c = 0; c < PW; ++c) {
        if (c < w) {
#pragma unroll
          for (int i = 0; i < c; ++i) fnma_(x[c], x[i], sm.D[i][c]);
          x[c] = x[c] / sm.D[c][c];
        }
      }
#pragma unroll
      for (int c = 0; c < PW; ++c)
        if (c < w) Fn[r + (long long)(k0 + c) * m] = x[c];
    }
    __syncthreads();  // the front (global memory) is consistent for the whole CTA
    // ---- (4) rank-32 update of the remaining columns of the pivot block, all rows below the panel
    const int r0 = k0 + w;
    if (r0 < ns) {
      for (int rb = r0; rb < m; rb += FS_TM)
        for (int cb = r0; cb < ns; cb += FS_TN) fs_gemm_tile<S>(sm, Fn, m, ns, k0, w, rb, cb, ns, false);
    }
    __syncthreads();
  }
  if (m == ns) return;
  // ---- (5) U strip of the boundary columns: X <- inv(L11) X with the unit-lower ns x ns pivot block, 32 rows at a time
  for (int s0 = 0; s0 < ns; s0 += PW) {
    const int w = min(PW, ns - s0);
    const int rows = ns - s0;
    __syncthreads();
    for (int idx = t; idx < rows * w; idx += FS_THREADS) {
      const int r = idx % rows, cc = idx / rows;
      sm.u.Ls[r][cc] = Fn[(s0 + r) + (long long)(s0 + cc) * m];
    }
    __syncthreads();
    for (int c = ns + t; c < m; c += FS_THREADS) {
      S* col = Fn + (long long)c * m;
      S x[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) x[j] = (j < w) ? col[s0 + j] : scalar_traits<S>::zero();
#pragma unroll
      for (int j = 1; j < PW; ++j) {
        if (j < w) {
#pragma unroll
          for (int i = 0; i < j; ++i) fnma_(x[j], sm.u.Ls[j][i], x[i]);
        }
      }
#pragma unroll
      for (int j = 0; j < PW; ++j)
        if (j < w) col[s0 + j] = x[j];
      int r = w;
      for (; r + 8 <= rows; r += 8) {
        S acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = col[s0 + r + u];
#pragma unroll
        for (int j = 0; j < PW; ++j)
          if (j < w) {
#pragma unroll
            for (int u = 0; u < 8; ++u) fnma_(acc[u], sm.u.Ls[r + u][j], x[j]);
          }
#pragma unroll
        for (int u = 0; u < 8; ++u) col[s0 + r + u] = acc[u];
      }
      for (; r < rows; ++r) {
        S acc = col[s0 + r];
#pragma unroll
        for (int j = 0; j < PW; ++j)
          if (j < w) fnma_(acc, sm.u.Ls[r][j], x[j]);
        col[s0 + r] = acc;
      }
    }
  }
  __syncthreads();
  // ---- (6) Schur complement S22 -= L21 U12 (inner dimension ns)
  for (int rb = ns; rb < m; rb += FS_TM)
    for (int cb = ns; cb < m; cb += FS_TN) fs_gemm_tile<S>(sm, Fn, m, ns, 0, ns, rb, cb, m, false);
}

// levels whose fronts fit the kernel: pivot blocks of at most FS_WIN rows
bool lu_front_small_applies(int max_ns, int max_m) {
  static const int max_m_lim = getenv("FW_FRONT_SMALL_MAX_M") ? atoi(getenv("FW_FRONT_SMALL_MAX_M")) : 400;
  // opt-in (FW_FRONT_SMALL=1): measured SLOWER than the level-by-level kernels at config 2 (leaf level 32.5 vs 26.5 ms,
  // levels 14-11 equal): ncu shows 213 k warp instructions per front for 0.8 M useful FMAs -- zero-padded 32-wide
  // triangular solves and mostly empty 64 x 32 tiles on 140-row fronts -- and an instruction-fetch stall of 8 per issue
  // (four fully unrolled 32 x 32 solves in one kernel); DRAM traffic does drop to 12.8 GB for the level.
  static const bool on = getenv("FW_FRONT_SMALL") && atoi(getenv("FW_FRONT_SMALL")) > 0;
  return on && max_ns <= FS_WIN && max_m <= max_m_lim;
}

template <class S>
void lu_front_small_launch(FrontTab T, const int* nodes, int nfront, S* F, int* piv, double thr, int* counters,
                           cudaStream_t s) {
  const size_t smem = sizeof(FsSmem<S>);
  static bool configured[64] = {};
  if (first_use_on_device(configured))
    FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_front_small<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_lu_front_small<S><<<nfront, FS_THREADS, smem, s>>>(T, nodes, F, piv, thr, counters);
}
template void lu_front_small_launch<double>(FrontTab, const int*, int, double*, int*, double, int*, cudaStream_t);
template void lu_front_small_launch<cplx>(FrontTab, const int*, int, cplx*, int*, double, int*, cudaStream_t);

}  // namespace fw
