// Register-tiled FP64 / C128 GEMM for the trailing updates of the multifrontal LU (large fronts).
//   C[r0:, r0:] -= L[r0:, kbeg:kbeg+klen] * U[kbeg:kbeg+klen, r0:]       (column-major front, ld = m)
// FP64 has no tcgen05 path (the 5th-gen tensor cores stop at TF32); the f64 update runs on the FP64 tensor
// instruction of the legacy path (mma.sync m8n8k4 = DMMA, second half of this file), c128 on this SIMT kernel: CTA tile 16*MT x 16*MT (MT = 8 for f64 -> 128 x 128, MT = 4 for c128 -> 64 x 64), 256 threads, every
// thread an MT x MT register tile split into 2 x 2 sub-blocks of MT/2 so that shared-memory reads are 128-bit
// and conflict free, k-chunks of 16 double-buffered through registers (global loads of chunk i+1 are in
// flight while chunk i is multiplied).  Shared-memory bandwidth and the FP64 pipe are balanced at this tile
// shape (8 LDS.128 per 64 DFMA per thread and k).
#include "lu_dev.h"

namespace fw {

template <class S>
struct Gemm2Cfg {
  static constexpr int MT = 8;  // register tile per thread
};
template <>
struct Gemm2Cfg<cplx> {
  static constexpr int MT = 4;
};

constexpr int G2_KC = 16;

// SCHUR = false: per-panel rank-32 update, entries with both indices >= ns (the Schur block) are skipped.
// SCHUR = true : deferred Schur complement, S22 -= L21 U12 over the full pivot block.
template <class S, bool SCHUR>
__global__ void __launch_bounds__(256) k_lu_gemm_big(FrontTab T, const int* __restrict__ nodes, int kstep,
                                                     S* __restrict__ F, int kw) {
  constexpr int MT = Gemm2Cfg<S>::MT, H = MT / 2;
  constexpr int TM = 16 * MT;     // CTA tile edge
  constexpr int LDB = TM + 2;     // padded row length of Bs (keeps 16-byte alignment, 2-way store conflicts)
  // dynamic shared memory (2 x 16 x (TM + LDB) scalars = 65 KB > the 48 KB static limit)
  extern __shared__ __align__(16) unsigned char g2_smem[];
  S(*As)[G2_KC][TM] = reinterpret_cast<S(*)[G2_KC][TM]>(g2_smem);
  S(*Bs)[G2_KC][LDB] = reinterpret_cast<S(*)[G2_KC][LDB]>(g2_smem + sizeof(S) * 2 * G2_KC * TM);
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = SCHUR ? 0 : kstep * kw;  // kw: width of the column block (one panel, or an outer block)
  if (k0 >= ns) return;
  const int klen = SCHUR ? ns : min(kw, ns - k0);
  const int r0 = SCHUR ? ns : k0 + klen;
  const int nrem = m - r0;
  if (nrem <= 0) return;
  const int nt = (nrem + TM - 1) / TM;
  const int ti = blockIdx.y, tj = blockIdx.z;
  if (ti >= nt || tj >= nt) return;
  const int rbase = r0 + ti * TM, cbase = r0 + tj * TM;
  if (!SCHUR && rbase >= ns && cbase >= ns) return;  // tile entirely inside S22
  S* Fn = F + T.front_off[node];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // global -> register staging: A chunk = G2_KC x TM (rows contiguous), B chunk = G2_KC x TM (k contiguous)
  constexpr int PER = G2_KC * TM / 256;  // elements per thread per operand (8 for f64, 4 for c128)
  S ra[PER], rb[PER];
  auto fetch = [&](int kc) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int idx = tid + q * 256;
      {
        const int r = idx % TM, k = idx / TM;
        const int gr = rbase + r, gk = kc + k;
        ra[q] = (gr < m && gk < klen) ? Fn[gr + (long long)(k0 + gk) * m] : scalar_traits<S>::zero();
      }
      {
        const int k = idx % G2_KC, c = idx / G2_KC;
        const int gc = cbase + c, gk = kc + k;
        rb[q] = (gc < m && gk < klen) ? Fn[(k0 + gk) + (long long)gc * m] : scalar_traits<S>::zero();
      }
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int idx = tid + q * 256;
      As[buf][idx / TM][idx % TM] = ra[q];
      Bs[buf][idx % G2_KC][idx / G2_KC] = rb[q];
    }
  };

  S acc[MT][MT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j) acc[i][j] = scalar_traits<S>::zero();

  fetch(0);
  commit(0);
  __syncthreads();
  int buf = 0;
  for (int kc = 0; kc < klen; kc += G2_KC) {
    const bool more = kc + G2_KC < klen;
    if (more) fetch(kc + G2_KC);
#pragma unroll
    for (int k = 0; k < G2_KC; ++k) {
      S av[MT], bv[MT];
#pragma unroll
      for (int i = 0; i < H; ++i) {
        av[i] = As[buf][k][tx * H + i];
        av[H + i] = As[buf][k][TM / 2 + tx * H + i];
        bv[i] = Bs[buf][k][ty * H + i];
        bv[H + i] = Bs[buf][k][TM / 2 + ty * H + i];
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) fnma_(acc[i][j], av[i], bv[j]);
    }
    if (more) {
      commit(buf ^ 1);  // the other buffer was last read one iteration ago (barrier below)
      __syncthreads();
      buf ^= 1;
    }
  }
  // C += acc, two columns of the thread tile at a time: all loads of the pair before the first store (written as one
  // read-modify-write per entry the compiler must assume that a store aliases the next load, and every entry then
  // waits a full memory round trip: SASS showed LDG, STG, LDG, STG ...)
#pragma unroll
  for (int j0 = 0; j0 < MT; j0 += 2) {
    S old[2][MT];
    bool ok[2][MT];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = j0 + jj;
      const int gc = cbase + (j < H ? ty * H + j : TM / 2 + ty * H + (j - H));
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int gr = rbase + (i < H ? tx * H + i : TM / 2 + tx * H + (i - H));
        ok[jj][i] = gc < m && gr < m && (SCHUR || gr < ns || gc < ns);
        old[jj][i] = ok[jj][i] ? Fn[gr + (long long)gc * m] : scalar_traits<S>::zero();
      }
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = j0 + jj;
      const int gc = cbase + (j < H ? ty * H + j : TM / 2 + ty * H + (j - H));
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int gr = rbase + (i < H ? tx * H + i : TM / 2 + tx * H + (i - H));
        if (ok[jj][i]) Fn[gr + (long long)gc * m] = old[jj][i] + acc[i][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------ FP64 tensor-core variant (DMMA)
// The same update on mma.sync.aligned.m8n8k4.f64 (SASS: DMMA).  Measured on B200 (scripts/microbench/fp64_peak.cu):
// DMMA issues at 37.1 TFLOP/s from registers, DFMA tops out at 33.9 -- the ceiling is the same FP64 pipe, but one DMMA
// replaces 8 DFMA per lane and needs two operand registers instead of an 8 x 8 register tile: the SIMT kernel above
// needs 244 registers (one CTA of 8 warps per SM, FP64 pipe 16 % busy), this one 64 accumulator registers per lane.
//   CTA tile 128 x 128, 512 threads = 4 x 4 warps, warp tile 32 x 32 = 4 x 4 mma tiles, k-chunks of 16 through a
//   3-stage cp.async pipeline (8-byte copies with zero fill at the ragged edges: front rows are only 8-byte aligned).
//   Shared layouts make every fragment load conflict free: As[k][r] with row length 132 and Bs[c][k] with row
//   length 20 (both = 4 mod 16 doubles), so the 16 lanes of a half warp hit 16 distinct 8-byte banks.
constexpr int DM_KC = 16, DM_LDB = DM_KC + 4, DM_STAGES = 3;

__device__ __forceinline__ void cp_async8_zfill(double* smem, const double* gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// TM: CTA tile edge (128: 4 x 4 warps, 512 threads, one CTA per SM -- large fronts; 64: 2 x 2 warps, 128 threads,
// three CTAs per SM -- fronts whose trailing block is a few hundred rows, where 128-wide tiles waste up to half of
// the tensor work on the ragged edge: ncu on level 10 of config 2 showed the math pipe saturated at 8.9 useful TFLOP/s).
template <bool SCHUR, int TM>
__global__ void __launch_bounds__(TM * TM / 32, TM == 128 ? 1 : 3)
    k_lu_gemm_dmma(FrontTab T, const int* __restrict__ nodes, int kstep, double* __restrict__ F, int kw) {
  constexpr int WM = TM / 32, THREADS = TM * TM / 32, LDA = TM + 4;
  constexpr int STAGE_DOUBLES = DM_KC * LDA + TM * DM_LDB;
  extern __shared__ __align__(16) double dm_smem[];
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = SCHUR ? 0 : kstep * kw;  // kw: width of the column block (one panel, or an outer block)
  if (k0 >= ns) return;
  const int klen = SCHUR ? ns : min(kw, ns - k0);
  const int r0 = SCHUR ? ns : k0 + klen;
  const int nrem = m - r0;
  if (nrem <= 0) return;
  const int nt = (nrem + TM - 1) / TM;
  const int ti = blockIdx.y, tj = blockIdx.z;
  if (ti >= nt || tj >= nt) return;
  const int rbase = r0 + ti * TM, cbase = r0 + tj * TM;
  if (!SCHUR && rbase >= ns && cbase >= ns) return;  // tile entirely inside S22
  double* Fn = F + T.front_off[node];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, q = lane & 3;

  auto issue = [&](int chunk, int stage) {
    double* As = dm_smem + (size_t)stage * STAGE_DOUBLES;
    double* Bs = As + DM_KC * LDA;
    const int kc = chunk * DM_KC;
#pragma unroll
    for (int it = 0; it < DM_KC * TM / THREADS; ++it) {
      const int idx = tid + it * THREADS;
      {
        const int r = idx % TM, k = idx / TM;
        const int gr = rbase + r, gk = kc + k;
        const bool ok = gr < m && gk < klen;
        cp_async8_zfill(As + k * LDA + r, ok ? Fn + gr + (long long)(k0 + gk) * m : Fn, ok);
      }
      {
        const int k = idx % DM_KC, c = idx / DM_KC;
        const int gc = cbase + c, gk = kc + k;
        const bool ok = gc < m && gk < klen;
        cp_async8_zfill(Bs + c * DM_LDB + k, ok ? Fn + (k0 + gk) + (long long)gc * m : Fn, ok);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nchunk = (klen + DM_KC - 1) / DM_KC;
  issue(0, 0);
  if (nchunk > 1)
    issue(1, 1);
  else
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  for (int c = 0; c < nchunk; ++c) {
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();  // chunk c has landed for everybody; stage (c + 2) % 3 was last read in iteration c - 1
    if (c + 2 < nchunk)
      issue(c + 2, (c + 2) % DM_STAGES);
    else
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    const double* As = dm_smem + (size_t)(c % DM_STAGES) * STAGE_DOUBLES;
    const double* Bs = As + DM_KC * LDA;
#pragma unroll
    for (int kk = 0; kk < DM_KC / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(kk * 4 + q) * LDA + wm * 32 + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[(wn * 32 + j * 8 + g) * DM_LDB + kk * 4 + q];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  // C -= acc: lane holds rows g of every 8 x 8 tile, columns 2 q and 2 q + 1
  // (JG 8-column groups at a time -- 16 loads before the first store, 8 in the 512-thread variant with its
  // 128-register budget; see the SIMT kernel above)
  constexpr int JG = TM == 128 ? 1 : 2;
#pragma unroll
  for (int j0 = 0; j0 < 4; j0 += JG) {
    double old[JG][2][4];
    bool ok[JG][2][4];
#pragma unroll
    for (int jj = 0; jj < JG; ++jj)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gc = cbase + wn * 32 + (j0 + jj) * 8 + 2 * q + h;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gr = rbase + wm * 32 + i * 8 + g;
          ok[jj][h][i] = gc < m && gr < m && (SCHUR || gr < ns || gc < ns);
          old[jj][h][i] = ok[jj][h][i] ? Fn[gr + (long long)gc * m] : 0.0;
        }
      }
#pragma unroll
    for (int jj = 0; jj < JG; ++jj)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gc = cbase + wn * 32 + (j0 + jj) * 8 + 2 * q + h;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gr = rbase + wm * 32 + i * 8 + g;
          if (ok[jj][h][i]) Fn[gr + (long long)gc * m] = old[jj][h][i] - acc[i][j0 + jj][h];
        }
      }
  }
}

template <int TM>
static void lu_gemm_dmma_launch_tm(FrontTab T, const int* nodes, int nfront, int kstep, double* F, int max_rem,
                                   bool schur, cudaStream_t s, int kw) {
  constexpr size_t smem = sizeof(double) * DM_STAGES * (DM_KC * (TM + 4) + TM * DM_LDB);
  constexpr int THREADS = TM * TM / 32;
  const int nt = (max_rem + TM - 1) / TM;
  if (nt <= 0) return;
  static bool configured[64] = {};
  if (first_use_on_device(configured)) {
    FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_gemm_dmma<true, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_gemm_dmma<false, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (schur)
    k_lu_gemm_dmma<true, TM><<<dim3(nfront, nt, nt), THREADS, smem, s>>>(T, nodes, kstep, F, kw);
  else
    k_lu_gemm_dmma<false, TM><<<dim3(nfront, nt, nt), THREADS, smem, s>>>(T, nodes, kstep, F, kw);
}

static void lu_gemm_dmma_launch(FrontTab T, const int* nodes, int nfront, int kstep, double* F, int max_rem, bool schur,
                                cudaStream_t s, int kw) {
  // trailing blocks below ~900 rows: 64 x 64 tiles (less ragged-edge waste, three CTAs per SM)
  static const int small_max = getenv("FW_DMMA64_MAX") ? atoi(getenv("FW_DMMA64_MAX")) : 2000;
  if (max_rem <= small_max)
    lu_gemm_dmma_launch_tm<64>(T, nodes, nfront, kstep, F, max_rem, schur, s, kw);
  else
    lu_gemm_dmma_launch_tm<128>(T, nodes, nfront, kstep, F, max_rem, schur, s, kw);
}

template <class S>
static bool try_dmma(FrontTab, const int*, int, int, S*, int, bool, cudaStream_t, int) {
  return false;
}
template <>
bool try_dmma<double>(FrontTab T, const int* nodes, int nfront, int kstep, double* F, int max_rem, bool schur,
                      cudaStream_t s, int kw) {
  static const bool off = getenv("FW_NO_DMMA") != nullptr;
  if (off) return false;
  lu_gemm_dmma_launch(T, nodes, nfront, kstep, F, max_rem, schur, s, kw);
  return true;
}

template <class S>
void lu_gemm_big_launch(FrontTab T, const int* nodes, int nfront, int kstep, S* F, int max_rem, bool schur,
                        cudaStream_t s, int kw) {
  if (try_dmma<S>(T, nodes, nfront, kstep, F, max_rem, schur, s, kw)) return;
  constexpr int TM = 16 * Gemm2Cfg<S>::MT;
  constexpr size_t smem = sizeof(S) * 2 * G2_KC * (TM + TM + 2);
  const int nt = (max_rem + TM - 1) / TM;
  if (nt <= 0) return;
  static bool configured[64] = {};
  if (first_use_on_device(configured)) {
    FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_gemm_big<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FW_CHECK_CUDA(cudaFuncSetAttribute(k_lu_gemm_big<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (schur)
    k_lu_gemm_big<S, true><<<dim3(nfront, nt, nt), 256, smem, s>>>(T, nodes, kstep, F, kw);
  else
    k_lu_gemm_big<S, false><<<dim3(nfront, nt, nt), 256, smem, s>>>(T, nodes, kstep, F, kw);
}
template void lu_gemm_big_launch<double>(FrontTab, const int*, int, int, double*, int, bool, cudaStream_t, int);
template void lu_gemm_big_launch<cplx>(FrontTab, const int*, int, int, cplx*, int, bool, cudaStream_t, int);

}  // namespace fw
