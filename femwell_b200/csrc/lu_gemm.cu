// Register-tiled FP64 / C128 GEMM for the trailing updates of the multifrontal LU (large fronts).
//   C[r0:, r0:] -= L[r0:, kbeg:kbeg+klen] * U[kbeg:kbeg+klen, r0:]       (column-major front, ld = m)
// FP64 has no tcgen05 path (the 5th-gen tensor cores stop at TF32), so this is a SIMT kernel on the FP64
// pipe: CTA tile 16*MT x 16*MT (MT = 8 for f64 -> 128 x 128, MT = 4 for c128 -> 64 x 64), 256 threads, every
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
                                                     S* __restrict__ F) {
  constexpr int MT = Gemm2Cfg<S>::MT, H = MT / 2;
  constexpr int TM = 16 * MT;     // CTA tile edge
  constexpr int LDB = TM + 2;     // padded row length of Bs (keeps 16-byte alignment, 2-way store conflicts)
  // dynamic shared memory (2 x 16 x (TM + LDB) scalars = 65 KB > the 48 KB static limit)
  extern __shared__ __align__(16) unsigned char g2_smem[];
  S(*As)[G2_KC][TM] = reinterpret_cast<S(*)[G2_KC][TM]>(g2_smem);
  S(*Bs)[G2_KC][LDB] = reinterpret_cast<S(*)[G2_KC][LDB]>(g2_smem + sizeof(S) * 2 * G2_KC * TM);
  const int node = nodes[blockIdx.x];
  const int m = T.m[node], ns = T.ns[node];
  const int k0 = SCHUR ? 0 : kstep * PW;
  if (k0 >= ns) return;
  const int klen = SCHUR ? ns : min(PW, ns - k0);
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
#pragma unroll
  for (int j = 0; j < MT; ++j) {
    const int gc = cbase + (j < H ? ty * H + j : TM / 2 + ty * H + (j - H));
    if (gc >= m) continue;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const int gr = rbase + (i < H ? tx * H + i : TM / 2 + tx * H + (i - H));
      if (gr < m && (SCHUR || gr < ns || gc < ns)) {
        S* dst = Fn + gr + (long long)gc * m;
        *dst = *dst + acc[i][j];
      }
    }
  }
}

template <class S>
void lu_gemm_big_launch(FrontTab T, const int* nodes, int nfront, int kstep, S* F, int max_rem, bool schur,
                        cudaStream_t s) {
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
    k_lu_gemm_big<S, true><<<dim3(nfront, nt, nt), 256, smem, s>>>(T, nodes, kstep, F);
  else
    k_lu_gemm_big<S, false><<<dim3(nfront, nt, nt), 256, smem, s>>>(T, nodes, kstep, F);
}
template void lu_gemm_big_launch<double>(FrontTab, const int*, int, int, double*, int, bool, cudaStream_t);
template void lu_gemm_big_launch<cplx>(FrontTab, const int*, int, int, cplx*, int, bool, cudaStream_t);

}  // namespace fw
