// K4: CSR SpMV (f64 / c128 matrix, f64 / c128 vector), hand-written (no cuSPARSE).
// Replaces `M @ x` inside scipy's ARPACK reverse-communication loop (arpack.py:825-826) and
// `aform.assemble(basis) @ xs` of calculate_hfield (femwell/maxwell/waveguide.py:672).
// HBM-bound: nnz*(sv+4) + (n+1)*4 + 2*n*sv bytes.
#include "handle.h"

namespace fw {

__device__ inline double shfl_down_(double v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
__device__ inline cplx shfl_down_(cplx v, int o) {
  return {__shfl_down_sync(0xffffffffu, v.re, o), __shfl_down_sync(0xffffffffu, v.im, o)};
}

// TPR lanes cooperate on one row (rows of the order-2 operator hold 14..~110 entries)
template <class MT, class VT, int TPR>
__global__ void __launch_bounds__(256) k_spmv(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                                              const MT* __restrict__ vals, const VT* __restrict__ x,
                                              VT* __restrict__ y) {
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = (int)(gt / TPR);
  const int l = (int)(gt % TPR);
  VT acc = scalar_traits<VT>::zero();
  if (row < n) {
    const int a = indptr[row], b = indptr[row + 1];
    // batches of UNR entries per lane with predicated loads: the UNR index loads, value loads and x gathers of a
    // batch are independent, so a lane keeps UNR dependent (index -> x) chains in flight instead of one (the
    // average row holds ~23 entries = one batch of 8 lanes x 4)
    constexpr int UNR = 4;
    for (int k = a + l; k < b; k += UNR * TPR) {
      int c[UNR];
      MT v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int kk = k + u * TPR;
        const bool ok = kk < b;
        c[u] = ok ? indices[kk] : row;
        v[u] = ok ? vals[kk] : scalar_traits<MT>::zero();
      }
      VT xv[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) xv[u] = x[c[u]];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (k + u * TPR < b) fma_(acc, v[u], xv[u]);  // (predicated: a non-finite x[row] must not leak in as 0 * inf)
    }
  }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) acc = acc + shfl_down_(acc, o);
  if (l == 0 && row < n) y[row] = acc;
}

template <class MT, class VT>
void spmv(int n, const int* indptr, const int* indices, const MT* vals, const VT* x, VT* y, cudaStream_t s) {
  constexpr int TPR = 8;
  int64_t threads = (int64_t)n * TPR;
  k_spmv<MT, VT, TPR><<<cdiv(threads, 256), 256, 0, s>>>(n, indptr, indices, vals, x, y);
  FW_CHECK_CUDA(cudaGetLastError());
}
template void spmv<double, double>(int, const int*, const int*, const double*, const double*, double*, cudaStream_t);
template void spmv<double, cplx>(int, const int*, const int*, const double*, const cplx*, cplx*, cudaStream_t);
template void spmv<cplx, cplx>(int, const int*, const int*, const cplx*, const cplx*, cplx*, cudaStream_t);

}  // namespace fw
