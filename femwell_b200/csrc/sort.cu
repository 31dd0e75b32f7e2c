// K2: COO -> CSR.  Hand-written LSD radix sort-by-key (8-bit digits, stable), device-wide
// exclusive scan, segment-head detection, segmented reduce and eliminate_zeros compaction.
// Replaces scikit-fem COOData.tocsr -> scipy coo_matrix.eliminate_zeros().tocsr()
// (reference call site femwell/maxwell/waveguide.py:602-603).
// HBM-bound integer/byte work: no tensor cores; coalesced streaming + shared-memory staging.
#include <algorithm>

#include "handle.h"

namespace fw {

// ------------------------------------------------------------------------------ scan
constexpr int SC_THREADS = 512;
constexpr int SC_IPT = 8;
constexpr int SC_TILE = SC_THREADS * SC_IPT;

__device__ inline uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* sh /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    sh[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  uint32_t warp_off = warp ? sh[warp - 1] : 0;
  if (total) *total = sh[(blockDim.x >> 5) - 1];
  uint32_t r = warp_off + x - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ in, int64_t n,
                                                               uint32_t* __restrict__ sums) {
  __shared__ uint32_t sh[32];
  int64_t base = (int64_t)blockIdx.x * SC_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    int64_t g = base + (int64_t)k * SC_THREADS + threadIdx.x;
    if (g < n) s += in[g];
  }
  uint32_t tot;
  block_exclusive_scan(s, &tot, sh);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_single(uint32_t* data, int64_t n) {
  __shared__ uint32_t sh[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    int64_t g = base + threadIdx.x;
    uint32_t v = g < n ? data[g] : 0;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(v, &tot, sh);
    uint32_t c = carry;
    if (g < n) data[g] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                           int64_t n, const uint32_t* __restrict__ sums) {
  __shared__ uint32_t sh[32];
  // thread owns SC_IPT consecutive items (blocked arrangement) so that the scan is ordered
  int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_IPT;
  uint32_t v[SC_IPT];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    int64_t g = base + k;
    v[k] = g < n ? in[g] : 0;
    s += v[k];
  }
  uint32_t ex = block_exclusive_scan(s, nullptr, sh) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_IPT; ++k) {
    int64_t g = base + k;
    if (g < n) out[g] = ex;
    ex += v[k];
  }
}

// out[i] = sum_{j<i} in[j]; in == out allowed.  tmp holds the per-tile sums.
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, DevBuf<uint32_t>& tmp, cudaStream_t s) {
  if (n <= 0) return;
  int64_t tiles = (n + SC_TILE - 1) / SC_TILE;
  tmp.alloc((size_t)tiles);
  k_scan_tile_sums<<<(unsigned)tiles, SC_THREADS, 0, s>>>(in, n, tmp.p);
  k_scan_single<<<1, 1024, 0, s>>>(tmp.p, tiles);
  k_scan_apply<<<(unsigned)tiles, SC_THREADS, 0, s>>>(in, out, n, tmp.p);
  FW_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------ radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // 4096 keys per block, 512 consecutive per warp

__global__ void __launch_bounds__(RS_THREADS, 3) k_rs_hist(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                        uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
  uint64_t key[RS_IPT];  // all keys of the thread first: the ranking below would otherwise wait for one load at a time
#pragma unroll
  for (int k = 0; k < RS_IPT; ++k) {
    int64_t g = base + (int64_t)k * RS_THREADS + threadIdx.x;
    key[k] = g < n ? keys[g] : 0;
  }
#pragma unroll
  for (int k = 0; k < RS_IPT; ++k) {
    int64_t g = base + (int64_t)k * RS_THREADS + threadIdx.x;
    bool valid = g < n;
    unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
      unsigned m = __match_any_sync(vm, d);
      if (lane == __ffs(m) - 1) atomicAdd(&sh[d], (uint32_t)__popc(m));
    }
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// Stable scatter: warp w of a block owns the 512 consecutive keys [w*512, (w+1)*512) of the
// tile, walks them in rounds of 32, ranks equal digits with match_any.
__global__ void __launch_bounds__(RS_THREADS, 3) k_rs_scatter(const uint64_t* __restrict__ kin,
                                                           const uint32_t* __restrict__ vin,
                                                           uint64_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                           int64_t n, int shift, const uint32_t* __restrict__ offs,
                                                           int nblocks) {
  __shared__ uint32_t wh[RS_WARPS][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_TILE / RS_WARPS);
  uint64_t key[RS_IPT];
  uint32_t val[RS_IPT];
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {  // all loads first (see k_rs_hist)
    int64_t g = wbase + r * 32 + lane;
    key[r] = g < n ? kin[g] : 0;
    val[r] = (vin && g < n) ? vin[g] : (uint32_t)g;
  }
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    int64_t g = wbase + r * 32 + lane;
    bool valid = g < n;
    unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
      unsigned m = __match_any_sync(vm, d);
      if (lane == __ffs(m) - 1) wh[warp][d] += (uint32_t)__popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  {
    const int bin = threadIdx.x;  // RS_THREADS == 256 bins
    uint32_t run = offs[(size_t)bin * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = wh[w][bin];
      wh[w][bin] = run;
      run += c;
    }
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    int64_t g = wbase + r * 32 + lane;
    bool valid = g < n;
    unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
      unsigned m = __match_any_sync(vm, d);
      uint32_t pos = wh[warp][d] + (uint32_t)__popc(m & lt);
      kout[pos] = key[r];
      vout[pos] = val[r];
      __syncwarp(vm);
      if (lane == __ffs(m) - 1) wh[warp][d] += (uint32_t)__popc(m);
    }
    __syncwarp();
  }
}

// sorts (keys, payload) by the low `bits` bits of key; payload_in == nullptr means iota.
// Results end up in keys_a/pay_a (buffers are swapped internally as needed).
static void radix_sort_pairs(DevBuf<uint64_t>& keys_a, DevBuf<uint64_t>& keys_b, DevBuf<uint32_t>& pay_a,
                             DevBuf<uint32_t>& pay_b, int64_t n, int bits, bool payload_iota,
                             DevBuf<uint32_t>& hist, DevBuf<uint32_t>& scan_tmp, cudaStream_t s) {
  if (n <= 0) return;
  int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  hist.alloc((size_t)256 * nblocks);
  keys_b.alloc((size_t)n);
  pay_a.alloc((size_t)n);
  pay_b.alloc((size_t)n);
  int passes = (bits + 7) / 8;
  if (passes < 1) passes = 1;
  bool first = true;
  for (int pass = 0; pass < passes; ++pass) {
    int shift = pass * 8;
    k_rs_hist<<<nblocks, RS_THREADS, 0, s>>>(keys_a.p, n, shift, hist.p, nblocks);
    exclusive_scan_u32(hist.p, hist.p, (int64_t)256 * nblocks, scan_tmp, s);
    k_rs_scatter<<<nblocks, RS_THREADS, 0, s>>>(keys_a.p, (first && payload_iota) ? nullptr : pay_a.p, keys_b.p,
                                                pay_b.p, n, shift, hist.p, nblocks);
    std::swap(keys_a, keys_b);
    std::swap(pay_a, pay_b);
    first = false;
  }
  FW_CHECK_CUDA(cudaGetLastError());
}

// exported for the device-side nested-dissection analysis (lu_symbolic_gpu.cu)
void radix_sort_pairs_dev(Handle& h, DevBuf<uint64_t>& keys_a, DevBuf<uint64_t>& keys_b, DevBuf<uint32_t>& pay_a,
                          DevBuf<uint32_t>& pay_b, int64_t n, int bits, bool payload_iota) {
  DevBuf<uint32_t> hist;
  radix_sort_pairs(keys_a, keys_b, pay_a, pay_b, n, bits, payload_iota, hist, h.scan_tmp, h.stream);
}

// ------------------------------------------------------------------------------ pattern from sorted keys
__global__ void k_heads(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void k_fill_pattern(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ segid, int64_t n,
                               int col_bits, int n_rows, int64_t nnz, int* __restrict__ indptr,
                               int* __restrict__ indices, uint32_t* __restrict__ seg) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = keys[i];
  bool head = (i == 0) || (k != keys[i - 1]);
  if (head) {
    uint32_t s = segid[i];
    seg[s] = (uint32_t)i;
    indices[s] = (int)(k & ((1ull << col_bits) - 1ull));
    int row = (int)(k >> col_bits);
    int prev = (i == 0) ? -1 : (int)(keys[i - 1] >> col_bits);
    for (int r = prev + 1; r <= row; ++r) indptr[r] = (int)s;
  }
  if (i == n - 1) {
    int row = (int)(k >> col_bits);
    for (int r = row + 1; r <= n_rows; ++r) indptr[r] = (int)nnz;
    seg[nnz] = (uint32_t)n;
  }
}

__global__ void k_fill_empty_indptr(int* indptr, int n_rows) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_rows) indptr[i] = 0;
}

// keys: sorted on return (first n_coo are the valid ones).  payload becomes pat.perm.
void build_pattern_from_keys(Pattern& pat, DevBuf<uint64_t>& keys, DevBuf<uint32_t>& payload, int64_t n_coo,
                             int n_rows, int col_bits, Handle& h) {
  cudaStream_t s = h.stream;
  pat.n = n_rows;
  pat.n_coo = n_coo;
  pat.indptr.alloc((size_t)n_rows + 1);
  if (n_coo == 0) {
    pat.nnz = 0;
    k_fill_empty_indptr<<<cdiv(n_rows + 1, 256), 256, 0, s>>>(pat.indptr.p, n_rows);
    pat.seg.alloc(1);
    pat.seg.zero(s);
    pat.perm = std::move(payload);
    return;
  }
  DevBuf<uint32_t> flag;
  flag.alloc((size_t)n_coo);
  k_heads<<<cdiv(n_coo, 256), 256, 0, s>>>(keys.p, n_coo, flag.p);
  uint32_t last_flag = 0, last_scan = 0;
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_flag, flag.p + (n_coo - 1), 4, cudaMemcpyDeviceToHost, s));
  exclusive_scan_u32(flag.p, flag.p, n_coo, h.scan_tmp, s);
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_scan, flag.p + (n_coo - 1), 4, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  pat.nnz = (int64_t)last_scan + last_flag;
  pat.indices.alloc((size_t)pat.nnz);
  pat.seg.alloc((size_t)pat.nnz + 1);
  k_fill_pattern<<<cdiv(n_coo, 256), 256, 0, s>>>(keys.p, flag.p, n_coo, col_bits, n_rows, pat.nnz, pat.indptr.p,
                                                  pat.indices.p, pat.seg.p);
  FW_CHECK_CUDA(cudaGetLastError());
  pat.perm = std::move(payload);
}

static int bit_length(int64_t v) {
  int b = 0;
  while (v > 0) {
    ++b;
    v >>= 1;
  }
  return b < 1 ? 1 : b;
}

// ------------------------------------------------------------------------------ structural symbolic of the FE space
__global__ void k_space_keys(const int* __restrict__ edofs, int ne, int nb, int col_bits, uint64_t* __restrict__ keys) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)nb * nb * ne;
  if (idx >= total) return;
  int k = (int)(idx / ne);
  int e = (int)(idx - (int64_t)k * ne);
  int i = k / nb, j = k - i * nb;  // row = test dof i, col = trial dof j
  uint64_t row = (uint64_t)edofs[(size_t)i * ne + e];
  uint64_t col = (uint64_t)edofs[(size_t)j * ne + e];
  keys[idx] = (row << col_bits) | col;
}

// ------------------------------------------------------------------------------ pattern from the dof -> element incidence
// The structural pattern of the FE space does not need a sort of all nb^2 * Ne COO keys (130 M 12-byte pairs through six
// radix passes = 14 GB of traffic for a 3.4 GB job, 19 ms at config 2).  Row r only receives entries from the elements
// that contain dof r: the nb * Ne (dof, local index, element) incidences are sorted by dof (7 M keys, three radix
// passes), then ONE WARP PER ROW gathers its <= 512 candidate (column, source entry) pairs, sorts them in shared memory
// (bitonic, key = column << 32 | source index) and emits the unique columns, the segment starts and the source
// permutation.  The result is identical to the stable radix sort by (row, column): rows ascending, columns ascending,
// sources of a slot in ascending source index -- the summation order of the segmented reduce does not change.
constexpr int RP_WARPS = 8, RP_MAX = 512;

__global__ void k_inc_keys(const int* __restrict__ edofs, int64_t nq, uint64_t* __restrict__ keys) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq) keys[q] = (uint64_t)edofs[q];
}

// rstart[r] = first sorted incidence of dof r (rstart[n] = nq)
__global__ void k_row_starts(const uint64_t* __restrict__ keys, int64_t nq, int n, int* __restrict__ rstart) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nq) return;
  const int d = (int)keys[p];
  const int prev = p == 0 ? -1 : (int)keys[p - 1];
  if (d != prev)
    for (int r = prev + 1; r <= d; ++r) rstart[r] = (int)p;
  if (p == nq - 1)
    for (int r = d + 1; r <= n; ++r) rstart[r] = (int)nq;
}

template <bool FILL>
__global__ void __launch_bounds__(32 * RP_WARPS) k_row_pattern(int n, int ne, int nb, const int* __restrict__ edofs,
                                                               const int* __restrict__ rstart,
                                                               const uint32_t* __restrict__ inc,
                                                               uint32_t* __restrict__ count,
                                                               const uint32_t* __restrict__ slot0,
                                                               int* __restrict__ indptr, int* __restrict__ indices,
                                                               uint32_t* __restrict__ seg, uint32_t* __restrict__ perm,
                                                               int* __restrict__ overflow) {
  __shared__ unsigned long long sk[RP_WARPS][RP_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * RP_WARPS + warp;
  if (row >= n) return;  // (whole warps leave; no block-wide barrier below)
  const int p0 = rstart[row], p1 = rstart[row + 1];
  const int cand = (p1 - p0) * nb;
  if (cand > RP_MAX) {
    if (lane == 0) atomicExch(overflow, 1);
    if (!FILL && lane == 0) count[row] = 0;
    return;
  }
  int P = 32;
  while (P < cand) P <<= 1;
  unsigned long long* a = sk[warp];
  for (int t = lane; t < P; t += 32) {
    unsigned long long key = ~0ull;
    if (t < cand) {
      const int ii = t / nb, j = t - ii * nb;
      const uint32_t q = inc[p0 + ii];
      const int l = (int)(q / (uint32_t)ne), e = (int)(q - (uint32_t)l * (uint32_t)ne);
      const unsigned long long col = (unsigned long long)edofs[(size_t)j * ne + e];
      const unsigned long long src = (unsigned long long)(l * nb + j) * (unsigned long long)ne + (unsigned long long)e;
      key = (col << 32) | src;
    }
    a[t] = key;
  }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          if ((x > y) == ((i & k) == 0)) a[i] = y, a[ixj] = x;
        }
      }
      __syncwarp();
    }
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t base = (uint32_t)((unsigned long long)nb * (unsigned long long)p0);
  uint32_t hoff = 0;
  const uint32_t s0 = FILL ? slot0[row] : 0u;
  if (FILL && lane == 0) indptr[row] = (int)s0;
  for (int c0 = 0; c0 < cand; c0 += 32) {
    const int t = c0 + lane;
    const bool valid = t < cand;
    const unsigned long long kt = valid ? a[t] : 0ull;
    const bool head = valid && (t == 0 || (a[t - 1] >> 32) != (kt >> 32));
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (FILL) {
      if (valid) perm[base + (uint32_t)t] = (uint32_t)(kt & 0xffffffffull);
      if (head) {
        const uint32_t sidx = s0 + hoff + (uint32_t)__popc(m & lt);
        indices[sidx] = (int)(kt >> 32);
        seg[sidx] = base + (uint32_t)t;
      }
    }
    hoff += (uint32_t)__popc(m);
  }
  if (!FILL && lane == 0) count[row] = hoff;
}

__global__ void k_pattern_tail(int n, int64_t nnz, int64_t n_coo, int* __restrict__ indptr, uint32_t* __restrict__ seg) {
  indptr[n] = (int)nnz;
  seg[nnz] = (uint32_t)n_coo;
}

// returns false when a row has more than RP_MAX candidates (the caller falls back to the full sort)
static bool symbolic_space_incidence(Handle& h, int64_t n_coo) {
  cudaStream_t s = h.stream;
  const Space& sp = h.sp;
  Pattern& pat = h.pat;
  const int64_t nq = (int64_t)sp.nb * sp.ne;
  const int cb = [&] { int b = 0; for (int64_t v = sp.n; v > 0; v >>= 1) ++b; return b < 1 ? 1 : b; }();
  Trace tr(s);
  DevBuf<uint64_t> ka, kb;
  DevBuf<uint32_t> pa, pb, hist, count;
  DevBuf<int> rstart, overflow;
  ka.alloc((size_t)nq);
  k_inc_keys<<<cdiv(nq, 256), 256, 0, s>>>(sp.edofs.p, nq, ka.p);
  radix_sort_pairs(ka, kb, pa, pb, nq, cb, true, hist, h.scan_tmp, s);
  rstart.alloc((size_t)sp.n + 1);
  k_row_starts<<<cdiv(nq, 256), 256, 0, s>>>(ka.p, nq, sp.n, rstart.p);
  tr.mark("symbolic: dof -> element incidence");
  overflow.alloc(1);
  overflow.zero(s);
  count.alloc((size_t)sp.n + 1);
  const int grid = cdiv(sp.n, RP_WARPS);
  k_row_pattern<false><<<grid, 32 * RP_WARPS, 0, s>>>(sp.n, sp.ne, sp.nb, sp.edofs.p, rstart.p, pa.p, count.p, nullptr,
                                                     nullptr, nullptr, nullptr, nullptr, overflow.p);
  int ovf = 0;
  uint32_t last_cnt = 0, last_scan = 0;
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_cnt, count.p + (sp.n - 1), 4, cudaMemcpyDeviceToHost, s));
  exclusive_scan_u32(count.p, count.p, sp.n, h.scan_tmp, s);
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_scan, count.p + (sp.n - 1), 4, cudaMemcpyDeviceToHost, s));
  overflow.download(&ovf, 1, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  if (ovf) return false;
  pat.n = sp.n;
  pat.n_coo = n_coo;
  pat.nnz = (int64_t)last_scan + last_cnt;
  pat.indptr.alloc((size_t)sp.n + 1);
  pat.indices.alloc((size_t)std::max<int64_t>(1, pat.nnz));
  pat.seg.alloc((size_t)pat.nnz + 1);
  pat.perm.alloc((size_t)n_coo);
  tr.mark("symbolic: row counts");
  k_row_pattern<true><<<grid, 32 * RP_WARPS, 0, s>>>(sp.n, sp.ne, sp.nb, sp.edofs.p, rstart.p, pa.p, nullptr, count.p,
                                                    pat.indptr.p, pat.indices.p, pat.seg.p, pat.perm.p, overflow.p);
  k_pattern_tail<<<1, 1, 0, s>>>(sp.n, pat.nnz, n_coo, pat.indptr.p, pat.seg.p);
  FW_CHECK_CUDA(cudaGetLastError());
  tr.mark("symbolic: row patterns");
  return true;
}

void symbolic_space(Handle& h) {
  FW_REQUIRE(h.sp.valid, "fw_set_space must be called first");
  cudaStream_t s = h.stream;
  const Space& sp = h.sp;
  int64_t n_coo = (int64_t)sp.nb * sp.nb * sp.ne;
  FW_REQUIRE(n_coo < (int64_t)0xffffffffll, "too many COO entries for 32-bit payload");
  int cb = bit_length(sp.n);
  bool done = false;
  if (!getenv("FW_SYMBOLIC_FULL_SORT") && sp.n > 0) done = symbolic_space_incidence(h, n_coo);
  if (!done) {
    Trace tr(s);
    DevBuf<uint64_t> ka, kb;
    DevBuf<uint32_t> pa, pb, hist;
    ka.alloc((size_t)n_coo);
    kb.alloc((size_t)n_coo);
    pa.alloc((size_t)n_coo);
    pb.alloc((size_t)n_coo);
    tr.mark("symbolic: alloc");
    k_space_keys<<<cdiv(n_coo, 256), 256, 0, s>>>(sp.edofs.p, sp.ne, sp.nb, cb, ka.p);
    tr.mark("symbolic: keys");
    radix_sort_pairs(ka, kb, pa, pb, n_coo, 2 * cb, true, hist, h.scan_tmp, s);
    tr.mark("symbolic: radix sort");
    build_pattern_from_keys(h.pat, ka, pa, n_coo, sp.n, cb, h);
    tr.mark("symbolic: pattern");
  }
  h.has_symbolic = true;
  h.pat.n_tt = -1;
  h.A.valid = h.B.valid = h.Mass.valid = h.C0.valid = h.C1.valid = h.MassC.valid = h.Bc.valid = false;
}

// ------------------------------------------------------------------------------ segmented reduce
// One thread per SEG_ILP CSR slots; contributions are summed in sorted (= original COO) order, so the result
// is deterministic.  lut (optional) maps the local-block entry index k = src / stride to the compact entry
// index of a form that stores only some sub-blocks (or -1).  slots (optional) restricts the work to a list of
// slots (bform touches only the Nedelec x Nedelec third of the pattern).
// The kernel is a chain of dependent gathers (seg -> perm -> value); ncu of the one-slot-per-thread version
// showed 46 long-scoreboard stall cycles per issue at 29 % DRAM utilisation, hence SEG_ILP independent
// chains per thread: the loads of one stage are issued for all slots before the next stage starts.
constexpr int SEG_ILP = 4;
template <class S>
__global__ void __launch_bounds__(256) k_segreduce(const uint32_t* __restrict__ seg, const uint32_t* __restrict__ perm,
                                                   int64_t nwork, const int* __restrict__ slots,
                                                   const S* __restrict__ vals, const int* __restrict__ lut,
                                                   int64_t stride, S* __restrict__ out, uint8_t* __restrict__ flag) {
  const int64_t base = (int64_t)blockIdx.x * (256 * SEG_ILP) + threadIdx.x;
  int64_t sid[SEG_ILP];
  uint32_t a[SEG_ILP], b[SEG_ILP];
  bool valid[SEG_ILP];
#pragma unroll
  for (int u = 0; u < SEG_ILP; ++u) {
    const int64_t w = base + (int64_t)u * 256;
    valid[u] = w < nwork;
    sid[u] = valid[u] ? (slots ? (int64_t)slots[w] : w) : 0;
  }
#pragma unroll
  for (int u = 0; u < SEG_ILP; ++u) {
    a[u] = valid[u] ? seg[sid[u]] : 0;
    b[u] = valid[u] ? seg[sid[u] + 1] : 0;
  }
  uint32_t src[SEG_ILP];
#pragma unroll
  for (int u = 0; u < SEG_ILP; ++u) src[u] = (valid[u] && a[u] < b[u]) ? perm[a[u]] : 0;
  // address of the first contribution of every slot
  int64_t addr[SEG_ILP];
#pragma unroll
  for (int u = 0; u < SEG_ILP; ++u) {
    addr[u] = -1;
    if (valid[u] && a[u] < b[u]) {
      if (lut) {
        const uint32_t k = (uint32_t)(src[u] / stride);
        const int kk = lut[k];
        if (kk >= 0) addr[u] = (int64_t)kk * stride + (src[u] - (int64_t)k * stride);
      } else {
        addr[u] = src[u];
      }
    }
  }
  S v0[SEG_ILP];
#pragma unroll
  for (int u = 0; u < SEG_ILP; ++u) v0[u] = addr[u] >= 0 ? vals[addr[u]] : scalar_traits<S>::zero();
#pragma unroll
  for (int u = 0; u < SEG_ILP; ++u) {
    if (!valid[u]) continue;
    S acc = v0[u];
    bool any = addr[u] >= 0;
    bool nz = any && !is_zero_(v0[u]);
    for (uint32_t i = a[u] + 1; i < b[u]; ++i) {  // further contributions (vertex / edge dofs shared by elements)
      const uint32_t sr = perm[i];
      S v;
      if (lut) {
        const uint32_t k = (uint32_t)(sr / stride);
        const int kk = lut[k];
        if (kk < 0) continue;
        v = vals[(int64_t)kk * stride + (sr - (int64_t)k * stride)];
      } else {
        v = vals[sr];
      }
      any = true;
      nz |= !is_zero_(v);
      acc = acc + v;
    }
    out[sid[u]] = acc;
    // bit 0: structurally present, bit 1: at least one non-zero contribution (survives eliminate_zeros)
    flag[sid[u]] = (any ? 1 : 0) | (nz ? 2 : 0);
  }
}

template <class S>
void reduce_onto_pattern(const Pattern& pat, const S* coo_vals, const int* lut, int lut_n, int64_t stride, S* out,
                         uint8_t* flag, cudaStream_t s, const int* slots, int64_t nslots) {
  (void)lut_n;
  if (pat.nnz == 0) return;
  int64_t nwork = pat.nnz;
  if (slots) {
    // slots outside the list receive no contribution
    FW_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)pat.nnz * sizeof(S), s));
    FW_CHECK_CUDA(cudaMemsetAsync(flag, 0, (size_t)pat.nnz, s));
    nwork = nslots;
    if (nwork == 0) return;
  }
  k_segreduce<S><<<cdiv(nwork, 256 * SEG_ILP), 256, 0, s>>>(pat.seg.p, pat.perm.p, nwork, slots, coo_vals, lut, stride,
                                                            out, flag);
  FW_CHECK_CUDA(cudaGetLastError());
}
template void reduce_onto_pattern<double>(const Pattern&, const double*, const int*, int, int64_t, double*, uint8_t*,
                                          cudaStream_t, const int*, int64_t);
template void reduce_onto_pattern<cplx>(const Pattern&, const cplx*, const int*, int, int64_t, cplx*, uint8_t*,
                                        cudaStream_t, const int*, int64_t);

// slots whose contributions belong to the sub-blocks selected by lut (all contributions of a slot share the kinds
// of its row and column dof, so the first one decides)
__global__ void k_slot_in_lut(const uint32_t* __restrict__ seg, const uint32_t* __restrict__ perm, int64_t nnz,
                              const int* __restrict__ lut, int64_t stride, uint32_t* __restrict__ mark) {
  int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= nnz) return;
  const uint32_t a = seg[sidx], b = seg[sidx + 1];
  mark[sidx] = (a < b && lut[perm[a] / stride] >= 0) ? 1u : 0u;
}
__global__ void k_compact_slots(const uint32_t* __restrict__ mark_scan, const uint32_t* __restrict__ seg,
                                const uint32_t* __restrict__ perm, int64_t nnz, const int* __restrict__ lut,
                                int64_t stride, int* __restrict__ list) {
  int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= nnz) return;
  const uint32_t a = seg[sidx], b = seg[sidx + 1];
  if (a < b && lut[perm[a] / stride] >= 0) list[mark_scan[sidx]] = (int)sidx;
}

void build_slot_list(Handle& h, Pattern& pat, const int* lut_dev, int64_t stride) {
  cudaStream_t s = h.stream;
  if (pat.n_tt >= 0 || pat.nnz == 0) return;
  h.scan_tmp2.alloc((size_t)pat.nnz);
  k_slot_in_lut<<<cdiv(pat.nnz, 256), 256, 0, s>>>(pat.seg.p, pat.perm.p, pat.nnz, lut_dev, stride, h.scan_tmp2.p);
  uint32_t last_mark = 0, last_scan = 0;
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_mark, h.scan_tmp2.p + (pat.nnz - 1), 4, cudaMemcpyDeviceToHost, s));
  exclusive_scan_u32(h.scan_tmp2.p, h.scan_tmp2.p, pat.nnz, h.scan_tmp, s);
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_scan, h.scan_tmp2.p + (pat.nnz - 1), 4, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  pat.n_tt = (int64_t)last_scan + last_mark;
  pat.tt_slots.alloc((size_t)std::max<int64_t>(1, pat.n_tt));
  k_compact_slots<<<cdiv(pat.nnz, 256), 256, 0, s>>>(h.scan_tmp2.p, pat.seg.p, pat.perm.p, pat.nnz, lut_dev, stride,
                                                     pat.tt_slots.p);
  FW_CHECK_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------ eliminate_zeros compaction
__global__ void k_flag_to_u32(const uint8_t* __restrict__ f, int64_t n, uint8_t mask, uint32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (f[i] & mask) ? 1u : 0u;
}

template <class S>
__global__ void k_compact(const uint8_t* __restrict__ flag, uint8_t mask, const uint32_t* __restrict__ pos,
                          int64_t nnz, const int* __restrict__ indices, const S* __restrict__ vals,
                          int* __restrict__ oind, S* __restrict__ oval) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz || !(flag[i] & mask)) return;
  uint32_t p = pos[i];
  oind[p] = indices[i];
  oval[p] = vals[i];
}

__global__ void k_compact_indptr(const int* __restrict__ indptr, int n, const uint32_t* __restrict__ pos, int64_t nnz,
                                 int64_t kept, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int a = indptr[i];
  out[i] = (a >= nnz) ? (int)kept : (int)pos[a];
}

static int64_t count_kept_mask(Handle& h, const Pattern& pat, Values& val, uint8_t mask) {
  cudaStream_t s = h.stream;
  if (pat.nnz == 0) return 0;
  h.scan_tmp2.alloc((size_t)pat.nnz);
  k_flag_to_u32<<<cdiv(pat.nnz, 256), 256, 0, s>>>(val.flag.p, pat.nnz, mask, h.scan_tmp2.p);
  exclusive_scan_u32(h.scan_tmp2.p, h.scan_tmp2.p, pat.nnz, h.scan_tmp, s);
  uint32_t last_scan = 0;
  uint8_t last_flag = 0;
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_scan, h.scan_tmp2.p + (pat.nnz - 1), 4, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaMemcpyAsync(&last_flag, val.flag.p + (pat.nnz - 1), 1, cudaMemcpyDeviceToHost, s));
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  return (int64_t)last_scan + ((last_flag & mask) ? 1 : 0);
}

int64_t count_kept(Handle& h, const Pattern& pat, Values& val) {
  val.nnz_kept = count_kept_mask(h, pat, val, h.eliminate_zeros ? 2 : 1);
  return val.nnz_kept;
}

int64_t compact_csr_dev(Handle& h, const Pattern& pat, Values& val, uint8_t mask, DevBuf<int>& oip, DevBuf<int>& oind,
                        DevBuf<double>& oval) {
  cudaStream_t s = h.stream;
  const int64_t kept = count_kept_mask(h, pat, val, mask);  // leaves the scanned positions in scan_tmp2
  oip.alloc((size_t)pat.n + 1);
  oind.alloc((size_t)(kept ? kept : 1));
  const size_t w = val.is_complex ? 2 : 1;
  oval.alloc((size_t)(kept ? kept : 1) * w);
  if (pat.nnz) {
    if (val.is_complex)
      k_compact<cplx><<<cdiv(pat.nnz, 256), 256, 0, s>>>(val.flag.p, mask, h.scan_tmp2.p, pat.nnz, pat.indices.p,
                                                        (const cplx*)val.v.p, oind.p, (cplx*)oval.p);
    else
      k_compact<double><<<cdiv(pat.nnz, 256), 256, 0, s>>>(val.flag.p, mask, h.scan_tmp2.p, pat.nnz, pat.indices.p,
                                                          val.v.p, oind.p, oval.p);
    k_compact_indptr<<<cdiv(pat.n + 1, 256), 256, 0, s>>>(pat.indptr.p, pat.n, h.scan_tmp2.p, pat.nnz, kept, oip.p);
  } else {
    k_fill_empty_indptr<<<cdiv(pat.n + 1, 256), 256, 0, s>>>(oip.p, pat.n);
  }
  FW_CHECK_CUDA(cudaGetLastError());
  return kept;
}

void export_compacted(Handle& h, const Pattern& pat, Values& val, int* indptr, int* indices, void* data) {
  cudaStream_t s = h.stream;
  DevBuf<int> oip, oind;
  DevBuf<double> oval;
  const int64_t kept = compact_csr_dev(h, pat, val, h.eliminate_zeros ? 2 : 1, oip, oind, oval);
  val.nnz_kept = kept;
  const size_t w = val.is_complex ? 2 : 1;
  oip.download(indptr, (size_t)pat.n + 1, s);
  oind.download(indices, (size_t)kept, s);
  oval.download((double*)data, (size_t)kept * w, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
}

// ------------------------------------------------------------------------------ generic COO -> CSR
template <class S>
__global__ void k_coo_keys(const int* __restrict__ rows, const int* __restrict__ cols, const S* __restrict__ vals,
                           int64_t n, int col_bits, int n_rows, uint64_t* __restrict__ keys,
                           unsigned long long* __restrict__ n_zero) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool z = is_zero_(vals[i]);
  // exact zeros are dropped BEFORE duplicates are summed (eliminate_zeros precedes tocsr): they get a
  // sentinel key that sorts behind every real key
  keys[i] = z ? ((uint64_t)n_rows << col_bits) : (((uint64_t)rows[i] << col_bits) | (uint64_t)cols[i]);
  if (z) atomicAdd(n_zero, 1ull);
}

void coo_to_csr(Handle& h, int n_rows, int64_t n_entries, const int* rows, const int* cols, const void* data,
                bool is_complex) {
  cudaStream_t s = h.stream;
  FW_REQUIRE(n_entries < (int64_t)0xffffffffll, "too many COO entries");
  DevBuf<int> dr, dc;
  DevBuf<double> dv;
  size_t w = is_complex ? 2 : 1;
  dr.upload(rows, (size_t)n_entries, s);
  dc.upload(cols, (size_t)n_entries, s);
  dv.upload((const double*)data, (size_t)n_entries * w, s);
  coo_to_csr_dev(h, n_rows, n_entries, dr, dc, dv, is_complex);
}

// the same on device-resident triplets (result in h.coo_pat / h.coo_val)
void coo_to_csr_dev(Handle& h, int n_rows, int64_t n_entries, DevBuf<int>& dr, DevBuf<int>& dc, DevBuf<double>& dv,
                    bool is_complex) {
  cudaStream_t s = h.stream;
  FW_REQUIRE(n_entries < (int64_t)0xffffffffll, "too many COO entries");
  size_t w = is_complex ? 2 : 1;
  int cb = bit_length(n_rows);  // n_rows < 2^cb strictly, so the sentinel row n_rows fits
  DevBuf<uint64_t> ka, kb;
  DevBuf<uint32_t> pa, pb, hist;
  DevBuf<unsigned long long> nz;
  nz.alloc(1);
  nz.zero(s);
  ka.alloc((size_t)(n_entries ? n_entries : 1));
  if (n_entries) {
    if (is_complex)
      k_coo_keys<cplx><<<cdiv(n_entries, 256), 256, 0, s>>>(dr.p, dc.p, (const cplx*)dv.p, n_entries, cb, n_rows, ka.p,
                                                           nz.p);
    else
      k_coo_keys<double><<<cdiv(n_entries, 256), 256, 0, s>>>(dr.p, dc.p, dv.p, n_entries, cb, n_rows, ka.p, nz.p);
  }
  unsigned long long n_zero = 0;
  nz.download(&n_zero, 1, s);
  radix_sort_pairs(ka, kb, pa, pb, n_entries, 2 * cb, true, hist, h.scan_tmp, s);
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
  int64_t n_valid = n_entries - (int64_t)n_zero;
  build_pattern_from_keys(h.coo_pat, ka, pa, n_valid, n_rows, cb, h);
  Values& val = h.coo_val;
  val.is_complex = is_complex;
  val.v.alloc((size_t)(h.coo_pat.nnz ? h.coo_pat.nnz : 1) * w);
  val.flag.alloc((size_t)(h.coo_pat.nnz ? h.coo_pat.nnz : 1));
  if (is_complex)
    reduce_onto_pattern<cplx>(h.coo_pat, (const cplx*)dv.p, nullptr, 0, 1, (cplx*)val.v.p, val.flag.p, s);
  else
    reduce_onto_pattern<double>(h.coo_pat, dv.p, nullptr, 0, 1, val.v.p, val.flag.p, s);
  // every kept slot has at least one non-zero contribution by construction; a sum that cancels
  // to 0.0 stays as an explicit zero exactly like scipy's tocsr()
  if (h.coo_pat.nnz) FW_CHECK_CUDA(cudaMemsetAsync(val.flag.p, 3, (size_t)h.coo_pat.nnz, s));
  val.valid = true;
  FW_CHECK_CUDA(cudaStreamSynchronize(s));
}

}  // namespace fw
