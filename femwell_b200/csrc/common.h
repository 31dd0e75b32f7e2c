// Shared internals of libfemwell_b200: error handling, scalar types, device buffers, handle.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/femwell_b200.h"

namespace fw {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FW_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      throw fw::Error(FW_ERR_CUDA, std::string(#expr) + " -> " + cudaGetErrorString(_e) +    \
                                       " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

#define FW_REQUIRE(cond, msg)                                                      \
  do {                                                                             \
    if (!(cond)) throw fw::Error(FW_ERR_INVALID, std::string(msg) + " [" #cond "]"); \
  } while (0)

// ------------------------------------------------------------------------------ complex
struct cplx {
  double re, im;
};
__host__ __device__ inline cplx mk(double r, double i = 0.0) { return cplx{r, i}; }
__host__ __device__ inline cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__host__ __device__ inline cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__host__ __device__ inline cplx operator-(cplx a) { return {-a.re, -a.im}; }
__host__ __device__ inline cplx operator*(cplx a, cplx b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ inline cplx operator*(cplx a, double b) { return {a.re * b, a.im * b}; }
__host__ __device__ inline cplx operator*(double b, cplx a) { return {a.re * b, a.im * b}; }
__host__ __device__ inline cplx operator+(cplx a, double b) { return {a.re + b, a.im}; }
__host__ __device__ inline cplx operator/(cplx a, double b) { return {a.re / b, a.im / b}; }
__host__ __device__ inline cplx operator/(cplx a, cplx b) {
  // Smith's algorithm (avoids overflow, good accuracy)
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re, d = b.re + b.im * r;
    return {(a.re + a.im * r) / d, (a.im - a.re * r) / d};
  }
  double r = b.re / b.im, d = b.re * r + b.im;
  return {(a.re * r + a.im) / d, (a.im * r - a.re) / d};
}
__host__ __device__ inline cplx operator/(double a, cplx b) { return mk(a) / b; }
__host__ __device__ inline cplx& operator+=(cplx& a, cplx b) {
  a.re += b.re;
  a.im += b.im;
  return a;
}
__host__ __device__ inline cplx& operator-=(cplx& a, cplx b) {
  a.re -= b.re;
  a.im -= b.im;
  return a;
}
__host__ __device__ inline cplx conj_(cplx a) { return {a.re, -a.im}; }
__host__ __device__ inline double conj_(double a) { return a; }
__host__ __device__ inline double abs2_(cplx a) { return a.re * a.re + a.im * a.im; }
__host__ __device__ inline double abs2_(double a) { return a * a; }
__host__ __device__ inline double absv_(cplx a) { return hypot(a.re, a.im); }
__host__ __device__ inline double absv_(double a) { return fabs(a); }
// cheap magnitude used for pivot search (|re|+|im|, as LAPACK's cabs1)
__host__ __device__ inline double abs1_(cplx a) { return fabs(a.re) + fabs(a.im); }
__host__ __device__ inline double abs1_(double a) { return fabs(a); }
__host__ __device__ inline bool is_zero_(cplx a) { return a.re == 0.0 && a.im == 0.0; }
__host__ __device__ inline bool is_zero_(double a) { return a == 0.0; }
// acc += a*b with fused multiply-adds
__host__ __device__ inline void fma_(double& acc, double a, double b) { acc = fma(a, b, acc); }
__host__ __device__ inline void fma_(cplx& acc, cplx a, double b) {
  acc.re = fma(a.re, b, acc.re);
  acc.im = fma(a.im, b, acc.im);
}
__host__ __device__ inline void fma_(cplx& acc, double a, cplx b) { fma_(acc, b, a); }
__host__ __device__ inline void fma_(cplx& acc, cplx a, cplx b) {
  acc.re = fma(a.re, b.re, acc.re);
  acc.re = fma(-a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im);
  acc.im = fma(a.im, b.re, acc.im);
}
// acc -= a*b
__host__ __device__ inline void fnma_(double& acc, double a, double b) { acc = fma(-a, b, acc); }
__host__ __device__ inline void fnma_(cplx& acc, cplx a, cplx b) {
  acc.re = fma(-a.re, b.re, acc.re);
  acc.re = fma(a.im, b.im, acc.re);
  acc.im = fma(-a.re, b.im, acc.im);
  acc.im = fma(-a.im, b.re, acc.im);
}
__host__ __device__ inline void fnma_(cplx& acc, double a, cplx b) {
  acc.re = fma(-a, b.re, acc.re);
  acc.im = fma(-a, b.im, acc.im);
}
__host__ __device__ inline void fnma_(cplx& acc, cplx a, double b) { fnma_(acc, b, a); }

template <class T>
struct scalar_traits;
template <>
struct scalar_traits<double> {
  static constexpr bool is_complex = false;
  __host__ __device__ static double zero() { return 0.0; }
  __host__ __device__ static double one() { return 1.0; }
  __host__ __device__ static double from(double r, double) { return r; }
};
template <>
struct scalar_traits<cplx> {
  static constexpr bool is_complex = true;
  __host__ __device__ static cplx zero() { return {0.0, 0.0}; }
  __host__ __device__ static cplx one() { return {1.0, 0.0}; }
  __host__ __device__ static cplx from(double r, double i) { return {r, i}; }
};

// ------------------------------------------------------------------------------ device memory
// Device buffers are recycled through an exact-size free list owned by the handle the calling thread works for.
// A solve allocates and frees ~40 work buffers (sort keys 2.4 GB, Krylov basis 0.6 GB, PCG vectors 1.1 GB, ...);
// cudaMalloc / cudaFree of such blocks cost a device synchronisation plus physical (un)mapping -- measured as
// sporadic 100-800 ms stalls of exactly the stages that allocate (the factorisation, whose arena is persistent, never
// stalled).  In steady state every request finds the block of the previous solve: no driver call at all.
// One handle <-> one stream, so reuse is ordered by the stream.
struct DevPool {
  std::multimap<size_t, void*> free_;
  size_t cached = 0;
  static constexpr size_t LIMIT = (size_t)24 << 30;  // cached bytes above which everything is handed back
  void* get(size_t bytes) {
    auto it = free_.find(bytes);
    if (it != free_.end()) {
      void* p = it->second;
      free_.erase(it);
      cached -= bytes;
      return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim();
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess)
      throw Error(FW_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + ") -> " + cudaGetErrorString(e));
    return p;
  }
  void put(void* p, size_t bytes) {
    free_.emplace(bytes, p);
    cached += bytes;
    if (cached > LIMIT) trim();
  }
  void trim() {
    for (auto& kv : free_) cudaFree(kv.second);
    free_.clear();
    cached = 0;
  }
  ~DevPool() { trim(); }
};
inline DevPool*& current_pool() {
  static thread_local DevPool* p = nullptr;
  return p;
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevPool* pool = nullptr;  // where the block goes back to
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), pool(o.pool) { o.p = nullptr, o.n = 0, o.pool = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p, n = o.n, pool = o.pool;
      o.p = nullptr, o.n = 0, o.pool = nullptr;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      if (pool)
        pool->put(p, n * sizeof(T));
      else
        cudaFree(p);
    }
    p = nullptr;
    n = 0;
    pool = nullptr;
  }
  // (re)allocate when the capacity is insufficient; contents are NOT preserved
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    if (count == 0) return;
    pool = current_pool();
    if (pool)
      p = (T*)pool->get(count * sizeof(T));
    else
      FW_CHECK_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void upload(const T* host, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) FW_CHECK_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* host, size_t count, cudaStream_t s) const {
    if (count) FW_CHECK_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  void zero(cudaStream_t s, size_t count = (size_t)-1) {
    if (count == (size_t)-1) count = n;
    if (count) FW_CHECK_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s));
  }
};

// CUDA-event stage timer; a named timer also opens an NVTX range (visible in nsys / ncu --nvtx timelines) that
// closes with stop() or the destructor
struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  bool range = false;
  explicit Timer(cudaStream_t st, const char* name = nullptr) : s(st) {
    if (name) {
      nvtxRangePushA(name);
      range = true;
    }
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, s);
  }
  double stop() {
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (range) {
      nvtxRangePop();
      range = false;
    }
    return ms;
  }
  ~Timer() {
    if (range) nvtxRangePop();
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

// FW_TRACE=1: host wall-clock of the sub-stages (stream-synchronised), printed to stderr
struct Trace {
  cudaStream_t s;
  double t0;
  bool on;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  explicit Trace(cudaStream_t st) : s(st), on(getenv("FW_TRACE") != nullptr) {
    if (on) {
      cudaStreamSynchronize(s);
      t0 = now();
    }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(s);
    double t1 = now();
    fprintf(stderr, "[fw-trace] %-28s %9.3f ms\n", what, t1 - t0);
    t0 = t1;
  }
};

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// function attributes (dynamic shared memory opt-in) are per device: true the first time a call site runs on the
// current device
inline bool first_use_on_device(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

constexpr int MAXNB = 14;  // local basis functions (N2 x P2)
constexpr int MAXQ = 12;   // quadrature points

}  // namespace fw
