// Shared plumbing for the rsoptsc_b200 CUDA library: error handling, device buffers, the
// per-process context (stream + dense-solver handles) and small reduction helpers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

namespace rs {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* file, int line, const std::string& what) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s (%s:%d)", what.c_str(), file, line);
  throw Error(code, buf);
}

#define RS_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      ::rs::fail(2, __FILE__, __LINE__, std::string("CUDA: ") + cudaGetErrorString(_e)); \
  } while (0)
#define RS_CUBLAS(expr)                                                                  \
  do {                                                                                   \
    cublasStatus_t _e = (expr);                                                          \
    if (_e != CUBLAS_STATUS_SUCCESS)                                                     \
      ::rs::fail(3, __FILE__, __LINE__, "cuBLAS status " + std::to_string((int)_e));     \
  } while (0)
#define RS_CUSOLVER(expr)                                                                \
  do {                                                                                   \
    cusolverStatus_t _e = (expr);                                                        \
    if (_e != CUSOLVER_STATUS_SUCCESS)                                                   \
      ::rs::fail(4, __FILE__, __LINE__, "cuSOLVER status " + std::to_string((int)_e));   \
  } while (0)
#define RS_REQUIRE(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) ::rs::fail(1, __FILE__, __LINE__, (msg));      \
  } while (0)
#define RS_KERNEL_CHECK() RS_CUDA(cudaGetLastError())

// Size-bucketed caching allocator (capi.cu): cudaMalloc/cudaFree of the 0.1-1 GB temporaries of
// every SVT / Lanczos call cost tens of ms and synchronise the device; buffers are recycled
// instead and only returned to the driver by rsoptsc_finalize() / pool_trim().
void* pool_alloc(size_t bytes);
void pool_free(void* p, size_t bytes);
void pool_trim();

// Owning device allocation (library owns device memory; callers own host buffers).
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = static_cast<T*>(pool_alloc(count * sizeof(T)));
  }
  void release() {
    if (p) pool_free(p, n * sizeof(T));
    p = nullptr; n = 0;
  }
  void zero(cudaStream_t s) { if (n) RS_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) {
    RS_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* h, size_t count, cudaStream_t s) const {
    RS_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

// Either a borrowed device pointer (the *_dev entry points) or an uploaded copy of a host
// buffer (the drop-in entry points).  `on_device` selects which.
template <class T>
struct InBuf {
  const T* p = nullptr;
  DevBuf<T> own;
  void bind(const T* src, size_t count, bool on_device, cudaStream_t s) {
    if (on_device) { p = src; return; }
    own.alloc(count);
    own.upload(src, count, s);
    p = own.p;
  }
};

struct Context {
  int device = -1;
  cudaStream_t stream = nullptr;
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  int sm_count = 148;
  // event timing of library phases (filled when profiling is enabled)
  bool initialized = false;
};

Context& ctx();            // capi.cu; throws if rsoptsc_init was not called
void ensure_init(int dev); // lazy init on `dev` (or current device when dev < 0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace rs

#ifdef __CUDACC__
namespace rs {
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum; result valid in thread 0 (and broadcast to all when `bcast`).
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* smem /* >= THREADS/32 + 1 */, bool bcast = true) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < THREADS / 32 ? smem[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) smem[THREADS / 32] = t;
  }
  __syncthreads();
  return smem[THREADS / 32];
}
}  // namespace rs
#endif
