// CUDA plumbing shared by the kernels' host wrappers: error checks, the context, stream-ordered
// device buffers (cudaMallocAsync on the context's stream; the pool keeps freed blocks so steady-
// state proving does no driver allocation).
#pragma once
#include <cuda_runtime.h>

#include <utility>

#include "common.h"

#define LSK_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) lsk::fail(LSK_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

struct lsk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  void* ntt_tables = nullptr;   // lsk::NttTables (ntt.cu): twiddle / coset-scale tables, built lazily per size
};

namespace lsk {

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(lsk_ctx* ctx, size_t count) { alloc(ctx, count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; s = o.s;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(lsk_ctx* ctx, size_t count) {
    release();
    s = ctx->stream;
    n = count;
    if (count) LSK_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), s));
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  void upload(const T* host, size_t count) { LSK_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s)); }
  void download(T* host, size_t count) const { LSK_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s)); }
  void zero() { if (n) LSK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace lsk
