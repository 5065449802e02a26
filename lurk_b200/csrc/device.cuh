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

#include <vector>

namespace lsk {
// Per-kernel accounting: every launcher opens a ProfScope naming the kernel family, the number of
// kernels it launches and the ALGORITHMIC bytes they move (each array read / written once).  The
// launch counters are always on; CUDA-event timing (on the context's stream) only when enabled
// through lsk_profile_enable, which is how bench.py measures per-kernel achieved bandwidth live.
enum ProfTag : int {
  PT_TRACE = 0, PT_LDE_SMALL, PT_LDE_K1, PT_LDE_K2, PT_LDE_K3, PT_MERKLE_LEAF, PT_MERKLE_LAYER, PT_MERKLE_TOP,
  PT_FRI_LEAF, PT_PERM_ROWS, PT_SCAN, PT_QUOTIENT, PT_BARY, PT_OPEN, PT_REDUCE_OPEN, PT_FRI_FOLD, PT_GRIND,
  PT_GATHER, PT_POSEIDON2, PT_LAYOUT, PT_COUNT
};
static const char* const PROF_NAMES[PT_COUNT] = {
    "trace_gen", "lde_small", "lde_k1", "lde_k2", "lde_k3", "merkle_leaf", "merkle_layer", "merkle_top", "fri_leaf",
    "perm_rows", "scan", "quotient", "bary_weights", "open", "reduce_openings", "fri_fold", "grind", "gather",
    "poseidon2", "layout"};
struct Profiler {
  bool timing = false;
  struct Rec { int tag; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  unsigned long long launches[PT_COUNT] = {};
  double bytes[PT_COUNT] = {};
  double ms[PT_COUNT] = {};
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
}  // namespace lsk

struct lsk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  void* ntt_tables = nullptr;   // lsk::NttTables (ntt.cu): twiddle / coset-scale tables, built lazily per size
  lsk::Profiler prof;
};

namespace lsk {
struct ProfScope {
  lsk_ctx* ctx;
  int tag;
  cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(lsk_ctx* c, int t, unsigned nlaunch, double algorithmic_bytes) : ctx(c), tag(t) {
    Profiler& p = ctx->prof;
    p.launches[t] += nlaunch;
    p.bytes[t] += algorithmic_bytes;
    if (p.timing) {
      a = p.get();
      b = p.get();
      cudaEventRecord(a, ctx->stream);
    }
  }
  ~ProfScope() {
    if (a) {
      cudaEventRecord(b, ctx->stream);
      ctx->prof.recs.push_back({tag, a, b});
    }
  }
};
}  // namespace lsk

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
