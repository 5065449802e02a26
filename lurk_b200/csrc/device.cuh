// CUDA plumbing shared by the kernels' host wrappers: error checks, the context, the per-kernel
// profiler and the context-owned caching allocator behind DevBuf.
#pragma once
#include <cuda_runtime.h>

#include <map>
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
  PT_GATHER, PT_POSEIDON2, PT_LAYOUT, PT_COMM, PT_COUNT
};
static const char* const PROF_NAMES[PT_COUNT] = {
    "trace_gen", "lde_small", "lde_k1", "lde_k2", "lde_k3", "merkle_leaf", "merkle_layer", "merkle_top", "fri_leaf",
    "perm_rows", "scan", "quotient", "bary_weights", "open", "reduce_openings", "fri_fold", "grind", "gather",
    "poseidon2", "layout", "nccl"};
struct Profiler {
  bool timing = false;
  struct Rec { int tag; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  unsigned long long launches[PT_COUNT] = {};
  double bytes[PT_COUNT] = {};
  double work[PT_COUNT] = {};   // arithmetic work where a launcher states it: width-16 Poseidon2 permutations (hash kernels)
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

namespace lsk {
// Device memory of a context.  Every kernel and copy of a context runs on its ONE stream, so a block
// released by a buffer can be handed to the next allocation at once (whatever still reads it was queued
// earlier on the same stream).  Blocks are cached by size class (8 classes per power of two, <= 12.5 %
// slack) and never returned to the driver while the context lives: steady-state proving performs no
// driver allocation at all -- cudaMallocAsync / cudaFreeAsync of the multi-GB LDE buffers of a wide chip
// cost more than the kernels that fill them (measured: 245 ms of a 450 ms prove).
struct MemPool {
  std::multimap<size_t, void*> free_blocks;   // size class -> block
  size_t cached_bytes = 0, live_bytes = 0, driver_bytes = 0;
  static size_t size_class(size_t bytes) {
    if (bytes <= 512) return 512;
    size_t p = (size_t)1 << (63 - __builtin_clzll(bytes));   // highest power of two <= bytes
    size_t step = p >> 3;
    return (bytes + step - 1) / step * step;
  }
  void* alloc(size_t bytes, size_t* cls_out) {
    const size_t cls = size_class(bytes);
    *cls_out = cls;
    auto it = free_blocks.find(cls);
    if (it != free_blocks.end()) {
      void* p = it->second;
      free_blocks.erase(it);
      cached_bytes -= cls;
      live_bytes += cls;
      return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cls);
    if (e != cudaSuccess) {   // out of memory: give the cache back to the driver and retry once
      cudaGetLastError();
      trim();
      e = cudaMalloc(&p, cls);
      if (e != cudaSuccess) lsk::fail(LSK_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", cls, cudaGetErrorString(e));
    }
    driver_bytes += cls;
    live_bytes += cls;
    return p;
  }
  void release(void* p, size_t cls) {
    free_blocks.emplace(cls, p);
    cached_bytes += cls;
    live_bytes -= cls;
  }
  void trim() {
    cudaDeviceSynchronize();
    for (auto& kv : free_blocks) cudaFree(kv.second), driver_bytes -= kv.first;
    free_blocks.clear();
    cached_bytes = 0;
  }
};
}  // namespace lsk

namespace lsk { struct Comm; }
struct lsk_ctx {
  int device = 0;
  lsk::Comm* comm = nullptr;    // device group of a column-sharded prove (comm.cuh); null: this GPU works alone
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;      // second stream: the single-CTA launches of a shard's tiny chips run beside the big ones
  cudaEvent_t side_fork = nullptr, side_join = nullptr;
  void* pinned[2] = {nullptr, nullptr};   // page-locked host scratch (pinned_scratch): results the host consumes while kernels run
  size_t pinned_bytes[2] = {0, 0};
  int sm_count = 148;
  lsk::MemPool pool;
  void* ntt_tables = nullptr;   // lsk::NttTables (ntt.cu): twiddle / coset-scale tables, built lazily per size
  lsk::Profiler prof;
};

namespace lsk {
// page-locked host scratch `slot` of at least `bytes`, owned by the context (valid until a later call asks for more in the
// same slot)
inline void* pinned_scratch(lsk_ctx* ctx, int slot, size_t bytes) {
  if (bytes > ctx->pinned_bytes[slot]) {
    if (ctx->pinned[slot]) {
      cudaStreamSynchronize(ctx->stream);
      cudaFreeHost(ctx->pinned[slot]);
      ctx->pinned[slot] = nullptr;
      ctx->pinned_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 2;
    if (cudaHostAlloc(&ctx->pinned[slot], want, cudaHostAllocDefault) != cudaSuccess) lsk::fail(LSK_ERR_CUDA, "cudaHostAlloc(%zu) failed", want);
    ctx->pinned_bytes[slot] = want;
  }
  return ctx->pinned[slot];
}
// Work queued between fork() and join() on ctx->side runs concurrently with what the main stream does in between.
// While `active`, ctx->stream IS the side stream (every helper launches on ctx->stream), so the caller brackets exactly
// the calls that should go there.  The memory pool is ordered by the main stream only: nothing a side-stream kernel
// touches may be released before join().
struct SideStream {
  lsk_ctx* ctx;
  cudaStream_t main;
  bool forked = false;
  explicit SideStream(lsk_ctx* c) : ctx(c), main(c->stream) {}
  void enter() {   // subsequent launches go to the side stream (after everything queued on the main stream so far)
    if (!forked) {
      cudaEventRecord(ctx->side_fork, main);
      cudaStreamWaitEvent(ctx->side, ctx->side_fork, 0);
      forked = true;
    }
    ctx->stream = ctx->side;
  }
  void leave() { ctx->stream = main; }
  void join() {    // the main stream waits for the side stream's work
    ctx->stream = main;
    if (forked) {
      cudaEventRecord(ctx->side_join, ctx->side);
      cudaStreamWaitEvent(main, ctx->side_join, 0);
      forked = false;
    }
  }
  ~SideStream() { join(); }
};
struct ProfScope {
  lsk_ctx* ctx;
  int tag;
  cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(lsk_ctx* c, int t, unsigned nlaunch, double algorithmic_bytes, double work = 0.0) : ctx(c), tag(t) {
    Profiler& p = ctx->prof;
    p.launches[t] += nlaunch;
    p.bytes[t] += algorithmic_bytes;
    p.work[t] += work;
    // launches on the side stream overlap the main stream's: they are counted, not timed (their events would span
    // whatever the main stream ran meanwhile and the per-kernel times would no longer add up to the step)
    if (p.timing && ctx->stream != ctx->side) {
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
  lsk_ctx* owner = nullptr;
  size_t cls = 0;   // size class of the block (bytes)
  DevBuf() = default;
  DevBuf(lsk_ctx* ctx, size_t count) { alloc(ctx, count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s), owner(o.owner), cls(o.cls) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; s = o.s; owner = o.owner; cls = o.cls;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(lsk_ctx* ctx, size_t count) {
    release();
    s = ctx->stream;
    owner = ctx;
    n = count;
    if (count) p = (T*)ctx->pool.alloc(count * sizeof(T), &cls);
  }
  void release() {
    if (p) owner->pool.release(p, cls);
    p = nullptr;
    n = 0;
  }
  void upload(const T* host, size_t count) { LSK_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s)); }
  void download(T* host, size_t count) const { LSK_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s)); }
  void zero() { if (n) LSK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace lsk
