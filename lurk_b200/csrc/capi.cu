// C-ABI entry points (include/lurk_stark.h): context management, Poseidon2 batch hashing and the
// stage-level trace-generation calls.  The prover entry points live in prover.cu.
#include <cstring>
#include <mutex>

#include "comm.cuh"
#include "ntt.cuh"
#include "trace.cuh"

namespace lsk {
static thread_local std::string g_last_error;
void set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }

void init_poseidon2_kernels();
void poseidon2_hash_dev(lsk_ctx*, int, const uint32_t*, size_t, uint32_t*);
void poseidon2_permute_dev(lsk_ctx*, int, const uint32_t*, size_t, uint32_t*);
int poseidon2_witness_cols(int);
void poseidon2_witness_dev(lsk_ctx*, int, const uint32_t*, size_t, uint32_t*);
void init_prover_kernels();
}  // namespace lsk

using namespace lsk;

extern "C" {

const char* lsk_last_error(void) { return g_last_error.c_str(); }
int lsk_version(void) { return 100; }

int lsk_ctx_create(int device, lsk_ctx** out) {
  LSK_API_BEGIN
  if (!out) fail(LSK_ERR_ARG, "null out");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail(LSK_ERR_CUDA, "no CUDA device available (%s); liblurkstark has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) fail(LSK_ERR_ARG, "device %d out of range (%d devices)", device, count);
  LSK_CUDA(cudaSetDevice(device));
  auto* ctx = new lsk_ctx;
  ctx->device = device;
  LSK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  LSK_CUDA(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
  LSK_CUDA(cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming));
  LSK_CUDA(cudaEventCreateWithFlags(&ctx->side_join, cudaEventDisableTiming));
  cudaDeviceProp prop;
  LSK_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  init_poseidon2_kernels();
  init_trace_kernels();
  init_prover_kernels();
  *out = ctx;
  return LSK_OK;
  LSK_API_END
}
void lsk_ctx_destroy(lsk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  ntt_ctx_destroy(ctx);
  for (auto& r : ctx->prof.recs) cudaEventDestroy(r.a), cudaEventDestroy(r.b);
  for (auto e : ctx->prof.pool) cudaEventDestroy(e);
  ctx->pool.trim();
  cudaStreamDestroy(ctx->stream);
  for (void* h : ctx->pinned) if (h) cudaFreeHost(h);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->side_fork) cudaEventDestroy(ctx->side_fork);
  if (ctx->side_join) cudaEventDestroy(ctx->side_join);
  ctx->stream = nullptr;
  // objects that still own device buffers (a machine freed after its context) release into this pool later:
  // keep the (small) context object alive for them rather than leave them a dangling pointer
  if (ctx->pool.live_bytes == 0) delete ctx;
}
void* lsk_ctx_stream(lsk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int lsk_ctx_sync(lsk_ctx* ctx) {
  LSK_API_BEGIN
  if (!ctx) fail(LSK_ERR_ARG, "null ctx");
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}

int lsk_profile_enable(lsk_ctx* ctx, int on) {
  LSK_API_BEGIN
  if (!ctx) fail(LSK_ERR_ARG, "null ctx");
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  Profiler& p = ctx->prof;
  for (auto& r : p.recs) p.pool.push_back(r.a), p.pool.push_back(r.b);
  p.recs.clear();
  for (int i = 0; i < PT_COUNT; i++) p.launches[i] = 0, p.bytes[i] = 0, p.ms[i] = 0, p.work[i] = 0;
  if (ctx->comm) ctx->comm->bytes_received = 0;
  p.timing = on != 0;
  return LSK_OK;
  LSK_API_END
}
int lsk_profile_read(lsk_ctx* ctx, double* ms, double* bytes, uint64_t* launches, int cap) {
  LSK_API_BEGIN
  if (!ctx) fail(LSK_ERR_ARG, "null ctx");
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  Profiler& p = ctx->prof;
  for (auto& r : p.recs) {
    float t = 0;
    LSK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    p.ms[r.tag] += t;
    p.pool.push_back(r.a), p.pool.push_back(r.b);
  }
  p.recs.clear();
  for (int i = 0; i < PT_COUNT && i < cap; i++) {
    if (ms) ms[i] = p.ms[i];
    if (bytes) bytes[i] = p.bytes[i];
    if (launches) launches[i] = p.launches[i];
  }
  return PT_COUNT;
  LSK_API_END
}
int lsk_profile_read_work(lsk_ctx* ctx, double* work, int cap) {
  LSK_API_BEGIN
  if (!ctx || !work) fail(LSK_ERR_ARG, "null argument");
  for (int i = 0; i < PT_COUNT && i < cap; i++) work[i] = ctx->prof.work[i];
  return PT_COUNT;
  LSK_API_END
}
const char* lsk_profile_tag_name(int tag) { return tag >= 0 && tag < PT_COUNT ? PROF_NAMES[tag] : ""; }

static void check_width(int width, bool allow16) {
  if (!(width == 24 || width == 32 || width == 40 || (allow16 && width == 16)))
    fail(LSK_ERR_ARG, "unsupported Poseidon2 width %d", width);
}

int lsk_poseidon2_hash_dev(lsk_ctx* ctx, int width, const uint32_t* in_dev, size_t n, uint32_t* out8_dev) {
  LSK_API_BEGIN
  if (!ctx) fail(LSK_ERR_ARG, "null ctx");
  check_width(width, true);
  poseidon2_hash_dev(ctx, width, in_dev, n, out8_dev);
  return LSK_OK;
  LSK_API_END
}
int lsk_poseidon2_hash(lsk_ctx* ctx, int width, const uint32_t* in, size_t n, uint32_t* out8) {
  LSK_API_BEGIN
  if (!ctx || (n && (!in || !out8))) fail(LSK_ERR_ARG, "null argument");
  check_width(width, true);
  DevBuf<uint32_t> d_in(ctx, n * width), d_out(ctx, n * 8);
  if (n) d_in.upload(in, n * width);
  poseidon2_hash_dev(ctx, width, d_in.p, n, d_out.p);
  if (n) d_out.download(out8, n * 8);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_poseidon2_permute(lsk_ctx* ctx, int width, const uint32_t* in, size_t n, uint32_t* out) {
  LSK_API_BEGIN
  if (!ctx || (n && (!in || !out))) fail(LSK_ERR_ARG, "null argument");
  check_width(width, true);
  DevBuf<uint32_t> d_in(ctx, n * width), d_out(ctx, n * width);
  if (n) d_in.upload(in, n * width);
  poseidon2_permute_dev(ctx, width, d_in.p, n, d_out.p);
  if (n) d_out.download(out, n * width);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_poseidon2_witness_cols(int width) { return poseidon2_witness_cols(width); }
int lsk_poseidon2_witness(lsk_ctx* ctx, int width, const uint32_t* in, size_t n, uint32_t* out) {
  LSK_API_BEGIN
  if (!ctx || (n && (!in || !out))) fail(LSK_ERR_ARG, "null argument");
  check_width(width, false);
  size_t wc = (size_t)poseidon2_witness_cols(width);
  DevBuf<uint32_t> d_in(ctx, n * width), d_out(ctx, n * wc);
  if (n) d_in.upload(in, n * width);
  poseidon2_witness_dev(ctx, width, d_in.p, n, d_out.p);
  if (n) d_out.download(out, n * wc);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}

int lsk_func_trace(lsk_ctx* ctx, const lsk_program* p, const lsk_record* r, uint32_t func, uint32_t shard,
                   uint32_t shard_size, uint32_t* out) {
  LSK_API_BEGIN
  if (!ctx || !p || !r || !out || func >= p->nfuncs || !shard_size) fail(LSK_ERR_ARG, "bad argument");
  size_t n = r->funcs[func].rows();
  size_t s = (size_t)shard * shard_size, e = std::min<size_t>((size_t)(shard + 1) * shard_size, n);
  uint32_t real = e > s ? (uint32_t)(e - s) : 0;
  uint32_t height = (uint32_t)lsk_func_trace_height(r, func, shard, shard_size);
  uint32_t width = p->width(func);
  DeviceRecord drec;
  drec.upload(ctx, *p, *r);
  DevBuf<uint32_t> d_out(ctx, (size_t)height * width);
  func_trace_dev(ctx, *p, drec, func, (uint32_t)s, real, height, TraceLayout::RowMajorCanonical, d_out.p);
  d_out.download(out, (size_t)height * width);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_mem_trace(lsk_ctx* ctx, const lsk_record* r, uint32_t mem_idx, uint32_t* out) {
  LSK_API_BEGIN
  if (!ctx || !r || !out || mem_idx >= NUM_MEM_TABLES) fail(LSK_ERR_ARG, "bad argument");
  lsk_program empty;
  empty.words = {PROGRAM_MAGIC, 0, 0};
  DeviceRecord drec;
  drec.upload(ctx, empty, *r);
  uint32_t height = (uint32_t)lsk_mem_trace_height(r, mem_idx);
  size_t words = (size_t)height * (4 + r->mems[mem_idx].len);
  DevBuf<uint32_t> d_out(ctx, words);
  mem_trace_dev(ctx, drec, mem_idx, height, TraceLayout::RowMajorCanonical, d_out.p);
  d_out.download(out, words);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_bytes_trace(lsk_ctx* ctx, const lsk_record* r, uint32_t shard, uint32_t* out) {
  LSK_API_BEGIN
  if (!ctx || !r || !out) fail(LSK_ERR_ARG, "bad argument");
  lsk_program empty;
  empty.words = {PROGRAM_MAGIC, 0, 0};
  DeviceRecord drec;
  drec.upload(ctx, empty, *r);
  DevBuf<uint32_t> d_out(ctx, (size_t)65536 * 13);
  bytes_trace_dev(ctx, drec, shard == 0 && r->bytes_used, TraceLayout::RowMajorCanonical, d_out.p);
  d_out.download(out, (size_t)65536 * 13);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}
int lsk_bytes_preprocessed(lsk_ctx* ctx, uint32_t* out) {
  LSK_API_BEGIN
  if (!ctx || !out) fail(LSK_ERR_ARG, "bad argument");
  DevBuf<uint32_t> d_out(ctx, (size_t)65536 * 6);
  bytes_preprocessed_dev(ctx, TraceLayout::RowMajorCanonical, d_out.p);
  d_out.download(out, (size_t)65536 * 6);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return LSK_OK;
  LSK_API_END
}

}  // extern "C"
