// NCCL plumbing of the column-sharded prove (comm.cuh).  The library has no link-time dependency on
// NCCL: the seven entry points are looked up with dlopen/dlsym when a group is first created.
#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>
#include <mutex>


namespace lsk {
namespace {

// the part of nccl.h this file needs (stable since NCCL 2.0)
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { NCCL_SUM = 0, NCCL_UINT8 = 1, NCCL_UINT32 = 3 };
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  void* h = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return;
#define SYM(field, name)                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  if (!g_nccl.field) return;
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(AllGather, "ncclAllGather")
  SYM(Broadcast, "ncclBroadcast")
  SYM(AllReduce, "ncclAllReduce")
#undef SYM
  g_nccl.ok = true;
}
NcclApi& nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) fail(LSK_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded or lacks a required symbol");
  return g_nccl;
}
#define LSK_NCCL(expr)                                                                            \
  do {                                                                                            \
    int _r = (expr);                                                                              \
    if (_r != 0) fail(LSK_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString(_r));           \
  } while (0)

Comm& group(lsk_ctx* ctx) {
  if (!comm_active(ctx)) fail(LSK_ERR_ARG, "the context has no device group");
  return *ctx->comm;
}

}  // namespace

void comm_allgather(lsk_ctx* ctx, void* buf, size_t bytes_per_rank) {
  comm_allgather_strided(ctx, buf, 0, 1, bytes_per_rank);
}

void comm_allgather_strided(lsk_ctx* ctx, void* base, size_t stride_bytes, unsigned nseg, size_t bytes_per_rank) {
  Comm& c = group(ctx);
  if (!bytes_per_rank || !nseg) return;
  ProfScope ps(ctx, PT_COMM, 1, (double)nseg * bytes_per_rank * c.size);
  c.bytes_received += (double)nseg * bytes_per_rank * (c.size - 1);
  NcclApi& n = nccl();
  if (nseg > 1) LSK_NCCL(n.GroupStart());
  for (unsigned k = 0; k < nseg; k++) {
    char* b = (char*)base + (size_t)k * stride_bytes;
    LSK_NCCL(n.AllGather(b + (size_t)c.rank * bytes_per_rank, b, bytes_per_rank, NCCL_UINT8, (NcclComm)c.nccl, ctx->stream));
  }
  if (nseg > 1) LSK_NCCL(n.GroupEnd());
}

void comm_allgather_columns(lsk_ctx* ctx, uint32_t* cols, size_t stride_words, uint32_t w) {
  Comm& c = group(ctx);
  if (!w || !stride_words) return;
  ProfScope ps(ctx, PT_COMM, 1, 4.0 * w * stride_words);
  NcclApi& n = nccl();
  LSK_NCCL(n.GroupStart());
  for (int r = 0; r < c.size; r++) {
    size_t c0, c1;
    comm_range_of(r, c.size, w, &c0, &c1);
    if (c1 == c0) continue;
    uint32_t* b = cols + c0 * stride_words;
    const size_t bytes = (c1 - c0) * stride_words * 4;
    if (r != c.rank) c.bytes_received += (double)bytes;
    LSK_NCCL(n.Broadcast(b, b, bytes, NCCL_UINT8, r, (NcclComm)c.nccl, ctx->stream));
  }
  LSK_NCCL(n.GroupEnd());
}

void comm_broadcast_many(lsk_ctx* ctx, const std::vector<std::pair<void*, size_t>>& bufs, int root) {
  Comm& c = group(ctx);
  if (bufs.empty()) return;
  double bytes = 0;
  for (auto& b : bufs) bytes += (double)b.second;
  ProfScope ps(ctx, PT_COMM, 1, bytes);
  if (c.rank != root) c.bytes_received += bytes;
  NcclApi& n = nccl();
  LSK_NCCL(n.GroupStart());
  for (auto& b : bufs) LSK_NCCL(n.Broadcast(b.first, b.first, b.second, NCCL_UINT8, root, (NcclComm)c.nccl, ctx->stream));
  LSK_NCCL(n.GroupEnd());
}

void comm_allreduce_u32(lsk_ctx* ctx, uint32_t* buf, size_t count) {
  Comm& c = group(ctx);
  if (!count) return;
  ProfScope ps(ctx, PT_COMM, 1, 4.0 * count * c.size);
  c.bytes_received += 4.0 * count * (c.size - 1);
  LSK_NCCL(nccl().AllReduce(buf, buf, count, NCCL_UINT32, NCCL_SUM, (NcclComm)c.nccl, ctx->stream));
}

void comm_destroy(lsk_ctx* ctx) {
  if (!ctx->comm) return;
  if (ctx->comm->nccl && g_nccl.ok) {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy((NcclComm)ctx->comm->nccl);
  }
  delete ctx->comm;
  ctx->comm = nullptr;
}

}  // namespace lsk

using namespace lsk;

extern "C" {

int lsk_comm_unique_id(uint8_t out[LSK_COMM_ID_BYTES]) {
  LSK_API_BEGIN
  if (!out) fail(LSK_ERR_ARG, "null out");
  static_assert(LSK_COMM_ID_BYTES == sizeof(NcclUniqueId), "id size");
  NcclUniqueId id;
  LSK_NCCL(nccl().GetUniqueId(&id));
  memcpy(out, id.internal, sizeof id.internal);
  return LSK_OK;
  LSK_API_END
}

int lsk_ctx_comm_init(lsk_ctx* ctx, int rank, int size, const uint8_t id[LSK_COMM_ID_BYTES]) {
  LSK_API_BEGIN
  if (!ctx || !id) fail(LSK_ERR_ARG, "null argument");
  if (size < 1 || size > 64 || (size & (size - 1))) fail(LSK_ERR_ARG, "group size %d is not a power of two <= 64", size);
  if (rank < 0 || rank >= size) fail(LSK_ERR_ARG, "rank %d out of range (group of %d)", rank, size);
  if (ctx->comm) fail(LSK_ERR_ARG, "the context already has a device group");
  LSK_CUDA(cudaSetDevice(ctx->device));
  auto c = new Comm;
  c->rank = rank;
  c->size = size;
  while ((1 << c->log_size) < size) c->log_size++;
  if (size > 1) {
    NcclUniqueId nid;
    memcpy(nid.internal, id, sizeof nid.internal);
    NcclComm comm = nullptr;
    int r = nccl().CommInitRank(&comm, size, nid, rank);
    if (r != 0) {
      delete c;
      fail(LSK_ERR_CUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString(r));
    }
    c->nccl = comm;
  }
  ctx->comm = c;
  return LSK_OK;
  LSK_API_END
}

int lsk_ctx_comm_destroy(lsk_ctx* ctx) {
  LSK_API_BEGIN
  if (!ctx) fail(LSK_ERR_ARG, "null ctx");
  comm_destroy(ctx);
  return LSK_OK;
  LSK_API_END
}

int lsk_ctx_comm_info(const lsk_ctx* ctx, int* rank, int* size, double* bytes_received) {
  LSK_API_BEGIN
  if (!ctx) fail(LSK_ERR_ARG, "null ctx");
  if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
  if (size) *size = ctx->comm ? ctx->comm->size : 1;
  if (bytes_received) *bytes_received = ctx->comm ? ctx->comm->bytes_received : 0.0;
  return LSK_OK;
  LSK_API_END
}

int lsk_comm_partition(uint64_t n, int rank, int size, uint64_t* begin, uint64_t* end) {
  LSK_API_BEGIN
  if (size < 1 || rank < 0 || rank >= size || !begin || !end) fail(LSK_ERR_ARG, "bad partition arguments");
  size_t b, e;
  comm_range_of(rank, size, (size_t)n, &b, &e);
  *begin = b, *end = e;
  return LSK_OK;
  LSK_API_END
}

}  // extern "C"
