// Device group of a column-sharded prove: G processes, one GPU each, prove ONE shard
// together.  Every rank runs the same prover with the same transcript on full-size buffers; the
// throughput-bound stages compute only the rank's slice (LDE: a block of columns; Merkle leaves and
// layers, quotient, reduced openings: a contiguous range of rows; openings: a block of columns) and an
// NCCL all-gather over NVLink fills in the rest, so every rank ends with the identical proof.  The
// bottom layers of a Merkle tree are never gathered: each rank keeps its range and contributes the
// sibling digests it owns when the queries are answered (one small all-reduce of the query words).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): inside a PyTorch process that is the copy torch
// already loaded, in a plain C++ host the system one.  A context without a group runs exactly as before.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

#include "device.cuh"

namespace lsk {

constexpr unsigned COMM_MIN_LOG = 16;  // stages over fewer than 2^16 rows are computed by every rank

struct Comm {
  int rank = 0, size = 1;
  unsigned log_size = 0;
  void* nccl = nullptr;  // ncclComm_t
  double bytes_received = 0;  // since the last lsk_profile_enable (algorithmic: what the collectives deliver)
};

inline bool comm_active(const lsk_ctx* ctx) { return ctx->comm && ctx->comm->size > 1; }
// is a stage over 2^log_rows rows split across the group?
inline bool comm_splits(const lsk_ctx* ctx, unsigned log_rows) { return comm_active(ctx) && log_rows >= COMM_MIN_LOG; }
// this rank's contiguous share of n items (n a multiple of the group size whenever comm_splits said yes)
inline void comm_range(const lsk_ctx* ctx, size_t n, size_t* begin, size_t* end) {
  const Comm* c = ctx->comm;
  if (!c || c->size <= 1) { *begin = 0, *end = n; return; }
  *begin = n * (size_t)c->rank / (size_t)c->size;
  *end = n * (size_t)(c->rank + 1) / (size_t)c->size;
}
inline void comm_range_of(int rank, int size, size_t n, size_t* begin, size_t* end) {
  *begin = n * (size_t)rank / (size_t)size;
  *end = n * (size_t)(rank + 1) / (size_t)size;
}

// buf holds `size` equal blocks of `bytes_per_rank`; block r is valid on rank r; afterwards all are, everywhere
void comm_allgather(lsk_ctx* ctx, void* buf, size_t bytes_per_rank);
// the same for `nseg` arrays at once (one NCCL group): array k starts at base + k * stride_bytes
void comm_allgather_strided(lsk_ctx* ctx, void* base, size_t stride_bytes, unsigned nseg, size_t bytes_per_rank);
// a column-major matrix whose columns were computed in blocks: rank r owns columns comm_range_of(r, size, w)
void comm_allgather_columns(lsk_ctx* ctx, uint32_t* cols, size_t stride_words, uint32_t w);
// every buffer of the list: rank `root`'s contents to all ranks (one NCCL group)
void comm_broadcast_many(lsk_ctx* ctx, const std::vector<std::pair<void*, size_t>>& bufs, int root);
// element-wise sum of 32-bit words over the group (used where exactly one rank holds a non-zero value)
void comm_allreduce_u32(lsk_ctx* ctx, uint32_t* buf, size_t count);
void comm_destroy(lsk_ctx* ctx);

}  // namespace lsk
