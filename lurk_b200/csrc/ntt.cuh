// A9 (first half): coset low-degree extension of trace columns.
//
// Replaces the third-party `Radix2DitParallel::coset_lde_batch` + `bit_reverse_rows` that sphinx's
// prover reaches from `machine.prove` (call sites: src/core/cli/repl.rs:196, benches/fib.rs:158);
// restated by the oracle in oracle/stark.h (coset_lde_col).
//
// Contract: `in` holds `ncols` columns of N = 2^logn evaluations over in_shift * H_N, column-major,
// rows in BIT-REVERSED order, Montgomery form (exactly what the trace / LogUp / quotient kernels
// emit).  `out` receives `ncols` columns of 2N evaluations over g * H_2N, column-major, rows in
// bit-reversed order (the commitment layout: leaf r of the Merkle tree is row r), so that rows
// [0, N) are the coset g*H_N and rows [N, 2N) the coset g*w_2N*H_N.
#pragma once
#include <vector>

#include "device.cuh"

namespace lsk {

enum class LdeShift : int {
  Trace = 0,       // in_shift = 1               (main / permutation / preprocessed traces)
  QuotChunk0 = 1,  // in_shift = g               (quotient chunk 0 lives on g * H_N)
  QuotChunk1 = 2,  // in_shift = g * w_2N        (quotient chunk 1 lives on g * w_2N * H_N)
};

// Columns of small matrices (N <= 2^12: one CTA each) collected over a prover phase and extended by ONE
// launch at flush time: a shard has many small chips and their launches are pure latency.
struct LdeSmallJob {
  const uint32_t* in;
  uint32_t* out;
  const uint2* scale;   // the (logn, kind) coset-scale table
  uint32_t logn;
};
struct LdeBatch {
  std::vector<LdeSmallJob> jobs;
  unsigned max_logn = 0;
};
void lde_flush(lsk_ctx* ctx, LdeBatch& batch);

void init_ntt_kernels();
void ntt_ctx_destroy(lsk_ctx* ctx);
// in_stride / out_stride: distance in words between consecutive columns.
// With `batch`, matrices of N <= 2^12 are only queued (lde_flush extends them); larger ones launch at once.
void lde_dev(lsk_ctx* ctx, const uint32_t* in, size_t in_stride, uint32_t* out, size_t out_stride, unsigned logn,
             uint32_t ncols, LdeShift kind, LdeBatch* batch = nullptr);

}  // namespace lsk
