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
#include "device.cuh"

namespace lsk {

enum class LdeShift : int {
  Trace = 0,       // in_shift = 1               (main / permutation / preprocessed traces)
  QuotChunk0 = 1,  // in_shift = g               (quotient chunk 0 lives on g * H_N)
  QuotChunk1 = 2,  // in_shift = g * w_2N        (quotient chunk 1 lives on g * w_2N * H_N)
};

void init_ntt_kernels();
void ntt_ctx_destroy(lsk_ctx* ctx);
// in_stride / out_stride: distance in words between consecutive columns.
void lde_dev(lsk_ctx* ctx, const uint32_t* in, size_t in_stride, uint32_t* out, size_t out_stride, unsigned logn,
             uint32_t ncols, LdeShift kind);

}  // namespace lsk
