// A1: batch Poseidon2 hashing (hash3/4/5 of content-addressed data) and wide-witness population.
// See poseidon2.cuh for the reference map.  One permutation per thread; rows are staged through
// padded shared memory so global loads stay coalesced although each thread owns a whole row.
#include "device.cuh"
#include "poseidon2.cuh"

namespace lsk {

constexpr int P2_BLOCK = 128;

// mode 0: digest (first 8 words), mode 1: full state
template <int T, int MODE>
__global__ void __launch_bounds__(P2_BLOCK) poseidon2_rows_kernel(const uint32_t* __restrict__ in, size_t n,
                                                                   uint32_t* __restrict__ out) {
  constexpr int STRIDE = T + 1;  // odd stride: conflict-free row reads
  __shared__ uint32_t tile[P2_BLOCK * STRIDE];
  size_t base = (size_t)blockIdx.x * P2_BLOCK;
  size_t rows = n - base < (size_t)P2_BLOCK ? n - base : (size_t)P2_BLOCK;
  const uint32_t* src = in + base * T;
  for (unsigned i = threadIdx.x; i < rows * T; i += P2_BLOCK) tile[(i / T) * STRIDE + (i % T)] = to_mont(src[i]);
  __syncthreads();
  uint32_t s[T];
  if (threadIdx.x < rows) {
#pragma unroll
    for (int j = 0; j < T; j++) s[j] = tile[threadIdx.x * STRIDE + j];
    p2_permute<T>(s);
  }
  constexpr int OUTW = MODE == 0 ? 8 : T;
  constexpr int OSTRIDE = OUTW + 1;
  __syncthreads();
  if (threadIdx.x < rows) {
#pragma unroll
    for (int j = 0; j < OUTW; j++) tile[threadIdx.x * OSTRIDE + j] = from_mont(s[j]);
  }
  __syncthreads();
  uint32_t* dst = out + base * OUTW;
  for (unsigned i = threadIdx.x; i < rows * OUTW; i += P2_BLOCK) dst[i] = tile[(i / OUTW) * OSTRIDE + (i % OUTW)];
}

// Wide witness: row = digest[8] ++ Poseidon2Cols (repr(C) order of src/poseidon/wide/columns.rs:17-32)
struct RowWitness {
  uint32_t* row;  // points at the Poseidon2Cols part
  int t, rp;
  LSK_D void ext_state(int r, int i, uint32_t v) { row[r * t + i] = from_mont(v); }
  LSK_D void ext_sbox(int r, int i, uint32_t v) { row[8 * t + r * t + i] = from_mont(v); }
  LSK_D void int_init(int i, uint32_t v) { row[16 * t + i] = from_mont(v); }
  LSK_D void int_state0(int r, uint32_t v) { row[17 * t + r] = from_mont(v); }
  LSK_D void int_sbox(int r, uint32_t v) { row[17 * t + (rp - 1) + r] = from_mont(v); }
};

template <int T>
__global__ void __launch_bounds__(P2_BLOCK) poseidon2_witness_kernel(const uint32_t* __restrict__ in, size_t n,
                                                                      uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * P2_BLOCK + threadIdx.x;
  if (i >= n) return;
  constexpr int WC = 17 * T + 2 * P2Rounds<T>::RP - 1;
  uint32_t s[T];
#pragma unroll
  for (int j = 0; j < T; j++) s[j] = to_mont(in[i * T + j]);
  uint32_t* row = out + i * (size_t)(8 + WC);
  RowWitness wit{row + 8, T, P2Rounds<T>::RP};
  p2_permute<T, RowWitness>(s, wit);
#pragma unroll
  for (int j = 0; j < 8; j++) row[j] = from_mont(s[j]);
}

void init_poseidon2_kernels() { LSK_CUDA(p2_upload_constants()); }

template <int MODE>
static void launch_rows(lsk_ctx* ctx, int width, const uint32_t* in_dev, size_t n, uint32_t* out_dev) {
  if (!n) return;
  unsigned grid = div_up(n, P2_BLOCK);
  ProfScope ps(ctx, PT_POSEIDON2, 1, 4.0 * n * (width + (MODE == 0 ? 8 : width)));
  switch (width) {
    case 16: poseidon2_rows_kernel<16, MODE><<<grid, P2_BLOCK, 0, ctx->stream>>>(in_dev, n, out_dev); break;
    case 24: poseidon2_rows_kernel<24, MODE><<<grid, P2_BLOCK, 0, ctx->stream>>>(in_dev, n, out_dev); break;
    case 32: poseidon2_rows_kernel<32, MODE><<<grid, P2_BLOCK, 0, ctx->stream>>>(in_dev, n, out_dev); break;
    case 40: poseidon2_rows_kernel<40, MODE><<<grid, P2_BLOCK, 0, ctx->stream>>>(in_dev, n, out_dev); break;
    default: fail(LSK_ERR_ARG, "unsupported Poseidon2 width %d", width);
  }
  LSK_CUDA(cudaGetLastError());
}

void poseidon2_hash_dev(lsk_ctx* ctx, int width, const uint32_t* in_dev, size_t n, uint32_t* out_dev) {
  launch_rows<0>(ctx, width, in_dev, n, out_dev);
}
void poseidon2_permute_dev(lsk_ctx* ctx, int width, const uint32_t* in_dev, size_t n, uint32_t* out_dev) {
  launch_rows<1>(ctx, width, in_dev, n, out_dev);
}
int poseidon2_witness_cols(int width) {
  switch (width) {
    case 24: return 8 + 17 * 24 + 2 * 21 - 1;
    case 32: return 8 + 17 * 32 + 2 * 30 - 1;
    case 40: return 8 + 17 * 40 + 2 * 38 - 1;
  }
  return -1;
}
void poseidon2_witness_dev(lsk_ctx* ctx, int width, const uint32_t* in_dev, size_t n, uint32_t* out_dev) {
  if (!n) return;
  unsigned grid = div_up(n, P2_BLOCK);
  ProfScope ps(ctx, PT_POSEIDON2, 1, 4.0 * n * (width + poseidon2_witness_cols(width)));
  switch (width) {
    case 24: poseidon2_witness_kernel<24><<<grid, P2_BLOCK, 0, ctx->stream>>>(in_dev, n, out_dev); break;
    case 32: poseidon2_witness_kernel<32><<<grid, P2_BLOCK, 0, ctx->stream>>>(in_dev, n, out_dev); break;
    case 40: poseidon2_witness_kernel<40><<<grid, P2_BLOCK, 0, ctx->stream>>>(in_dev, n, out_dev); break;
    default: fail(LSK_ERR_ARG, "unsupported Poseidon2 witness width %d", width);
  }
  LSK_CUDA(cudaGetLastError());
}

}  // namespace lsk
