// Coset LDE kernels (see ntt.cuh for the contract and the reference map).
//
// Algorithm.  Per column: inverse NTT (decimation in time: bit-reversed in, natural out), scale
// coefficient m by s_j^m / N for the two target cosets s_0 = g/in_shift, s_1 = s_0 * w_2N, forward
// NTT (decimation in frequency: natural in, bit-reversed out).  Neither direction needs a
// permutation pass.  For N <= 2^12 one CTA does the whole column in shared memory.  Larger N is
// split Bailey-style at bit 12, N = 2^12 * 2^t:
//   K1  DIT stages 1..12 on contiguous 4096-blocks                      (4R + 4W bytes / element)
//   K2  per group of 2^c consecutive low indices: DIT stages 13..logn across the 2^t blocks ->
//       coefficients (shared memory), then for both cosets: scale, DIF stages logn..13 across the
//       blocks, store                                                    (4R + 8W)
//       These are the ordinary stages of the size-N transform (stage s pairs elements 2^(s-1) apart with
//       w_{2^s}^k, k = index mod 2^(s-1)), so no separate inter-pass (Bailey) twiddle exists: their twiddles
//       depend on the low index as well as on the block and come from one level-packed table of N entries
//       (w_{2^s}^k at 2^(s-1) + k, shared by every size up to N, read through the L2).  Round 1 split the
//       transform Bailey-style and paid two more table products per element at every global load and store
//       of K2 (22 Shoup products per element for N = 2^20; now 16).
//   K3  DIF stages 12..1 on contiguous 4096-blocks of both cosets        (8R + 8W)
// i.e. 36 B of HBM traffic per input element against the 12 B algorithmic minimum (4 read, 8
// written).  Butterflies run radix-16 in REGISTERS: a thread owns the 16 elements that differ in four
// index bits, does four stages without touching memory, and the 4096-element tile is re-dealt
// through (padded) shared memory between such passes -- 3 passes and 2 barriers for 12 stages
// instead of 12 shared-memory round trips.  Twiddles are (w, floor(w 2^32 / p)) pairs so every
// twiddle product is a Shoup multiplication (3 multiply-pipe instructions on constants instead of
// the 3 + 2 of a Montgomery product); the level-packed table w_{2^s}^k is 32 KB and L1-resident.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <vector>

#include "field.cuh"
#include "ntt.cuh"

namespace lsk {

constexpr unsigned LOG_BLOCK = 12;
constexpr unsigned BLOCK_ELEMS = 1u << LOG_BLOCK;

// level-packed twiddles: tw[2^(s-1) + k] = w_{2^s}^(+-k), k < 2^(s-1), s = 1..12  (index 0 unused)
// every table entry is a Shoup pair {c, floor(c * 2^32 / p)} of a plain residue c
struct NttTables {
  DevBuf<uint2> tw_fwd, tw_inv;
  // level-packed twiddles of the stages above 12, w_{2^s}^(+-k) at 2^(s-1) + k for 13 <= s <= big_log (entries below
  // 2^12 unused): a prefix of the table of any larger size, so one table serves every logn <= big_log
  DevBuf<uint2> big_f, big_i;
  unsigned big_log = 0;
  struct Scale {  // per (logn, kind): a[j][lo] = s_j^lo / N (lo < min(N, 4096)); b[j][k2] = s_j^(4096 k2)
    DevBuf<uint2> a, b;
  };
  std::map<std::pair<unsigned, int>, std::unique_ptr<Scale>> scale;
};

static std::vector<uint2> powers(uint32_t base, size_t n, uint32_t first = R1) {  // Montgomery in, Shoup pairs out
  std::vector<uint2> v(n);
  uint32_t p = first;
  for (size_t i = 0; i < n; i++) {
    uint32_t c = from_mont(p);
    v[i] = make_uint2(c, shoup_precompute(c));
    p = fmul(p, base);
  }
  return v;
}

static NttTables* tables(lsk_ctx* ctx) {
  if (!ctx->ntt_tables) {
    auto* t = new NttTables;
    std::vector<uint2> f(BLOCK_ELEMS, make_uint2(0, 0)), inv(BLOCK_ELEMS, make_uint2(0, 0));
    for (unsigned s = 1; s <= LOG_BLOCK; s++) {
      uint32_t w = two_adic_generator(s), wi = finv(w);
      auto pf = powers(w, (size_t)1 << (s - 1)), pi = powers(wi, (size_t)1 << (s - 1));
      std::copy(pf.begin(), pf.end(), f.begin() + ((size_t)1 << (s - 1)));
      std::copy(pi.begin(), pi.end(), inv.begin() + ((size_t)1 << (s - 1)));
    }
    t->tw_fwd.alloc(ctx, BLOCK_ELEMS);
    t->tw_inv.alloc(ctx, BLOCK_ELEMS);
    t->tw_fwd.upload(f.data(), BLOCK_ELEMS);
    t->tw_inv.upload(inv.data(), BLOCK_ELEMS);
    LSK_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ntt_tables = t;
  }
  return (NttTables*)ctx->ntt_tables;
}
void ntt_ctx_destroy(lsk_ctx* ctx) {
  delete (NttTables*)ctx->ntt_tables;
  ctx->ntt_tables = nullptr;
}

// T[idx] for idx in [2^12, 2^logn): level s = floor(log2 idx) + 1, k = idx - 2^(s-1), entry = Shoup pair of w_{2^s}^(+-k)
__global__ void big_twiddle_fill_kernel(uint2* __restrict__ t, unsigned logn, bool inverse) {
  const size_t idx = ((size_t)1 << LOG_BLOCK) + (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >> logn) return;
  const unsigned s = 64 - __clzll((unsigned long long)idx);
  const size_t k = idx - ((size_t)1 << (s - 1));
  uint32_t w = two_adic_generator(s);
  if (inverse) w = finv(w);
  const uint32_t c = from_mont(fpow(w, k));
  t[idx] = make_uint2(c, (uint32_t)(((uint64_t)c << 32) / P));
}
static void big_tables(lsk_ctx* ctx, unsigned logn) {
  NttTables* t = tables(ctx);
  if (logn <= t->big_log) return;
  // buffers of the old (smaller) table go back to the pool; the pool is stream-ordered, so kernels already queued keep
  // reading valid memory until they have run
  DevBuf<uint2> f(ctx, (size_t)1 << logn), inv(ctx, (size_t)1 << logn);
  const size_t n = ((size_t)1 << logn) - BLOCK_ELEMS;
  big_twiddle_fill_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(f.p, logn, false);
  big_twiddle_fill_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(inv.p, logn, true);
  LSK_CUDA(cudaGetLastError());
  t->big_f = std::move(f);
  t->big_i = std::move(inv);
  t->big_log = logn;
}

static NttTables::Scale* scale_tables(lsk_ctx* ctx, unsigned logn, LdeShift kind) {
  NttTables* t = tables(ctx);
  auto key = std::make_pair(logn, (int)kind);
  auto it = t->scale.find(key);
  if (it != t->scale.end()) return it->second.get();
  auto sc = std::make_unique<NttTables::Scale>();
  uint32_t g = to_mont(GEN), w2n = two_adic_generator(logn + 1);
  uint32_t base = kind == LdeShift::Trace ? g : kind == LdeShift::QuotChunk0 ? R1 : finv(w2n);
  uint32_t s[2] = {base, fmul(base, w2n)};
  uint32_t ninv = finv(to_mont((uint32_t)1 << logn));
  size_t na = (size_t)1 << (logn < LOG_BLOCK ? logn : LOG_BLOCK);
  size_t nb = logn > LOG_BLOCK ? (size_t)1 << (logn - LOG_BLOCK) : 1;
  std::vector<uint2> a, b;
  for (int j = 0; j < 2; j++) {
    auto pa = powers(s[j], na, ninv);
    auto pb = powers(fpow(s[j], BLOCK_ELEMS), nb);
    a.insert(a.end(), pa.begin(), pa.end());
    b.insert(b.end(), pb.begin(), pb.end());
  }
  sc->a.alloc(ctx, a.size()); sc->a.upload(a.data(), a.size());
  sc->b.alloc(ctx, b.size()); sc->b.upload(b.data(), b.size());
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));
  return (t->scale[key] = std::move(sc)).get();
}

// ---- radix-2^R butterflies in registers ----------------------------------------------------------------
// x[jj] is the element whose index is (high << (s0 + R)) | (jj << s0) | low; stages s0+1 .. s0+R of a
// transform only pair elements inside such a group.  tw: level-packed table, w_{2^s}^(+-k) at 2^(s-1) + k.
LSK_D uint32_t shoup(uint32_t a, uint2 c) { return fmul_shoup(a, c.x, c.y); }
template <int R>
LSK_D void dit_local(uint32_t (&x)[1 << R], unsigned s0, unsigned low, const uint2* __restrict__ tw) {
#pragma unroll
  for (int u = 1; u <= R; u++) {
    const unsigned h = 1u << (u - 1), H = 1u << (s0 + u - 1);
#pragma unroll
    for (unsigned b = 0; b < (1u << (R - 1)); b++) {
      const unsigned kl = b & (h - 1);
      const unsigned j0 = ((b >> (u - 1)) << u) | kl, j1 = j0 + h;
      uint32_t v = shoup(x[j1], __ldg(tw + H + low + (kl << s0)));
      uint32_t a = x[j0];
      x[j0] = fadd(a, v);
      x[j1] = fsub(a, v);
    }
  }
}
template <int R>
LSK_D void dif_local(uint32_t (&x)[1 << R], unsigned s0, unsigned low, const uint2* __restrict__ tw) {
#pragma unroll
  for (int u = R; u >= 1; u--) {
    const unsigned h = 1u << (u - 1), H = 1u << (s0 + u - 1);
#pragma unroll
    for (unsigned b = 0; b < (1u << (R - 1)); b++) {
      const unsigned kl = b & (h - 1);
      const unsigned j0 = ((b >> (u - 1)) << u) | kl, j1 = j0 + h;
      uint32_t a = x[j0], c = x[j1];
      x[j0] = fadd(a, c);
      x[j1] = shoup(fsub(a, c), __ldg(tw + H + low + (kl << s0)));
    }
  }
}

// The same register passes for the stages ABOVE 12 (K2): x[jj] is the element of block q = (high << (s0 + R)) | (jj << s0)
// | low at low index glo; stage 12 + s0 + u pairs blocks differing in bit s0 + u - 1 with the twiddle w_{2^s}^k,
// k = ((q mod 2^(s0+u-1)) << 12) | glo -- the level-packed index of the small passes shifted left by 12, plus glo.
template <int R>
LSK_D void dit_local_big(uint32_t (&x)[1 << R], unsigned s0, unsigned low, unsigned glo, const uint2* __restrict__ big) {
#pragma unroll
  for (int u = 1; u <= R; u++) {
    const unsigned h = 1u << (u - 1), H = 1u << (s0 + u - 1);
#pragma unroll
    for (unsigned b = 0; b < (1u << (R - 1)); b++) {
      const unsigned kl = b & (h - 1);
      const unsigned j0 = ((b >> (u - 1)) << u) | kl, j1 = j0 + h;
      uint32_t v = shoup(x[j1], __ldg(big + (((H + low + (kl << s0)) << LOG_BLOCK) | glo)));
      uint32_t a = x[j0];
      x[j0] = fadd(a, v);
      x[j1] = fsub(a, v);
    }
  }
}
template <int R>
LSK_D void dif_local_big(uint32_t (&x)[1 << R], unsigned s0, unsigned low, unsigned glo, const uint2* __restrict__ big) {
#pragma unroll
  for (int u = R; u >= 1; u--) {
    const unsigned h = 1u << (u - 1), H = 1u << (s0 + u - 1);
#pragma unroll
    for (unsigned b = 0; b < (1u << (R - 1)); b++) {
      const unsigned kl = b & (h - 1);
      const unsigned j0 = ((b >> (u - 1)) << u) | kl, j1 = j0 + h;
      uint32_t a = x[j0], c = x[j1];
      x[j0] = fadd(a, c);
      x[j1] = shoup(fsub(a, c), __ldg(big + (((H + low + (kl << s0)) << LOG_BLOCK) | glo)));
    }
  }
}

// ---- N <= 4096: whole column in one CTA (small chips; launch-latency bound) -----------------------------
LSK_D void dit_stages(uint32_t* s, unsigned t, const uint2* __restrict__ tw) {
  const unsigned half_tile = 1u << (t - 1);
  for (unsigned st = 1; st <= t; st++) {
    const unsigned h = 1u << (st - 1);
    for (unsigned p = threadIdx.x; p < half_tile; p += blockDim.x) {
      unsigned k = p & (h - 1);
      unsigned i0 = ((p >> (st - 1)) << st) | k, i1 = i0 + h;
      uint32_t u = s[i0], v = shoup(s[i1], __ldg(tw + h + k));
      s[i0] = fadd(u, v);
      s[i1] = fsub(u, v);
    }
    __syncthreads();
  }
}
LSK_D void dif_stages(uint32_t* s, unsigned t, const uint2* __restrict__ tw) {
  const unsigned half_tile = 1u << (t - 1);
  for (unsigned st = t; st >= 1; st--) {
    const unsigned h = 1u << (st - 1);
    for (unsigned p = threadIdx.x; p < half_tile; p += blockDim.x) {
      unsigned k = p & (h - 1);
      unsigned i0 = ((p >> (st - 1)) << st) | k, i1 = i0 + h;
      uint32_t u = s[i0], v = s[i1];
      s[i0] = fadd(u, v);
      s[i1] = shoup(fsub(u, v), __ldg(tw + h + k));
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(1024) lde_small_kernel(const uint32_t* __restrict__ in, size_t in_stride,
                                                          uint32_t* __restrict__ out, size_t out_stride, unsigned logn,
                                                          const uint2* __restrict__ tw_fwd, const uint2* __restrict__ tw_inv,
                                                          const uint2* __restrict__ scale_a) {
  extern __shared__ uint32_t smem[];
  const unsigned n = 1u << logn;
  uint32_t* coef = smem;
  uint32_t* work = smem + n;
  const uint32_t* src = in + (size_t)blockIdx.x * in_stride;
  uint32_t* dst = out + (size_t)blockIdx.x * out_stride;
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) coef[i] = src[i];
  __syncthreads();
  if (logn) dit_stages(coef, logn, tw_inv);
  for (int j = 0; j < 2; j++) {
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) work[i] = shoup(coef[i], __ldg(scale_a + j * n + i));
    __syncthreads();
    if (logn) dif_stages(work, logn, tw_fwd);
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[(size_t)j * n + i] = work[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) lde_small_batch_kernel(const LdeSmallJob* __restrict__ jobs,
                                                                const uint2* __restrict__ tw_fwd,
                                                                const uint2* __restrict__ tw_inv) {
  extern __shared__ uint32_t smem[];
  const LdeSmallJob job = jobs[blockIdx.x];
  const unsigned logn = job.logn, n = 1u << logn;
  uint32_t* coef = smem;
  uint32_t* work = smem + n;
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) coef[i] = job.in[i];
  __syncthreads();
  if (logn) dit_stages(coef, logn, tw_inv);
  for (int j = 0; j < 2; j++) {
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) work[i] = shoup(coef[i], __ldg(job.scale + j * n + i));
    __syncthreads();
    if (logn) dif_stages(work, logn, tw_fwd);
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) job.out[(size_t)j * n + i] = work[i];
    __syncthreads();
  }
}

// ---- N > 4096 ---------------------------------------------------------------------------------------
// 4096-element tile, 256 threads x 16 registers.  Shared-memory index of element i is i + (i >> 4) (one
// pad word per 16) so that both "16 consecutive per thread" and "stride 16 / 256 per thread" are
// conflict-free.
constexpr int K13_THREADS = 256;
constexpr unsigned K13_SMEM = BLOCK_ELEMS + BLOCK_ELEMS / 16;
LSK_D unsigned pad16(unsigned i) { return i + (i >> 4); }

// K1: DIT stages 1..12 on each contiguous 4096-block; in (N per column) -> out rows [0, N)
__global__ void __launch_bounds__(K13_THREADS) lde_k1_kernel(const uint32_t* __restrict__ in, size_t in_stride,
                                                              uint32_t* __restrict__ out, size_t out_stride,
                                                              const uint2* __restrict__ tw_inv) {
  __shared__ uint32_t s[K13_SMEM];
  const uint32_t* src = in + (size_t)blockIdx.y * in_stride + (size_t)blockIdx.x * BLOCK_ELEMS;
  uint32_t* dst = out + (size_t)blockIdx.y * out_stride + (size_t)blockIdx.x * BLOCK_ELEMS;
  const unsigned t = threadIdx.x, hi = t >> 4, lo = t & 15;
  uint32_t x[16];
#pragma unroll
  for (int k = 0; k < 4; k++) {  // elements 16 t .. 16 t + 15: index bits 0-3 in registers
    uint4 v = reinterpret_cast<const uint4*>(src)[t * 4 + k];
    x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
  }
  dit_local<4>(x, 0, 0, tw_inv);
#pragma unroll
  for (int j = 0; j < 16; j++) s[pad16(t * 16 + j)] = x[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = s[pad16(hi * 256 + j * 16 + lo)];  // bits 4-7 in registers
  dit_local<4>(x, 4, lo, tw_inv);
#pragma unroll
  for (int j = 0; j < 16; j++) s[pad16(hi * 256 + j * 16 + lo)] = x[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = s[pad16(j * 256 + t)];             // bits 8-11 in registers
  dit_local<4>(x, 8, t, tw_inv);
#pragma unroll
  for (int j = 0; j < 16; j++) dst[j * 256 + t] = x[j];
}

// K3: DIF stages 12..1 in place on each contiguous 4096-block of the 2N outputs
__global__ void __launch_bounds__(K13_THREADS) lde_k3_kernel(uint32_t* __restrict__ out, size_t out_stride,
                                                              const uint2* __restrict__ tw_fwd) {
  __shared__ uint32_t s[K13_SMEM];
  uint32_t* p = out + (size_t)blockIdx.y * out_stride + (size_t)blockIdx.x * BLOCK_ELEMS;
  const unsigned t = threadIdx.x, hi = t >> 4, lo = t & 15;
  uint32_t x[16];
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = p[j * 256 + t];
  dif_local<4>(x, 8, t, tw_fwd);
#pragma unroll
  for (int j = 0; j < 16; j++) s[pad16(j * 256 + t)] = x[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = s[pad16(hi * 256 + j * 16 + lo)];
  dif_local<4>(x, 4, lo, tw_fwd);
#pragma unroll
  for (int j = 0; j < 16; j++) s[pad16(hi * 256 + j * 16 + lo)] = x[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = s[pad16(t * 16 + j)];
  dif_local<4>(x, 0, 0, tw_fwd);
#pragma unroll
  for (int k = 0; k < 4; k++)
    reinterpret_cast<uint4*>(p)[t * 4 + k] = make_uint4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
}


// K3 with the tile staged by a bulk-asynchronous copy (cp.async.bulk + mbarrier: the 1-D form of TMA, SASS UBLKCP):
// persistent CTAs, two 16 KB input buffers; while the CTA transforms tile i the copy engine fetches tile i + 1, so the
// loads that lde_k3_kernel waits for (ncu: long_scoreboard is its top stall) are off the critical path.  Selected
// with LSK_LDE_BULK=1; measured against the plain kernel in DESIGN.md section 10.
LSK_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(K13_THREADS) lde_k3_bulk_kernel(uint32_t* __restrict__ out, size_t out_stride, unsigned tiles_per_col,
                                                                   unsigned ntiles, const uint2* __restrict__ tw_fwd) {
  extern __shared__ __align__(128) uint32_t k3b_smem[];   // in[2][4096] ++ s[K13_SMEM]: 50 KB, above the static limit
  uint32_t(*in)[BLOCK_ELEMS] = reinterpret_cast<uint32_t(*)[BLOCK_ELEMS]>(k3b_smem);
  uint32_t* s = k3b_smem + 2 * BLOCK_ELEMS;
  __shared__ __align__(8) uint64_t bar[2];
  const unsigned t = threadIdx.x, hi = t >> 4, lo = t & 15;
  auto tile_ptr = [&](unsigned tile) { return out + (size_t)(tile / tiles_per_col) * out_stride + (size_t)(tile % tiles_per_col) * BLOCK_ELEMS; };
  auto issue = [&](unsigned tile, unsigned b) {   // one thread: arm the barrier with the byte count, start the copy
    const uint32_t bytes = BLOCK_ELEMS * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(in[b])),
                 "l"(tile_ptr(tile)), "r"(bytes), "r"(smem_u32(&bar[b]))
                 : "memory");
  };
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (t == 0 && blockIdx.x < ntiles) issue(blockIdx.x, 0);
  unsigned it = 0;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    const unsigned b = it & 1, parity = (it >> 1) & 1;
    if (t == 0 && tile + gridDim.x < ntiles) issue(tile + gridDim.x, b ^ 1);
    {  // wait for this tile's bytes
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(smem_u32(&bar[b])), "r"(parity)
                     : "memory");
    }
    uint32_t* p = tile_ptr(tile);
    uint32_t x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = in[b][j * 256 + t];
    dif_local<4>(x, 8, t, tw_fwd);
#pragma unroll
    for (int j = 0; j < 16; j++) s[pad16(j * 256 + t)] = x[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = s[pad16(hi * 256 + j * 16 + lo)];
    dif_local<4>(x, 4, lo, tw_fwd);
#pragma unroll
    for (int j = 0; j < 16; j++) s[pad16(hi * 256 + j * 16 + lo)] = x[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = s[pad16(t * 16 + j)];
    dif_local<4>(x, 0, 0, tw_fwd);
#pragma unroll
    for (int k = 0; k < 4; k++)
      reinterpret_cast<uint4*>(p)[t * 4 + k] = make_uint4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    __syncthreads();   // `s` and in[b] are free again (in[b] is refilled by the copy issued in the next iteration)
  }
}

// K2: the middle pass over a tile of 2^t blocks (q) x 2^c consecutive low indices (lo), element (q, lo)
// at shared index (q << c) | lo.  Stages 13..logn over q run as ceil(t/4) register passes of 2^r (r <= 4)
// elements each; the first DIT pass reads global memory, the last DIF pass of each coset writes it.
struct K2Args {
  uint32_t* out;
  size_t out_stride;
  const uint2 *big_f, *big_i, *scale_a, *scale_b;
};
enum { K2_GLOBAL = 0, K2_COEF = 1, K2_WORK = 2 };
// Everything about a launch is fixed by T = logn - 12, so the kernel is instantiated per T (12 instances): tile shape,
// pass plan, loop trip counts, shifts and masks are compile-time constants (round 1 decoded them at run time: ~310
// instructions per element, a third of them index arithmetic).
template <int T>
struct K2Plan {
  static constexpr int LOG_TILE = T <= 8 ? 12 : (T <= 10 ? T + 4 : 14);   // 2^T blocks x 2^C low indices
  static constexpr int C = LOG_TILE - T;
  static constexpr int THREADS = 1 << (LOG_TILE - 4);                     // 16 elements per thread
  static constexpr int NP = (T + 3) / 4, BASE = T / NP, REM = T % NP;     // passes as even as possible (10 -> 4, 3, 3)
  static constexpr int r(int p) { return BASE + (p < REM ? 1 : 0); }
  static constexpr int s0(int p) { return p * BASE + (p < REM ? p : REM); }
};
template <int T, int R, int S0, bool DIF, int SRC, int DST>
LSK_D void k2_pass(const K2Args& a, uint32_t* __restrict__ col, uint32_t* __restrict__ coef, uint32_t* __restrict__ work,
                   int coset) {
  using Pl = K2Plan<T>;
  constexpr unsigned c = Pl::C, cmask = (1u << c) - 1;
  constexpr unsigned nitems = 1u << (T - R + c);
  constexpr size_t n = (size_t)1 << (T + LOG_BLOCK);
  const unsigned lo_base = blockIdx.x << c;
#pragma unroll
  for (unsigned it0 = 0; it0 < nitems; it0 += Pl::THREADS) {
    const unsigned it = it0 + threadIdx.x;
    const unsigned lo = it & cmask, rest = it >> c;
    const unsigned ql = rest & ((1u << S0) - 1), qh = rest >> S0;
    const unsigned qbase = (qh << (S0 + R)) | ql, glo = lo_base + lo;
    uint32_t x[1 << R];
    if (SRC == K2_GLOBAL) {
#pragma unroll
      for (int jj = 0; jj < (1 << R); jj++) x[jj] = col[((size_t)(qbase | (jj << S0)) << LOG_BLOCK) | glo];
    } else if (SRC == K2_COEF && DIF) {  // first DIF pass of a coset: coefficient m = q * 4096 + lo times s_j^m / N
      const uint2 sa = __ldg(a.scale_a + coset * BLOCK_ELEMS + glo);
#pragma unroll
      for (int jj = 0; jj < (1 << R); jj++) {
        const unsigned q = qbase | (jj << S0);
        x[jj] = shoup(shoup(coef[(q << c) | lo], sa), __ldg(a.scale_b + ((size_t)coset << T) + q));
      }
    } else {
      const uint32_t* sp = SRC == K2_COEF ? coef : work;
#pragma unroll
      for (int jj = 0; jj < (1 << R); jj++) x[jj] = sp[((qbase | (jj << S0)) << c) | lo];
    }
    if (DIF) dif_local_big<R>(x, S0, ql, glo, a.big_f);
    else dit_local_big<R>(x, S0, ql, glo, a.big_i);
    if (DST == K2_GLOBAL) {
#pragma unroll
      for (int jj = 0; jj < (1 << R); jj++)
        col[(size_t)coset * n + (((size_t)(qbase | (jj << S0)) << LOG_BLOCK) | glo)] = x[jj];
    } else {
      uint32_t* dp = DST == K2_COEF ? coef : work;
#pragma unroll
      for (int jj = 0; jj < (1 << R); jj++) dp[((qbase | (jj << S0)) << c) | lo] = x[jj];
    }
  }
}
template <int T, int P>
LSK_D void k2_dit_passes(const K2Args& a, uint32_t* col, uint32_t* coef, uint32_t* work) {   // stages 13 .. logn
  using Pl = K2Plan<T>;
  if constexpr (P < Pl::NP) {
    k2_pass<T, Pl::r(P), Pl::s0(P), false, P == 0 ? K2_GLOBAL : K2_COEF, K2_COEF>(a, col, coef, work, 0);
    __syncthreads();
    k2_dit_passes<T, P + 1>(a, col, coef, work);
  }
}
template <int T, int P>
LSK_D void k2_dif_passes(const K2Args& a, uint32_t* col, uint32_t* coef, uint32_t* work, int coset) {   // stages logn .. 13
  using Pl = K2Plan<T>;
  if constexpr (P >= 0) {
    k2_pass<T, Pl::r(P), Pl::s0(P), true, P == Pl::NP - 1 ? K2_COEF : K2_WORK, P == 0 ? K2_GLOBAL : K2_WORK>(a, col, coef, work,
                                                                                                         coset);
    __syncthreads();
    k2_dif_passes<T, P - 1>(a, col, coef, work, coset);
  }
}
// 256-thread tiles: two resident CTAs per SM (128 registers, next to no spilling in the fully unrolled passes) measured
// faster than three (80 registers, 192 B of spills at T = 8): 5.39 vs 5.97 ms for the LDEs of a Lurk fib 2^20 prove.
template <int T>
__global__ void __launch_bounds__(K2Plan<T>::THREADS, K2Plan<T>::THREADS == 256 ? 2 : 1) lde_k2_kernel(K2Args a) {
  extern __shared__ uint32_t smem[];
  uint32_t* coef = smem;
  uint32_t* work = smem + ((size_t)1 << K2Plan<T>::LOG_TILE);
  uint32_t* col = a.out + (size_t)blockIdx.y * a.out_stride;
  k2_dit_passes<T, 0>(a, col, coef, work);
  for (int coset = 0; coset < 2; coset++) k2_dif_passes<T, K2Plan<T>::NP - 1>(a, col, coef, work, coset);
}
template <int T>
static void k2_set_smem() {
  LSK_CUDA(cudaFuncSetAttribute(lde_k2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * 4) << K2Plan<T>::LOG_TILE));
}
template <int T>
static void k2_launch(const K2Args& a, uint32_t nc, cudaStream_t st) {
  using Pl = K2Plan<T>;
  lde_k2_kernel<T><<<dim3(BLOCK_ELEMS >> Pl::C, nc), Pl::THREADS, (2 * 4) << Pl::LOG_TILE, st>>>(a);
}

constexpr size_t K3B_SMEM_BYTES = (2 * BLOCK_ELEMS + K13_SMEM) * sizeof(uint32_t);
void init_ntt_kernels() {
  LSK_CUDA(cudaFuncSetAttribute(lde_k3_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K3B_SMEM_BYTES));
  k2_set_smem<1>(); k2_set_smem<2>(); k2_set_smem<3>(); k2_set_smem<4>(); k2_set_smem<5>(); k2_set_smem<6>();
  k2_set_smem<7>(); k2_set_smem<8>(); k2_set_smem<9>(); k2_set_smem<10>(); k2_set_smem<11>(); k2_set_smem<12>();
}

void lde_flush(lsk_ctx* ctx, LdeBatch& batch) {
  if (batch.jobs.empty()) return;
  NttTables* t = tables(ctx);
  DevBuf<LdeSmallJob> d_jobs(ctx, batch.jobs.size());
  d_jobs.upload(batch.jobs.data(), batch.jobs.size());
  unsigned n = 1u << batch.max_logn;
  unsigned threads = n / 2 < 32 ? 32 : (n / 2 > 1024 ? 1024 : n / 2);
  double bytes = 0;
  for (auto& j : batch.jobs) bytes += 12.0 * (double)(1u << j.logn);
  ProfScope ps(ctx, PT_LDE_SMALL, 1, bytes);
  lde_small_batch_kernel<<<(unsigned)batch.jobs.size(), threads, 2 * n * sizeof(uint32_t), ctx->stream>>>(
      d_jobs.p, t->tw_fwd.p, t->tw_inv.p);
  LSK_CUDA(cudaGetLastError());
  batch.jobs.clear();
  batch.max_logn = 0;
}

void lde_dev(lsk_ctx* ctx, const uint32_t* in, size_t in_stride, uint32_t* out, size_t out_stride, unsigned logn,
             uint32_t ncols, LdeShift kind, LdeBatch* batch) {
  if (!ncols) return;
  if (logn > 24) fail(LSK_ERR_ARG, "trace height 2^%u exceeds the supported 2^24", logn);
  NttTables* t = tables(ctx);
  NttTables::Scale* sc = scale_tables(ctx, logn, kind);
  cudaStream_t st = ctx->stream;
  if (logn <= LOG_BLOCK && batch) {
    for (uint32_t c = 0; c < ncols; c++)
      batch->jobs.push_back(LdeSmallJob{in + (size_t)c * in_stride, out + (size_t)c * out_stride, sc->a.p, logn});
    batch->max_logn = std::max(batch->max_logn, logn);
    return;
  }
  if (logn <= LOG_BLOCK) {
    unsigned n = 1u << logn;
    unsigned threads = n / 2 < 32 ? 32 : (n / 2 > 1024 ? 1024 : n / 2);
    ProfScope ps(ctx, PT_LDE_SMALL, 1, 12.0 * n * ncols);
    lde_small_kernel<<<ncols, threads, 2 * n * sizeof(uint32_t), st>>>(in, in_stride, out, out_stride, logn,
                                                                      t->tw_fwd.p, t->tw_inv.p, sc->a.p);
    LSK_CUDA(cudaGetLastError());
    return;
  }
  big_tables(ctx, logn);
  unsigned tt = logn - LOG_BLOCK;
  unsigned blocks = 1u << tt;
  if (ncols > 65535) fail(LSK_ERR_ARG, "too many columns");
  // tile = 2^t blocks x 2^c low indices = 2^12 .. 2^14 elements (K2Plan), two copies in shared memory
  // (coefficients + work), 16 elements per thread
  // Column chunking (LSK_LDE_L2_MB, default off).  The three passes of a column exchange 4N + 8N + 8N bytes through
  // `out`; run over a chunk of columns whose 12N-byte footprints fit the L2 together, K2 and K3 find what the previous
  // pass wrote still on chip.  MEASURED AND REJECTED (B200, Lurk fib 2^20, 80 MB chunks): K1+K2+K3 19.4 ms against 16.3 ms
  // with all columns per launch -- the passes are bound by instruction issue / the multiply pipe (ncu: K2 executes ~310
  // instructions per element, fmaheavy 68 % busy, math_pipe_throttle the top stall, DRAM at 0.9 TB/s), not by HBM, and
  // the smaller grids only add partial waves.  profiles/r02_ncu_lde.csv.
  static const size_t l2_budget = [] {
    const char* e = getenv("LSK_LDE_L2_MB");
    return (size_t)(e ? std::max(0L, atol(e)) : 0) << 20;
  }();
  const size_t col_bytes = (size_t)12 << logn;
  const uint32_t chunk = l2_budget ? (uint32_t)std::min<size_t>(ncols, std::max<size_t>(1, l2_budget / col_bytes)) : ncols;
  for (uint32_t c0 = 0; c0 < ncols; c0 += chunk) {
    const uint32_t nc = std::min(chunk, ncols - c0);
    const uint32_t* cin = in + (size_t)c0 * in_stride;
    uint32_t* cout = out + (size_t)c0 * out_stride;
    const double cells = (double)((size_t)1 << logn) * nc;
    {
      ProfScope ps(ctx, PT_LDE_K1, 1, 8.0 * cells);
      lde_k1_kernel<<<dim3(blocks, nc), K13_THREADS, 0, st>>>(cin, in_stride, cout, out_stride, t->tw_inv.p);
    }
    {
      ProfScope ps(ctx, PT_LDE_K2, 1, 12.0 * cells);
      K2Args a{cout, out_stride, t->big_f.p, t->big_i.p, sc->a.p, sc->b.p};
      switch (tt) {
        case 1: k2_launch<1>(a, nc, st); break;
        case 2: k2_launch<2>(a, nc, st); break;
        case 3: k2_launch<3>(a, nc, st); break;
        case 4: k2_launch<4>(a, nc, st); break;
        case 5: k2_launch<5>(a, nc, st); break;
        case 6: k2_launch<6>(a, nc, st); break;
        case 7: k2_launch<7>(a, nc, st); break;
        case 8: k2_launch<8>(a, nc, st); break;
        case 9: k2_launch<9>(a, nc, st); break;
        case 10: k2_launch<10>(a, nc, st); break;
        case 11: k2_launch<11>(a, nc, st); break;
        default: k2_launch<12>(a, nc, st); break;
      }
    }
    {
      ProfScope ps(ctx, PT_LDE_K3, 1, 16.0 * cells);
      static const bool bulk = [] { const char* e = getenv("LSK_LDE_BULK"); return e && e[0] == '1'; }();
      if (bulk) {
        const unsigned ntiles = 2 * blocks * nc;
        lde_k3_bulk_kernel<<<std::min(ntiles, 148u * 4u), K13_THREADS, K3B_SMEM_BYTES, st>>>(cout, out_stride, 2 * blocks, ntiles,
                                                                                             t->tw_fwd.p);
      } else
      lde_k3_kernel<<<dim3(2 * blocks, nc), K13_THREADS, 0, st>>>(cout, out_stride, t->tw_fwd.p);
    }
  }
  LSK_CUDA(cudaGetLastError());
}

}  // namespace lsk
