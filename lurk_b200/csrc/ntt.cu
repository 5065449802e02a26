// Coset LDE kernels (see ntt.cuh for the contract and the reference map).
//
// Algorithm.  Per column: inverse NTT (decimation in time: bit-reversed in, natural out), scale
// coefficient m by s_j^m / N for the two target cosets s_0 = g/in_shift, s_1 = s_0 * w_2N, forward
// NTT (decimation in frequency: natural in, bit-reversed out).  Neither direction needs a
// permutation pass.  For N <= 2^12 one CTA does the whole column in shared memory.  Larger N is
// split Bailey-style at bit 12, N = 2^12 * 2^t:
//   K1  DIT stages 1..12 on contiguous 4096-blocks                      (4R + 4W bytes / element)
//   K2  per group of C consecutive low indices: inter-pass twiddle, DIT over the 2^t blocks ->
//       coefficients (kept in registers), then for both cosets: scale, DIF over the blocks,
//       inter-pass twiddle, store                                        (4R + 8W)
//   K3  DIF stages 12..1 on contiguous 4096-blocks of both cosets        (8R + 8W)
// i.e. 36 B of HBM traffic per input element against the 12 B algorithmic minimum (4 read, 8
// written), and ~3*logn/2 modmuls per element: at B200's ~5 instructions per HBM byte the
// butterflies, not HBM, bound this stage.  Small stage twiddles w_{2^s}^k are read from a 16 KB
// level-packed table (L1-resident); they do not depend on N.
#include <map>
#include <memory>
#include <vector>

#include "field.cuh"
#include "ntt.cuh"

namespace lsk {

constexpr unsigned LOG_BLOCK = 12;
constexpr unsigned BLOCK_ELEMS = 1u << LOG_BLOCK;

// level-packed twiddles: tw[2^(s-1) + k] = w_{2^s}^(+-k), k < 2^(s-1), s = 1..12  (index 0 unused)
struct NttTables {
  DevBuf<uint32_t> tw_fwd, tw_inv;
  struct Big {  // per logn > 12
    DevBuf<uint32_t> lo_f, hi_f, lo_i, hi_i;  // W_N^(+-x), x < 4096; W_N^(+-4096 y), y < N / 4096
  };
  struct Scale {  // per (logn, kind): a[j][lo] = s_j^lo / N (lo < min(N, 4096)); b[j][k2] = s_j^(4096 k2)
    DevBuf<uint32_t> a, b;
  };
  std::map<unsigned, std::unique_ptr<Big>> big;
  std::map<std::pair<unsigned, int>, std::unique_ptr<Scale>> scale;
};

static std::vector<uint32_t> powers(uint32_t base, size_t n, uint32_t first = R1) {
  std::vector<uint32_t> v(n);
  uint32_t p = first;
  for (size_t i = 0; i < n; i++) v[i] = p, p = fmul(p, base);
  return v;
}

static NttTables* tables(lsk_ctx* ctx) {
  if (!ctx->ntt_tables) {
    auto* t = new NttTables;
    std::vector<uint32_t> f(BLOCK_ELEMS, 0), inv(BLOCK_ELEMS, 0);
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

static NttTables::Big* big_tables(lsk_ctx* ctx, unsigned logn) {
  NttTables* t = tables(ctx);
  auto it = t->big.find(logn);
  if (it != t->big.end()) return it->second.get();
  auto b = std::make_unique<NttTables::Big>();
  uint32_t w = two_adic_generator(logn), wi = finv(w);
  size_t nhi = (size_t)1 << (logn - LOG_BLOCK);
  auto lo_f = powers(w, BLOCK_ELEMS), lo_i = powers(wi, BLOCK_ELEMS);
  auto hi_f = powers(fpow(w, BLOCK_ELEMS), nhi), hi_i = powers(fpow(wi, BLOCK_ELEMS), nhi);
  b->lo_f.alloc(ctx, BLOCK_ELEMS); b->lo_f.upload(lo_f.data(), BLOCK_ELEMS);
  b->lo_i.alloc(ctx, BLOCK_ELEMS); b->lo_i.upload(lo_i.data(), BLOCK_ELEMS);
  b->hi_f.alloc(ctx, nhi); b->hi_f.upload(hi_f.data(), nhi);
  b->hi_i.alloc(ctx, nhi); b->hi_i.upload(hi_i.data(), nhi);
  LSK_CUDA(cudaStreamSynchronize(ctx->stream));  // the host vectors die here
  return (t->big[logn] = std::move(b)).get();
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
  std::vector<uint32_t> a, b;
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

// ---- shared-memory butterfly stages ---------------------------------------------------------------
// The tile holds 2^t "rows" of 2^c consecutive elements (element (q, j) at s[(q << c) | j]); the
// transform runs over q.  c = 0 is a plain contiguous transform.
LSK_D void dit_stages(uint32_t* s, unsigned t, unsigned c, const uint32_t* __restrict__ tw) {
  const unsigned half_tile = 1u << (t + c - 1);
  for (unsigned st = 1; st <= t; st++) {
    const unsigned h = 1u << (st - 1);
    for (unsigned b = threadIdx.x; b < half_tile; b += blockDim.x) {
      unsigned j = b & ((1u << c) - 1), p = b >> c;
      unsigned k = p & (h - 1);
      unsigned q0 = ((p >> (st - 1)) << st) | k;
      unsigned i0 = (q0 << c) | j, i1 = i0 + (h << c);
      uint32_t u = s[i0], v = fmul(s[i1], __ldg(tw + h + k));
      s[i0] = fadd(u, v);
      s[i1] = fsub(u, v);
    }
    __syncthreads();
  }
}
LSK_D void dif_stages(uint32_t* s, unsigned t, unsigned c, const uint32_t* __restrict__ tw) {
  const unsigned half_tile = 1u << (t + c - 1);
  for (unsigned st = t; st >= 1; st--) {
    const unsigned h = 1u << (st - 1);
    for (unsigned b = threadIdx.x; b < half_tile; b += blockDim.x) {
      unsigned j = b & ((1u << c) - 1), p = b >> c;
      unsigned k = p & (h - 1);
      unsigned q0 = ((p >> (st - 1)) << st) | k;
      unsigned i0 = (q0 << c) | j, i1 = i0 + (h << c);
      uint32_t u = s[i0], v = s[i1];
      s[i0] = fadd(u, v);
      s[i1] = fmul(fsub(u, v), __ldg(tw + h + k));
    }
    __syncthreads();
  }
}

// ---- N <= 4096: whole column in one CTA -------------------------------------------------------------
__global__ void __launch_bounds__(1024) lde_small_kernel(const uint32_t* __restrict__ in, size_t in_stride,
                                                          uint32_t* __restrict__ out, size_t out_stride, unsigned logn,
                                                          const uint32_t* __restrict__ tw_fwd,
                                                          const uint32_t* __restrict__ tw_inv,
                                                          const uint32_t* __restrict__ scale_a) {
  extern __shared__ uint32_t smem[];
  const unsigned n = 1u << logn;
  uint32_t* coef = smem;
  uint32_t* work = smem + n;
  const uint32_t* src = in + (size_t)blockIdx.x * in_stride;
  uint32_t* dst = out + (size_t)blockIdx.x * out_stride;
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) coef[i] = src[i];
  __syncthreads();
  if (logn) dit_stages(coef, logn, 0, tw_inv);
  for (int j = 0; j < 2; j++) {
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) work[i] = fmul(coef[i], __ldg(scale_a + j * n + i));
    __syncthreads();
    if (logn) dif_stages(work, logn, 0, tw_fwd);
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[(size_t)j * n + i] = work[i];
    __syncthreads();
  }
}

// ---- N > 4096 ---------------------------------------------------------------------------------------
// K1: DIT stages 1..12 on each contiguous 4096-block; in (N per column) -> out rows [0, N)
__global__ void __launch_bounds__(1024) lde_k1_kernel(const uint32_t* __restrict__ in, size_t in_stride,
                                                       uint32_t* __restrict__ out, size_t out_stride,
                                                       const uint32_t* __restrict__ tw_inv) {
  __shared__ uint32_t s[BLOCK_ELEMS];
  const uint32_t* src = in + (size_t)blockIdx.y * in_stride + (size_t)blockIdx.x * BLOCK_ELEMS;
  uint32_t* dst = out + (size_t)blockIdx.y * out_stride + (size_t)blockIdx.x * BLOCK_ELEMS;
  for (unsigned i = threadIdx.x; i < BLOCK_ELEMS / 4; i += blockDim.x)
    reinterpret_cast<uint4*>(s)[i] = reinterpret_cast<const uint4*>(src)[i];
  __syncthreads();
  dit_stages(s, LOG_BLOCK, 0, tw_inv);
  for (unsigned i = threadIdx.x; i < BLOCK_ELEMS / 4; i += blockDim.x)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(s)[i];
}

// K2: the middle pass.  EPT elements per thread; tile = EPT * 1024 = 2^t blocks x 2^c low indices.
template <int EPT>
__global__ void __launch_bounds__(1024) lde_k2_kernel(uint32_t* __restrict__ out, size_t out_stride, unsigned logn,
                                                       unsigned c, const uint32_t* __restrict__ tw_fwd,
                                                       const uint32_t* __restrict__ tw_inv,
                                                       const uint32_t* __restrict__ lo_f, const uint32_t* __restrict__ hi_f,
                                                       const uint32_t* __restrict__ lo_i, const uint32_t* __restrict__ hi_i,
                                                       const uint32_t* __restrict__ scale_a,
                                                       const uint32_t* __restrict__ scale_b) {
  extern __shared__ uint32_t smem[];
  const unsigned t = logn - LOG_BLOCK;
  const unsigned cmask = (1u << c) - 1;
  const unsigned lo_base = blockIdx.x << c;
  const size_t n = (size_t)1 << logn;
  uint32_t* col = out + (size_t)blockIdx.y * out_stride;
  uint32_t keep[EPT];
  // load + inverse inter-pass twiddle W_N^(-bitrev_t(q) * lo)
#pragma unroll
  for (int r = 0; r < EPT; r++) {
    unsigned e = threadIdx.x + r * 1024;
    unsigned q = e >> c, lo = lo_base + (e & cmask);
    unsigned ex = bitrev32(q, t) * lo;
    uint32_t w = fmul(__ldg(lo_i + (ex & (BLOCK_ELEMS - 1))), __ldg(hi_i + (ex >> LOG_BLOCK)));
    smem[e] = fmul(col[((size_t)q << LOG_BLOCK) | lo], w);
  }
  __syncthreads();
  dit_stages(smem, t, c, tw_inv);
#pragma unroll
  for (int r = 0; r < EPT; r++) keep[r] = smem[threadIdx.x + r * 1024];
  for (int j = 0; j < 2; j++) {
    const uint32_t* sa = scale_a + j * BLOCK_ELEMS;
    const uint32_t* sb = scale_b + ((size_t)j << t);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < EPT; r++) {
      unsigned e = threadIdx.x + r * 1024;
      unsigned q = e >> c, lo = lo_base + (e & cmask);
      smem[e] = fmul(keep[r], fmul(__ldg(sa + lo), __ldg(sb + q)));
    }
    __syncthreads();
    dif_stages(smem, t, c, tw_fwd);
#pragma unroll
    for (int r = 0; r < EPT; r++) {
      unsigned e = threadIdx.x + r * 1024;
      unsigned q = e >> c, lo = lo_base + (e & cmask);
      unsigned ex = bitrev32(q, t) * lo;
      uint32_t w = fmul(__ldg(lo_f + (ex & (BLOCK_ELEMS - 1))), __ldg(hi_f + (ex >> LOG_BLOCK)));
      col[(size_t)j * n + (((size_t)q << LOG_BLOCK) | lo)] = fmul(smem[e], w);
    }
  }
}

// K3: DIF stages 12..1 in place on each contiguous 4096-block of the 2N outputs
__global__ void __launch_bounds__(1024) lde_k3_kernel(uint32_t* __restrict__ out, size_t out_stride,
                                                       const uint32_t* __restrict__ tw_fwd) {
  __shared__ uint32_t s[BLOCK_ELEMS];
  uint32_t* p = out + (size_t)blockIdx.y * out_stride + (size_t)blockIdx.x * BLOCK_ELEMS;
  for (unsigned i = threadIdx.x; i < BLOCK_ELEMS / 4; i += blockDim.x)
    reinterpret_cast<uint4*>(s)[i] = reinterpret_cast<const uint4*>(p)[i];
  __syncthreads();
  dif_stages(s, LOG_BLOCK, 0, tw_fwd);
  for (unsigned i = threadIdx.x; i < BLOCK_ELEMS / 4; i += blockDim.x)
    reinterpret_cast<uint4*>(p)[i] = reinterpret_cast<const uint4*>(s)[i];
}

void init_ntt_kernels() {
  LSK_CUDA(cudaFuncSetAttribute(lde_k2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 1024 * 4));
  LSK_CUDA(cudaFuncSetAttribute(lde_k2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 4));
  LSK_CUDA(cudaFuncSetAttribute(lde_k2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4));
  LSK_CUDA(cudaFuncSetAttribute(lde_k2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4));
}

void lde_dev(lsk_ctx* ctx, const uint32_t* in, size_t in_stride, uint32_t* out, size_t out_stride, unsigned logn,
             uint32_t ncols, LdeShift kind) {
  if (!ncols) return;
  if (logn > 24) fail(LSK_ERR_ARG, "trace height 2^%u exceeds the supported 2^24", logn);
  NttTables* t = tables(ctx);
  NttTables::Scale* sc = scale_tables(ctx, logn, kind);
  cudaStream_t st = ctx->stream;
  if (logn <= LOG_BLOCK) {
    unsigned n = 1u << logn;
    unsigned threads = n / 2 < 32 ? 32 : (n / 2 > 1024 ? 1024 : n / 2);
    ProfScope ps(ctx, PT_LDE_SMALL, 1, 12.0 * n * ncols);
    lde_small_kernel<<<ncols, threads, 2 * n * sizeof(uint32_t), st>>>(in, in_stride, out, out_stride, logn,
                                                                      t->tw_fwd.p, t->tw_inv.p, sc->a.p);
    LSK_CUDA(cudaGetLastError());
    return;
  }
  NttTables::Big* bg = big_tables(ctx, logn);
  unsigned tt = logn - LOG_BLOCK;
  unsigned blocks = 1u << tt;
  if (ncols > 65535) fail(LSK_ERR_ARG, "too many columns");
  const double cells = (double)((size_t)1 << logn) * ncols;
  {
    ProfScope ps(ctx, PT_LDE_K1, 1, 8.0 * cells);
    lde_k1_kernel<<<dim3(blocks, ncols), 1024, 0, st>>>(in, in_stride, out, out_stride, t->tw_inv.p);
  }
  // tile = max(4096, 2^t * 16) elements, except 2^t * 8 for t >= 11
  unsigned log_tile = tt <= 8 ? 12 : (tt <= 10 ? tt + 4 : tt + 3);
  unsigned c = log_tile - tt;
  dim3 grid(BLOCK_ELEMS >> c, ncols);
  size_t smem = ((size_t)1 << log_tile) * sizeof(uint32_t);
#define K2(EPT)                                                                                                   \
  lde_k2_kernel<EPT><<<grid, 1024, smem, st>>>(out, out_stride, logn, c, t->tw_fwd.p, t->tw_inv.p, bg->lo_f.p,   \
                                               bg->hi_f.p, bg->lo_i.p, bg->hi_i.p, sc->a.p, sc->b.p)
  {
    ProfScope ps(ctx, PT_LDE_K2, 1, 12.0 * cells);
    switch (log_tile) {
      case 12: K2(4); break;
      case 13: K2(8); break;
      case 14: K2(16); break;
      case 15: K2(32); break;
      default: fail(LSK_ERR_ARG, "bad LDE tile");
    }
  }
#undef K2
  {
    ProfScope ps(ctx, PT_LDE_K3, 1, 16.0 * cells);
    lde_k3_kernel<<<dim3(2 * blocks, ncols), 1024, 0, st>>>(out, out_stride, t->tw_fwd.p);
  }
  LSK_CUDA(cudaGetLastError());
}

}  // namespace lsk
