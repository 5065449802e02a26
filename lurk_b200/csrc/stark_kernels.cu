// Device stages of the shard prover (see stark_kernels.cuh for the reference map).
//
// Roofline notes: the LogUp trace, quotient, opening and FRI-reduce kernels stream
// column-major LDE cells once with fully coalesced row-parallel access (thread = row, so a warp
// reads 128 contiguous bytes of each column) and are HBM-bound; the per-row scalar work (one
// extension-field inversion, a few dozen modmuls) rides under the loads.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "fri.cuh"
#include "poseidon2.cuh"
#include "stark_kernels.cuh"
#include "quotient_kernel.cuh"

#include "comm.cuh"

namespace lsk {

// ---- LogUp permutation trace --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) perm_rows_kernel(const uint32_t* __restrict__ inter, uint32_t nint,
                                                         const uint32_t* __restrict__ main, size_t mstride,
                                                         const uint32_t* __restrict__ prep, size_t pstride, unsigned logn,
                                                         Ext alpha, const Ext* __restrict__ beta_pows,
                                                         uint32_t* __restrict__ perm, size_t perm_stride,
                                                         size_t row_begin, size_t row_end, bool scatter,
                                                         Ext* __restrict__ rowsum_out) {
  // rows [row_begin, row_end) in storage (bit-reversed) order; the row sums go to natural order for the scan
  // (scatter) or, when a device group splits the rows, stay in storage order until they have been all-gathered
  size_t r = row_begin + (size_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= row_end) return;
  Cells cells{main, mstride, prep, pstride, r};
  const uint32_t* s = inter;
  Ext rowsum = ext_zero();
  for (uint32_t b = 0; 2 * b < nint; b++) {
    Ext r0, v;
    uint32_t m0;
    eval_interaction(s, cells, alpha, beta_pows, r0, m0);
    if (2 * b + 1 < nint) {
      Ext r1;
      uint32_t m1;
      eval_interaction(s, cells, alpha, beta_pows, r1, m1);
      Ext num = emul_base2(r1, m0, r0, m1);
      v = (m0 | m1) ? emul(num, einv(emul(r0, r1))) : ext_zero();
    } else {
      v = m0 ? emul_base(einv(r0), m0) : ext_zero();
    }
#pragma unroll
    for (int k = 0; k < 4; k++) perm[(size_t)(4 * b + k) * perm_stride + r] = v.c[k];
    rowsum = eadd(rowsum, v);
  }
  rowsum_out[scatter ? bitrev32((uint32_t)r, logn) : r] = rowsum;
}
__global__ void bitrev_permute_ext_kernel(const Ext* __restrict__ in, unsigned logn, Ext* __restrict__ out) {
  size_t r = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (r >> logn) return;
  out[bitrev32((uint32_t)r, logn)] = in[r];
}

// inclusive scan of extension elements, 1024 per block
__global__ void __launch_bounds__(256) scan_block_kernel(Ext* __restrict__ data, size_t n, Ext* __restrict__ block_sums) {
  __shared__ Ext warp_tot[8];
  size_t base = (size_t)blockIdx.x * 1024 + threadIdx.x * 4;
  Ext v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = base + k < n ? data[base + k] : ext_zero();
#pragma unroll
  for (int k = 1; k < 4; k++) v[k] = eadd(v[k], v[k - 1]);
  Ext inc = v[3];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Ext t;
#pragma unroll
    for (int k = 0; k < 4; k++) t.c[k] = __shfl_up_sync(0xffffffffu, inc.c[k], d);
    if (lane >= d) inc = eadd(inc, t);
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  Ext off = esub(inc, v[3]);  // exclusive within the warp
  for (unsigned w = 0; w < warp; w++) off = eadd(off, warp_tot[w]);
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (base + k < n) data[base + k] = eadd(v[k], off);
  if (threadIdx.x == 255 && block_sums) block_sums[blockIdx.x] = eadd(v[3], off);
}
__global__ void scan_add_kernel(Ext* __restrict__ data, size_t n, const Ext* __restrict__ block_sums) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x + 1024;  // block 0 needs no offset
  if (i < n) data[i] = eadd(data[i], block_sums[i / 1024 - 1]);
}
static void scan_ext_dev(lsk_ctx* ctx, Ext* data, size_t n) {
  size_t nb = (n + 1023) / 1024;
  if (nb <= 1) {
    ProfScope ps(ctx, PT_SCAN, 1, 32.0 * n);
    scan_block_kernel<<<1, 256, 0, ctx->stream>>>(data, n, nullptr);
    return;
  }
  DevBuf<Ext> sums(ctx, nb);
  {
    ProfScope ps(ctx, PT_SCAN, 1, 32.0 * n);
    scan_block_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(data, n, sums.p);
  }
  scan_ext_dev(ctx, sums.p, nb);
  ProfScope ps(ctx, PT_SCAN, 1, 32.0 * n);
  scan_add_kernel<<<div_up(n - 1024, 256), 256, 0, ctx->stream>>>(data, n, sums.p);
}
__global__ void phi_scatter_kernel(const Ext* __restrict__ phi_nat, unsigned logn, uint32_t* __restrict__ perm,
                                   size_t perm_stride, uint32_t first_col, Ext* __restrict__ cum_sum) {
  size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (k >> logn) return;
  Ext v = phi_nat[k];
  size_t r = bitrev32((uint32_t)k, logn);
#pragma unroll
  for (int e = 0; e < 4; e++) perm[(size_t)(first_col + e) * perm_stride + r] = v.c[e];
  if (k + 1 == ((size_t)1 << logn)) *cum_sum = v;
}

void perm_trace_dev(lsk_ctx* ctx, const ChipDev& chip, const uint32_t* main, size_t mstride, const uint32_t* prep,
                    size_t pstride, unsigned logn, Ext chal_alpha, const Ext* beta_pows_dev, uint32_t* perm_out,
                    size_t perm_stride, Ext* scratch, Ext* cum_sum_out) {
  size_t n = (size_t)1 << logn, b = 0, e = n;
  // in a device group each rank computes the batch columns and row sums of a range of rows (one extension-field
  // inversion per batch and row); columns and sums are all-gathered, the scan below runs on every rank
  const bool split = comm_splits(ctx, logn);
  if (split) comm_range(ctx, n, &b, &e);
  DevBuf<Ext> sums_pos;
  if (split) sums_pos.alloc(ctx, n);
  {
    ProfScope ps(ctx, PT_PERM_ROWS, 1, (double)(e - b) * (4.0 * (chip.width + chip.prep_width) + 16.0 * chip.pw));
    if (chip.jit) {   // the chip's interactions compiled to straight-line code (jit.cu)
      PermRowsArgs pa{main, mstride, prep, pstride, logn, chal_alpha, beta_pows_dev, perm_out, perm_stride, b, e,
                      split ? 0u : 1u, split ? sums_pos.p : scratch};
      jit_launch_perm(ctx, chip.jit.get(), pa);
    } else
    perm_rows_kernel<<<div_up(e - b, 256), 256, 0, ctx->stream>>>(chip.inter.p, chip.nint, main, mstride, prep, pstride,
                                                                  logn, chal_alpha, beta_pows_dev, perm_out, perm_stride,
                                                                  b, e, !split, split ? sums_pos.p : scratch);
    LSK_CUDA(cudaGetLastError());
  }
  if (split) {
    if (chip.pw > 1) comm_allgather_strided(ctx, perm_out, perm_stride * 4, 4 * (chip.pw - 1), (e - b) * 4);
    comm_allgather(ctx, sums_pos.p, (e - b) * sizeof(Ext));
    ProfScope ps(ctx, PT_SCAN, 1, 32.0 * n);
    bitrev_permute_ext_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(sums_pos.p, logn, scratch);
  }
  scan_ext_dev(ctx, scratch, n);
  ProfScope ps(ctx, PT_SCAN, 1, 32.0 * n);
  phi_scatter_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(scratch, logn, perm_out, perm_stride, 4 * (chip.pw - 1),
                                                              cum_sum_out);
  LSK_CUDA(cudaGetLastError());
}

// ---- quotient -----------------------------------------------------------------------------------------
void quotient_dev(lsk_ctx* ctx, const ChipDev& chip, const QuotientArgs& args) {
  QuotientArgs a = args;
  QuotientConsts qc;
  qc.g = to_mont(GEN);
  qc.w2n = two_adic_generator(a.logn + 1);
  qc.wn_inv = finv(two_adic_generator(a.logn));
  uint32_t gn = fpow(qc.g, (uint64_t)1 << a.logn);
  qc.zh[0] = fsub(gn, R1);
  qc.zh[1] = fsub(fneg(gn), R1);
  qc.inv_zh[0] = finv(qc.zh[0]);
  qc.inv_zh[1] = finv(qc.zh[1]);
  size_t n2 = (size_t)2 << a.logn;
  // in a device group each rank evaluates a contiguous range of LDE rows; the 4 output columns are all-gathered
  const bool split = comm_splits(ctx, a.logn + 1);
  a.row_begin = 0, a.row_end = n2;
  if (split) comm_range(ctx, n2, &a.row_begin, &a.row_end);
  const size_t nrows = a.row_end - a.row_begin;
  const uint4* code = reinterpret_cast<const uint4*>(chip.code.p);
  if (chip.jit) {   // the program compiled to machine code (one row per thread)
    ProfScope ps(ctx, PT_QUOTIENT, 1, (double)nrows * (4.0 * (chip.width + chip.prep_width) + 16.0 * chip.pw + 16.0));
    jit_launch(ctx, chip.jit.get(), a, qc, chip.inter.p, chip.nint, chip.pw);
  } else {
  ProfScope ps(ctx, PT_QUOTIENT, 1, (double)nrows * (4.0 * (chip.width + chip.prep_width) + 16.0 * chip.pw + 16.0));
#define QK(R, ROWS)                                                                                              \
  quotient_kernel<R, ROWS><<<div_up(nrows, 256 * ROWS), 256, 0, ctx->stream>>>(a, qc, code, chip.ninstr, chip.consts.p, \
                                                                               chip.inter.p, chip.nint, chip.pw)
  // two rows per thread for the long programs (hasher chips) once there are enough rows to fill the machine
  // (measured on the 655-wide hash4 chip, 2^21 LDE rows: 35.8 ms with one row per thread, 24.1 with two, 29.2 with four)
  const bool two = chip.ninstr >= 1024 && n2 >= ((size_t)1 << 19);
  if (chip.n_regs <= 8) QK(8, 1);
  else if (chip.n_regs <= 32) QK(32, 1);
  else if (chip.n_regs <= 128) { if (two) QK(128, 2); else QK(128, 1); }
  else if (chip.n_regs <= 1024) QK(1024, 1);
  else fail(LSK_ERR_ARG, "constraint program needs %u registers (> 1024)", chip.n_regs);
#undef QK
  LSK_CUDA(cudaGetLastError());
  }
  if (split) comm_allgather_strided(ctx, a.out, a.ostride * 4, 4, nrows * 4);
}

// ---- openings -------------------------------------------------------------------------------------------
__global__ void bary_weights_kernel(unsigned logn, uint32_t g, uint32_t wn, Ext z, Ext scale, size_t row_begin,
                                    size_t row_end, Ext* __restrict__ wts) {
  size_t r = row_begin + (size_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= row_end) return;
  uint32_t x = fmul(g, fpow(wn, bitrev32((uint32_t)r, logn)));
  Ext d = einv(esub_base(z, x));   // z - x
  wts[r] = emul(emul_base(d, x), scale);
}
void bary_weights_dev(lsk_ctx* ctx, unsigned logn, Ext z, Ext* wts) {
  size_t n = (size_t)1 << logn;
  uint32_t g = to_mont(GEN);
  uint32_t gn = fpow(g, n);
  // scale = (z^n - g^n) / (g^n * n)
  Ext zn = epow(z, n);
  Ext scale = emul_base(esub_base(zn, gn), finv(fmul(gn, to_mont((uint32_t)n))));
  size_t b = 0, e = n;
  const bool split = comm_splits(ctx, logn);   // one Ext inversion per row: a range of rows per rank, all-gathered
  if (split) comm_range(ctx, n, &b, &e);
  {
    ProfScope ps(ctx, PT_BARY, 1, 16.0 * (e - b));
    bary_weights_kernel<<<div_up(e - b, 256), 256, 0, ctx->stream>>>(logn, g, two_adic_generator(logn), z, scale, b, e, wts);
    LSK_CUDA(cudaGetLastError());
  }
  if (split) comm_allgather(ctx, wts, (e - b) * sizeof(Ext));
}

constexpr int OPEN_COLS = 4;
struct OpenGroup { uint32_t mat, c0; };
// One CTA: OPEN_ROWS rows x OPEN_COLS columns of one matrix.  The dot products are accumulated as raw 64-bit
// products kept below P * 2^32 (one conditional subtraction per term, one Montgomery reduction at the end: a third
// of the multiply-pipe work of a Montgomery product per term); rows are predicated, not broken out of, so the
// loop unrolls and the loads of two rows are in flight together.
__global__ void __launch_bounds__(256, 2) open_partial_batch_kernel(const OpenMat* __restrict__ mats,
                                                                  const OpenGroup* __restrict__ groups, size_t n,
                                                                  const Ext* __restrict__ wts0, const Ext* __restrict__ wts1,
                                                                  Ext* __restrict__ partial, uint32_t nchunks) {
  __shared__ uint32_t red[8][OPEN_COLS * 8];
  const OpenGroup g = groups[blockIdx.y];
  const OpenMat mt = mats[g.mat];
  const uint32_t c0 = g.c0, w = mt.w;
  const uint32_t* __restrict__ m = mt.p;
  const size_t stride = mt.stride;
  const size_t row0 = (size_t)blockIdx.x * OPEN_ROWS;
  const bool two = !mt.one_point;
  uint64_t acc[OPEN_COLS][2][4];
#pragma unroll
  for (int c = 0; c < OPEN_COLS; c++)
#pragma unroll
    for (int e = 0; e < 4; e++) acc[c][0][e] = acc[c][1][e] = 0;
  // two rows per step: acc < P * 2^32 plus two products < P^2 stays below 2^64, so ONE conditional subtraction
  // serves two terms
#pragma unroll 1
  for (unsigned k = 0; k < OPEN_ROWS / 256; k += 2) {
    const size_t ra = row0 + k * 256 + threadIdx.x, rb = ra + 256;
    const bool oka = ra < n, okb = rb < n;
    Ext w0a = ext_zero(), w1a = ext_zero(), w0b = ext_zero(), w1b = ext_zero();
    if (oka) {
      w0a = wts0[ra];
      if (two) w1a = wts1[ra];
    }
    if (okb) {
      w0b = wts0[rb];
      if (two) w1b = wts1[rb];
    }
    uint32_t va[OPEN_COLS], vb[OPEN_COLS];
#pragma unroll
    for (int c = 0; c < OPEN_COLS; c++) {
      va[c] = (oka && c0 + c < w) ? m[(size_t)(c0 + c) * stride + ra] : 0u;
      vb[c] = (okb && c0 + c < w) ? m[(size_t)(c0 + c) * stride + rb] : 0u;
    }
#pragma unroll
    for (int c = 0; c < OPEN_COLS; c++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        acc[c][0][e] = acc_reduce64(acc[c][0][e] + (uint64_t)w0a.c[e] * va[c] + (uint64_t)w0b.c[e] * vb[c]);
        if (two) acc[c][1][e] = acc_reduce64(acc[c][1][e] + (uint64_t)w1a.c[e] * va[c] + (uint64_t)w1b.c[e] * vb[c]);
      }
  }
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < OPEN_COLS; c++)
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        uint32_t v = mont_reduce(acc[c][p][e]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v = fadd(v, __shfl_xor_sync(0xffffffffu, v, d));
        if (lane == 0) red[warp][(c * 2 + p) * 4 + e] = v;
      }
  __syncthreads();
  if (threadIdx.x < OPEN_COLS * 8) {
    uint32_t cp = threadIdx.x / 4, e = threadIdx.x % 4;  // cp = c * 2 + p
    if (mt.one_point && (cp & 1)) cp ^= 1;               // the second output repeats the first
    uint32_t v = 0;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) v = fadd(v, red[wv][cp * 4 + e]);
    uint32_t c = c0 + (threadIdx.x / 4) / 2;
    if (c < w) partial[(size_t)(mt.out_off + c * 2 + ((threadIdx.x / 4) & 1)) * nchunks + blockIdx.x].c[e] = v;
  }
}
__global__ void __launch_bounds__(128) open_reduce_batch_kernel(const Ext* __restrict__ partial, uint32_t nchunks,
                                                                 const uint32_t* __restrict__ outs, Ext* __restrict__ out) {
  __shared__ Ext red[4];
  const uint32_t o = outs[blockIdx.x];
  const Ext* p = partial + (size_t)o * nchunks;
  Ext acc = ext_zero();
  for (uint32_t k = threadIdx.x; k < nchunks; k += 128) acc = eadd(acc, p[k]);
#pragma unroll
  for (int e = 0; e < 4; e++) {
    uint32_t v = acc.c[e];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = fadd(v, __shfl_xor_sync(0xffffffffu, v, d));
    acc.c[e] = v;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) out[o] = eadd(eadd(red[0], red[1]), eadd(red[2], red[3]));
}
void open_batch_dev(lsk_ctx* ctx, const std::vector<OpenMat>& mats, unsigned logn, const Ext* wts0, const Ext* wts1,
                    Ext* out) {
  if (mats.empty()) return;
  const size_t n = (size_t)1 << logn;
  const uint32_t nchunks = div_up(n, OPEN_ROWS);
  std::vector<OpenGroup> groups;
  std::vector<uint32_t> outs;
  uint32_t max_out = 0;
  double bytes = 0;
  for (uint32_t i = 0; i < mats.size(); i++) {
    for (uint32_t c0 = 0; c0 < mats[i].w; c0 += OPEN_COLS) groups.push_back({i, c0});
    for (uint32_t k = 0; k < 2 * mats[i].w; k++) outs.push_back(mats[i].out_off + k);
    max_out = std::max(max_out, mats[i].out_off + 2 * mats[i].w);
    bytes += (double)n * 4.0 * mats[i].w;
  }
  if (groups.empty()) return;
  DevBuf<OpenMat> d_mats(ctx, mats.size());
  DevBuf<OpenGroup> d_groups(ctx, groups.size());
  DevBuf<uint32_t> d_outs(ctx, outs.size());
  DevBuf<Ext> partial(ctx, (size_t)max_out * nchunks);
  d_mats.upload(mats.data(), mats.size());
  d_groups.upload(groups.data(), groups.size());
  d_outs.upload(outs.data(), outs.size());
  ProfScope ps(ctx, PT_OPEN, 2, bytes + 32.0 * n);
  open_partial_batch_kernel<<<dim3(nchunks, (unsigned)groups.size()), 256, 0, ctx->stream>>>(d_mats.p, d_groups.p, n, wts0,
                                                                                            wts1, partial.p, nchunks);
  open_reduce_batch_kernel<<<(unsigned)outs.size(), 128, 0, ctx->stream>>>(partial.p, nchunks, d_outs.p, out);
  LSK_CUDA(cudaGetLastError());
}

// ---- FRI: reduced openings, fold, grind --------------------------------------------------------------------
__global__ void __launch_bounds__(128) reduce_y_kernel(ReduceMat* __restrict__ mats, const Ext* __restrict__ opened,
                                                       const Ext* __restrict__ apow) {
  __shared__ Ext red[2][4];
  ReduceMat& m = mats[blockIdx.x];
  Ext acc[2] = {ext_zero(), ext_zero()};
  for (uint32_t c = threadIdx.x; c < m.w; c += 128) {
    const Ext a = apow[c];
    acc[0] = eadd(acc[0], emul(a, opened[m.open_off + 2 * c]));
    if (m.npoints > 1) acc[1] = eadd(acc[1], emul(a, opened[m.open_off + 2 * c + 1]));
  }
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < 2; p++) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      uint32_t v = acc[p].c[e];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) v = fadd(v, __shfl_xor_sync(0xffffffffu, v, d));
      if (lane == 0) red[p][warp].c[e] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) m.y[threadIdx.x] = eadd(eadd(red[threadIdx.x][0], red[threadIdx.x][1]), eadd(red[threadIdx.x][2], red[threadIdx.x][3]));
}
void reduce_y_dev(lsk_ctx* ctx, ReduceMat* mats_dev, uint32_t nmats, const Ext* opened_dev, const Ext* alpha_pows_dev) {
  if (!nmats) return;
  ProfScope ps(ctx, PT_REDUCE_OPEN, 1, 0.0);
  reduce_y_kernel<<<nmats, 128, 0, ctx->stream>>>(mats_dev, opened_dev, alpha_pows_dev);
  LSK_CUDA(cudaGetLastError());
}

template <int COLS>   // columns per loop iteration (each pair of columns shares one conditional subtraction)
__global__ void __launch_bounds__(256) reduce_openings_kernel(const ReduceMat* __restrict__ mats, uint32_t nmats,
                                                               unsigned lh, uint32_t g, uint32_t w, Ext z0, Ext z1,
                                                               const Ext* __restrict__ apow, size_t row_begin,
                                                               size_t row_end, Ext* __restrict__ ro) {
  size_t r = row_begin + (size_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= row_end) return;
  uint32_t x = fmul(g, fpow(w, bitrev32((uint32_t)r, lh)));
  Ext acc0 = ext_zero(), acc1 = ext_zero();
  bool any1 = false;
  for (uint32_t mi = 0; mi < nmats; mi++) {
    const ReduceMat& m = mats[mi];
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;   // raw 64-bit products below P * 2^32, one reduction per matrix
    const uint32_t* p = m.p + r;
    uint32_t c = 0;
    for (; c + COLS <= m.w; c += COLS) {   // the loads of COLS columns in flight together
      uint32_t v[COLS];
#pragma unroll
      for (int k = 0; k < COLS; k++) v[k] = __ldg(p + (size_t)(c + k) * m.stride);
#pragma unroll
      for (int k = 0; k < COLS; k += 2) {   // two terms per conditional subtraction (acc + 2 products < 2^64)
        const Ext ap = apow[c + k], aq = apow[c + k + 1];
        s0 = acc_reduce64(s0 + (uint64_t)ap.c[0] * v[k] + (uint64_t)aq.c[0] * v[k + 1]);
        s1 = acc_reduce64(s1 + (uint64_t)ap.c[1] * v[k] + (uint64_t)aq.c[1] * v[k + 1]);
        s2 = acc_reduce64(s2 + (uint64_t)ap.c[2] * v[k] + (uint64_t)aq.c[2] * v[k + 1]);
        s3 = acc_reduce64(s3 + (uint64_t)ap.c[3] * v[k] + (uint64_t)aq.c[3] * v[k + 1]);
      }
    }
#pragma unroll 1
    for (; c + 2 <= m.w; c += 2) {
      const Ext ap = apow[c], aq = apow[c + 1];
      const uint32_t v = p[(size_t)c * m.stride], u = p[(size_t)(c + 1) * m.stride];
      s0 = acc_reduce64(s0 + (uint64_t)ap.c[0] * v + (uint64_t)aq.c[0] * u);
      s1 = acc_reduce64(s1 + (uint64_t)ap.c[1] * v + (uint64_t)aq.c[1] * u);
      s2 = acc_reduce64(s2 + (uint64_t)ap.c[2] * v + (uint64_t)aq.c[2] * u);
      s3 = acc_reduce64(s3 + (uint64_t)ap.c[3] * v + (uint64_t)aq.c[3] * u);
    }
    if (c < m.w) {
      const Ext ap = apow[c];
      const uint32_t v = p[(size_t)c * m.stride];
      s0 = acc_reduce64(s0 + (uint64_t)ap.c[0] * v);
      s1 = acc_reduce64(s1 + (uint64_t)ap.c[1] * v);
      s2 = acc_reduce64(s2 + (uint64_t)ap.c[2] * v);
      s3 = acc_reduce64(s3 + (uint64_t)ap.c[3] * v);
    }
    const Ext s{{mont_reduce(s0), mont_reduce(s1), mont_reduce(s2), mont_reduce(s3)}};
    acc0 = eadd(acc0, emul(m.off[0], esub(m.y[0], s)));
    if (m.npoints > 1) {
      acc1 = eadd(acc1, emul(m.off[1], esub(m.y[1], s)));
      any1 = true;
    }
  }
  Ext out = emul(acc0, einv(esub_base(z0, x)));
  if (any1) out = eadd(out, emul(acc1, einv(esub_base(z1, x))));
  ro[r] = out;
}
void reduce_openings_dev(lsk_ctx* ctx, const ReduceMat* mats_dev, uint32_t nmats, unsigned lh, Ext zeta,
                         const Ext* alpha_pows_dev, Ext* ro, double algorithmic_bytes) {
  size_t n = (size_t)1 << lh, b = 0, e = n;
  const bool split = comm_splits(ctx, lh);   // each rank reduces a range of rows, the vector is all-gathered
  if (split) comm_range(ctx, n, &b, &e);
  {
    ProfScope ps(ctx, PT_REDUCE_OPEN, 1, algorithmic_bytes * (double)(e - b) / (double)n);
    Ext z1 = emul_base(zeta, two_adic_generator(lh - 1));
    // columns in flight per loop iteration, measured on B200 (Lurk fib 2^20, ms per prove): 2 -> 3.50, 4 -> 2.76, 8 -> 3.42
    // (round 1's loop, two columns and generic loads: 4.27); LSK_RO_COLS overrides
    static const int cols = [] { const char* e = getenv("LSK_RO_COLS"); return e ? atoi(e) : 4; }();
    if (cols >= 8) reduce_openings_kernel<8><<<div_up(e - b, 256), 256, 0, ctx->stream>>>(mats_dev, nmats, lh, to_mont(GEN),
                                                                        two_adic_generator(lh), zeta, z1, alpha_pows_dev,
                                                                        b, e, ro);
    else if (cols >= 4) reduce_openings_kernel<4><<<div_up(e - b, 256), 256, 0, ctx->stream>>>(mats_dev, nmats, lh, to_mont(GEN),
                                                                        two_adic_generator(lh), zeta, z1, alpha_pows_dev,
                                                                        b, e, ro);
    else reduce_openings_kernel<2><<<div_up(e - b, 256), 256, 0, ctx->stream>>>(mats_dev, nmats, lh, to_mont(GEN),
                                                                        two_adic_generator(lh), zeta, z1, alpha_pows_dev,
                                                                        b, e, ro);
    LSK_CUDA(cudaGetLastError());
  }
  if (split) comm_allgather(ctx, ro, (e - b) * sizeof(Ext));
}

__global__ void __launch_bounds__(256) fri_fold_kernel(const Ext* __restrict__ cur, unsigned logh, uint32_t ginv,
                                                        uint32_t one_half, const Ext* __restrict__ beta,
                                                        const Ext* __restrict__ add, Ext* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >> logh) return;
  Ext half_beta = emul_base(*beta, one_half);
  Ext v = fri_fold_pair(cur[2 * i], cur[2 * i + 1], half_beta, one_half, ginv, (uint32_t)i, logh);
  if (add) v = eadd(v, add[i]);
  out[i] = v;
}
void fri_fold_dev(lsk_ctx* ctx, const Ext* cur, unsigned logh, const Ext* beta_dev, const Ext* add, Ext* out) {
  size_t half = (size_t)1 << logh;
  uint32_t one_half = finv(to_mont(2));
  ProfScope ps(ctx, PT_FRI_FOLD, 1, (double)half * (48.0 + (add ? 16.0 : 0.0)));
  fri_fold_kernel<<<div_up(half, 256), 256, 0, ctx->stream>>>(cur, logh, finv(two_adic_generator(logh + 1)), one_half,
                                                              beta_dev, add, out);
  LSK_CUDA(cudaGetLastError());
}

struct GrindState {
  uint32_t s[16];  // Montgomery, pending inputs already written; slot `pos` takes the witness
  uint32_t pos, mask;
};
__global__ void __launch_bounds__(128) grind_kernel(GrindState gs, uint32_t base, uint32_t* __restrict__ best) {
  uint32_t w = base + blockIdx.x * 128 + threadIdx.x;
  if (w >= P) return;
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = gs.s[i];
  uint32_t wm = to_mont(w);
#pragma unroll
  for (int i = 0; i < 8; i++)
    if ((uint32_t)i == gs.pos) s[i] = wm;
  p2_permute<16>(s);
  if ((from_mont(s[7]) & gs.mask) == 0) atomicMin(best, w);
}
uint32_t grind_dev(lsk_ctx* ctx, const uint32_t state[16], const uint32_t* pending, unsigned n_pending, unsigned bits) {
  if (n_pending >= 8 || bits > 30) fail(LSK_ERR_ARG, "bad grind arguments");
  GrindState gs;
  for (int i = 0; i < 16; i++) gs.s[i] = to_mont(state[i]);
  for (unsigned i = 0; i < n_pending; i++) gs.s[i] = to_mont(pending[i]);
  gs.pos = n_pending;
  gs.mask = (1u << bits) - 1;
  DevBuf<uint32_t> best(ctx, 1);
  // the expected smallest witness is ~2^bits: start with a batch of that order, then widen
  uint32_t batch = std::max(1u << 14, std::min(1u << 20, 4u << bits));
  for (uint64_t base = 0; base < P; base += batch, batch = std::min(1u << 22, batch * 4)) {
    uint32_t h = 0xffffffffu;
    best.upload(&h, 1);
    ProfScope ps(ctx, PT_GRIND, 1, 0.0);
    grind_kernel<<<batch / 128, 128, 0, ctx->stream>>>(gs, (uint32_t)base, best.p);
    LSK_CUDA(cudaGetLastError());
    best.download(&h, 1);
    LSK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h != 0xffffffffu) return h;
  }
  fail(LSK_ERR_EXEC, "proof-of-work grinding found no witness");
}

// ---- query gathers -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) query_assemble_kernel(const QuerySeg* __restrict__ segs, uint32_t nsegs,
                                                              const uint32_t* __restrict__ indices, uint32_t wpq,
                                                              uint32_t* __restrict__ out) {
  const uint32_t index = indices[blockIdx.x];
  uint32_t* o = out + (size_t)blockIdx.x * wpq;
  const unsigned lane = threadIdx.x & 31;
  for (uint32_t s = threadIdx.x >> 5; s < nsegs; s += 8) {  // one warp per segment
    const QuerySeg g = segs[s];
    const uint32_t node = (index >> g.shift) ^ g.xor1;
    if (g.own_rank != QSEG_ALL && (g.own_rank == QSEG_NONE || (node >> g.own_shift) != (uint32_t)g.own_rank)) {
      for (uint32_t k = lane; k < g.count; k += 32) o[g.dst + k] = 0;   // another rank of the group supplies these
    } else if (!g.p) {
      for (uint32_t k = lane; k < g.count; k += 32) o[g.dst + k] = g.lit;
    } else {
      const size_t first = (size_t)node * g.mul;
      for (uint32_t k = lane; k < g.count; k += 32) o[g.dst + k] = from_mont(g.p[first + (size_t)k * g.stride]);
    }
  }
}
void query_assemble_dev(lsk_ctx* ctx, const QuerySeg* segs_dev, uint32_t nsegs, const uint32_t* indices_dev,
                        uint32_t nqueries, uint32_t words_per_query, uint32_t* out) {
  if (!nqueries || !nsegs) return;
  ProfScope ps(ctx, PT_GATHER, 1, 4.0 * nqueries * words_per_query);
  query_assemble_kernel<<<nqueries, 256, 0, ctx->stream>>>(segs_dev, nsegs, indices_dev, words_per_query, out);
  LSK_CUDA(cudaGetLastError());
}

void init_stark_kernels() { LSK_CUDA(p2_upload_constants()); }

}  // namespace lsk
