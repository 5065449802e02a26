// Quotient evaluation (A6 + A10): the kernel that evaluates a chip's constraints on the LDE rows.
//
// This header is compiled twice: by nvcc into liblurkstark.so (the INTERPRETER: quotient_kernel<MAXREGS, ROWS>
// walks the chip's constraint program instruction by instruction) and, at run time, by NVRTC for the chips whose
// programs are long enough to pay for it (jit.cu: the program is emitted as straight-line CUDA -- 128-statement
// __noinline__ chunks over a register file in local memory, constants and alpha-power indices as literals -- and
// included as "lsk_program.cuh"; measured on the 655-wide hash4 chip, 2^21 rows: 24.1 ms interpreted,
// 10.7 ms compiled).  Row set-up (selectors, vanishing polynomial) and the LogUp constraints + division are
// shared.  Only "field.cuh" may be included here: NVRTC has no system headers.
#pragma once
#include "field.cuh"

namespace lsk {

// device opcodes of the constraint program: 0..26 = 9 * {add, sub, mul} + 3 * kind(a) + kind(b) with raw
// indices (kind 0 register, 1 main-trace cell of the row, 2 constant); the rest keep tagged operands
enum : uint32_t { QOP_NEG = 27, QOP_ASSERT_REG = 28, QOP_ASSERT = 29, QOP_ADD = 30, QOP_SUB = 31, QOP_MUL = 32 };

struct QuotientArgs {
  const uint32_t* main; size_t mstride;   // LDEs, 2N rows
  const uint32_t* prep; size_t pstride;
  const uint32_t* perm; size_t permstride;
  const uint32_t* pv;       // public values, Montgomery, device
  const Ext* apow;          // apow[k] = alpha^k; constraint i (AIR first, then LogUp) is weighted alpha^(n_total-1-i)
  uint32_t n_total;
  const Ext* beta_pows;     // beta^(v+1)
  const Ext* cum_sum;
  Ext chal_alpha;
  unsigned logn;
  uint32_t* out; size_t ostride;  // 4 columns of 2N rows
  size_t row_begin, row_end;      // set by quotient_dev: the LDE rows this launch evaluates
};
struct QuotientConsts {
  uint32_t g, w2n, wn_inv, zh[2], inv_zh[2];
};
struct PermRowsArgs {   // LogUp permutation-trace rows of one chip (stark_kernels.cu::perm_rows_kernel and its compiled form)
  const uint32_t* main; size_t mstride;
  const uint32_t* prep; size_t pstride;
  unsigned logn;
  Ext alpha;
  const Ext* beta_pows;
  uint32_t* perm; size_t perm_stride;
  size_t row_begin, row_end;
  uint32_t scatter;       // row sums to natural order (for the scan) or left in storage order
  Ext* rowsum_out;
};

// ---- device helpers shared with the LogUp trace kernel ---------------------------------------------------------
struct Cells {  // local-row view of a chip's main + preprocessed matrices
  const uint32_t* main; size_t mstride;
  const uint32_t* prep; size_t pstride;
  size_t r;
  LSK_D uint32_t get(uint32_t is_prep, uint32_t col) const {
    return is_prep ? prep[(size_t)col * pstride + r] : main[(size_t)col * mstride + r];
  }
};
LSK_D uint32_t eval_affine(const uint32_t* __restrict__& s, const Cells& c) {
  uint32_t n = s[0], acc = s[1];
  s += 2;
  for (uint32_t i = 0; i < n; i++, s += 3) acc = fadd(acc, fmul(c.get(s[0], s[1]), s[2]));
  return acc;
}
// one interaction: rlc = alpha + kind + sum_v beta^(v+1) value_v ; m = +-multiplicity
LSK_D void eval_interaction(const uint32_t* __restrict__& s, const Cells& c, const Ext& alpha,
                            const Ext* __restrict__ beta_pows, Ext& rlc, uint32_t& m) {
  uint32_t kind = s[0], sign = s[1], nv = s[2];
  s += 3;
  m = eval_affine(s, c);
  if (sign) m = fneg(m);
  ExtAcc sum = ext_acc_zero();
  for (uint32_t v = 0; v < nv; v++) {
    uint32_t val = eval_affine(s, c);
    sum.add_reduce(beta_pows[v], val);
  }
  rlc = eadd(eadd_base(alpha, kind), sum.value());
}
// a * x + b * y for extension a, b and base x, y (the numerator of a LogUp batch of two)
LSK_D Ext emul_base2(const Ext& a, uint32_t x, const Ext& b, uint32_t y) {
  ExtAcc s = ext_acc_zero();
  s.add(a, x);
  s.add_reduce(b, y);
  return s.value();
}
LSK_D Ext ldg_ext(const Ext* __restrict__ p) {   // 16-byte aligned (arrays of Ext in device buffers)
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  return Ext{{v.x, v.y, v.z, v.w}};
}
LSK_D Ext load_ext_cols(const uint32_t* __restrict__ m, size_t stride, uint32_t first_col, size_t r) {
  Ext e;
#pragma unroll
  for (int k = 0; k < 4; k++) e.c[k] = m[(size_t)(first_col + k) * stride + r];
  return e;
}


// ---- one LDE row: position, selectors ---------------------------------------------------------------------------
struct QuotientRow {
  size_t r, rn;        // storage position (bit-reversed) of the row and of the next row of the trace domain
  uint32_t ii;         // natural index on the quotient domain
  uint32_t sel[3];     // is_first, is_last, is_transition
};
LSK_D void quotient_row_setup(const QuotientArgs& a, const QuotientConsts& qc, size_t row, QuotientRow& q) {
  const unsigned l2 = a.logn + 1;
  const size_t n2 = (size_t)1 << l2;
  q.r = row;
  q.ii = bitrev32((uint32_t)row, l2);
  q.rn = bitrev32((uint32_t)((q.ii + 2) & (n2 - 1)), l2);
  uint32_t x = fmul(qc.g, fpow(qc.w2n, q.ii));
  uint32_t zh = qc.zh[q.ii & 1];
  uint32_t xm1 = fsub(x, R1), xml = fsub(x, qc.wn_inv);
  uint32_t inv = finv(fmul(xm1, xml));
  q.sel[0] = fmul(zh, fmul(inv, xml));  // is_first
  q.sel[1] = fmul(zh, fmul(inv, xm1));  // is_last
  q.sel[2] = xml;                       // is_transition
}
LSK_D void quotient_row_tail(const QuotientArgs& a, const QuotientConsts& qc, const QuotientRow& q, Ext ac, uint32_t cj,
                             const Ext& sum_local, const Ext& sum_next, uint32_t pw, bool live);
// LogUp constraints (sphinx eval_permutation_constraints) on top of the folded AIR constraints `ac` (the first
// `ci` alpha powers are spent), division by the vanishing polynomial, store
LSK_D void quotient_row_finish(const QuotientArgs& a, const QuotientConsts& qc, const QuotientRow& q, Ext ac, uint32_t ci,
                               const uint32_t* __restrict__ inter, uint32_t nint, uint32_t pw, bool live) {
  Cells cells{a.main, a.mstride, a.prep, a.pstride, q.r};
  const uint32_t* s = inter;
  uint32_t cj = ci;
  Ext sum_local = ext_zero(), sum_next = ext_zero();
  for (uint32_t b = 0; 2 * b < nint; b++) {
    Ext r0, lhs, rhs;
    uint32_t m0;
    eval_interaction(s, cells, a.chal_alpha, a.beta_pows, r0, m0);
    Ext pl = load_ext_cols(a.perm, a.permstride, 4 * b, q.r);
    if (2 * b + 1 < nint) {
      Ext r1;
      uint32_t m1;
      eval_interaction(s, cells, a.chal_alpha, a.beta_pows, r1, m1);
      lhs = emul(emul(pl, r0), r1);
      rhs = emul_base2(r1, m0, r0, m1);
    } else {
      lhs = emul(pl, r0);
      rhs = ext_from_base(m0);
    }
    ac = eadd(ac, emul(a.apow[a.n_total - 1 - (cj++)], esub(lhs, rhs)));
    sum_local = eadd(sum_local, pl);
    sum_next = eadd(sum_next, load_ext_cols(a.perm, a.permstride, 4 * b, q.rn));
  }
  quotient_row_tail(a, qc, q, ac, cj, sum_local, sum_next, pw, live);
}
LSK_D void quotient_row_tail(const QuotientArgs& a, const QuotientConsts& qc, const QuotientRow& q, Ext ac, uint32_t cj,
                             const Ext& sum_local, const Ext& sum_next, uint32_t pw, bool live) {
  Ext phi_local = load_ext_cols(a.perm, a.permstride, 4 * (pw - 1), q.r);
  Ext phi_next = load_ext_cols(a.perm, a.permstride, 4 * (pw - 1), q.rn);
  ac = eadd(ac, emul(a.apow[a.n_total - 1 - (cj++)], emul_base(esub(esub(phi_next, phi_local), sum_next), q.sel[2])));
  ac = eadd(ac, emul(a.apow[a.n_total - 1 - (cj++)], emul_base(esub(phi_local, sum_local), q.sel[0])));
  ac = eadd(ac, emul(a.apow[a.n_total - 1 - (cj++)], emul_base(esub(phi_local, *a.cum_sum), q.sel[1])));
  Ext quot = emul_base(ac, qc.inv_zh[q.ii & 1]);
  if (live) {
#pragma unroll
    for (int e = 0; e < 4; e++) a.out[(size_t)e * a.ostride + q.r] = quot.c[e];
  }
}

#if defined(LSK_JIT_PROGRAM)
// ---- compiled form: jit.cu defines LSK_JIT_REGS and LSK_JIT_ASSERTS and supplies "lsk_program.cuh" with the
// chunk functions lsk_chunk_<k>(R, acc, a, q) plus lsk_program(R, acc, a, q) that calls them in order
#if LSK_JIT_SMEM   // register file in shared memory, [register][thread]: R points at this thread's column
#define LSK_R(i) R[(i) * 256]
#else
#define LSK_R(i) R[i]
#endif
// trace cells, public values and challenge powers are read-only for the whole launch: __ldg makes the loads
// LDG.CONSTANT -- through the argument structs they were GENERIC loads (LD.E), which the compiler must keep in order with
// every store to the register file / the stack and cannot hoist (ncu: long_scoreboard was the top stall)
#define LSK_ML(i) __ldg(a.main + ((size_t)(i) * a.mstride + q.r))
#define LSK_MN(i) __ldg(a.main + ((size_t)(i) * a.mstride + q.rn))
#define LSK_PL(i) __ldg(a.prep + ((size_t)(i) * a.pstride + q.r))
#define LSK_PN(i) __ldg(a.prep + ((size_t)(i) * a.pstride + q.rn))
#define LSK_PV(i) __ldg(a.pv + (i))
#define LSK_SEL(i) q.sel[i]
// constraints of one chunk: sum_i alpha^(n_total-1-ci) * v_i as four 64-bit sums of raw products, each kept below
// P * 2^32 (what mont_reduce accepts) by one conditional subtraction per term; reduced once when the chunk ends
#define LSK_ASSERT_BEGIN uint64_t lsk_s0 = 0, lsk_s1 = 0, lsk_s2 = 0, lsk_s3 = 0
// LSK_ASSERT_A adds a term without reducing (sum < P * 2^32 + P^2), LSK_ASSERT_B adds the next one and reduces
// (P * 2^32 + 2 P^2 < 2^64): jit.cu alternates them; LSK_ASSERT_END reduces whatever is pending
#define LSK_ASSERT_A(ci, v)                                \
  {                                                        \
    const Ext lsk_ap = ldg_ext(a.apow + (a.n_total - 1 - (ci))); \
    const uint32_t lsk_x = (v);                            \
    lsk_s0 += (uint64_t)lsk_ap.c[0] * lsk_x;               \
    lsk_s1 += (uint64_t)lsk_ap.c[1] * lsk_x;               \
    lsk_s2 += (uint64_t)lsk_ap.c[2] * lsk_x;               \
    lsk_s3 += (uint64_t)lsk_ap.c[3] * lsk_x;               \
  }
#define LSK_ASSERT_B(ci, v)                                             \
  {                                                                     \
    const Ext lsk_ap = ldg_ext(a.apow + (a.n_total - 1 - (ci)));        \
    const uint32_t lsk_x = (v);                                         \
    lsk_s0 = acc_reduce64(lsk_s0 + (uint64_t)lsk_ap.c[0] * lsk_x);      \
    lsk_s1 = acc_reduce64(lsk_s1 + (uint64_t)lsk_ap.c[1] * lsk_x);      \
    lsk_s2 = acc_reduce64(lsk_s2 + (uint64_t)lsk_ap.c[2] * lsk_x);      \
    lsk_s3 = acc_reduce64(lsk_s3 + (uint64_t)lsk_ap.c[3] * lsk_x);      \
  }
#define LSK_ASSERT_END                                                                                     \
  acc = eadd(acc, Ext{{mont_reduce(acc_reduce64(lsk_s0)), mont_reduce(acc_reduce64(lsk_s1)),               \
                       mont_reduce(acc_reduce64(lsk_s2)), mont_reduce(acc_reduce64(lsk_s3))}})
#include "lsk_program.cuh"
// "lsk_interactions.cuh" (jit.cu::interactions_source): lsk_interactions<V>(cells, alpha, beta_pows, v) evaluates the
// chip's interactions as straight-line code and hands every LogUp batch to v.pair(b, r0, m0, r1, m1) / v.single(b, r0, m0)
#define LSK_CM(i) __ldg(c.main + ((size_t)(i) * c.mstride + c.r))
#define LSK_CP(i) __ldg(c.prep + ((size_t)(i) * c.pstride + c.r))
#include "lsk_interactions.cuh"
// the batch bodies are real calls: inlined into every one of a chip's 10-30 batches they multiply NVVM's work (the
// 41 chips of the Lurk machine took 19 minutes to compile that way, 40 s like this) for no measurable gain
__device__ __noinline__ Ext lsk_quot_pair(Ext pl, Ext r0, uint32_t m0, Ext r1, uint32_t m1) {
  return esub(emul(emul(pl, r0), r1), emul_base2(r1, m0, r0, m1));
}
__device__ __noinline__ Ext lsk_fold(Ext ac, Ext ap, Ext x) { return eadd(ac, emul(ap, x)); }
struct LskQuotientBatches {   // quotient_row_finish's loop body
  const QuotientArgs& a;
  const QuotientRow& q;
  Ext ac, sum_local, sum_next;
  uint32_t cj;
  LSK_D void fold(uint32_t b, const Ext& pl, const Ext& diff) {
    ac = lsk_fold(ac, ldg_ext(a.apow + (a.n_total - 1 - (cj++))), diff);
    sum_local = eadd(sum_local, pl);
    sum_next = eadd(sum_next, load_ext_cols(a.perm, a.permstride, 4 * b, q.rn));
  }
  LSK_D void pair(uint32_t b, const Ext& r0, uint32_t m0, const Ext& r1, uint32_t m1) {
    const Ext pl = load_ext_cols(a.perm, a.permstride, 4 * b, q.r);
    fold(b, pl, lsk_quot_pair(pl, r0, m0, r1, m1));
  }
  LSK_D void single(uint32_t b, const Ext& r0, uint32_t m0) {
    const Ext pl = load_ext_cols(a.perm, a.permstride, 4 * b, q.r);
    fold(b, pl, esub(emul(pl, r0), ext_from_base(m0)));
  }
};
// perm_rows_kernel's loop body, with the inversions of up to four consecutive batches shared (Montgomery's trick: one
// einv -- ~85 base products, 62 of them a dependent chain -- and 3 (n - 1) extension products instead of n einv).  The
// values are the same field elements; only a denominator that is exactly zero (probability 2^-124 per batch) would
// behave differently from the one-inversion-per-batch form, which maps 1/0 to 0.
constexpr uint32_t LSK_PERM_GROUP = 4;
__device__ __noinline__ Ext lsk_perm_flush(const PermRowsArgs& a, size_t r, uint32_t b0, uint32_t n, const Ext* __restrict__ den,
                                           const Ext* __restrict__ num) {
  Ext pre[LSK_PERM_GROUP];
  pre[0] = den[0];
  for (uint32_t i = 1; i < n; i++) pre[i] = emul(pre[i - 1], den[i]);
  Ext inv = einv(pre[n - 1]), rowsum = ext_zero();
  for (uint32_t i = n; i-- > 0;) {
    const Ext v = emul(num[i], i ? emul(inv, pre[i - 1]) : inv);
    if (i) inv = emul(inv, den[i]);
#pragma unroll
    for (int k = 0; k < 4; k++) a.perm[(size_t)(4 * (b0 + i) + k) * a.perm_stride + r] = v.c[k];
    rowsum = eadd(rowsum, v);
  }
  return rowsum;
}
__device__ __noinline__ void lsk_perm_zero(const PermRowsArgs& a, size_t r, uint32_t b0, uint32_t n) {
  for (uint32_t i = 0; i < 4 * n; i++) a.perm[(size_t)(4 * b0 + i) * a.perm_stride + r] = 0u;
}
struct LskPermBatches {
  const PermRowsArgs& a;
  size_t r;
  Ext rowsum;
  Ext den[LSK_PERM_GROUP], num[LSK_PERM_GROUP];
  uint32_t any;   // OR of the multiplicities of the open group: a group of padding rows costs no inversion at all
  LSK_D void close(uint32_t b) {
    const uint32_t slot = b % LSK_PERM_GROUP;
    if (slot + 1 != LSK_PERM_GROUP && b + 1 != (LSK_JIT_NINT + 1) / 2) return;
    if (any) rowsum = eadd(rowsum, lsk_perm_flush(a, r, b - slot, slot + 1, den, num));
    else lsk_perm_zero(a, r, b - slot, slot + 1);
    any = 0;
  }
  LSK_D void pair(uint32_t b, const Ext& r0, uint32_t m0, const Ext& r1, uint32_t m1) {
    den[b % LSK_PERM_GROUP] = emul(r0, r1);
    num[b % LSK_PERM_GROUP] = emul_base2(r1, m0, r0, m1);
    any |= m0 | m1;
    close(b);
  }
  LSK_D void single(uint32_t b, const Ext& r0, uint32_t m0) {
    den[b % LSK_PERM_GROUP] = r0;
    num[b % LSK_PERM_GROUP] = ext_from_base(m0);
    any |= m0;
    close(b);
  }
};
#ifndef LSK_JIT_PTHREADS
#define LSK_JIT_PTHREADS 256
#endif
extern "C" __global__ void __launch_bounds__(LSK_JIT_PTHREADS) lsk_perm_rows_jit(PermRowsArgs a) {
  const size_t r = a.row_begin + (size_t)blockIdx.x * LSK_JIT_PTHREADS + threadIdx.x;
  if (r >= a.row_end) return;
  const Cells cells{a.main, a.mstride, a.prep, a.pstride, r};
  LskPermBatches v{a, r, ext_zero(), {}, {}, 0u};
  lsk_interactions(cells, a.alpha, a.beta_pows, v);
  a.rowsum_out[a.scatter ? bitrev32((uint32_t)r, a.logn) : r] = v.rowsum;
}
// threads per CTA of the compiled quotient (jit.cu defines it; LSK_JIT_QTHREADS in the environment at compile time).  The
// warps of a CTA walk the straight-line program in step (a barrier per chunk), so a larger CTA means fewer instruction
// streams per SM.
#ifndef LSK_JIT_QTHREADS
#define LSK_JIT_QTHREADS 768
#endif
extern "C" __global__ void __launch_bounds__(LSK_JIT_QTHREADS) lsk_quotient_jit(QuotientArgs a, QuotientConsts qc,
                                                                   const uint32_t* __restrict__ inter, uint32_t nint,
                                                                   uint32_t pw) {
  // no early exit: the chunks are separated by __syncthreads() so that the warps of a CTA walk the (megabyte-sized,
  // straight-line) program together and share what the instruction cache fetched -- ncu showed "no instruction"
  // as the second stall reason without it.  Threads past the end recompute a valid row and do not store.
  size_t row = a.row_begin + (size_t)blockIdx.x * LSK_JIT_QTHREADS + threadIdx.x;
  const bool live = row < a.row_end;
  if (!live) row = a.row_begin;
  QuotientRow q;
  quotient_row_setup(a, qc, row, q);
#if LSK_JIT_SMEM
  extern __shared__ uint32_t lsk_regs[];
  uint32_t* R = lsk_regs + threadIdx.x;
#else
  uint32_t R[LSK_JIT_REGS];
#endif
  LSK_R(0) = 0;
  Ext acc = ext_zero();
  lsk_program(R, acc, a, q);
  const Cells cells{a.main, a.mstride, a.prep, a.pstride, q.r};
  LskQuotientBatches v{a, q, acc, ext_zero(), ext_zero(), LSK_JIT_ASSERTS};
  lsk_interactions(cells, a.chal_alpha, a.beta_pows, v);
  quotient_row_tail(a, qc, q, v.ac, v.cj, v.sum_local, v.sum_next, pw, live);
}
#else
// ---- interpreted form.  ROWS rows per thread (rows t, t + 256, ... of the CTA's slab, so every access stays
// coalesced): the long constraint programs of the hasher chips are bound by instruction fetch / decode and by the
// latency of the register file in local memory, both of which a second independent row amortises.
template <int MAXREGS, int ROWS>
__global__ void __launch_bounds__(256) quotient_kernel(QuotientArgs a, QuotientConsts qc,
                                                        const uint4* __restrict__ code, uint32_t ninstr,
                                                        const uint32_t* __restrict__ consts,
                                                        const uint32_t* __restrict__ inter, uint32_t nint, uint32_t pw) {
  const size_t base = a.row_begin + (size_t)blockIdx.x * (256 * ROWS) + threadIdx.x;
  if (base >= a.row_end) return;
  QuotientRow q[ROWS];
  bool live[ROWS];
  uint32_t regs[ROWS][MAXREGS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    size_t row = base + (size_t)k * 256;
    live[k] = row < a.row_end;
    if (!live[k]) row = base;  // recompute a valid row, never stored
    quotient_row_setup(a, qc, row, q[k]);
    regs[k][0] = 0;
  }
  auto operand = [&](int k, uint32_t w) -> uint32_t {
    uint32_t idx = w & 0x0fffffffu;
    switch (w >> 28) {
      case 0: return regs[k][idx];
      case 1: return consts[idx];
      case 2: return a.main[(size_t)idx * a.mstride + q[k].r];
      case 3: return a.main[(size_t)idx * a.mstride + q[k].rn];
      case 4: return a.prep[(size_t)idx * a.pstride + q[k].r];
      case 5: return a.prep[(size_t)idx * a.pstride + q[k].rn];
      case 6: return a.pv[idx];
      default: return q[k].sel[idx];
    }
  };
  Ext acc[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) acc[k] = ext_zero();
  uint32_t ci = 0;
  // Device form of the constraint program (prover.cu::compile_chip): the operand kinds of the common
  // instructions are folded into the opcode -- op = 9 * {add, sub, mul} + 3 * kind(a) + kind(b) with kind
  // 0 register, 1 main-trace cell of this row, 2 constant -- so that one indexed branch replaces the
  // opcode switch plus two operand-tag switches; everything else takes the generic tagged path.
  auto R = [&](int k, uint32_t i) { return regs[k][i]; };
  auto M = [&](int k, uint32_t i) { return a.main[(size_t)i * a.mstride + q[k].r]; };
  auto C = [&](int, uint32_t i) { return __ldg(consts + i); };
#define LSK_QROWS(EXPR)              \
  _Pragma("unroll") for (int k = 0; k < ROWS; k++) { EXPR; } \
  break;
#define LSK_Q3(BASE, FN, KA, LA)                                                              \
  case BASE + 3 * KA + 0: LSK_QROWS(regs[k][ins.y] = FN(LA(k, ins.z), R(k, ins.w)))           \
  case BASE + 3 * KA + 1: LSK_QROWS(regs[k][ins.y] = FN(LA(k, ins.z), M(k, ins.w)))           \
  case BASE + 3 * KA + 2: LSK_QROWS(regs[k][ins.y] = FN(LA(k, ins.z), C(k, ins.w)))
#define LSK_Q9(BASE, FN) LSK_Q3(BASE, FN, 0, R) LSK_Q3(BASE, FN, 1, M) LSK_Q3(BASE, FN, 2, C)
  for (uint32_t pc = 0; pc < ninstr; pc++) {
    const uint4 ins = __ldg(code + pc);
    switch (ins.x) {
      LSK_Q9(0, fadd)
      LSK_Q9(9, fsub)
      LSK_Q9(18, fmul)
      case QOP_NEG: LSK_QROWS(regs[k][ins.y] = fneg(operand(k, ins.z)))
      case QOP_ASSERT_REG: {
        const Ext ap = a.apow[a.n_total - 1 - (ci++)];
        LSK_QROWS(acc[k] = eadd(acc[k], emul_base(ap, regs[k][ins.z])))
      }
      case QOP_ASSERT: {
        const Ext ap = a.apow[a.n_total - 1 - (ci++)];
        LSK_QROWS(acc[k] = eadd(acc[k], emul_base(ap, operand(k, ins.z))))
      }
      case QOP_ADD: LSK_QROWS(regs[k][ins.y] = fadd(operand(k, ins.z), operand(k, ins.w)))
      case QOP_SUB: LSK_QROWS(regs[k][ins.y] = fsub(operand(k, ins.z), operand(k, ins.w)))
      default: LSK_QROWS(regs[k][ins.y] = fmul(operand(k, ins.z), operand(k, ins.w)))
    }
  }
#undef LSK_Q9
#undef LSK_Q3
#undef LSK_QROWS
#pragma unroll 1
  for (int k = 0; k < ROWS; k++) quotient_row_finish(a, qc, q[k], acc[k], ci, inter, nint, pw, live[k]);
}
#endif  // LSK_JIT_PROGRAM

}  // namespace lsk
