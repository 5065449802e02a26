// BabyBear (p = 2^31 - 2^27 + 1) and its quartic extension F[x]/(x^4 - 11) for sm_100a.
//
// Device representation: Montgomery form (x * 2^32 mod p), always reduced to [0, p).  Field
// elements cross the C ABI in canonical form (the reference's from_canonical_u32 /
// as_canonical_u32 view, e.g. src/lair/mod.rs:33-46); conversion happens in the first/last kernel
// that touches boundary data.  The arithmetic itself lives in the third-party p3-baby-bear /
// p3-field crates in the reference, so this is a from-scratch implementation of
// the published field, not a translation.
//
// A Montgomery product costs 1 IMAD.WIDE + 1 IMAD (lo) + 1 IMAD.HI + sub/compare/select on the
// fma/alu pipes; there are no tensor-core shaped operations anywhere on this path.
#pragma once
#if defined(__CUDACC_RTC__)   // NVRTC (jit.cu) has no system headers
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <stdint.h>
#endif

#if defined(__CUDACC__)
#define LSK_HD __host__ __device__ __forceinline__
#define LSK_D __device__ __forceinline__
#else
#define LSK_HD inline
#define LSK_D inline
#endif

namespace lsk {

constexpr uint32_t P = 0x78000001u;
constexpr uint32_t PINV = 0x88000001u;      // P * PINV == 1 (mod 2^32)
constexpr uint32_t R1 = 268435454u;         // 2^32 mod P  (Montgomery one)
constexpr uint32_t R2 = 1172168163u;        // 2^64 mod P
constexpr uint32_t GEN = 31u;               // multiplicative generator (coset shift)
constexpr uint32_t TWO_ADIC_GEN_27 = 0x1a427a41u;  // 31^((p-1)/2^27)

// Pipe placement of the PLAIN additions (tools/perm_variants.cu).  ptxas emits a + b either as IADD3 (ALU pipe) or as
// IMAD.IADD (multiply pipe) by its own balancing; on B200 the multiply pipe is the bound of every hashing kernel, so
// LSK_ALU_ADD forces chosen additions onto the ALU pipe by writing them as min(a + b, g_lsk_ones): an add + min pair with
// an operand ptxas cannot fold (a __constant__ it does not know the value of), which it fuses into one VIADDMNMX.
// bit 0: fadd, bit 1: fsub, bit 2: the high-word subtraction of mont_reduce.
#if defined(__CUDACC__) && !defined(__CUDACC_RTC__) && defined(LSK_ALU_ADD)
static __constant__ uint32_t g_lsk_ones = 0xffffffffu;
#define LSK_ALU_PLUS(a, b) min((uint32_t)((a) + (b)), g_lsk_ones)
#define LSK_ALU_MINUS(a, b) min((uint32_t)((a) - (b)), g_lsk_ones)
#endif
LSK_HD uint32_t mont_reduce(uint64_t x) {
  uint32_t t = (uint32_t)x * PINV;
#if defined(__CUDA_ARCH__)
  uint32_t hu = __umulhi(t, P);
#else
  uint32_t hu = (uint32_t)(((uint64_t)t * P) >> 32);
#endif
  uint32_t hx = (uint32_t)(x >> 32);
#if defined(__CUDA_ARCH__) && defined(LSK_ALU_PLUS) && (LSK_ALU_ADD & 4)
  uint32_t r = LSK_ALU_MINUS(hx, hu);
#else
  uint32_t r = hx - hu;
#endif
#if defined(__CUDA_ARCH__)
  return min(r, r + P);  // r wrapped iff hx < hu; one fused add+min (VIADDMNMX) on the ALU pipe
#else
  return hx < hu ? r + P : r;
#endif
}
// Integer multiplies issue on the half-rate "fmaheavy" pipe only (measured on B200: IMAD.lo 64,
// IMAD.HI 32, IMAD.WIDE ~26 lanes/clk/SM), so a Montgomery product costs ~11 pipe cycles per warp;
// corrections use the min() forms (one fused VIADDMNMX).  Additions are not free either: ptxas emits
// about half of the plain adds as IMAD.IADD on that same pipe and the rest on the ALU pipe, and
// forcing all of them onto the ALU pipe (three-input IADD3 with an opaque zero) made the hashing
// kernels 3 % SLOWER on B200 -- the ALU pipe is then the bound.  The two integer pipes together,
// ~19 k pipe-cycles per width-16 permutation per warp, are the roofline of every hashing kernel.
LSK_HD uint32_t fmul(uint32_t a, uint32_t b) { return mont_reduce((uint64_t)a * b); }
LSK_HD uint32_t fadd(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && defined(LSK_ALU_PLUS) && (LSK_ALU_ADD & 1)
  uint32_t s = LSK_ALU_PLUS(a, b);
#else
  uint32_t s = a + b;
#endif
#if defined(__CUDA_ARCH__)
  return min(s, s - P);
#else
  return s >= P ? s - P : s;
#endif
}
LSK_HD uint32_t fsub(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && defined(LSK_ALU_PLUS) && (LSK_ALU_ADD & 2)
  uint32_t s = LSK_ALU_MINUS(a, b);
#else
  uint32_t s = a - b;
#endif
#if defined(__CUDA_ARCH__)
  return min(s, s + P);
#else
  return a < b ? s + P : s;
#endif
}
// Product by a constant c known in advance (Shoup): c_std is the plain residue, c_shoup =
// floor(c_std * 2^32 / P).  a may be in either representation (the result keeps a's); costs
// IMAD.HI + 2 IMAD.lo = 8 pipe cycles instead of 11.
LSK_HD uint32_t shoup_precompute(uint32_t c_std) { return (uint32_t)(((uint64_t)c_std << 32) / P); }
LSK_HD uint32_t fmul_shoup(uint32_t a, uint32_t c_std, uint32_t c_shoup) {
#if defined(__CUDA_ARCH__)
  uint32_t q = __umulhi(a, c_shoup);
  uint32_t r = a * c_std - q * P;
  return min(r, r - P);
#else
  uint32_t q = (uint32_t)(((uint64_t)a * c_shoup) >> 32);
  uint32_t r = a * c_std - q * P;
  return r >= P ? r - P : r;
#endif
}
LSK_HD uint32_t fneg(uint32_t a) { return a ? P - a : 0u; }
LSK_HD uint32_t fdbl(uint32_t a) { return fadd(a, a); }
LSK_HD uint32_t to_mont(uint32_t canonical) { return fmul(canonical, R2); }
LSK_HD uint32_t from_mont(uint32_t m) { return mont_reduce((uint64_t)m); }
LSK_HD uint32_t fsqr(uint32_t a) { return fmul(a, a); }

// a^(p-2): p - 2 = 7 * 2^28 + (2^27 - 1).  ~62 products.  Returns 0 for 0.
LSK_HD uint32_t finv(uint32_t a) {
  uint32_t x1 = a;
  uint32_t x2 = fmul(fsqr(x1), x1);
  uint32_t x3 = fmul(fsqr(x2), x1);
  uint32_t t = x3;
  for (int i = 0; i < 3; i++) t = fsqr(t);
  uint32_t x6 = fmul(t, x3);
  t = x6;
  for (int i = 0; i < 6; i++) t = fsqr(t);
  uint32_t x12 = fmul(t, x6);
  t = x12;
  for (int i = 0; i < 12; i++) t = fsqr(t);
  uint32_t x24 = fmul(t, x12);
  t = x24;
  for (int i = 0; i < 3; i++) t = fsqr(t);
  uint32_t x27 = fmul(t, x3);  // a^(2^27 - 1)
  t = x3;                      // a^7
  for (int i = 0; i < 28; i++) t = fsqr(t);
  return fmul(t, x27);
}
LSK_HD uint32_t fpow(uint32_t b, uint64_t e) {
  uint32_t r = R1;
  while (e) {
    if (e & 1) r = fmul(r, b);
    b = fsqr(b);
    e >>= 1;
  }
  return r;
}
// generator of the subgroup of order 2^bits, Montgomery form
LSK_HD uint32_t two_adic_generator(unsigned bits) {
  uint32_t g = to_mont(TWO_ADIC_GEN_27);
  for (unsigned i = bits; i < 27; i++) g = fsqr(g);
  return g;
}
LSK_HD uint32_t fmul11(uint32_t a) { return mont_reduce((uint64_t)a * 939524073u); }   // 11 in Montgomery form (M11)

// ---- quartic extension ------------------------------------------------------------------------
struct __align__(16) Ext {
  uint32_t c[4];
};
LSK_HD Ext ext_zero() { return Ext{{0, 0, 0, 0}}; }
LSK_HD Ext ext_one() { return Ext{{R1, 0, 0, 0}}; }
LSK_HD Ext ext_from_base(uint32_t a) { return Ext{{a, 0, 0, 0}}; }
LSK_HD bool ext_eq(const Ext& a, const Ext& b) {
  return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3];
}
LSK_HD Ext eadd(const Ext& a, const Ext& b) {
  return Ext{{fadd(a.c[0], b.c[0]), fadd(a.c[1], b.c[1]), fadd(a.c[2], b.c[2]), fadd(a.c[3], b.c[3])}};
}
LSK_HD Ext esub(const Ext& a, const Ext& b) {
  return Ext{{fsub(a.c[0], b.c[0]), fsub(a.c[1], b.c[1]), fsub(a.c[2], b.c[2]), fsub(a.c[3], b.c[3])}};
}
LSK_HD Ext eneg(const Ext& a) { return Ext{{fneg(a.c[0]), fneg(a.c[1]), fneg(a.c[2]), fneg(a.c[3])}}; }
LSK_HD Ext emul_base(const Ext& a, uint32_t b) {
  return Ext{{fmul(a.c[0], b), fmul(a.c[1], b), fmul(a.c[2], b), fmul(a.c[3], b)}};
}
LSK_HD Ext eadd_base(const Ext& a, uint32_t b) { return Ext{{fadd(a.c[0], b), a.c[1], a.c[2], a.c[3]}}; }
LSK_HD Ext esub_base(const Ext& a, uint32_t b) { return Ext{{fsub(a.c[0], b), a.c[1], a.c[2], a.c[3]}}; }

// Schoolbook product with the 64-bit partial sums reduced once per coefficient: each a_i*b_j is
// < P^2 < 2^62, so the products of one coefficient are accumulated modulo M = P*2^32 (the bound mont_reduce
// accepts) before a single Montgomery reduction.  acc_reduce64 takes x < 2M: M has a zero low word, so the
// conditional subtraction only touches the HIGH word -- one VIADDMNMX on the device, where the 64-bit compare /
// subtract / select the portable form compiles to is six instructions (cuobjdump, sm_100a).
constexpr uint32_t M11 = 939524073u;  // 11 in Montgomery form: the non-residue W of x^4 = W
LSK_HD uint64_t acc_reduce64(uint64_t x) {  // x < 2 * P * 2^32  ->  x mod P * 2^32
#if defined(__CUDA_ARCH__)
  const uint32_t hi = (uint32_t)(x >> 32);
  return ((uint64_t)min(hi, hi - P) << 32) | (uint32_t)x;
#else
  const uint64_t M = (uint64_t)P << 32;
  return x >= M ? x - M : x;
#endif
}
// The part of the product that wraps around x^4 = 11 is reduced to a field element h and re-enters the low
// coefficient's 64-bit sum as the raw product h * M11 (one more IMAD.WIDE) instead of three doublings and two
// additions of reduced values: 19 wide products, 7 Montgomery reductions and 6 one-instruction corrections per
// extension product.
LSK_HD Ext emul(const Ext& a, const Ext& b) {
  const uint32_t h0 = mont_reduce(acc_reduce64(acc_reduce64((uint64_t)a.c[1] * b.c[3] + (uint64_t)a.c[2] * b.c[2]) +
                                               (uint64_t)a.c[3] * b.c[1]));
  const uint32_t h1 = mont_reduce((uint64_t)a.c[2] * b.c[3] + (uint64_t)a.c[3] * b.c[2]);   // 2 P^2 < P * 2^32
  const uint32_t h2 = fmul(a.c[3], b.c[3]);
  const uint32_t c0 = mont_reduce((uint64_t)a.c[0] * b.c[0] + (uint64_t)h0 * M11);
  const uint32_t c1 = mont_reduce(acc_reduce64((uint64_t)a.c[0] * b.c[1] + (uint64_t)a.c[1] * b.c[0] + (uint64_t)h1 * M11));
  const uint32_t c2 = mont_reduce(acc_reduce64(
      acc_reduce64((uint64_t)a.c[0] * b.c[2] + (uint64_t)a.c[1] * b.c[1] + (uint64_t)a.c[2] * b.c[0]) + (uint64_t)h2 * M11));
  const uint32_t c3 = mont_reduce(acc_reduce64(
      acc_reduce64((uint64_t)a.c[0] * b.c[3] + (uint64_t)a.c[1] * b.c[2] + (uint64_t)a.c[2] * b.c[1]) +
      (uint64_t)a.c[3] * b.c[0]));
  return Ext{{c0, c1, c2, c3}};
}
// sum_k a_k * b_k of extension elements a_k by base elements b_k (random linear combinations: LogUp messages, folded
// constraints, openings) as four 64-bit sums of raw products.  add() leaves a sum below M + P^2, add_reduce() takes
// such a sum back below M (M + 2 P^2 < 2M), so the two alternate; value() reduces once.  Same field elements as
// eadd(acc, emul_base(a, b)) term by term, at 4 IMAD.WIDE + 2 corrections per term instead of 4 Montgomery products
// + 4 modular additions (32 instructions).
struct ExtAcc {
  uint64_t s[4];
  LSK_HD void add(const Ext& a, uint32_t b) {          // precondition: every s < M
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) s[k] += (uint64_t)a.c[k] * b;
  }
  LSK_HD void add_reduce(const Ext& a, uint32_t b) {   // precondition: every s < M + P^2
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) s[k] = acc_reduce64(s[k] + (uint64_t)a.c[k] * b);
  }
  LSK_HD Ext value() const {
    return Ext{{mont_reduce(acc_reduce64(s[0])), mont_reduce(acc_reduce64(s[1])), mont_reduce(acc_reduce64(s[2])),
                mont_reduce(acc_reduce64(s[3]))}};
  }
};
LSK_HD ExtAcc ext_acc_zero() { return ExtAcc{{0, 0, 0, 0}}; }
LSK_HD Ext esqr(const Ext& a) { return emul(a, a); }

// 1/a via the tower F[y]/(y^2-11) -> F[x]/(x^2-y): a = A + xB, 1/a = (A - xB)/(A^2 - yB^2).
LSK_HD Ext einv(const Ext& a) {
  uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
  uint32_t a1a3 = fmul(a1, a3), a0a2 = fmul(a0, a2);
  uint32_t d0 = fsub(fadd(fsqr(a0), fmul11(fsqr(a2))), fmul11(fdbl(a1a3)));
  uint32_t d1 = fsub(fsub(fdbl(a0a2), fsqr(a1)), fmul11(fsqr(a3)));
  uint32_t n = fsub(fsqr(d0), fmul11(fsqr(d1)));
  uint32_t ni = finv(n);
  uint32_t e0 = fmul(d0, ni), e1 = fneg(fmul(d1, ni));
  uint32_t r0 = fadd(fmul(a0, e0), fmul11(fmul(a2, e1)));
  uint32_t r2 = fadd(fmul(a0, e1), fmul(a2, e0));
  uint32_t r1 = fneg(fadd(fmul(a1, e0), fmul11(fmul(a3, e1))));
  uint32_t r3 = fneg(fadd(fmul(a1, e1), fmul(a3, e0)));
  return Ext{{r0, r1, r2, r3}};
}
LSK_HD Ext epow(Ext b, uint64_t e) {
  Ext r = ext_one();
  while (e) {
    if (e & 1) r = emul(r, b);
    b = esqr(b);
    e >>= 1;
  }
  return r;
}

LSK_HD uint32_t bitrev32(uint32_t x, unsigned bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
  uint32_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}
LSK_HD unsigned log2_exact(uint64_t n) {
  unsigned l = 0;
  while ((1ull << l) < n) l++;
  return l;
}

}  // namespace lsk
