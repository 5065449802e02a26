// ORACLE (test infrastructure, NOT product code).  CPU restatement of the BabyBear field and
// its quartic binomial extension as used by the reference through the third-party crates
// p3-baby-bear / p3-field (Plonky3 `sp1` branch @ a0b9287, not vendored under /root/reference;
// see SURVEY.md F1).  Call sites in the reference: src/lair/mod.rs:33-46 (from_canonical_u32),
// src/air/builder.rs:159-168 (inverse), every `F::` in src/lair/trace.rs.
//
// p = 2^31 - 2^27 + 1 = 2013265921; multiplicative generator 31; two-adic generator of order
// 2^27 = 31^((p-1)/2^27) = 0x1a427a41; extension F[x]/(x^4 - 11)  (the
// generator / extension choices are *recollection* of Plonky3 -- parity unpinned for anything
// that depends on them, i.e. everything below the trace/hash level).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything under oracle/.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace orc {

static const uint32_t P = 0x78000001u;
static const uint32_t MONTY_PINV = 0x88000001u;  // P * PINV == 1 mod 2^32
static const uint32_t R2 = 1172168163u;          // 2^64 mod P

// Montgomery representation: value v stored as v * 2^32 mod P, always reduced to [0, P).
static inline uint32_t monty_reduce(uint64_t x) {
  uint32_t t = (uint32_t)x * MONTY_PINV;
  uint64_t u = (uint64_t)t * P;
  uint32_t hx = (uint32_t)(x >> 32), hu = (uint32_t)(u >> 32);
  uint32_t r = hx - hu;
  return hx < hu ? r + P : r;
}

struct Fp {
  uint32_t v;  // Montgomery form
  static inline Fp from_u32(uint32_t x) { return Fp{monty_reduce((uint64_t)(x % P) * R2)}; }
  static inline Fp from_i64(int64_t x) {
    int64_t m = x % (int64_t)P;
    if (m < 0) m += P;
    return from_u32((uint32_t)m);
  }
  static inline Fp zero() { return Fp{0}; }
  static inline Fp one() { return from_u32(1); }
  inline uint32_t to_u32() const { return monty_reduce(v); }
  inline bool is_zero() const { return v == 0; }
};
static inline bool operator==(Fp a, Fp b) { return a.v == b.v; }
static inline bool operator!=(Fp a, Fp b) { return a.v != b.v; }
static inline Fp operator+(Fp a, Fp b) {
  uint32_t s = a.v + b.v;
  return Fp{s >= P ? s - P : s};
}
static inline Fp operator-(Fp a, Fp b) {
  uint32_t s = a.v - b.v;
  return Fp{a.v < b.v ? s + P : s};
}
static inline Fp operator-(Fp a) { return Fp{a.v ? P - a.v : 0}; }
static inline Fp operator*(Fp a, Fp b) { return Fp{monty_reduce((uint64_t)a.v * b.v)}; }
static inline Fp& operator+=(Fp& a, Fp b) { return a = a + b; }
static inline Fp& operator-=(Fp& a, Fp b) { return a = a - b; }
static inline Fp& operator*=(Fp& a, Fp b) { return a = a * b; }

static inline Fp fp_pow(Fp b, uint64_t e) {
  Fp r = Fp::one();
  while (e) {
    if (e & 1) r = r * b;
    b = b * b;
    e >>= 1;
  }
  return r;
}
// p3 `Field::inverse` panics on zero; `try_inverse` returns None.  We return 0 for 0 and let
// the callers that must panic check explicitly.
static inline Fp fp_inv(Fp a) { return fp_pow(a, (uint64_t)P - 2); }

static inline Fp fp_generator() { return Fp::from_u32(31); }
// generator of the order-2^bits subgroup (p3 `TwoAdicField::two_adic_generator`)
static inline Fp two_adic_generator(unsigned bits) {
  Fp g = Fp::from_u32(0x1a427a41u);
  for (unsigned i = bits; i < 27; i++) g = g * g;
  return g;
}

// ---- quartic extension F[x]/(x^4 - W), W = 11 -------------------------------------------
struct Ef {
  Fp c[4];
  static inline Ef zero() { return Ef{{Fp::zero(), Fp::zero(), Fp::zero(), Fp::zero()}}; }
  static inline Ef one() { return Ef{{Fp::one(), Fp::zero(), Fp::zero(), Fp::zero()}}; }
  static inline Ef from_base(Fp a) { return Ef{{a, Fp::zero(), Fp::zero(), Fp::zero()}}; }
  inline bool is_zero() const { return c[0].v == 0 && c[1].v == 0 && c[2].v == 0 && c[3].v == 0; }
};
static inline bool operator==(const Ef& a, const Ef& b) {
  return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3];
}
static inline bool operator!=(const Ef& a, const Ef& b) { return !(a == b); }
static inline Ef operator+(const Ef& a, const Ef& b) {
  return Ef{{a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2], a.c[3] + b.c[3]}};
}
static inline Ef operator-(const Ef& a, const Ef& b) {
  return Ef{{a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2], a.c[3] - b.c[3]}};
}
static inline Ef operator-(const Ef& a) { return Ef{{-a.c[0], -a.c[1], -a.c[2], -a.c[3]}}; }
static inline Ef operator*(const Ef& a, Fp b) { return Ef{{a.c[0] * b, a.c[1] * b, a.c[2] * b, a.c[3] * b}}; }
static inline Ef operator+(const Ef& a, Fp b) { return Ef{{a.c[0] + b, a.c[1], a.c[2], a.c[3]}}; }
static inline Ef operator-(const Ef& a, Fp b) { return Ef{{a.c[0] - b, a.c[1], a.c[2], a.c[3]}}; }
static inline Fp mul_w(Fp a) {  // * 11
  Fp a2 = a + a, a4 = a2 + a2, a8 = a4 + a4;
  return a8 + a2 + a;
}
static inline Ef operator*(const Ef& a, const Ef& b) {
  Fp c0 = a.c[0] * b.c[0] + mul_w(a.c[1] * b.c[3] + a.c[2] * b.c[2] + a.c[3] * b.c[1]);
  Fp c1 = a.c[0] * b.c[1] + a.c[1] * b.c[0] + mul_w(a.c[2] * b.c[3] + a.c[3] * b.c[2]);
  Fp c2 = a.c[0] * b.c[2] + a.c[1] * b.c[1] + a.c[2] * b.c[0] + mul_w(a.c[3] * b.c[3]);
  Fp c3 = a.c[0] * b.c[3] + a.c[1] * b.c[2] + a.c[2] * b.c[1] + a.c[3] * b.c[0];
  return Ef{{c0, c1, c2, c3}};
}
static inline Ef& operator+=(Ef& a, const Ef& b) { return a = a + b; }
static inline Ef& operator-=(Ef& a, const Ef& b) { return a = a - b; }
static inline Ef& operator*=(Ef& a, const Ef& b) { return a = a * b; }

// a = A + x B with A = a0 + a2 y, B = a1 + a3 y, y = x^2, y^2 = W.
// 1/a = (A - x B) / (A^2 - y B^2); the denominator D = d0 + d1 y lives in F[y]/(y^2 - W) and
// 1/D = (d0 - d1 y) / (d0^2 - W d1^2).
static inline Ef ef_inv(const Ef& a) {
  Fp a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
  // A^2 = (a0^2 + W a2^2) + (2 a0 a2) y ;  B^2 = (a1^2 + W a3^2) + (2 a1 a3) y
  // y B^2 = W (2 a1 a3) + (a1^2 + W a3^2) y
  Fp d0 = a0 * a0 + mul_w(a2 * a2) - mul_w((a1 * a3) + (a1 * a3));
  Fp d1 = (a0 * a2) + (a0 * a2) - (a1 * a1) - mul_w(a3 * a3);
  Fp n = d0 * d0 - mul_w(d1 * d1);
  Fp ni = fp_inv(n);
  Fp e0 = d0 * ni, e1 = -(d1 * ni);  // 1/D = e0 + e1 y
  // (A - xB) * (e0 + e1 y):  A*(e0+e1 y) = (a0 e0 + W a2 e1) + (a0 e1 + a2 e0) y
  //                          B*(e0+e1 y) = (a1 e0 + W a3 e1) + (a1 e1 + a3 e0) y
  Fp r0 = a0 * e0 + mul_w(a2 * e1);
  Fp r2 = a0 * e1 + a2 * e0;
  Fp r1 = -(a1 * e0 + mul_w(a3 * e1));
  Fp r3 = -(a1 * e1 + a3 * e0);
  return Ef{{r0, r1, r2, r3}};
}
static inline Ef ef_pow(Ef b, uint64_t e) {
  Ef r = Ef::one();
  while (e) {
    if (e & 1) r = r * b;
    b = b * b;
    e >>= 1;
  }
  return r;
}

static inline unsigned log2_strict(size_t n) {
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
static inline size_t bitrev(size_t x, unsigned bits) {
  size_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

}  // namespace orc
