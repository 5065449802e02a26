// ORACLE (test infrastructure, NOT product code).  CPU restatement of the Poseidon2 permutation
// over BabyBear exactly as the reference drives it:
//   round order        src/poseidon/wide/trace.rs:12-28   (== p3_poseidon2::Poseidon2::permute)
//   external round     src/poseidon/wide/trace.rs:30-54   (add RC, x^7 as x * (x^3)^2, M_E)
//   internal round     src/poseidon/wide/trace.rs:56-82   (x0 += rc, x0^7, M_I)
//   M_I                src/poseidon/config.rs:104-119     (x_i * diag_i + sum)
//   M_E                third-party Poseidon2ExternalMatrixGeneral (p3-poseidon2 0.1.0 @ a0b9287,
//                      absent from /root/reference): M4 = circ(2,3,1,1) on every 4-chunk, then
//                      x_i += sum_{j = i mod 4} x_j; structure described at src/poseidon/air.rs:133-138;
//                      pinned by KAT-1 (src/core/zstore.rs:1008-1019) and KAT-2
//                      (src/core/cli/tests/prove.lurk:12-17) -- see tests/test_oracle_poseidon2.py.
//   wide witness       src/poseidon/wide/columns.rs:17-32 and src/core/poseidon.rs:65-72
// Constants: poseidon2_constants.h, generated from src/poseidon/constants.rs.
#pragma once
#include "bb.h"
#include "poseidon2_constants.h"

namespace orc {

struct P2Params {
  int t, rp;
  const uint32_t *diag, *full_rc, *part_rc;
};
static inline P2Params p2_params(int t) {
  switch (t) {
    case 16: return {16, P2_RP_16, P2_DIAG_16, P2_FULL_RC_16, P2_PART_RC_16};
    case 24: return {24, P2_RP_24, P2_DIAG_24, P2_FULL_RC_24, P2_PART_RC_24};
    case 32: return {32, P2_RP_32, P2_DIAG_32, P2_FULL_RC_32, P2_PART_RC_32};
    case 40: return {40, P2_RP_40, P2_DIAG_40, P2_FULL_RC_40, P2_PART_RC_40};
  }
  return {0, 0, nullptr, nullptr, nullptr};
}

// Montgomery-form copies of the constants, built once per width.
struct P2Consts {
  int t = 0, rp = 0;
  Fp diag[40], full_rc[8 * 40], part_rc[40];
};
static inline const P2Consts& p2_consts(int t) {
  static P2Consts tab[4];
  int slot = t == 16 ? 0 : t == 24 ? 1 : t == 32 ? 2 : 3;
  P2Consts& c = tab[slot];
  if (c.t != t) {  // benign race: every thread writes identical values
    P2Params p = p2_params(t);
    for (int i = 0; i < t; i++) c.diag[i] = Fp::from_u32(p.diag[i]);
    for (int i = 0; i < 8 * t; i++) c.full_rc[i] = Fp::from_u32(p.full_rc[i]);
    for (int i = 0; i < p.rp; i++) c.part_rc[i] = Fp::from_u32(p.part_rc[i]);
    c.rp = p.rp;
    c.t = t;
  }
  return c;
}

static inline void p2_external_linear(Fp* s, int t) {
  for (int i = 0; i < t; i += 4) {
    Fp x0 = s[i], x1 = s[i + 1], x2 = s[i + 2], x3 = s[i + 3];
    s[i] = x0 + x0 + x1 + x1 + x1 + x2 + x3;
    s[i + 1] = x0 + x1 + x1 + x2 + x2 + x2 + x3;
    s[i + 2] = x0 + x1 + x2 + x2 + x3 + x3 + x3;
    s[i + 3] = x0 + x0 + x0 + x1 + x2 + x3 + x3;
  }
  Fp sums[4] = {Fp::zero(), Fp::zero(), Fp::zero(), Fp::zero()};
  for (int i = 0; i < t; i++) sums[i & 3] += s[i];
  for (int i = 0; i < t; i++) s[i] += sums[i & 3];
}

static inline void p2_internal_linear(Fp* s, const P2Consts& c) {
  Fp sum = Fp::zero();
  for (int i = 0; i < c.t; i++) sum += s[i];
  for (int i = 0; i < c.t; i++) s[i] = s[i] * c.diag[i] + sum;
}

// Permutation with optional wide-witness capture.  `wit` (may be null) receives, in the
// repr(C) order of Poseidon2Cols: external_rounds_state[8][t], external_rounds_sbox[8][t],
// internal_rounds_state_init[t], internal_rounds_state0[rp-1], internal_rounds_sbox[rp].
static inline void p2_permute(Fp* s, int t, Fp* wit = nullptr) {
  const P2Consts& c = p2_consts(t);
  Fp* w_ext_state = wit;
  Fp* w_ext_sbox = wit ? wit + 8 * t : nullptr;
  Fp* w_int_init = wit ? wit + 16 * t : nullptr;
  Fp* w_int_s0 = wit ? wit + 17 * t : nullptr;
  Fp* w_int_sbox = wit ? wit + 17 * t + (c.rp - 1) : nullptr;
  p2_external_linear(s, t);
  auto ext_round = [&](int r) {
    for (int i = 0; i < t; i++) {
      if (wit) w_ext_state[r * t + i] = s[i];
      Fp x = s[i] + c.full_rc[r * t + i];
      Fp x3 = x * x * x;
      if (wit) w_ext_sbox[r * t + i] = x3;
      s[i] = x * (x3 * x3);
    }
    p2_external_linear(s, t);
  };
  for (int r = 0; r < 4; r++) ext_round(r);
  for (int r = 0; r < c.rp; r++) {
    if (wit) {
      if (r == 0)
        for (int i = 0; i < t; i++) w_int_init[i] = s[i];
      else
        w_int_s0[r - 1] = s[0];
    }
    Fp x = s[0] + c.part_rc[r];
    Fp x3 = x * x * x;
    if (wit) w_int_sbox[r] = x3;
    s[0] = x * (x3 * x3);
    p2_internal_linear(s, c);
  }
  for (int r = 4; r < 8; r++) ext_round(r);
}

static inline int p2_witness_cols(int t) { return 17 * t + 2 * p2_params(t).rp - 1; }

// ---- width-16 symmetric primitives of the commitment scheme (third party: p3-symmetric) ----
// PaddingFreeSponge<Perm,16,8,8>::hash_iter: overwrite-mode absorption, rate 8, squeeze 8.
struct Sponge16 {
  Fp st[16];
  int pos = 0;
  Sponge16() {
    for (auto& x : st) x = Fp::zero();
  }
  inline void absorb(Fp x) {
    st[pos++] = x;
    if (pos == 8) {
      p2_permute(st, 16);
      pos = 0;
    }
  }
  inline void finish(Fp out[8]) {
    if (pos != 0) p2_permute(st, 16);
    for (int i = 0; i < 8; i++) out[i] = st[i];
  }
};
// TruncatedPermutation<Perm,2,8,16>::compress
static inline void compress16(const Fp l[8], const Fp r[8], Fp out[8]) {
  Fp st[16];
  for (int i = 0; i < 8; i++) st[i] = l[i], st[8 + i] = r[i];
  p2_permute(st, 16);
  for (int i = 0; i < 8; i++) out[i] = st[i];
}

}  // namespace orc
