// Host-side Fiat-Shamir transcript: the third-party `DuplexChallenger<Val, Perm16, 16, 8>` that
// sphinx's `machine.config().challenger()` returns (call sites src/core/cli/repl.rs:195,
// benches/fib.rs:157).  Width-16 Poseidon2 over BabyBear with the in-tree BabyBearConfig16
// constants (src/poseidon/config.rs:189-198; see DESIGN.md "unpinned assumptions"), rate 8,
// samples pop from the back of the output buffer.  It absorbs a few hundred words per shard, so it
// runs on the host between kernel launches; only proof-of-work grinding goes to the device.
#pragma once
#include <vector>

#include "field.cuh"
#include "poseidon2_constants.h"

namespace lsk {

struct HostPerm16 {
  uint32_t diag[16], full_rc[8 * 16], part_rc[13];
  HostPerm16() {
    for (int i = 0; i < 16; i++) diag[i] = to_mont(P2_DIAG_16[i]);
    for (int i = 0; i < 8 * 16; i++) full_rc[i] = to_mont(P2_FULL_RC_16[i]);
    for (int i = 0; i < 13; i++) part_rc[i] = to_mont(P2_PART_RC_16[i]);
  }
  static void external(uint32_t* s) {
    for (int i = 0; i < 16; i += 4) {
      uint32_t x0 = s[i], x1 = s[i + 1], x2 = s[i + 2], x3 = s[i + 3];
      uint32_t sum = fadd(fadd(x0, x1), fadd(x2, x3));
      s[i] = fadd(fadd(sum, x0), fdbl(x1));        // 2 3 1 1
      s[i + 1] = fadd(fadd(sum, x1), fdbl(x2));    // 1 2 3 1
      s[i + 2] = fadd(fadd(sum, x2), fdbl(x3));    // 1 1 2 3
      s[i + 3] = fadd(fadd(sum, x3), fdbl(x0));    // 3 1 1 2
    }
    uint32_t sums[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++) sums[i & 3] = fadd(sums[i & 3], s[i]);
    for (int i = 0; i < 16; i++) s[i] = fadd(s[i], sums[i & 3]);
  }
  static uint32_t sbox(uint32_t x) {
    uint32_t x3 = fmul(fsqr(x), x);
    return fmul(x, fsqr(x3));
  }
  void permute(uint32_t* s) const {  // Montgomery in / out
    external(s);
    for (int r = 0; r < 4; r++) {
      for (int i = 0; i < 16; i++) s[i] = sbox(fadd(s[i], full_rc[r * 16 + i]));
      external(s);
    }
    for (int r = 0; r < 13; r++) {
      s[0] = sbox(fadd(s[0], part_rc[r]));
      uint32_t sum = 0;
      for (int i = 0; i < 16; i++) sum = fadd(sum, s[i]);
      for (int i = 0; i < 16; i++) s[i] = fadd(fmul(s[i], diag[i]), sum);
    }
    for (int r = 4; r < 8; r++) {
      for (int i = 0; i < 16; i++) s[i] = sbox(fadd(s[i], full_rc[r * 16 + i]));
      external(s);
    }
  }
};

// all values Montgomery form
struct Challenger {
  const HostPerm16* perm;
  uint32_t state[16];
  std::vector<uint32_t> in, out;
  explicit Challenger(const HostPerm16* p) : perm(p) {
    for (auto& s : state) s = 0;
  }
  void duplex() {
    for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
    in.clear();
    perm->permute(state);
    out.assign(state, state + 8);
  }
  void observe(uint32_t x) {
    out.clear();
    in.push_back(x);
    if (in.size() == 8) duplex();
  }
  void observe_n(const uint32_t* x, size_t n) {
    for (size_t i = 0; i < n; i++) observe(x[i]);
  }
  void observe(const Ext& e) { observe_n(e.c, 4); }
  uint32_t sample() {
    if (!in.empty() || out.empty()) duplex();
    uint32_t x = out.back();
    out.pop_back();
    return x;
  }
  Ext sample_ext() {
    Ext e;
    for (int i = 0; i < 4; i++) e.c[i] = sample();
    return e;
  }
  uint32_t sample_bits(unsigned bits) { return from_mont(sample()) & ((1u << bits) - 1u); }
  bool check_witness(unsigned bits, uint32_t w_mont) {
    observe(w_mont);
    return sample_bits(bits) == 0;
  }
};

}  // namespace lsk
