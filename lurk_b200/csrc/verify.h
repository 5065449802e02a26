// Verifier state: what `machine.setup` returns as the verifying key (src/core/cli/repl.rs:175-176) plus the
// chip vector and FRI parameters the machine itself holds.  Host-only.
#pragma once
#include <array>
#include <vector>

#include "challenger.h"
#include "machine.h"

namespace lsk {

struct VerifyingKey {
  std::vector<ChipH> chips;
  uint32_t num_pv = 0;
  std::array<uint32_t, 8> prep_commit{};   // Montgomery form
  bool has_prep = false;
  uint32_t num_queries = 100, pow_bits = 16;
  HostPerm16 perm;
};

// throws lsk::Error(LSK_ERR_VERIFY, reason) when the proof is rejected
void verify_proof(const VerifyingKey& vk, const uint32_t* words, size_t n);

}  // namespace lsk
