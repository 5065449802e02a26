// Host-side description of the STARK machine: the chip vector of
// build_lair_chip_vector / build_chip_vector (src/lair/lair_chip.rs:196-231) after each chip's
// `Air::eval` has been captured symbolically by the host layer (lurk_b200/air.py), i.e. what
// sphinx-core's `Chip::new` + `StarkMachine::new` hold (call site src/core/stark_machine.rs:20-30).
//
// Blob format ('MACH', all u32 words; produced by lurk_b200/air.py::serialize_machine):
//   MAGIC nchips num_pv { len kind index width prep_width n_constraints log_quotient_degree n_regs
//                         name_len name_words name.. nconsts consts.. ninstr (op dst a b)*ninstr
//                         nsends interaction* nreceives interaction* }*
//   interaction := kind nvalues affine*nvalues affine(multiplicity)
//   affine      := nterms const (is_prep col coeff)*nterms
// Constraint program: straight-line register code over base-field values;
//   op 1 add, 2 sub, 3 mul, 4 neg (dst = reg), 5 assert_zero(a);
//   operand = tag << 28 | idx, tag 0 reg, 1 const, 2 main local, 3 main next, 4 prep local,
//   5 prep next, 6 public value, 7 selector (0 is_first, 1 is_last, 2 is_transition).
#pragma once
#include <array>
#include <string>
#include <vector>

#include "common.h"

namespace lsk {

constexpr uint32_t MACHINE_MAGIC = 0x4d414348u;
enum : uint32_t { CHIP_FUNC = 0, CHIP_MEM = 1, CHIP_BYTES = 2, CHIP_ENTRY = 3 };
enum : uint32_t { COP_ADD = 1, COP_SUB = 2, COP_MUL = 3, COP_NEG = 4, COP_ASSERT = 5 };

struct AffineH {
  uint32_t cst = 0;                                // canonical
  std::vector<std::array<uint32_t, 3>> terms;      // (is_prep, col, coeff canonical)
};
struct InteractionH {
  uint32_t kind = 0;
  std::vector<AffineH> values;
  AffineH mult;
};
struct ChipH {
  std::string name;
  uint32_t kind = 0, index = 0, width = 0, prep_width = 0, n_constraints = 0, lqd = 0, n_regs = 0;
  std::vector<uint32_t> consts, code;
  std::vector<InteractionH> sends, receives;
  size_t n_interactions() const { return sends.size() + receives.size(); }
  size_t perm_width() const {  // extension columns: one per batch of 2^lqd interactions + running sum
    size_t batch = (size_t)1 << lqd;
    return (n_interactions() + batch - 1) / batch + 1;
  }
  size_t n_perm_constraints() const { return perm_width() - 1 + 3; }
  size_t max_values() const {
    size_t m = 0;
    for (auto& i : sends) m = std::max(m, i.values.size());
    for (auto& i : receives) m = std::max(m, i.values.size());
    return m;
  }
};

inline AffineH parse_affine(const uint32_t* w, size_t n, size_t& p) {
  if (p + 2 > n) fail(LSK_ERR_ARG, "machine blob truncated");
  AffineH a;
  uint32_t k = w[p++];
  a.cst = w[p++];
  if (p + 3ull * k > n) fail(LSK_ERR_ARG, "machine blob truncated");
  for (uint32_t i = 0; i < k; i++, p += 3) a.terms.push_back({w[p], w[p + 1], w[p + 2]});
  return a;
}

inline std::vector<ChipH> parse_machine_blob(const uint32_t* w, size_t n, uint32_t* num_pv) {
  if (n < 3 || w[0] != MACHINE_MAGIC) fail(LSK_ERR_ARG, "bad machine blob magic");
  if (w[1] > n) fail(LSK_ERR_ARG, "machine blob truncated");
  std::vector<ChipH> chips(w[1]);
  *num_pv = w[2];
  size_t p = 3;
  for (auto& c : chips) {
    if (p + 10 > n) fail(LSK_ERR_ARG, "machine blob truncated");
    size_t len = w[p++], end = p + len;
    if (len > n || end > n) fail(LSK_ERR_ARG, "machine blob truncated");
    auto rd = [&]() -> uint32_t {   // every read of this chip's descriptor stays inside it
      if (p >= end) fail(LSK_ERR_ARG, "machine blob truncated");
      return w[p++];
    };
    c.kind = rd(); c.index = rd(); c.width = rd(); c.prep_width = rd();
    c.n_constraints = rd(); c.lqd = rd(); c.n_regs = rd();
    uint32_t name_len = rd(), name_words = rd();
    if (name_words > end - p || name_len > 4ull * name_words) fail(LSK_ERR_ARG, "bad chip name");
    c.name.assign((const char*)(w + p), name_len);
    p += name_words;
    uint32_t nc = rd();
    if (nc > end - p) fail(LSK_ERR_ARG, "machine blob truncated");
    c.consts.assign(w + p, w + p + nc);
    p += nc;
    uint32_t ni = rd();
    if (4ull * ni > end - p) fail(LSK_ERR_ARG, "machine blob truncated");
    c.code.assign(w + p, w + p + 4ull * ni);
    p += 4ull * ni;
    for (int s = 0; s < 2; s++) {
      uint32_t cnt = rd();
      auto& lst = s == 0 ? c.sends : c.receives;
      for (uint32_t i = 0; i < cnt; i++) {
        InteractionH it;
        it.kind = rd();
        uint32_t nv = rd();
        for (uint32_t k = 0; k < nv; k++) it.values.push_back(parse_affine(w, end, p));
        it.mult = parse_affine(w, end, p);
        lst.push_back(std::move(it));
      }
    }
    if (p != end) fail(LSK_ERR_ARG, "chip descriptor length mismatch");
    // interaction operands are walked unchecked by the prover (perm_rows) and the verifier (eval_affine)
    for (int s = 0; s < 2; s++)
      for (auto& it : (s == 0 ? c.sends : c.receives)) {
        auto check_affine = [&](const AffineH& a) {
          for (auto& t : a.terms)
            if (t[0] > 1 || t[1] >= (t[0] ? c.prep_width : c.width))
              fail(LSK_ERR_ARG, "an interaction of chip %s references a column outside its trace", c.name.c_str());
        };
        check_affine(it.mult);
        for (auto& v : it.values) check_affine(v);
      }
    if (c.lqd != 1) fail(LSK_ERR_ARG, "only log_quotient_degree = 1 is supported (chip %s)", c.name.c_str());
    if (c.width > (1u << 20) || c.prep_width > (1u << 20) || c.n_regs > (1u << 20))
      fail(LSK_ERR_ARG, "implausible sizes in the descriptor of chip %s", c.name.c_str());
  }
  return chips;
}

}  // namespace lsk
