// ORACLE (test infrastructure, NOT product code).  Row-by-row evaluation of every chip's AIR on
// CONCRETE values, restated from the reference's `Air::eval` implementations directly -- it never
// reads the product's captured constraint program ('MACH' blob).  It is the reference's own semantic
// checker (src/air/debug.rs) plus the lookup expansion the real prover sees (src/air/builder.rs):
//
//   DebugConstraintBuilder / TraceQueries    src/air/debug.rs:17-116,161-206,209-406
//   LookupBuilder::{provide, require}        src/air/builder.rs:42-104 (receive + send per access,
//                                            `count * count_inv = 1` per require)
//   Func::eval / Block / Op / Ctrl           src/lair/air.rs:158-566, eval_depth :105-131
//   relations                                src/lair/relations.rs:6-59, src/gadgets/bytes/relation.rs:7,72-135
//   BytesAirRecordWithContext                src/gadgets/bytes/builder.rs:17-90, bytes/mod.rs:129-160
//   MemChip::eval                            src/lair/memory.rs:71-109
//   BytesChip::eval                          src/gadgets/bytes/trace.rs:116-143
//   Entrypoint                               src/lair/lair_chip.rs:166-191
//   hasher chips                             src/core/poseidon.rs:79-93, src/poseidon/wide/air.rs:11-124
//   U64 chipset                              src/core/u64.rs:170-222 over src/gadgets/unsigned/{add,mul,
//                                            div_rem,cmp,is_zero,less_than}.rs
//   BigNum::LessThan                         src/core/big_num.rs:75-97, src/gadgets/big_num/cmp.rs:50-115,
//                                            src/gadgets/unsigned/field.rs:38-140
//
// Two uses (tests/test_oracle_air.py):
//   strict mode  = debug.rs: every constraint must vanish, every multiplicity must be 0 or 1, memoset
//                  accesses are collected and `TraceQueries::verify` is run over all chips and shards;
//   raw mode     = the constraint VALUES and the send / receive messages of arbitrary (random) rows, in
//                  evaluation order -- compared with what the product's captured program computes on the
//                  same rows, which pins the capture to this restatement.
// Pinned by the reference's own tests: every golden trace passes, `lair_range_test`'s three memoset
// keys (src/lair/air.rs:967-977) are present.
#pragma once
#include <map>
#include <sstream>

#include "lair.h"

namespace orc {

struct AirMsg {
  std::vector<Fp> values;
  Fp mult;
};
struct MemoRec { uint32_t prev_nonce, prev_count, nonce; };  // debug.rs:17-24

struct TraceQueries {  // debug.rs:26-116
  std::map<std::vector<uint32_t>, long> balance;  // sends - receives per message (canonical words)
  std::map<std::vector<uint32_t>, std::map<uint32_t, MemoRec>> memoset;
  size_t n_sends = 0, n_receives = 0;
  static std::vector<uint32_t> key(const std::vector<Fp>& v) {
    std::vector<uint32_t> k(v.size());
    for (size_t i = 0; i < v.size(); i++) k[i] = v[i].to_u32();
    return k;
  }
  void send(const std::vector<Fp>& m) { balance[key(m)]++; n_sends++; }
  void receive(const std::vector<Fp>& m) { balance[key(m)]--; n_receives++; }
  void access(const std::vector<Fp>& q, uint32_t count, MemoRec r) {
    auto& recs = memoset[key(q)];
    if (!recs.emplace(count, r).second) throw std::runtime_error("memoset record already accessed");
  }
  void merge(const TraceQueries& o) {  // debug.rs:58-89
    for (auto& kv : o.balance) balance[kv.first] += kv.second;
    n_sends += o.n_sends;
    n_receives += o.n_receives;
    for (auto& kv : o.memoset)
      for (auto& r : kv.second)
        if (!memoset[kv.first].emplace(r.first, r.second).second) throw std::runtime_error("memoset record already accessed");
  }
  void verify() const {  // debug.rs:91-107
    for (auto& kv : balance)
      if (kv.second != 0) {
        std::ostringstream s;
        s << "sends != receives for message [";
        for (uint32_t x : kv.first) s << x << " ";
        s << "] (balance " << kv.second << ")";
        throw std::runtime_error(s.str());
      }
    for (auto& kv : memoset) {
      const auto& recs = kv.second;
      uint32_t prev_count = recs.rbegin()->first;
      MemoRec prev = recs.rbegin()->second;
      size_t i = 0;
      for (auto& r : recs) {
        if (r.first != i) throw std::runtime_error("count should be increasing");
        if (r.second.prev_count != prev_count) throw std::runtime_error("memoset: prev_count does not chain");
        if (r.second.prev_nonce != prev.nonce) throw std::runtime_error("memoset: prev_nonce does not chain");
        prev_count = r.first;
        prev = r.second;
        i++;
      }
    }
  }
};

struct AirRowOut {
  std::vector<Fp> constraints;
  std::vector<AirMsg> sends, receives;
  void clear() { constraints.clear(); sends.clear(); receives.clear(); }
};

// The concrete-value builder.  `q` (strict mode) collects the memoset accesses and the message balance;
// `out` (raw mode) collects values in evaluation order.
struct DebugBuilder {
  const Fp *local = nullptr, *next = nullptr, *prep_local = nullptr, *prep_next = nullptr, *pv = nullptr;
  Fp is_first, is_last, is_transition;
  size_t row = 0;
  bool strict = true;
  TraceQueries* q = nullptr;
  AirRowOut* out = nullptr;
  size_t n_constraints = 0;

  void assert_zero(Fp x) {
    if (out) out->constraints.push_back(x);
    if (strict && !x.is_zero()) {
      std::ostringstream s;
      s << "constraint " << n_constraints << " failed at row " << row << ": " << x.to_u32() << " != 0";
      throw std::runtime_error(s.str());
    }
    n_constraints++;
  }
  bool real(Fp is_real) {
    if (!strict) return true;
    if (is_real.is_zero()) return false;
    if (is_real != Fp::one()) throw std::runtime_error("lookup multiplicity is not boolean");
    return true;
  }
  void receive(const std::vector<Fp>& m, Fp is_real) {
    if (out) out->receives.push_back({m, is_real});
    if (q && strict && !is_real.is_zero()) q->receive(m);
  }
  void send(const std::vector<Fp>& m, Fp is_real) {
    if (out) out->sends.push_back({m, is_real});
    if (q && strict && !is_real.is_zero()) q->send(m);
  }
  static std::vector<Fp> chain(Fp a, Fp b, const std::vector<Fp>& rel) {
    std::vector<Fp> m{a, b};
    m.insert(m.end(), rel.begin(), rel.end());
    return m;
  }
  // builder.rs:42-72 (+ debug.rs:246-271)
  void provide(const std::vector<Fp>& rel, Fp last_nonce, Fp last_count, Fp is_real) {
    if (!real(is_real)) {
      if (out) { out->receives.push_back({chain(last_nonce, last_count, rel), is_real}); out->sends.push_back({chain(Fp::zero(), Fp::zero(), rel), is_real}); }
      return;
    }
    receive(chain(last_nonce, last_count, rel), is_real);
    send(chain(Fp::zero(), Fp::zero(), rel), is_real);
    if (q && strict) q->access(rel, 0, MemoRec{last_nonce.to_u32(), last_count.to_u32(), 0});
  }
  // builder.rs:75-104 (+ debug.rs:273-303)
  void require(const std::vector<Fp>& rel, Fp nonce, Fp prev_nonce, Fp prev_count, Fp count_inv, Fp is_real) {
    Fp count = prev_count + Fp::one();
    assert_zero(is_real * (count * count_inv - Fp::one()));
    if (!real(is_real)) {
      if (out) { out->receives.push_back({chain(prev_nonce, prev_count, rel), is_real}); out->sends.push_back({chain(nonce, count, rel), is_real}); }
      return;
    }
    receive(chain(prev_nonce, prev_count, rel), is_real);
    send(chain(nonce, count, rel), is_real);
    if (q && strict) q->access(rel, count.to_u32(), MemoRec{prev_nonce.to_u32(), prev_count.to_u32(), nonce.to_u32()});
  }
};

// FilteredAirBuilder: `builder.when(c)` multiplies every asserted value by c; nests.
struct When {
  DebugBuilder& b;
  Fp c;
  When when(Fp c2) const { return When{b, c * c2}; }
  void assert_zero(Fp x) const { b.assert_zero(c * x); }
  void assert_eq(Fp x, Fp y) const { b.assert_zero(c * (x - y)); }
  void assert_one(Fp x) const { b.assert_zero(c * (x - Fp::one())); }
  void assert_bool(Fp x) const { b.assert_zero(c * (x * (x - Fp::one()))); }
};
static inline When top(DebugBuilder& b) { return When{b, Fp::one()}; }

static inline Fp fu(uint32_t x) { return Fp::from_u32(x); }

// ---- byte relations --------------------------------------------------------------------------------
struct ByteRel { uint32_t tag; std::vector<Fp> vals; Fp is_real; };   // bytes/relation.rs:7,72-81,121-135
struct BytesAirRecord {  // bytes/builder.rs:17-90
  std::vector<ByteRel> records;
  void range_check_u8_pair(Fp i1, Fp i2, Fp r) { records.push_back({1, {i1, i2}, r}); }
  void range_check_u8_iter(const std::vector<Fp>& xs, Fp r) {  // bytes/mod.rs:136-146
    for (size_t i = 0; i < xs.size(); i += 2) range_check_u8_pair(xs[i], i + 1 < xs.size() ? xs[i + 1] : Fp::zero(), r);
  }
  void range_check_u16(Fp i, Fp r) { records.push_back({2, {i}, r}); }
  void less_than(Fp i1, Fp i2, Fp lt, Fp r) { records.push_back({3, {i1, i2, lt}, r}); }
  static std::vector<Fp> values(const ByteRel& r) {
    std::vector<Fp> v{fu(3), fu(r.tag)};
    v.insert(v.end(), r.vals.begin(), r.vals.end());
    return v;
  }
};
struct Require { Fp prev_nonce, prev_count, count_inv; };
static inline void require_all(BytesAirRecord& rec, DebugBuilder& b, Fp nonce, const std::vector<Require>& reqs) {
  if (rec.records.size() != reqs.size()) throw std::runtime_error("zip_eq: byte records and requires differ in length");
  for (size_t i = 0; i < reqs.size(); i++)
    b.require(BytesAirRecord::values(rec.records[i]), nonce, reqs[i].prev_nonce, reqs[i].prev_count, reqs[i].count_inv,
              rec.records[i].is_real);
}

// ---- gadgets (src/gadgets/unsigned, src/gadgets/big_num) ------------------------------------------------
typedef std::vector<Fp> Word;

// less_than.rs:44-101: witness = is_comp[W], lhs_comp_limb, rhs_comp_limb
static inline void assert_less_than(const Fp* wit, size_t W, DebugBuilder& b, const Word& lhs, const Word& rhs,
                                    BytesAirRecord& rec, Fp is_real) {
  When w{b, is_real};
  Fp is_equal = Fp::zero();
  for (size_t i = 0; i < W; i++) {
    if (i > 0) w.when(is_equal).assert_eq(lhs[i], rhs[i]);
    w.assert_bool(wit[i]);
    is_equal += wit[i];
  }
  w.assert_one(is_equal);
  Fp sl = Fp::zero(), sr = Fp::zero();
  for (size_t i = 0; i < W; i++) sl += lhs[i] * wit[i], sr += rhs[i] * wit[i];
  w.assert_eq(sl, wit[W]);
  w.assert_eq(sr, wit[W + 1]);
  rec.less_than(wit[W], wit[W + 1], Fp::one(), is_real);
}
struct CompareResult { Fp is_less_than, is_equal; };
// cmp.rs:50-127: witness = is_comp[W], lhs_comp_limb, rhs_comp_limb, comp_diff_inv, is_less_than
static inline CompareResult compare_eval(const Fp* wit, size_t W, DebugBuilder& b, const Word& lhs, const Word& rhs,
                                         BytesAirRecord& rec, Fp is_real) {
  When w{b, is_real};
  Fp is_equal = Fp::one();
  for (size_t i = W; i-- > 0;) {
    w.assert_bool(wit[i]);
    is_equal -= wit[i];
    w.when(is_equal).assert_eq(lhs[i], rhs[i]);
  }
  w.assert_bool(is_equal);
  Fp sl = Fp::zero(), sr = Fp::zero();
  for (size_t i = 0; i < W; i++) sl += lhs[i] * wit[i], sr += rhs[i] * wit[i];
  w.assert_eq(sl, wit[W]);
  w.assert_eq(sr, wit[W + 1]);
  Fp is_different = Fp::one() - is_equal;
  w.assert_eq((wit[W] - wit[W + 1]) * wit[W + 2], is_different);
  rec.less_than(wit[W], wit[W + 1], wit[W + 3], is_real);
  return {wit[W + 3], is_equal};
}
// add.rs:17-38
static inline Fp assert_add(DebugBuilder& b, const Word& lhs, const Word& rhs, const Word& out, Fp is_real) {
  When w{b, is_real};
  Fp base_inv = fp_inv(fu(256));
  Fp carry = Fp::zero();
  for (size_t i = 0; i < out.size(); i++) {
    Fp sum = lhs[i] + rhs[i] + carry;
    carry = (sum - out[i]) * base_inv;
    w.assert_bool(carry);
  }
  return carry;
}
static inline Word into_checked(const Fp* wit, size_t W, BytesAirRecord& rec, Fp is_real) {  // unsigned/mod.rs:123-135
  Word v(wit, wit + W);
  rec.range_check_u8_iter(v, is_real);
  return v;
}
// mul.rs:59-96 (witness carry[W]) and :113-137 (Product = MulWitness + result)
static inline Word product_eval(const Fp* wit, size_t W, DebugBuilder& b, const Word& lhs, const Word& rhs,
                                BytesAirRecord& rec, Fp is_real) {
  const Fp* carry = wit;
  Word out(wit + W, wit + 2 * W);
  Fp base = fu(256);
  std::vector<Fp> products(W, Fp::zero());
  for (size_t i = 0; i < lhs.size(); i++)
    for (size_t j = 0; j < rhs.size(); j++)
      if (i + j < W) products[i + j] += lhs[i] * rhs[j];
  Fp carry_prev = Fp::zero();
  for (size_t k = 0; k < W; k++) {
    rec.range_check_u16(carry[k], is_real);
    Fp o = out[k] + carry[k] * base;
    When{b, is_real}.assert_eq(products[k] + carry_prev, o);
    carry_prev = carry[k];
  }
  return into_checked(wit + W, W, rec, is_real);
}
// is_zero.rs:46-62
static inline void assert_non_zero(const Fp* inverses, DebugBuilder& b, const Word& input, Fp is_real) {
  Fp lc = Fp::zero();
  for (size_t i = 0; i < input.size(); i++) lc += input[i] * inverses[i];
  When{b, is_real}.assert_one(lc);
}
// is_zero.rs:64-85 and :139-155 (IsZero = inverses[W], result)
static inline Fp eval_is_zero(const Fp* wit, size_t W, DebugBuilder& b, const Word& input, Fp is_real) {
  Fp is_zero = wit[W];
  When w{b, is_real};
  w.assert_bool(is_zero);
  Fp lc = Fp::zero();
  for (size_t i = 0; i < W; i++) {
    w.when(is_zero).assert_zero(input[i]);
    lc += input[i] * wit[i];
  }
  w.assert_eq(lc, Fp::one() - is_zero);
  return is_zero;
}
// field.rs:38-84 (FieldWitness::assert_eq) and :112-140 (FieldToWord32 = is_msb_less_than, result[4])
static inline Word field_to_word32(const Fp* wit, DebugBuilder& b, Fp field, BytesAirRecord& rec, Fp is_real) {
  Fp is_msb_lt = wit[0];
  Word word(wit + 1, wit + 5);
  Fp base = fu(256), msb = fu(0x78);
  When w{b, is_real};
  w.assert_bool(is_msb_lt);
  Fp recomposed = Fp::zero();
  for (int i = 3; i >= 0; i--) recomposed = recomposed * base + word[i];
  w.assert_eq(field, recomposed);
  rec.less_than(word[3], msb, is_msb_lt, is_real);
  When we{b, is_real * (Fp::one() - is_msb_lt)};
  we.assert_eq(word[3], msb);
  for (int i = 0; i < 3; i++) we.assert_eq(word[i], Fp::zero());
  return into_checked(wit + 1, 4, rec, is_real);
}
// big_num/cmp.rs:50-115: is_comp[8], lhs_comp_limb, rhs_comp_limb, 2 x FieldToWord32 (5), CompareWitness<4> (8)
static inline CompareResult bignum_compare_eval(const Fp* wit, DebugBuilder& b, const Word& lhs, const Word& rhs,
                                                BytesAirRecord& rec, Fp is_real) {
  When w{b, is_real};
  Fp is_equal = Fp::one();
  for (int i = 7; i >= 0; i--) {
    w.assert_bool(wit[i]);
    is_equal -= wit[i];
    w.when(is_equal).assert_eq(lhs[i], rhs[i]);
  }
  w.assert_bool(is_equal);
  Fp sl = Fp::zero(), sr = Fp::zero();
  for (int i = 0; i < 8; i++) sl += lhs[i] * wit[i], sr += rhs[i] * wit[i];
  w.assert_eq(sl, wit[8]);
  w.assert_eq(sr, wit[9]);
  Word lw = field_to_word32(wit + 10, b, wit[8], rec, is_real);
  Word rw = field_to_word32(wit + 15, b, wit[9], rec, is_real);
  CompareResult r = compare_eval(wit + 20, 4, b, lw, rw, rec, is_real);
  w.assert_eq(is_equal, r.is_equal);
  return r;
}

// ---- Poseidon2 wide AIR (poseidon/wide/air.rs:11-124) ----------------------------------------------------
static inline void poseidon2_wide_eval(int t, const Fp* cols, DebugBuilder& b, const Word& input, const Fp* output,
                                       Fp is_real) {
  const P2Consts& c = p2_consts(t);
  const Fp* ext_state = cols;
  const Fp* ext_sbox = cols + 8 * t;
  const Fp* int_init = cols + 16 * t;
  const Fp* int_s0 = cols + 17 * t;
  const Fp* int_sbox = cols + 17 * t + (c.rp - 1);
  When w = top(b);
  std::vector<Fp> state(t);
  for (int i = 0; i < t; i++) state[i] = is_real * input[i];
  p2_external_linear(state.data(), t);
  auto external_round = [&](int round) {
    for (int i = 0; i < t; i++) {
      w.assert_eq(state[i], ext_state[round * t + i]);
      state[i] = ext_state[round * t + i];
    }
    for (int i = 0; i < t; i++) state[i] += is_real * c.full_rc[round * t + i];
    for (int i = 0; i < t; i++) {
      Fp s3 = ext_sbox[round * t + i];
      w.assert_eq(state[i] * state[i] * state[i], s3);
      state[i] *= s3 * s3;
    }
    p2_external_linear(state.data(), t);
  };
  for (int r = 0; r < 4; r++) external_round(r);
  for (int r = 0; r < c.rp; r++) {
    if (r == 0) {
      for (int i = 0; i < t; i++) {
        w.assert_eq(state[i], int_init[i]);
        state[i] = int_init[i];
      }
    } else {
      w.assert_eq(state[0], int_s0[r - 1]);
      state[0] = int_s0[r - 1];
    }
    state[0] += is_real * c.part_rc[r];
    Fp s3 = int_sbox[r];
    w.assert_eq(state[0] * state[0] * state[0], s3);
    state[0] *= s3 * s3;
    p2_internal_linear(state.data(), c);
  }
  for (int r = 4; r < 8; r++) external_round(r);
  for (int i = 0; i < 8; i++) w.assert_eq(state[i], is_real * output[i]);
}

// Chipset::eval for every chip of lurk_chip_map_native (src/core/chipset.rs:34-52); kinds as in lair.h:
// 1..3 hasher3/4/5, 4 u64 add, 5 sub, 6 less-than, 7 is-zero, 8 mul, 9 div-rem, 10 big-num less-than.
static inline std::vector<Fp> chipset_eval(const ChipDef& c, DebugBuilder& b, Fp is_real, const Word& ins, const Fp* wit,
                                           Fp nonce, const std::vector<Require>& requires_) {
  if (c.kind >= 1 && c.kind <= 3) {  // core/poseidon.rs:79-93
    int t = chip_width(c.kind);
    poseidon2_wide_eval(t, wit + 8, b, ins, wit, is_real);
    return std::vector<Fp>(wit, wit + 8);
  }
  BytesAirRecord rec;
  std::vector<Fp> out;
  if (c.kind == 10) {  // core/big_num.rs:75-97
    Word in1(ins.begin(), ins.begin() + 8), in2(ins.begin() + 8, ins.begin() + 16);
    out = {bignum_compare_eval(wit, b, in1, in2, rec, is_real).is_less_than};
  } else if (c.kind >= 4 && c.kind <= 9) {  // core/u64.rs:170-222
    Word in1(ins.begin(), ins.begin() + 8), in2(8, Fp::zero());
    if (c.kind != 7) in2.assign(ins.begin() + 8, ins.begin() + 16);
    switch (c.kind) {
      case 4: {  // Sum::eval, add.rs:59-79
        Word r = into_checked(wit, 8, rec, is_real);
        assert_add(b, in1, in2, r, is_real);
        out = r;
      } break;
      case 5: {  // Diff::eval, add.rs:117-135
        Word r = into_checked(wit, 8, rec, is_real);
        assert_add(b, r, in2, in1, is_real);
        out = r;
      } break;
      case 6: out = {compare_eval(wit, 8, b, in1, in2, rec, is_real).is_less_than}; break;
      case 7: out = {eval_is_zero(wit, 8, b, in1, is_real)}; break;
      case 8: out = product_eval(wit, 8, b, in1, in2, rec, is_real); break;
      default: {  // DivRem::eval, div_rem.rs:52-90: b_non_zero[8], q[8], qb (16), r[8], r_lt_b (10), qb_cmp_a (12)
        assert_non_zero(wit, b, in2, is_real);
        Word q = into_checked(wit + 8, 8, rec, is_real);
        Word qb = product_eval(wit + 16, 8, b, q, in2, rec, is_real);
        Word r = into_checked(wit + 32, 8, rec, is_real);  // Diff::eval(a, qb)
        assert_add(b, r, qb, in1, is_real);
        assert_less_than(wit + 40, 8, b, r, in2, rec, is_real);
        CompareResult cr = compare_eval(wit + 50, 8, b, qb, in1, rec, is_real);
        When{b, is_real}.assert_one(cr.is_less_than + cr.is_equal);
        out = q;
        out.insert(out.end(), r.begin(), r.end());
      }
    }
  } else {
    throw std::runtime_error("unknown extern chip kind");
  }
  require_all(rec, b, nonce, requires_);
  return out;
}

// ---- Func::eval (src/lair/air.rs:158-566) -----------------------------------------------------------------
struct FuncAir {
  const Toplevel& tl;
  const FuncDef& f;
  DebugBuilder& b;
  const Fp *nonce, *input, *output, *aux, *sel;   // ColumnSlice::from_slice, air.rs:40-55
  size_t i_aux = 0, i_out = 0;                    // ColumnIndex
  struct Val { Fp v; uint32_t deg; };             // degree as tracked by the layout (func_chip.rs:183-277)
  std::vector<Val> map;
  Word depth;

  FuncAir(const Toplevel& t, const FuncDef& fd, DebugBuilder& bd) : tl(t), f(fd), b(bd) {
    nonce = b.local;
    input = nonce + 1;
    output = input + f.input_size;
    aux = output + f.output_size;
    sel = aux + f.aux;
  }
  Fp next_aux() {
    if (i_aux >= f.aux) throw std::runtime_error("aux columns exhausted");
    return aux[i_aux++];
  }
  const Fp* next_n_aux(size_t n) {
    if (i_aux + n > f.aux) throw std::runtime_error("aux columns exhausted");
    const Fp* p = aux + i_aux;
    i_aux += n;
    return p;
  }
  Require next_require() {
    Fp a = next_aux(), c = next_aux(), d = next_aux();
    return {a, c, d};
  }
  Fp return_sel(const Block& blk) const {
    Fp s = Fp::zero();
    for (uint32_t i : blk.return_idents) s += sel[i];
    return s;
  }
  static std::vector<Fp> call_relation(uint32_t idx, const Word& inp, const Word& out) {  // relations.rs:13-31
    std::vector<Fp> v{fu(0), fu(idx)};
    v.insert(v.end(), inp.begin(), inp.end());
    v.insert(v.end(), out.begin(), out.end());
    return v;
  }
  static std::vector<Fp> memory_relation(Fp ptr, const Word& vals) {  // relations.rs:45-59
    std::vector<Fp> v{fu(1), ptr};
    v.insert(v.end(), vals.begin(), vals.end());
    return v;
  }

  void eval() {  // air.rs:158-235
    top(b).when(b.is_transition).assert_eq(b.next[0], *nonce + Fp::one());
    Word call_inp;
    for (uint32_t i = 0; i < f.input_size; i++) {
      map.push_back({input[i], 1});
      call_inp.push_back(input[i]);
    }
    Fp toplevel_sel = return_sel(f.body);
    top(b).assert_bool(toplevel_sel);
    Fp last_nonce = next_aux(), last_count = next_aux();
    Word out(output, output + f.output_size);
    if (f.partial) {
      Word d;
      for (int i = 0; i < 4; i++) d.push_back(next_aux());
      std::vector<Require> reqs{next_require(), next_require()};
      BytesAirRecord rec;
      rec.range_check_u8_iter(d, toplevel_sel);
      require_all(rec, b, *nonce, reqs);
      out.insert(out.end(), d.begin(), d.end());
      depth = d;
    }
    b.provide(call_relation(f.index, call_inp, out), last_nonce, last_count, toplevel_sel);
    eval_block(f.body, toplevel_sel);
  }

  void eval_depth(Fp s, Word& out) {  // air.rs:105-131
    Word dep;
    for (int i = 0; i < 4; i++) dep.push_back(next_aux());
    const Fp* wit = next_n_aux(6);
    BytesAirRecord rec;
    assert_less_than(wit, 4, b, dep, depth, rec, s);
    std::vector<Require> reqs{next_require()};
    require_all(rec, b, *nonce, reqs);
    out.insert(out.end(), dep.begin(), dep.end());
  }

  void eval_block(const Block& blk, Fp s) {  // air.rs:249-273
    for (const OpNode& op : blk.ops) eval_op(op, s);
    eval_ctrl(blk.ctrl);
  }

  void eval_op(const OpNode& op, Fp s) {  // air.rs:276-492
    When w{b, s};
    switch (op.code) {
      case OP_ASSERT_NE: {  // + constrain_inequality_witness, air.rs:554-566
        Fp acc = Fp::zero();
        std::vector<Fp> coeffs;
        for (size_t i = 0; i < op.xs.size(); i++) coeffs.push_back(next_aux());
        for (size_t i = 0; i < op.xs.size(); i++) acc += coeffs[i] * (map[op.xs[i]].v - map[op.ys[i]].v);
        w.assert_one(acc);
      } break;
      case OP_ASSERT_EQ:
        for (size_t i = 0; i < op.xs.size(); i++) w.assert_eq(map[op.xs[i]].v, map[op.ys[i]].v);
        break;
      case OP_CONTAINS: {
        Fp bv = map[op.b].v;
        Fp acc = map[op.xs[0]].v - bv;
        for (size_t i = 1; i < op.xs.size(); i++) {
          Fp a = next_aux();
          acc = acc * (map[op.xs[i]].v - bv);
          w.assert_eq(acc, a);
          acc = a;
        }
        w.assert_zero(acc);
      } break;
      case OP_CONST: map.push_back({op.cst, 0}); break;
      case OP_ADD: map.push_back({map[op.a].v + map[op.b].v, std::max(map[op.a].deg, map[op.b].deg)}); break;
      case OP_SUB: map.push_back({map[op.a].v - map[op.b].v, std::max(map[op.a].deg, map[op.b].deg)}); break;
      case OP_MUL: {
        // A witness column exists iff the layout gave it one: deg(a) + deg(b) >= 2 (func_chip.rs:202-210, trace.rs:291-302).
        // air.rs:345-358 keys on "both constant" instead; the two rules coincide for every program whose trace the
        // reference can generate and check, since a column air.rs reads that trace.rs never wrote would misalign the row.
        Val x = map[op.a], y = map[op.b];
        uint32_t deg = x.deg + y.deg;
        if (deg < 2) {
          map.push_back({x.v * y.v, deg});
        } else {
          Fp c = next_aux();
          w.assert_eq(x.v * y.v, c);
          map.push_back({c, 1});
        }
      } break;
      case OP_INV: {
        Val x = map[op.a];
        if (x.deg == 0) {
          map.push_back({fp_inv(x.v), 0});
        } else {
          Fp c = next_aux();
          w.assert_one(x.v * c);
          map.push_back({c, 1});
        }
      } break;
      case OP_NOT: {
        Val x = map[op.a];
        if (x.deg == 0) {
          map.push_back({x.v.is_zero() ? Fp::one() : Fp::zero(), 0});
        } else {
          Fp d = next_aux(), nx = next_aux();
          w.assert_zero(x.v * nx);
          w.assert_one(x.v * d + nx);
          map.push_back({nx, 1});
        }
      } break;
      case OP_CALL: {
        const FuncDef& callee = tl.funcs[op.a];
        Word out;
        for (uint32_t i = 0; i < callee.output_size; i++) {
          Fp o = next_aux();
          map.push_back({o, 1});
          out.push_back(o);
        }
        Word inp;
        for (uint32_t x : op.xs) inp.push_back(map[x].v);
        Require r = next_require();
        if (callee.partial) eval_depth(s, out);
        b.require(call_relation(op.a, inp, out), *nonce, r.prev_nonce, r.prev_count, r.count_inv, s);
      } break;
      case OP_PREIMG: {
        const FuncDef& callee = tl.funcs[op.a];
        Word inp;
        for (uint32_t i = 0; i < callee.input_size; i++) {
          Fp v = next_aux();
          map.push_back({v, 1});
          inp.push_back(v);
        }
        Word out;
        for (uint32_t x : op.xs) out.push_back(map[x].v);
        Require r = next_require();
        if (callee.partial) eval_depth(s, out);
        b.require(call_relation(op.a, inp, out), *nonce, r.prev_nonce, r.prev_count, r.count_inv, s);
      } break;
      case OP_STORE: {
        Fp ptr = next_aux();
        map.push_back({ptr, 1});
        Word vals;
        for (uint32_t x : op.xs) vals.push_back(map[x].v);
        Require r = next_require();
        b.require(memory_relation(ptr, vals), *nonce, r.prev_nonce, r.prev_count, r.count_inv, s);
      } break;
      case OP_LOAD: {
        Fp ptr = map[op.b].v;
        Word vals;
        for (uint32_t i = 0; i < op.a; i++) {
          Fp o = next_aux();
          map.push_back({o, 1});
          vals.push_back(o);
        }
        Require r = next_require();
        b.require(memory_relation(ptr, vals), *nonce, r.prev_nonce, r.prev_count, r.count_inv, s);
      } break;
      case OP_EXTERN: {
        const ChipDef& c = tl.chips[op.a];
        Word in;
        for (uint32_t x : op.xs) in.push_back(map[x].v);
        const Fp* wit = next_n_aux(c.witness_size);
        std::vector<Require> reqs;
        for (uint32_t i = 0; i < c.require_size; i++) reqs.push_back(next_require());
        for (Fp o : chipset_eval(c, b, s, in, wit, *nonce, reqs)) map.push_back({o, 1});
      } break;
      case OP_RANGE_U8: {
        std::vector<Require> reqs;
        for (size_t i = 0; i < (op.xs.size() + 1) / 2; i++) reqs.push_back(next_require());
        BytesAirRecord rec;
        Word xs;
        for (uint32_t x : op.xs) xs.push_back(map[x].v);
        rec.range_check_u8_iter(xs, s);
        require_all(rec, b, *nonce, reqs);
      } break;
      default: throw std::runtime_error("bad op");
    }
  }

  void eval_ctrl(const CtrlNode& c) {  // air.rs:495-551
    if (c.kind == CTRL_RETURN) {
      Fp s = sel[c.ident];
      for (uint32_t v : c.vars) {
        if (i_out >= f.output_size) throw std::runtime_error("output columns exhausted");
        Fp out_var = output[i_out++];
        When{b, s}.assert_eq(map[v].v, out_var);
      }
      return;
    }
    size_t map_len = map.size(), s_aux = i_aux, s_out = i_out;
    auto process = [&](const Block& blk) {
      Fp s = return_sel(blk);
      eval_block(blk, s);
      map.resize(map_len);
      i_aux = s_aux;
      i_out = s_out;
    };
    for (auto& br : c.branches) process(*br.block);
    if (c.dflt) process(*c.dflt);
  }
};

// ---- the other chips -----------------------------------------------------------------------------------------
static inline void mem_chip_eval(uint32_t len, DebugBuilder& b) {  // memory.rs:71-109
  const Fp *l = b.local, *n = b.next;
  Fp is_real = l[0], ptr_local = l[1], last_nonce = l[2], last_count = l[3];
  When w = top(b);
  w.assert_bool(is_real);
  Fp is_real_transition = n[0] * b.is_transition;
  w.when(is_real_transition).assert_one(is_real);
  w.when(b.is_first).when(is_real).assert_one(ptr_local);
  w.when(is_real_transition).assert_eq(ptr_local + Fp::one(), n[1]);
  b.provide(FuncAir::memory_relation(ptr_local, Word(l + 4, l + 4 + len)), last_nonce, last_count, is_real);
}
static inline void bytes_chip_eval(DebugBuilder& b) {  // bytes/trace.rs:116-143; main = is_real, 6 x (last_nonce, last_count)
  const Fp *p = b.prep_local, *m = b.local;
  top(b).assert_bool(m[0]);
  Fp u16v = p[0] + p[1] * fu(256);
  std::vector<ByteRel> rels = {{1, {p[0], p[1]}, m[0]}, {2, {u16v}, m[0]},          {3, {p[0], p[1], p[2]}, m[0]},
                               {4, {p[0], p[1], p[3]}, m[0]}, {5, {p[0], p[1], p[4]}, m[0]}, {6, {p[0], p[1], p[5]}, m[0]}};
  for (int k = 0; k < 6; k++) b.provide(BytesAirRecord::values(rels[k]), m[1 + 2 * k], m[2 + 2 * k], m[0]);
}
static inline void entry_chip_eval(uint32_t func_idx, uint32_t num_pv, DebugBuilder& b) {  // lair_chip.rs:166-191
  Word pvs(b.local, b.local + num_pv);
  for (uint32_t i = 0; i < num_pv; i++) top(b).assert_eq(pvs[i], b.pv[i]);
  std::vector<Fp> rel{fu(0), fu(func_idx)};  // OuterCallRelation, relations.rs:33-43
  rel.insert(rel.end(), pvs.begin(), pvs.end());
  b.require(rel, Fp::zero(), Fp::zero(), Fp::zero(), Fp::one(), Fp::one());
}

enum { AIR_FUNC = 0, AIR_MEM = 1, AIR_BYTES = 2, AIR_ENTRY = 3 };
// One row of one chip.  kind / index as in the chip vector (build_lair_chip_vector, lair_chip.rs:196-231):
// func index, memory length, -, entry func index.
static inline void air_eval_row(const Toplevel& tl, uint32_t kind, uint32_t index, uint32_t num_pv, DebugBuilder& b) {
  switch (kind) {
    case AIR_FUNC: {
      FuncAir fa(tl, tl.funcs.at(index), b);
      fa.eval();
    } break;
    case AIR_MEM: mem_chip_eval(index, b); break;
    case AIR_BYTES: bytes_chip_eval(b); break;
    case AIR_ENTRY: entry_chip_eval(index, num_pv, b); break;
    default: throw std::runtime_error("bad chip kind");
  }
}

// debug_constraints_collecting_queries (debug.rs:161-206): every row of a trace, strict mode.
static inline void air_check_trace(const Toplevel& tl, uint32_t kind, uint32_t index, const std::vector<Fp>& main, size_t h,
                                   size_t w, const std::vector<Fp>* prep, size_t pw, const std::vector<Fp>& pv,
                                   TraceQueries& q) {
  for (size_t r = 0; r < h; r++) {
    size_t rn = (r + 1) % h;
    DebugBuilder b;
    b.local = &main[r * w];
    b.next = &main[rn * w];
    b.prep_local = prep ? &(*prep)[r * pw] : nullptr;
    b.prep_next = prep ? &(*prep)[rn * pw] : nullptr;
    b.pv = pv.data();
    b.is_first = r == 0 ? Fp::one() : Fp::zero();
    b.is_last = r == h - 1 ? Fp::one() : Fp::zero();
    b.is_transition = r != h - 1 ? Fp::one() : Fp::zero();
    b.row = r;
    b.q = &q;
    air_eval_row(tl, kind, index, (uint32_t)pv.size(), b);
  }
}

}  // namespace orc
