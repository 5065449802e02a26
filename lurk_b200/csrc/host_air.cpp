// Symbolic capture of every chip's AIR into the flat chip descriptor ('MACH' blob, format in machine.h).
//
// In the reference each chip's `Air::eval` is executed by sphinx-core with different builders: once
// symbolically in `Chip::new` to collect the lookup interactions (call sites src/lair/lair_chip.rs:222-231)
// and then per row of the quotient domain with a folding builder.  Here `eval` runs once, symbolically,
// and is compiled into (a) a register-allocated straight-line constraint program over base-field values
// and (b) the send / receive interaction list with affine column combinations -- both interpreted by the
// CUDA quotient / LogUp kernels, by the verifier (verify.cu) and by the CPU oracle.
//
// Mirrored reference code:
//   Func::eval / Op::eval / Ctrl::eval, eval_depth   src/lair/air.rs:105-566
//   MemChip::eval                                    src/lair/memory.rs:71-109
//   BytesChip::eval                                  src/gadgets/bytes/trace.rs:116-143
//   Entrypoint eval                                  src/lair/lair_chip.rs:166-191
//   LookupBuilder::{provide, require}                src/air/builder.rs:42-104
//   relations                                        src/lair/relations.rs:6-59, src/gadgets/bytes/relation.rs:7-135
//   LessThanWitness::assert_less_than                src/gadgets/unsigned/less_than.rs:49-103
//   Poseidon2Cols::eval, PoseidonChipset::eval       src/poseidon/wide/air.rs:11-124, src/core/poseidon.rs:78-93
//   build_lair_chip_vector                           src/lair/lair_chip.rs:196-215
// Third-party behaviour assumed (sphinx-core): interactions are InteractionKind::Memory
// (argument index 1), `when(c).assert_zero(x)` is the constraint c * x,
// log_quotient_degree = log2_ceil(max(max_constraint_degree, 3) - 1).
// lurk_b200/air.py is the same pass in Python; tests/test_host_cpp.py holds the two to identical output.
#include <algorithm>
#include <stdexcept>

#include "host_lair.h"
#include "poseidon2_constants.h"

namespace lurk {
namespace host {
namespace {

constexpr uint32_t Pm = FIELD_P;
constexpr uint32_t CALL_TAG = 0, MEMORY_TAG = 1, BYTE_TAG = 3, INTERACTION_KIND_MEMORY = 1;
enum : int { L_CONST = 0, L_MAIN, L_MAIN_NEXT, L_PREP, L_PREP_NEXT, L_PV, L_SPECIAL };
enum : int { SP_IS_FIRST = 0, SP_IS_LAST = 1, SP_IS_TRANSITION = 2 };

inline uint32_t mulp(uint64_t a, uint64_t b) { return (uint32_t)(a * b % Pm); }
inline uint32_t addp(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + b) % Pm); }
inline uint32_t negp(uint32_t a) { return a ? Pm - a : 0; }

// Symbolic polynomial over trace cells: affine expressions in normal form, everything else a DAG node
using Key = std::pair<int, int>;
using E = int;  // handle into Graph::nodes
struct Node {
  enum Kind { Aff, Add, Sub, Mul, Neg } kind = Aff;
  E a = -1, b = -1;
  std::map<Key, uint32_t> lin;
  uint32_t cst = 0, degree = 0;
};
struct Graph {
  std::vector<Node> nodes;
  E make(Node n) {
    nodes.push_back(std::move(n));
    return (E)nodes.size() - 1;
  }
  bool is_affine(E e) const { return nodes[e].kind == Node::Aff; }
  bool is_const(E e) const { return nodes[e].kind == Node::Aff && nodes[e].lin.empty(); }
  E affine(std::map<Key, uint32_t> lin, uint32_t cst) {
    Node n;
    n.degree = lin.empty() ? 0 : 1;
    n.lin = std::move(lin);
    n.cst = cst;
    return make(std::move(n));
  }
  E cst(uint32_t c) { return affine({}, c % Pm); }
  E leaf(int kind, int idx) { return affine({{Key(kind, idx), 1u}}, 0); }
  E op(Node::Kind k, E a, E b, uint32_t degree) {
    Node n;
    n.kind = k; n.a = a; n.b = b; n.degree = degree;
    return make(std::move(n));
  }
  E add(E x, E y) {
    if (is_affine(x) && is_affine(y)) {
      std::map<Key, uint32_t> lin = nodes[x].lin;
      for (auto& kv : nodes[y].lin) {
        uint32_t v = addp(lin.count(kv.first) ? lin[kv.first] : 0, kv.second);
        if (v) lin[kv.first] = v;
        else lin.erase(kv.first);
      }
      return affine(std::move(lin), addp(nodes[x].cst, nodes[y].cst));
    }
    return op(Node::Add, x, y, std::max(nodes[x].degree, nodes[y].degree));
  }
  E neg(E x) {
    if (is_affine(x)) {
      std::map<Key, uint32_t> lin;
      for (auto& kv : nodes[x].lin) lin[kv.first] = negp(kv.second);
      return affine(std::move(lin), negp(nodes[x].cst));
    }
    return op(Node::Neg, x, -1, nodes[x].degree);
  }
  E sub(E x, E y) {
    if (is_affine(x) && is_affine(y)) return add(x, neg(y));
    return op(Node::Sub, x, y, std::max(nodes[x].degree, nodes[y].degree));
  }
  E mul(E x, E y) {
    if (is_const(y)) std::swap(x, y);
    if (is_const(x)) {
      uint32_t c = nodes[x].cst;
      if (c == 0) return cst(0);
      if (c == 1) return y;
      if (is_affine(y)) {
        std::map<Key, uint32_t> lin;
        for (auto& kv : nodes[y].lin) lin[kv.first] = mulp(kv.second, c);
        return affine(std::move(lin), mulp(nodes[y].cst, c));
      }
    }
    return op(Node::Mul, x, y, nodes[x].degree + nodes[y].degree);
  }
};

struct Interaction {
  std::vector<E> values;
  E mult;
  uint32_t kind = INTERACTION_KIND_MEMORY;
};

// sphinx's interaction-collecting symbolic builder (`Chip::new`) and the constraint folder at once
struct Builder {
  Graph g;
  uint32_t width, prep_width, num_pv;
  std::vector<E> constraints, cond;
  std::vector<Interaction> sends, receives;
  Builder(uint32_t w, uint32_t pw, uint32_t npv) : width(w), prep_width(pw), num_pv(npv) {}
  E local(int c) { return g.leaf(L_MAIN, c); }
  E next(int c) { return g.leaf(L_MAIN_NEXT, c); }
  E prep(int c) { return g.leaf(L_PREP, c); }
  E public_value(int i) { return g.leaf(L_PV, i); }
  E is_first_row() { return g.leaf(L_SPECIAL, SP_IS_FIRST); }
  E is_transition() { return g.leaf(L_SPECIAL, SP_IS_TRANSITION); }
  E k(uint32_t c) { return g.cst(c); }
  struct When {
    Builder& b;
    When(Builder& bb, E c) : b(bb) { b.cond.push_back(c); }
    ~When() { b.cond.pop_back(); }
  };
  void assert_zero(E x) {
    for (size_t i = cond.size(); i-- > 0;) x = g.mul(cond[i], x);
    constraints.push_back(x);
  }
  void assert_eq(E a, E b) { assert_zero(g.sub(a, b)); }
  void assert_one(E a) { assert_zero(g.sub(a, k(1))); }
  void assert_bool(E a) { assert_zero(g.mul(a, g.sub(a, k(1)))); }
  // LookupBuilder (src/air/builder.rs:34-133)
  void provide(std::vector<E> relation, E last_nonce, E last_count, E is_real) {
    std::vector<E> r{last_nonce, last_count};
    r.insert(r.end(), relation.begin(), relation.end());
    receives.push_back({r, is_real});
    std::vector<E> s{k(0), k(0)};
    s.insert(s.end(), relation.begin(), relation.end());
    sends.push_back({s, is_real});
  }
  void require(std::vector<E> relation, E nonce, E prev_nonce, E prev_count, E count_inv, E is_real) {
    E count = g.add(prev_count, k(1));
    {
      When w(*this, is_real);
      assert_one(g.mul(count, count_inv));
    }
    std::vector<E> r{prev_nonce, prev_count};
    r.insert(r.end(), relation.begin(), relation.end());
    receives.push_back({r, is_real});
    std::vector<E> s{nonce, count};
    s.insert(s.end(), relation.begin(), relation.end());
    sends.push_back({s, is_real});
  }
};

// ColumnSlice + ColumnIndex (src/lair/air.rs:24-103)
struct Cols {
  Builder& b;
  int in0, out0, aux0, sel0, i_in = 0, i_aux = 0, i_out = 0;
  Cols(Builder& bb, const LayoutSizes& ls)
      : b(bb), in0(1), out0(1 + ls.input), aux0(1 + ls.input + ls.output), sel0(1 + ls.input + ls.output + ls.aux) {}
  E nonce() { return b.local(0); }
  E next_input() { return b.local(in0 + i_in++); }
  E next_aux() { return b.local(aux0 + i_aux++); }
  E next_output() { return b.local(out0 + i_out++); }
  E output(int i) { return b.local(out0 + i); }
  E sel(int i) { return b.local(sel0 + i); }
  struct Req { E pn, pc, ci; };
  Req next_require() {
    Req r;
    r.pn = next_aux(); r.pc = next_aux(); r.ci = next_aux();
    return r;
  }
};

// BytesAirRecordWithContext (src/gadgets/bytes/builder.rs:18-89)
struct ByteAirRecord {
  std::vector<std::pair<std::vector<E>, E>> records;
  void range_check_u8_iter(Builder& b, const std::vector<E>& xs, E is_real) {  // src/gadgets/bytes/mod.rs:127-138
    for (size_t i = 0; i < xs.size(); i += 2)
      records.push_back({{b.k(1), xs[i], i + 1 < xs.size() ? xs[i + 1] : b.k(0)}, is_real});
  }
  void less_than(Builder& b, E i1, E i2, E r, E is_real) { records.push_back({{b.k(3), i1, i2, r}, is_real}); }
  void range_check_u16(Builder& b, E i, E is_real) { records.push_back({{b.k(2), i}, is_real}); }
  void require_all(Builder& b, E nonce, const std::vector<Cols::Req>& requires_) {
    if (requires_.size() != records.size()) throw std::runtime_error("byte record / require count mismatch");
    for (size_t i = 0; i < records.size(); i++) {
      std::vector<E> rel{b.k(BYTE_TAG)};
      rel.insert(rel.end(), records[i].first.begin(), records[i].first.end());
      b.require(rel, nonce, requires_[i].pn, requires_[i].pc, requires_[i].ci, records[i].second);
    }
  }
};

E return_sel(const Block& block, Cols& cols) {
  E e = cols.b.k(0);
  for (uint32_t i : block.return_idents) e = cols.b.g.add(e, cols.sel((int)i));
  return e;
}

// LessThanWitness::assert_less_than; wit = is_comp[W] ++ [lhs_comp_limb, rhs_comp_limb]
void assert_less_than(Builder& b, const std::vector<E>& wit, const std::vector<E>& lhs, const std::vector<E>& rhs,
                      ByteAirRecord& rec, E is_real) {
  Graph& g = b.g;
  const size_t w = lhs.size();
  E lhs_limb = wit[w], rhs_limb = wit[w + 1];
  {
    Builder::When wr(b, is_real);
    E is_equal = b.k(0);
    for (size_t i = 0; i < w; i++) {
      if (i > 0) {
        Builder::When we(b, is_equal);
        b.assert_eq(lhs[i], rhs[i]);
      }
      b.assert_bool(wit[i]);
      is_equal = g.add(is_equal, wit[i]);
    }
    b.assert_one(is_equal);
    E sel_l = b.k(0), sel_r = b.k(0);
    for (size_t i = 0; i < w; i++) {
      sel_l = g.add(sel_l, g.mul(lhs[i], wit[i]));
      sel_r = g.add(sel_r, g.mul(rhs[i], wit[i]));
    }
    b.assert_eq(sel_l, lhs_limb);
    b.assert_eq(sel_r, rhs_limb);
  }
  rec.less_than(b, lhs_limb, rhs_limb, b.k(1), is_real);
}

// eval_depth (src/lair/air.rs:105-131): the callee's depth bytes, proven smaller than the row's own
void eval_depth(Builder& b, Cols& cols, const std::vector<E>& depth, E sel, std::vector<E>& out) {
  std::vector<E> dep_depth, witness;
  for (uint32_t i = 0; i < DEPTH_W; i++) dep_depth.push_back(cols.next_aux());
  for (uint32_t i = 0; i < DEPTH_LESS_THAN_SIZE; i++) witness.push_back(cols.next_aux());
  ByteAirRecord rec;
  assert_less_than(b, witness, dep_depth, depth, rec, sel);
  std::vector<Cols::Req> reqs;
  for (uint32_t i = 0; i < DEPTH_LESS_THAN_REQUIRES; i++) reqs.push_back(cols.next_require());
  rec.require_all(b, cols.nonce(), reqs);
  out.insert(out.end(), dep_depth.begin(), dep_depth.end());
}

// Poseidon2ExternalMatrixGeneral: M4 = circ(2,3,1,1) on each 4-chunk, then the sums of the residue classes mod 4
std::vector<E> external_linear(Graph& g, const std::vector<E>& s) {
  std::vector<E> out;
  for (size_t i = 0; i < s.size(); i += 4) {
    E x0 = s[i], x1 = s[i + 1], x2 = s[i + 2], x3 = s[i + 3];
    E t01 = g.add(x0, x1), t23 = g.add(x2, x3);
    E t = g.add(t01, t23);
    E t1 = g.add(t, x1), t3 = g.add(t, x3);
    out.push_back(g.add(t1, t01));
    out.push_back(g.add(t1, g.mul(x2, g.cst(2))));
    out.push_back(g.add(t3, t23));
    out.push_back(g.add(t3, g.mul(x0, g.cst(2))));
  }
  E sums[4] = {-1, -1, -1, -1};
  for (size_t i = 0; i < out.size(); i++) sums[i & 3] = sums[i & 3] < 0 ? out[i] : g.add(sums[i & 3], out[i]);
  std::vector<E> res;
  for (size_t i = 0; i < out.size(); i++) res.push_back(g.add(out[i], sums[i & 3]));
  return res;
}

struct P2Params { const uint32_t *diag, *full, *part; int rp; };
P2Params p2_params(int width) {
  switch (width) {
    case 24: return {P2_DIAG_24, P2_FULL_RC_24, P2_PART_RC_24, P2_RP_24};
    case 32: return {P2_DIAG_32, P2_FULL_RC_32, P2_PART_RC_32, P2_RP_32};
    case 40: return {P2_DIAG_40, P2_FULL_RC_40, P2_PART_RC_40, P2_RP_40};
  }
  throw std::runtime_error("unsupported hasher width");
}

// PoseidonChipset::eval + Poseidon2Cols::eval; witness = output[8] ++ Poseidon2Cols (wide/columns.rs:17-32)
std::vector<E> eval_poseidon_chipset(Builder& b, int width, E is_real, const std::vector<E>& inp, const std::vector<E>& witness) {
  Graph& g = b.g;
  const P2Params pp = p2_params(width);
  const int rp = pp.rp;
  std::vector<E> output(witness.begin(), witness.begin() + 8), w(witness.begin() + 8, witness.end());
  if ((int)inp.size() != width || (int)w.size() != 17 * width + 2 * rp - 1) throw std::runtime_error("bad hasher witness size");
  auto ext_state = [&](int r, int i) { return w[r * width + i]; };
  auto ext_sbox = [&](int r, int i) { return w[(8 + r) * width + i]; };
  const int o = 16 * width;
  auto int_init = [&](int i) { return w[o + i]; };
  auto int_state0 = [&](int r) { return w[o + width + r]; };
  auto int_sbox = [&](int r) { return w[o + width + rp - 1 + r]; };

  std::vector<E> start;
  for (E x : inp) start.push_back(g.mul(is_real, x));
  std::vector<E> state = external_linear(g, start);
  auto external_round = [&](std::vector<E> st, int r) {
    for (int i = 0; i < width; i++) b.assert_eq(st[i], ext_state(r, i));
    for (int i = 0; i < width; i++) st[i] = g.add(ext_state(r, i), g.mul(is_real, g.cst(pp.full[r * width + i])));
    for (int i = 0; i < width; i++) {
      b.assert_eq(g.mul(g.mul(st[i], st[i]), st[i]), ext_sbox(r, i));
      st[i] = g.mul(st[i], g.mul(ext_sbox(r, i), ext_sbox(r, i)));
    }
    return external_linear(g, st);
  };
  for (int r = 0; r < 4; r++) state = external_round(state, r);
  for (int r = 0; r < rp; r++) {
    if (r == 0) {
      for (int i = 0; i < width; i++) b.assert_eq(state[i], int_init(i));
      for (int i = 0; i < width; i++) state[i] = int_init(i);
    } else {
      b.assert_eq(state[0], int_state0(r - 1));
      state[0] = int_state0(r - 1);
    }
    E x = g.add(state[0], g.mul(is_real, g.cst(pp.part[r])));
    b.assert_eq(g.mul(g.mul(x, x), x), int_sbox(r));
    state[0] = g.mul(x, g.mul(int_sbox(r), int_sbox(r)));
    E total = state[0];
    for (int i = 1; i < width; i++) total = g.add(total, state[i]);
    for (int i = 0; i < width; i++) state[i] = g.add(g.mul(state[i], g.cst(pp.diag[i] % Pm)), total);  // config.rs:104-119
  }
  for (int r = 4; r < 8; r++) state = external_round(state, r);
  for (int i = 0; i < 8; i++) b.assert_eq(state[i], g.mul(is_real, output[i]));
  return output;
}

// AddWitness::assert_add (src/gadgets/unsigned/add.rs:19-60): byte-wise lhs + rhs = out with boolean carries
void assert_add(Builder& b, const std::vector<E>& lhs, const std::vector<E>& rhs, const std::vector<E>& out, E is_real) {
  Graph& g = b.g;
  uint64_t base = 256, r = 1, e = Pm - 2;
  while (e) {
    if (e & 1) r = r * base % Pm;
    base = base * base % Pm;
    e >>= 1;
  }
  const uint32_t base_inv = (uint32_t)r;
  Builder::When w(b, is_real);
  E carry = b.k(0);
  for (size_t i = 0; i < out.size(); i++) {
    E total = g.add(g.add(lhs[i], rhs[i]), carry);
    carry = g.mul(g.sub(total, out[i]), b.k(base_inv));
    b.assert_bool(carry);
  }
}

// CompareWitness::eval (src/gadgets/unsigned/cmp.rs:49-128) -> (is_less_than, is_equal);
// wit = is_comp[W] ++ [lhs_comp_limb, rhs_comp_limb, comp_diff_inv, is_less_than]
std::pair<E, E> eval_compare(Builder& b, const std::vector<E>& wit, const std::vector<E>& lhs, const std::vector<E>& rhs,
                             ByteAirRecord& rec, E is_real) {
  Graph& g = b.g;
  const int w = (int)lhs.size();
  E lhs_limb = wit[w], rhs_limb = wit[w + 1], diff_inv = wit[w + 2], is_lt = wit[w + 3], is_equal;
  {
    Builder::When wr(b, is_real);
    is_equal = b.k(1);
    for (int i = w - 1; i >= 0; i--) {
      b.assert_bool(wit[i]);
      is_equal = g.sub(is_equal, wit[i]);
      Builder::When we(b, is_equal);
      b.assert_eq(lhs[i], rhs[i]);
    }
    b.assert_bool(is_equal);
    E sel_l = b.k(0), sel_r = b.k(0);
    for (int i = 0; i < w; i++) {
      sel_l = g.add(sel_l, g.mul(lhs[i], wit[i]));
      sel_r = g.add(sel_r, g.mul(rhs[i], wit[i]));
    }
    b.assert_eq(sel_l, lhs_limb);
    b.assert_eq(sel_r, rhs_limb);
    E is_different = g.sub(b.k(1), is_equal);
    b.assert_eq(g.mul(g.sub(lhs_limb, rhs_limb), diff_inv), is_different);
  }
  rec.less_than(b, lhs_limb, rhs_limb, is_lt, is_real);
  return {is_lt, is_equal};
}

// Product::eval = MulWitness::assert_mul + range-checked result (src/gadgets/unsigned/mul.rs:68-164);
// wit = carry[W] ++ result[W]
std::vector<E> eval_product(Builder& b, const std::vector<E>& wit, const std::vector<E>& lhs, const std::vector<E>& rhs,
                            ByteAirRecord& rec, E is_real) {
  Graph& g = b.g;
  const size_t w = lhs.size();
  std::vector<E> result(wit.begin() + w, wit.begin() + 2 * w), products;
  for (size_t k = 0; k < w; k++) products.push_back(b.k(0));
  for (size_t i = 0; i < w; i++)
    for (size_t j = 0; j < w; j++)
      if (i + j < w) products[i + j] = g.add(products[i + j], g.mul(lhs[i], rhs[j]));
  E carry_prev = b.k(0);
  for (size_t k = 0; k < w; k++) {
    rec.range_check_u16(b, wit[k], is_real);
    E out = g.add(result[k], g.mul(wit[k], b.k(256)));
    {
      Builder::When wr(b, is_real);
      b.assert_eq(g.add(products[k], carry_prev), out);
    }
    carry_prev = wit[k];
  }
  rec.range_check_u8_iter(b, result, is_real);
  return result;
}

// U64::eval (src/core/u64.rs:170-222), chip kinds 4..9 = Add, Sub, LessThan, IsZero, Mul, DivRem: Sum64 / Diff64
// (src/gadgets/unsigned/add.rs:63-175), Compare64 (cmp.rs:49-128), IsZero64 (is_zero.rs:46-156), Product64
// (mul.rs:68-164), DivRem64 (div_rem.rs:62-103)
std::vector<E> eval_u64_chipset(Builder& b, uint32_t kind, E is_real, const std::vector<E>& inp, const std::vector<E>& witness,
                                E nonce, const std::vector<Cols::Req>& requires_) {
  Graph& g = b.g;
  std::vector<E> in1(inp.begin(), inp.begin() + 8), in2, out;
  if (inp.size() >= 16) in2.assign(inp.begin() + 8, inp.begin() + 16);
  ByteAirRecord rec;
  auto slice = [&](size_t from, size_t to) { return std::vector<E>(witness.begin() + from, witness.begin() + to); };
  if (kind == 4 || kind == 5) {
    std::vector<E> result = slice(0, 8);
    rec.range_check_u8_iter(b, result, is_real);  // UncheckedWord::into_checked
    if (kind == 4) assert_add(b, in1, in2, result, is_real);
    else assert_add(b, result, in2, in1, is_real);  // result + rhs = lhs
    out = result;
  } else if (kind == 6) {
    out = {eval_compare(b, slice(0, 12), in1, in2, rec, is_real).first};
  } else if (kind == 7) {  // IsZeroOrEqual::eval_is_zero
    E is_zero = witness[8];
    {
      Builder::When wr(b, is_real);
      b.assert_bool(is_zero);
    }
    {
      Builder::When wr(b, is_real);
      E lc = b.k(0);
      for (int i = 0; i < 8; i++) {
        {
          Builder::When wz(b, is_zero);
          b.assert_zero(in1[i]);
        }
        lc = g.add(lc, g.mul(in1[i], witness[i]));
      }
      b.assert_eq(lc, g.sub(b.k(1), is_zero));
    }
    out = {is_zero};
  } else if (kind == 8) {
    out = eval_product(b, slice(0, 16), in1, in2, rec, is_real);
  } else {  // DivRem::eval, (a, b) = (in1, in2)
    std::vector<E> q = slice(8, 16), w_r = slice(32, 40);
    {
      Builder::When wr(b, is_real);  // IsZeroWitness::assert_non_zero(b)
      E lc = b.k(0);
      for (int i = 0; i < 8; i++) lc = g.add(lc, g.mul(in2[i], witness[i]));
      b.assert_one(lc);
    }
    rec.range_check_u8_iter(b, q, is_real);
    std::vector<E> qb = eval_product(b, slice(16, 32), q, in2, rec, is_real);
    rec.range_check_u8_iter(b, w_r, is_real);  // Diff::eval: r + qb = a
    assert_add(b, w_r, qb, in1, is_real);
    assert_less_than(b, slice(40, 50), w_r, in2, rec, is_real);
    auto cmp = eval_compare(b, slice(50, 62), qb, in1, rec, is_real);
    {
      Builder::When wr(b, is_real);
      b.assert_one(g.add(cmp.first, cmp.second));
    }
    out = q;
    out.insert(out.end(), w_r.begin(), w_r.end());
  }
  rec.require_all(b, nonce, requires_);
  return out;
}

// FieldToWord32::eval (src/gadgets/unsigned/field.rs:36-140): wit = [is_msb_less_than] ++ result[4]
std::vector<E> eval_field_to_word32(Builder& b, const std::vector<E>& wit, E field, ByteAirRecord& rec, E is_real) {
  Graph& g = b.g;
  E is_lt = wit[0];
  std::vector<E> word(wit.begin() + 1, wit.begin() + 5);
  {
    Builder::When w(b, is_real);
    b.assert_bool(is_lt);
  }
  E recomposed = b.k(0);
  for (int i = 3; i >= 0; i--) recomposed = g.add(g.mul(recomposed, b.k(256)), word[i]);
  {
    Builder::When w(b, is_real);
    b.assert_eq(field, recomposed);
  }
  rec.less_than(b, word[3], b.k(0x78), is_lt, is_real);
  {
    Builder::When w(b, g.mul(is_real, g.sub(b.k(1), is_lt)));
    b.assert_eq(word[3], b.k(0x78));
    for (int i = 0; i < 3; i++) b.assert_eq(word[i], b.k(0));
  }
  rec.range_check_u8_iter(b, word, is_real);
  return word;
}

// BigNum::LessThan (src/core/big_num.rs:75-97) = BigNumCompareWitness::eval (src/gadgets/big_num/cmp.rs:48-124)
std::vector<E> eval_big_num_less_than(Builder& b, E is_real, const std::vector<E>& inp, const std::vector<E>& witness, E nonce,
                                      const std::vector<Cols::Req>& requires_) {
  Graph& g = b.g;
  std::vector<E> lhs(inp.begin(), inp.begin() + 8), rhs(inp.begin() + 8, inp.begin() + 16);
  E lhs_limb = witness[8], rhs_limb = witness[9], is_equal;
  ByteAirRecord rec;
  auto slice = [&](size_t from, size_t to) { return std::vector<E>(witness.begin() + from, witness.begin() + to); };
  {
    Builder::When wr(b, is_real);
    is_equal = b.k(1);
    for (int i = 7; i >= 0; i--) {
      b.assert_bool(witness[i]);
      is_equal = g.sub(is_equal, witness[i]);
      Builder::When we(b, is_equal);
      b.assert_eq(lhs[i], rhs[i]);
    }
    b.assert_bool(is_equal);
    E sel_l = b.k(0), sel_r = b.k(0);
    for (int i = 0; i < 8; i++) {
      sel_l = g.add(sel_l, g.mul(lhs[i], witness[i]));
      sel_r = g.add(sel_r, g.mul(rhs[i], witness[i]));
    }
    b.assert_eq(sel_l, lhs_limb);
    b.assert_eq(sel_r, rhs_limb);
  }
  std::vector<E> lhs_word = eval_field_to_word32(b, slice(10, 15), lhs_limb, rec, is_real);
  std::vector<E> rhs_word = eval_field_to_word32(b, slice(15, 20), rhs_limb, rec, is_real);
  auto cmp = eval_compare(b, slice(20, 28), lhs_word, rhs_word, rec, is_real);
  {
    Builder::When wr(b, is_real);
    b.assert_eq(is_equal, cmp.second);
  }
  rec.require_all(b, nonce, requires_);
  return {cmp.first};
}

std::vector<E> eval_chip(Builder& b, const ChipInfo& chip, E is_real, const std::vector<E>& inp, const std::vector<E>& witness,
                         E nonce, const std::vector<Cols::Req>& requires_) {
  if (chip.kind >= 1 && chip.kind <= 3) return eval_poseidon_chipset(b, chip.kind == 1 ? 24 : chip.kind == 2 ? 32 : 40, is_real, inp, witness);
  if (chip.kind >= 4 && chip.kind <= 9) return eval_u64_chipset(b, chip.kind, is_real, inp, witness, nonce, requires_);
  if (chip.kind == 10) return eval_big_num_less_than(b, is_real, inp, witness, nonce, requires_);
  throw std::runtime_error("AIR capture of extern chip " + chip.name + " is not implemented");
}

using VMap = std::vector<std::pair<E, uint32_t>>;  // (expr, layout degree as in func_chip.rs)

void eval_block(Builder& b, const Block& block, Cols& cols, E sel, VMap& vmap, const Toplevel& top, const std::vector<E>& depth);

void eval_op(Builder& b, const Op& op, Cols& cols, E sel, VMap& vmap, const Toplevel& top, const std::vector<E>& depth) {
  Graph& g = b.g;
  E nonce = cols.nonce();
  switch (op.kind) {
    case OK::AssertNe: {
      std::vector<E> coeffs;
      for (size_t i = 0; i < op.xs.size(); i++) coeffs.push_back(cols.next_aux());
      E acc = b.k(0);
      for (size_t i = 0; i < op.xs.size(); i++) acc = g.add(acc, g.mul(coeffs[i], g.sub(vmap[op.xs[i]].first, vmap[op.ys[i]].first)));
      Builder::When w(b, sel);
      b.assert_one(acc);
    } break;
    case OK::AssertEq:
      for (size_t i = 0; i < op.xs.size(); i++) {
        Builder::When w(b, sel);
        b.assert_eq(vmap[op.xs[i]].first, vmap[op.ys[i]].first);
      }
      break;
    case OK::Contains: {
      E bv = vmap[op.b].first;
      E acc = g.sub(vmap[op.xs[0]].first, bv);
      for (size_t i = 1; i < op.xs.size(); i++) {
        E aux = cols.next_aux();
        {
          Builder::When w(b, sel);
          b.assert_eq(g.mul(acc, g.sub(vmap[op.xs[i]].first, bv)), aux);
        }
        acc = aux;
      }
      Builder::When w(b, sel);
      b.assert_zero(acc);
    } break;
    case OK::Const: vmap.push_back({b.k(op.a), 0}); break;
    case OK::Add: case OK::Sub: {
      auto x = vmap[op.a], y = vmap[op.b];
      vmap.push_back({op.kind == OK::Add ? g.add(x.first, y.first) : g.sub(x.first, y.first), std::max(x.second, y.second)});
    } break;
    case OK::Mul: {
      auto x = vmap[op.a], y = vmap[op.b];
      // a witness column exists iff the layout allocated one (func_chip.rs:199-209)
      if (x.second + y.second < 2) {
        vmap.push_back({g.mul(x.first, y.first), x.second + y.second});
      } else {
        E c = cols.next_aux();
        {
          Builder::When w(b, sel);
          b.assert_eq(g.mul(x.first, y.first), c);
        }
        vmap.push_back({c, 1});
      }
    } break;
    case OK::Inv: {
      auto x = vmap[op.a];
      if (x.second == 0) {
        uint64_t base = g.nodes[x.first].cst, r = 1, e = Pm - 2;
        while (e) {
          if (e & 1) r = r * base % Pm;
          base = base * base % Pm;
          e >>= 1;
        }
        vmap.push_back({b.k((uint32_t)r), 0});
      } else {
        E c = cols.next_aux();
        {
          Builder::When w(b, sel);
          b.assert_one(g.mul(x.first, c));
        }
        vmap.push_back({c, 1});
      }
    } break;
    case OK::Not: {
      auto x = vmap[op.a];
      if (x.second == 0) {
        vmap.push_back({b.k(g.nodes[x.first].cst == 0 ? 1 : 0), 0});
      } else {
        E d = cols.next_aux(), r = cols.next_aux();
        {
          Builder::When w(b, sel);
          b.assert_zero(g.mul(x.first, r));
        }
        {
          Builder::When w(b, sel);
          b.assert_one(g.add(g.mul(x.first, d), r));
        }
        vmap.push_back({r, 1});
      }
    } break;
    case OK::Call: {
      const Func& callee = top.func_by_index(op.a);
      std::vector<E> out, inp;
      for (uint32_t i = 0; i < callee.output_size; i++) {
        E o = cols.next_aux();
        vmap.push_back({o, 1});
        out.push_back(o);
      }
      for (uint32_t i : op.xs) inp.push_back(vmap[i].first);
      auto rq = cols.next_require();
      if (callee.partial) eval_depth(b, cols, depth, sel, out);
      std::vector<E> rel{b.k(CALL_TAG), b.k(op.a)};
      rel.insert(rel.end(), inp.begin(), inp.end());
      rel.insert(rel.end(), out.begin(), out.end());
      b.require(rel, nonce, rq.pn, rq.pc, rq.ci, sel);
    } break;
    case OK::PreImg: {
      const Func& callee = top.func_by_index(op.a);
      std::vector<E> inp, out;
      for (uint32_t i = 0; i < callee.input_size; i++) {
        E x = cols.next_aux();
        vmap.push_back({x, 1});
        inp.push_back(x);
      }
      for (uint32_t i : op.xs) out.push_back(vmap[i].first);
      auto rq = cols.next_require();
      if (callee.partial) eval_depth(b, cols, depth, sel, out);
      std::vector<E> rel{b.k(CALL_TAG), b.k(op.a)};
      rel.insert(rel.end(), inp.begin(), inp.end());
      rel.insert(rel.end(), out.begin(), out.end());
      b.require(rel, nonce, rq.pn, rq.pc, rq.ci, sel);
    } break;
    case OK::Store: {
      E ptr = cols.next_aux();
      vmap.push_back({ptr, 1});
      std::vector<E> rel{b.k(MEMORY_TAG), ptr};
      for (uint32_t i : op.xs) rel.push_back(vmap[i].first);
      auto rq = cols.next_require();
      b.require(rel, nonce, rq.pn, rq.pc, rq.ci, sel);
    } break;
    case OK::Load: {
      std::vector<E> rel{b.k(MEMORY_TAG), vmap[op.b].first};
      for (uint32_t i = 0; i < op.a; i++) {
        E o = cols.next_aux();
        vmap.push_back({o, 1});
        rel.push_back(o);
      }
      auto rq = cols.next_require();
      b.require(rel, nonce, rq.pn, rq.pc, rq.ci, sel);
    } break;
    case OK::RangeU8: {
      std::vector<E> xs;
      for (uint32_t i : op.xs) xs.push_back(vmap[i].first);
      const size_t n_req = (xs.size() + 1) / 2;
      std::vector<Cols::Req> reqs;
      for (size_t j = 0; j < n_req; j++) reqs.push_back(cols.next_require());
      for (size_t j = 0; j < n_req; j++) {
        E i1 = xs[2 * j], i2 = 2 * j + 1 < xs.size() ? xs[2 * j + 1] : b.k(0);
        b.require({b.k(BYTE_TAG), b.k(1), i1, i2}, nonce, reqs[j].pn, reqs[j].pc, reqs[j].ci, sel);
      }
    } break;
    case OK::ExternCall: {  // Op::ExternCall (src/lair/air.rs:461-482): witness columns first, then the require triples
      const ChipInfo& chip = top.chips()[op.a];
      std::vector<E> inp, witness;
      for (uint32_t i : op.xs) inp.push_back(vmap[i].first);
      for (uint32_t i = 0; i < chip.witness_size; i++) witness.push_back(cols.next_aux());
      std::vector<Cols::Req> reqs;
      for (uint32_t i = 0; i < chip.require_size; i++) reqs.push_back(cols.next_require());
      for (E o : eval_chip(b, chip, sel, inp, witness, nonce, reqs)) vmap.push_back({o, 1});
    } break;
  }
}

void eval_block(Builder& b, const Block& block, Cols& cols, E sel, VMap& vmap, const Toplevel& top, const std::vector<E>& depth) {
  for (auto& op : block.ops) eval_op(b, op, cols, sel, vmap, top, depth);
  const Ctrl& c = block.ctrl;
  if (c.kind == Ctrl::Return) {
    E s = cols.sel((int)c.ident);
    for (uint32_t v : c.vars) {
      Builder::When w(b, s);
      b.assert_eq(vmap[v].first, cols.next_output());
    }
    return;
  }
  const size_t n = vmap.size();
  const int s_aux = cols.i_aux, s_out = cols.i_out;
  auto one = [&](const Block& blk) {
    eval_block(b, blk, cols, return_sel(blk, cols), vmap, top, depth);
    vmap.resize(n);
    cols.i_aux = s_aux;
    cols.i_out = s_out;
  };
  for (auto& br : c.branches) one(*br.second);
  if (c.dflt) one(*c.dflt);
}

// Func::eval (src/lair/air.rs:158-235)
void eval_func(Builder& b, const Func& f, const Toplevel& top) {
  LayoutSizes ls = top.compute_layout_sizes(f);
  Cols cols(b, ls);
  E nonce = cols.nonce();
  {
    Builder::When w(b, b.is_transition());
    b.assert_eq(b.next(0), b.g.add(nonce, b.k(1)));
  }
  VMap vmap;
  std::vector<E> call_inp;
  for (uint32_t i = 0; i < f.input_size; i++) {
    E x = cols.next_input();
    vmap.push_back({x, 1});
    call_inp.push_back(x);
  }
  E toplevel_sel = return_sel(f.body, cols);
  b.assert_bool(toplevel_sel);
  E last_nonce = cols.next_aux(), last_count = cols.next_aux();
  std::vector<E> out;
  for (uint32_t i = 0; i < f.output_size; i++) out.push_back(cols.output((int)i));
  std::vector<E> depth;
  if (f.partial) {  // src/lair/air.rs:210-224
    for (uint32_t i = 0; i < DEPTH_W; i++) depth.push_back(cols.next_aux());
    std::vector<Cols::Req> reqs;
    for (uint32_t i = 0; i < (DEPTH_W + 1) / 2; i++) reqs.push_back(cols.next_require());
    ByteAirRecord rec;
    rec.range_check_u8_iter(b, depth, toplevel_sel);
    rec.require_all(b, nonce, reqs);
    out.insert(out.end(), depth.begin(), depth.end());
  }
  std::vector<E> rel{b.k(CALL_TAG), b.k(f.index)};
  rel.insert(rel.end(), call_inp.begin(), call_inp.end());
  rel.insert(rel.end(), out.begin(), out.end());
  b.provide(rel, last_nonce, last_count, toplevel_sel);
  eval_block(b, f.body, cols, toplevel_sel, vmap, top, depth);
}

// MemChip::eval (src/lair/memory.rs:71-109)
void eval_mem(Builder& b, uint32_t length) {
  Graph& g = b.g;
  E is_real = b.local(0), ptr = b.local(1), last_nonce = b.local(2), last_count = b.local(3);
  std::vector<E> values;
  for (uint32_t i = 0; i < length; i++) values.push_back(b.local(4 + (int)i));
  E is_real_next = b.next(0), ptr_next = b.next(1);
  b.assert_bool(is_real);
  E is_real_transition = g.mul(is_real_next, b.is_transition());
  {
    Builder::When w(b, is_real_transition);
    b.assert_one(is_real);
  }
  {
    Builder::When w1(b, b.is_first_row());
    Builder::When w2(b, is_real);
    b.assert_one(ptr);
  }
  {
    Builder::When w(b, is_real_transition);
    b.assert_eq(g.add(ptr, b.k(1)), ptr_next);
  }
  std::vector<E> rel{b.k(MEMORY_TAG), ptr};
  rel.insert(rel.end(), values.begin(), values.end());
  b.provide(rel, last_nonce, last_count, is_real);
}

// BytesChip::eval (src/gadgets/bytes/trace.rs:116-143)
void eval_bytes(Builder& b) {
  Graph& g = b.g;
  E i1 = b.prep(0), i2 = b.prep(1), lt = b.prep(2), and_ = b.prep(3), xor_ = b.prep(4), or_ = b.prep(5);
  E is_real = b.local(0);
  b.assert_bool(is_real);
  E input_u16 = g.add(i1, g.mul(i2, b.k(256)));
  E tag = b.k(BYTE_TAG);
  std::vector<std::vector<E>> relations = {
      {tag, b.k(1), i1, i2},        {tag, b.k(2), input_u16},     {tag, b.k(3), i1, i2, lt},
      {tag, b.k(4), i1, i2, and_},  {tag, b.k(5), i1, i2, xor_},  {tag, b.k(6), i1, i2, or_},
  };
  for (size_t k = 0; k < relations.size(); k++) b.provide(relations[k], b.local(1 + 2 * (int)k), b.local(2 + 2 * (int)k), is_real);
}

// LairChip::Entrypoint eval (src/lair/lair_chip.rs:166-191)
void eval_entrypoint(Builder& b, uint32_t func_idx, uint32_t num_pv) {
  std::vector<E> pvs;
  for (uint32_t i = 0; i < num_pv; i++) pvs.push_back(b.local((int)i));
  for (uint32_t i = 0; i < num_pv; i++) b.assert_eq(pvs[i], b.public_value((int)i));
  std::vector<E> rel{b.k(CALL_TAG), b.k(func_idx)};
  rel.insert(rel.end(), pvs.begin(), pvs.end());
  b.require(rel, b.k(0), b.k(0), b.k(0), b.k(1), b.k(1));
}

// ---- compilation of the captured AIR into the flat descriptor ----------------------------------------------------
constexpr uint32_t I_ADD = 1, I_SUB = 2, I_MUL = 3, I_NEG = 4, I_ASSERT = 5;
inline uint32_t encode_leaf(int kind, uint32_t idx) { return ((uint32_t)(1 + kind) << 28) | idx; }

struct Compiled {
  std::vector<uint32_t> code, consts;
  uint32_t n_regs = 1;
};

// straight-line program: one pass for use counts, one pass emitting with a free-list register allocator
Compiled compile_constraints(const Graph& g, const std::vector<E>& constraints) {
  std::vector<uint32_t> uses(g.nodes.size(), 0);
  std::vector<char> seen(g.nodes.size(), 0);
  {
    // iterative form of: visit(e) { uses[e]++; if (seen) return; seen; visit(a); visit(b); }
    std::vector<E> stack;
    for (E c : constraints) {
      stack.push_back(c);
      while (!stack.empty()) {
        E e = stack.back();
        stack.pop_back();
        uses[e]++;
        if (seen[e]) continue;
        seen[e] = 1;
        const Node& n = g.nodes[e];
        if (n.kind != Node::Aff) {
          if (n.b >= 0) stack.push_back(n.b);
          stack.push_back(n.a);
        }
      }
    }
  }
  Compiled out;
  std::vector<uint32_t>& code = out.code;
  std::map<uint32_t, uint32_t> const_index;
  std::vector<uint32_t> const_order;
  auto const_idx = [&](uint32_t c) {
    c %= Pm;
    auto it = const_index.find(c);
    if (it != const_index.end()) return it->second;
    uint32_t i = (uint32_t)const_order.size();
    const_index[c] = i;
    const_order.push_back(c);
    return i;
  };
  std::vector<uint32_t> free_regs;
  uint32_t n_regs = 0;
  std::vector<uint32_t> remaining = uses;
  std::vector<uint32_t> val(g.nodes.size(), 0);
  std::vector<char> has_val(g.nodes.size(), 0);
  auto alloc = [&]() {
    if (!free_regs.empty()) {
      uint32_t r = free_regs.back();
      free_regs.pop_back();
      return r;
    }
    return n_regs++;
  };
  auto release = [&](E e) {
    if (--remaining[e] == 0) {
      has_val[e] = 0;
      if ((val[e] >> 28) == 0) free_regs.push_back(val[e]);
    }
  };
  auto emit_affine = [&](const Node& n) -> uint32_t {
    if (n.lin.size() == 1 && n.lin.begin()->second == 1 && n.cst == 0)
      return encode_leaf(n.lin.begin()->first.first, (uint32_t)n.lin.begin()->first.second);
    if (n.lin.empty()) return encode_leaf(L_CONST, const_idx(n.cst));
    uint32_t r = alloc();
    bool first = true;
    for (auto& kv : n.lin) {  // std::map iterates in key order = sorted(items)
      uint32_t leaf = encode_leaf(kv.first.first, (uint32_t)kv.first.second), cf = kv.second;
      if (cf == 1) {
        if (first) code.insert(code.end(), {I_ADD, r, leaf, encode_leaf(L_CONST, const_idx(n.cst))});
        else code.insert(code.end(), {I_ADD, r, r, leaf});
      } else if (cf == Pm - 1 && !first) {
        code.insert(code.end(), {I_SUB, r, r, leaf});
      } else if (first) {
        code.insert(code.end(), {I_MUL, r, leaf, encode_leaf(L_CONST, const_idx(cf))});
        if (n.cst) code.insert(code.end(), {I_ADD, r, r, encode_leaf(L_CONST, const_idx(n.cst))});
      } else {
        uint32_t t = alloc();
        code.insert(code.end(), {I_MUL, t, leaf, encode_leaf(L_CONST, const_idx(cf))});
        code.insert(code.end(), {I_ADD, r, r, t});
        free_regs.push_back(t);
      }
      first = false;
    }
    return r;
  };
  // gen(e): post-order over the DAG with an explicit stack (the Poseidon2 chains are thousands deep)
  struct Frame { E e; int stage; uint32_t wa; };
  auto gen = [&](E root) -> uint32_t {
    if (has_val[root]) return val[root];
    std::vector<Frame> st{{root, 0, 0}};
    uint32_t result = 0;
    while (!st.empty()) {
      Frame& f = st.back();
      const Node& n = g.nodes[f.e];
      if (f.stage == 0 && has_val[f.e]) {
        result = val[f.e];
        st.pop_back();
        continue;
      }
      if (n.kind == Node::Aff) {
        val[f.e] = emit_affine(n);
        has_val[f.e] = 1;
        result = val[f.e];
        st.pop_back();
        continue;
      }
      if (f.stage == 0) {
        f.stage = 1;
        st.push_back({n.a, 0, 0});
        continue;
      }
      if (f.stage == 1) {
        f.wa = result;
        f.stage = 2;
        if (n.b >= 0) {
          st.push_back({n.b, 0, 0});
          continue;
        }
        result = 0;
      }
      uint32_t wb = n.b >= 0 ? result : 0, wa = f.wa;
      E e = f.e;
      release(n.a);
      if (n.b >= 0) release(n.b);
      uint32_t r = alloc();
      uint32_t op = n.kind == Node::Add ? I_ADD : n.kind == Node::Sub ? I_SUB : n.kind == Node::Mul ? I_MUL : I_NEG;
      code.insert(code.end(), {op, r, wa, wb});
      val[e] = r;
      has_val[e] = 1;
      result = r;
      st.pop_back();
    }
    return result;
  };
  for (E c : constraints) {
    uint32_t w = gen(c);
    code.insert(code.end(), {I_ASSERT, 0u, w, 0u});
    release(c);
  }
  out.consts = const_order;
  out.n_regs = std::max(n_regs, 1u);
  return out;
}

void encode_affine(const Graph& g, E e, std::vector<uint32_t>& out) {
  const Node& n = g.nodes[e];
  if (n.kind != Node::Aff) throw std::runtime_error("interaction values must be affine in the trace columns");
  out.push_back((uint32_t)n.lin.size());
  out.push_back(n.cst % Pm);
  for (auto& kv : n.lin) {
    if (kv.first.first != L_MAIN && kv.first.first != L_PREP)
      throw std::runtime_error("interactions may only reference local main/preprocessed cells");
    out.push_back(kv.first.first == L_MAIN ? 0u : 1u);
    out.push_back((uint32_t)kv.first.second);
    out.push_back(kv.second);
  }
}

enum : uint32_t { KIND_FUNC = 0, KIND_MEM = 1, KIND_BYTES = 2, KIND_ENTRY = 3 };

std::vector<uint32_t> finish(const std::string& name, uint32_t kind, uint32_t index, Builder& b) {
  Compiled c = compile_constraints(b.g, b.constraints);
  uint32_t max_deg = 0;
  for (E e : b.constraints) max_deg = std::max(max_deg, b.g.nodes[e].degree);
  uint32_t d = std::max(max_deg, 3u), lqd = 0;
  for (uint32_t v = d - 2; v; v >>= 1) lqd++;  // (d - 2).bit_length() = log2_ceil(d - 1)
  std::vector<uint32_t> name_words((name.size() + 3) / 4, 0);
  for (size_t i = 0; i < name.size(); i++) name_words[i / 4] |= (uint32_t)(uint8_t)name[i] << (8 * (i % 4));
  std::vector<uint32_t> out{kind, index, b.width, b.prep_width, (uint32_t)b.constraints.size(), lqd, c.n_regs,
                            (uint32_t)name.size(), (uint32_t)name_words.size()};
  out.insert(out.end(), name_words.begin(), name_words.end());
  out.push_back((uint32_t)c.consts.size());
  out.insert(out.end(), c.consts.begin(), c.consts.end());
  out.push_back((uint32_t)(c.code.size() / 4));
  out.insert(out.end(), c.code.begin(), c.code.end());
  for (auto* lst : {&b.sends, &b.receives}) {
    out.push_back((uint32_t)lst->size());
    for (auto& it : *lst) {
      out.push_back(it.kind);
      out.push_back((uint32_t)it.values.size());
      for (E v : it.values) encode_affine(b.g, v, out);
      encode_affine(b.g, it.mult, out);
    }
  }
  return out;
}

}  // namespace

// build_lair_chip_vector + Chip::new: Entrypoint, every func of the toplevel, the six memory tables, the byte
// table (named "CPU", src/lair/lair_chip.rs:90-92)
std::vector<uint32_t> chip_vector_blob(const Toplevel& top, uint32_t entry_func) {
  const Func& f = top.func_by_index(entry_func);
  const uint32_t num_pv = f.input_size + f.output_size + (f.partial ? DEPTH_W : 0);
  std::vector<std::vector<uint32_t>> bodies;
  {
    Builder b(num_pv, 0, num_pv);
    eval_entrypoint(b, f.index, num_pv);
    bodies.push_back(finish("Entrypoint[" + std::to_string(f.index) + "]", KIND_ENTRY, f.index, b));
  }
  for (uint32_t i = 0; i < top.num_funcs(); i++) {
    const Func& fn = top.func_by_index(i);
    Builder b(top.compute_layout_sizes(fn).total(), 0, num_pv);
    eval_func(b, fn, top);
    bodies.push_back(finish("Func[" + fn.name + "]", KIND_FUNC, fn.index, b));
  }
  for (uint32_t i = 0; i < 6; i++) {
    Builder b(4 + MEM_TABLE_SIZES[i], 0, num_pv);
    eval_mem(b, MEM_TABLE_SIZES[i]);
    bodies.push_back(finish("Mem[" + std::to_string(MEM_TABLE_SIZES[i]) + "-wide]", KIND_MEM, i, b));
  }
  {
    Builder b(13, 6, num_pv);
    eval_bytes(b);
    bodies.push_back(finish("CPU", KIND_BYTES, 0, b));
  }
  std::vector<uint32_t> out{0x4d414348u, (uint32_t)bodies.size(), num_pv};  // 'MACH'
  for (auto& body : bodies) {
    out.push_back((uint32_t)body.size());
    out.insert(out.end(), body.begin(), body.end());
  }
  return out;
}

}  // namespace host
}  // namespace lurk
