// ORACLE (test infrastructure, NOT product code).  CPU restatement of the STARK machine that the
// reference obtains from third-party crates that are NOT under /root/reference:
//   sphinx-core 1.0.0 @ 8a39b951 (argumentcomputer/sphinx, branch dev): StarkMachine::{setup, prove,
//     verify}, LocalProver, LogUp permutation trace + constraints, quotient, ShardProof;
//   Plonky3 `sp1` branch @ a0b9287: Radix2Dit coset LDE, FieldMerkleTreeMmcs, TwoAdicFriPcs, FRI,
//     DuplexChallenger.
// Reference call sites this stands behind: src/core/cli/repl.rs:164-207, src/lair/lair_chip.rs:
// 246-276, benches/fib.rs:136-163, tests/fib.rs:73-98.
//
// *** PARITY UNPINNED ***  The reference pins no commitment root, transcript or proof bytes
// (SURVEY.md F4): its tests only check prove -> verify == Ok.  This file restates the published
// SP1-v1 / Plonky3-2024 protocol from recollection; every assumption is marked ASSUMPTION.  What it
// guarantees: (1) internal consistency (its verifier accepts its proofs and rejects tampering),
// (2) it is the bit-exact checker for the CUDA prover (same protocol, independent code).
//
// ASSUMPTIONS: Val = BabyBear, Challenge = F[x]/(x^4-11); commitment
// permutation = Poseidon2 width 16 -- we use the in-tree BabyBearConfig16 constants
// (src/poseidon/config.rs:189-198) because p3's RC_16_30 / Montgomery diagonal are unavailable;
// PaddingFreeSponge<16,8,8>, TruncatedPermutation<2,8,16>; mixed-height Merkle tree with
// injection; coset LDE with shift = generator (31), blowup 2, bit-reversed rows; DuplexChallenger
// rate 8 (samples pop from the back of the output buffer); FRI: arity 2, 100 queries, 16 PoW bits;
// chips ordered by descending height (stable); LogUp batch = 2; quotient degree 2; opened values
// are observed by the challenger before the FRI batching challenge is drawn (the sounder of the
// two historical variants).
#pragma once
#include <algorithm>
#include <cstring>
#include <map>

#include "lair.h"

namespace orc {

// ---- small containers -------------------------------------------------------------------------
struct Mat {  // column-major
  size_t h = 0, w = 0;
  std::vector<Fp> v;
  Mat() {}
  Mat(size_t h_, size_t w_) : h(h_), w(w_), v(h_ * w_, Fp::zero()) {}
  Fp& at(size_t r, size_t c) { return v[c * h + r]; }
  const Fp& at(size_t r, size_t c) const { return v[c * h + r]; }
  Fp* col(size_t c) { return &v[c * h]; }
  const Fp* col(size_t c) const { return &v[c * h]; }
};
static inline Mat from_row_major(const std::vector<Fp>& rows, size_t h, size_t w) {
  Mat m(h, w);
  for (size_t r = 0; r < h; r++)
    for (size_t c = 0; c < w; c++) m.at(r, c) = rows[r * w + c];
  return m;
}
struct Digest {
  Fp d[8];
};

// ---- machine description (decoded from lurk_b200/air.py::serialize_machine) ----------------------
struct Affine {
  Fp cst;
  struct Term { uint32_t prep, col; Fp coeff; };
  std::vector<Term> terms;
  template <class V, class GetMain, class GetPrep>
  V eval(GetMain gm, GetPrep gp) const {
    V acc = V::zero() + cst;
    for (auto& t : terms) acc = acc + (t.prep ? gp(t.col) : gm(t.col)) * t.coeff;
    return acc;
  }
};
struct Interaction {
  uint32_t kind;
  std::vector<Affine> values;
  Affine mult;
};
enum { KIND_FUNC = 0, KIND_MEM = 1, KIND_BYTES = 2, KIND_ENTRY = 3 };
struct Chip {
  std::string name;
  uint32_t kind, index, width, prep_width, n_constraints, lqd, n_regs;
  std::vector<Fp> consts;
  std::vector<uint32_t> code;
  std::vector<Interaction> sends, receives;
  size_t n_interactions() const { return sends.size() + receives.size(); }
  size_t perm_width() const {  // in extension columns
    size_t batch = (size_t)1 << lqd;
    return (n_interactions() + batch - 1) / batch + 1;
  }
  size_t n_perm_constraints() const { return perm_width() - 1 + 3; }
};

static inline Affine parse_affine(const uint32_t* w, size_t& p) {
  Affine a;
  uint32_t n = w[p++];
  a.cst = Fp::from_u32(w[p++]);
  for (uint32_t i = 0; i < n; i++, p += 3) a.terms.push_back({w[p], w[p + 1], Fp::from_u32(w[p + 2])});
  return a;
}
static inline std::vector<Chip> parse_machine(const uint32_t* w, size_t n, uint32_t* num_pv) {
  if (n < 3 || w[0] != 0x4d414348u) throw std::runtime_error("bad machine magic");
  std::vector<Chip> chips(w[1]);
  *num_pv = w[2];
  size_t p = 3;
  for (auto& c : chips) {
    size_t len = w[p++], end = p + len;
    c.kind = w[p++]; c.index = w[p++]; c.width = w[p++]; c.prep_width = w[p++];
    c.n_constraints = w[p++]; c.lqd = w[p++]; c.n_regs = w[p++];
    uint32_t name_len = w[p++], name_words = w[p++];
    c.name.assign((const char*)(w + p), name_len);
    p += name_words;
    uint32_t nc = w[p++];
    for (uint32_t i = 0; i < nc; i++) c.consts.push_back(Fp::from_u32(w[p++]));
    uint32_t ni = w[p++];
    c.code.assign(w + p, w + p + 4 * ni);
    p += 4 * ni;
    for (int s = 0; s < 2; s++) {
      uint32_t cnt = w[p++];
      auto& lst = s == 0 ? c.sends : c.receives;
      for (uint32_t i = 0; i < cnt; i++) {
        Interaction it;
        it.kind = w[p++];
        uint32_t nv = w[p++];
        for (uint32_t k = 0; k < nv; k++) it.values.push_back(parse_affine(w, p));
        it.mult = parse_affine(w, p);
        lst.push_back(std::move(it));
      }
    }
    if (p != end) throw std::runtime_error("chip descriptor length mismatch");
    if (c.lqd != 1) throw std::runtime_error("only log_quotient_degree = 1 is supported");
  }
  return chips;
}

// ---- constraint program interpreter ------------------------------------------------------------
// V is Fp (prover, on the quotient domain: all cells are base field) or Ef (verifier, at zeta).
template <class V>
struct RowView {
  const V *main_local, *main_next, *prep_local, *prep_next;
  const Fp* pv;
  V is_first, is_last, is_transition;
};
template <class V>
static inline V lift(Fp x);
template <> inline Fp lift<Fp>(Fp x) { return x; }
template <> inline Ef lift<Ef>(Fp x) { return Ef::from_base(x); }

// Evaluates the AIR constraints and folds them: acc = acc * alpha + C_i (first constraint gets the
// highest power), starting from `acc`.
template <class V>
static inline Ef fold_air_constraints(const Chip& c, const RowView<V>& row, Ef alpha, Ef acc, std::vector<V>& regs) {
  regs.resize(c.n_regs);
  auto operand = [&](uint32_t w) -> V {
    uint32_t tag = w >> 28, idx = w & 0x0fffffffu;
    switch (tag) {
      case 0: return regs[idx];
      case 1: return lift<V>(c.consts[idx]);
      case 2: return row.main_local[idx];
      case 3: return row.main_next[idx];
      case 4: return row.prep_local[idx];
      case 5: return row.prep_next[idx];
      case 6: return lift<V>(row.pv[idx]);
      default: return idx == 0 ? row.is_first : idx == 1 ? row.is_last : row.is_transition;
    }
  };
  for (size_t i = 0; i < c.code.size(); i += 4) {
    uint32_t op = c.code[i], dst = c.code[i + 1];
    switch (op) {
      case 1: regs[dst] = operand(c.code[i + 2]) + operand(c.code[i + 3]); break;
      case 2: regs[dst] = operand(c.code[i + 2]) - operand(c.code[i + 3]); break;
      case 3: regs[dst] = operand(c.code[i + 2]) * operand(c.code[i + 3]); break;
      case 4: regs[dst] = -operand(c.code[i + 2]); break;
      case 5: acc = acc * alpha + operand(c.code[i + 2]); break;
      default: throw std::runtime_error("bad constraint opcode");
    }
  }
  return acc;
}
static inline Ef to_ef(Fp x) { return Ef::from_base(x); }
static inline Ef to_ef(const Ef& x) { return x; }

// LogUp constraints (sphinx `eval_permutation_constraints`), folded like the AIR ones.
// perm_local / perm_next: perm_width extension values; main/prep local cells as V.
template <class V>
static inline Ef fold_perm_constraints(const Chip& c, const RowView<V>& row, const Ef* perm_local, const Ef* perm_next,
                                       Ef cumulative_sum, Ef chal_alpha, Ef chal_beta, Ef alpha, Ef acc) {
  size_t pw = c.perm_width(), batch = (size_t)1 << c.lqd, nint = c.n_interactions();
  auto gm = [&](uint32_t col) { return row.main_local[col]; };
  auto gp = [&](uint32_t col) { return row.prep_local[col]; };
  for (size_t b = 0; b * batch < nint; b++) {
    // entry * prod(rlc_i) - sum_i m_i * prod_{j != i} rlc_j = 0
    Ef rlcs[8], mults[8];
    size_t cnt = 0;
    for (size_t k = b * batch; k < std::min(nint, (b + 1) * batch); k++, cnt++) {
      bool is_send = k < c.sends.size();
      const Interaction& it = is_send ? c.sends[k] : c.receives[k - c.sends.size()];
      Ef rlc = chal_alpha + Fp::from_u32(it.kind);
      Ef bp = chal_beta;
      for (auto& v : it.values) {
        rlc = rlc + bp * to_ef(v.template eval<V>(gm, gp));
        bp = bp * chal_beta;
      }
      Ef m = to_ef(it.mult.template eval<V>(gm, gp));
      rlcs[cnt] = rlc;
      mults[cnt] = is_send ? m : -m;
    }
    Ef lhs = perm_local[b];
    for (size_t i = 0; i < cnt; i++) lhs = lhs * rlcs[i];
    Ef rhs = Ef::zero();
    for (size_t i = 0; i < cnt; i++) {
      Ef t = mults[i];
      for (size_t j = 0; j < cnt; j++)
        if (j != i) t = t * rlcs[j];
      rhs = rhs + t;
    }
    acc = acc * alpha + (lhs - rhs);
  }
  Ef sum_local = Ef::zero(), sum_next = Ef::zero();
  for (size_t i = 0; i + 1 < pw; i++) sum_local += perm_local[i], sum_next += perm_next[i];
  Ef phi_local = perm_local[pw - 1], phi_next = perm_next[pw - 1];
  acc = acc * alpha + to_ef(row.is_transition) * (phi_next - phi_local - sum_next);
  acc = acc * alpha + to_ef(row.is_first) * (phi_local - sum_local);
  acc = acc * alpha + to_ef(row.is_last) * (phi_local - cumulative_sum);
  return acc;
}

// ---- NTT / LDE -----------------------------------------------------------------------------------
static inline void ntt_inplace(Fp* a, unsigned logn, bool inverse) {  // natural -> natural
  size_t n = (size_t)1 << logn;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, logn);
    if (i < j) std::swap(a[i], a[j]);
  }
  for (unsigned s = 1; s <= logn; s++) {
    size_t half = (size_t)1 << (s - 1);
    Fp wn = two_adic_generator(s);
    if (inverse) wn = fp_inv(wn);
    std::vector<Fp> tw(half);
    Fp w = Fp::one();
    for (size_t k = 0; k < half; k++) tw[k] = w, w = w * wn;
    for (size_t base = 0; base < n; base += 2 * half)
      for (size_t k = 0; k < half; k++) {
        Fp u = a[base + k], t = a[base + k + half] * tw[k];
        a[base + k] = u + t;
        a[base + k + half] = u - t;
      }
  }
  if (inverse) {
    Fp ninv = fp_inv(Fp::from_u32((uint32_t)n));
    for (size_t i = 0; i < n; i++) a[i] = a[i] * ninv;
  }
}
// evals: natural-order evaluations over in_shift * H_N of a polynomial of degree < N.
// returns its evaluations over g * H_2N in bit-reversed order (p3 coset_lde_batch + bit_reverse_rows)
static inline void coset_lde_col(const Fp* evals, unsigned logn, Fp in_shift, Fp* out) {
  size_t n = (size_t)1 << logn;
  std::vector<Fp> coef(evals, evals + n), tmp(n);
  ntt_inplace(coef.data(), logn, true);
  Fp w2n = two_adic_generator(logn + 1);
  Fp base = fp_generator() * fp_inv(in_shift);
  for (int j = 0; j < 2; j++) {
    Fp s = j ? base * w2n : base, p = Fp::one();
    for (size_t m = 0; m < n; m++) tmp[m] = coef[m] * p, p = p * s;
    ntt_inplace(tmp.data(), logn, false);
    for (size_t k = 0; k < n; k++) out[(size_t)j * n + bitrev(k, logn)] = tmp[k];
  }
}
static inline Mat coset_lde(const Mat& m, Fp in_shift = Fp::one()) {
  unsigned logn = log2_strict(m.h);
  Mat out(2 * m.h, m.w);
  parallel_for(m.w, [&](size_t c) { coset_lde_col(m.col(c), logn, in_shift, out.col(c)); }, 1);
  return out;
}

// ---- Merkle tree over matrices of mixed heights (p3 FieldMerkleTreeMmcs) --------------------------
struct MerkleTree {
  std::vector<const Mat*> mats;              // committed (LDE) matrices, original order
  std::vector<size_t> order;                 // indices sorted by height desc (stable)
  std::vector<std::vector<Digest>> layers;   // layers[0] = leaves
  Digest root() const { return layers.back()[0]; }
  size_t max_height() const { return layers[0].size(); }
};
static inline void hash_rows(const std::vector<const Mat*>& ms, size_t row, Digest& out) {
  Sponge16 sp;
  for (const Mat* m : ms)
    for (size_t c = 0; c < m->w; c++) sp.absorb(m->at(row, c));
  sp.finish(out.d);
}
static inline MerkleTree merkle_commit(const std::vector<const Mat*>& mats) {
  MerkleTree t;
  t.mats = mats;
  t.order.resize(mats.size());
  for (size_t i = 0; i < mats.size(); i++) t.order[i] = i;
  std::stable_sort(t.order.begin(), t.order.end(), [&](size_t a, size_t b) { return mats[a]->h > mats[b]->h; });
  size_t pos = 0, maxh = mats[t.order[0]]->h;
  std::vector<const Mat*> cur;
  while (pos < t.order.size() && mats[t.order[pos]]->h == maxh) cur.push_back(mats[t.order[pos++]]);
  t.layers.emplace_back(maxh);
  parallel_for(maxh, [&](size_t r) { hash_rows(cur, r, t.layers[0][r]); }, 64);
  while (t.layers.back().size() > 1) {
    const auto& prev = t.layers.back();
    size_t nl = prev.size() / 2;
    std::vector<const Mat*> inj;
    while (pos < t.order.size() && mats[t.order[pos]]->h == nl) inj.push_back(mats[t.order[pos++]]);
    std::vector<Digest> next(nl);
    parallel_for(nl, [&](size_t i) {
      Digest d;
      compress16(prev[2 * i].d, prev[2 * i + 1].d, d.d);
      if (!inj.empty()) {
        Digest rd;
        hash_rows(inj, i, rd);
        compress16(d.d, rd.d, next[i].d);
      } else {
        next[i] = d;
      }
    }, 64);
    t.layers.push_back(std::move(next));
  }
  if (pos != t.order.size()) throw std::runtime_error("matrix heights must be powers of two");
  return t;
}
struct BatchOpening {
  std::vector<std::vector<Fp>> rows;  // per matrix (original order)
  std::vector<Digest> path;
};
static inline BatchOpening merkle_open(const MerkleTree& t, size_t index) {
  BatchOpening o;
  unsigned logmax = log2_strict(t.max_height());
  for (const Mat* m : t.mats) {
    size_t r = index >> (logmax - log2_strict(m->h));
    std::vector<Fp> row(m->w);
    for (size_t c = 0; c < m->w; c++) row[c] = m->at(r, c);
    o.rows.push_back(std::move(row));
  }
  for (unsigned i = 0; i < logmax; i++) o.path.push_back(t.layers[i][(index >> i) ^ 1]);
  return o;
}
static inline bool merkle_verify(const Digest& root, const std::vector<size_t>& heights, size_t index,
                                 const std::vector<std::vector<Fp>>& rows, const std::vector<Digest>& path) {
  std::vector<size_t> order(heights.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return heights[a] > heights[b]; });
  size_t pos = 0, cur_h = heights[order[0]];
  if (path.size() != log2_strict(cur_h)) return false;
  auto hash_group = [&](size_t h, Digest& out) {
    Sponge16 sp;
    while (pos < order.size() && heights[order[pos]] == h) {
      for (Fp x : rows[order[pos]]) sp.absorb(x);
      pos++;
    }
    sp.finish(out.d);
  };
  Digest cur;
  hash_group(cur_h, cur);
  for (const Digest& sib : path) {
    Digest nx;
    if (index & 1) compress16(sib.d, cur.d, nx.d);
    else compress16(cur.d, sib.d, nx.d);
    cur = nx;
    index >>= 1;
    cur_h >>= 1;
    if (pos < order.size() && heights[order[pos]] == cur_h) {
      Digest rd, nx2;
      hash_group(cur_h, rd);
      compress16(cur.d, rd.d, nx2.d);
      cur = nx2;
    }
  }
  if (pos != order.size()) return false;
  for (int i = 0; i < 8; i++)
    if (cur.d[i] != root.d[i]) return false;
  return true;
}

// ---- DuplexChallenger<Val, Perm16, 16, 8> -----------------------------------------------------------
struct Challenger {
  Fp state[16];
  std::vector<Fp> in, out;
  Challenger() {
    for (auto& s : state) s = Fp::zero();
  }
  void duplex() {
    for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
    in.clear();
    p2_permute(state, 16);
    out.assign(state, state + 8);
  }
  void observe(Fp x) {
    out.clear();
    in.push_back(x);
    if (in.size() == 8) duplex();
  }
  void observe(const Digest& d) { for (Fp x : d.d) observe(x); }
  void observe(const Ef& e) { for (Fp x : e.c) observe(x); }
  Fp sample() {
    if (!in.empty() || out.empty()) duplex();
    Fp x = out.back();
    out.pop_back();
    return x;
  }
  Ef sample_ext() {
    Ef e;
    for (int i = 0; i < 4; i++) e.c[i] = sample();
    return e;
  }
  uint32_t sample_bits(unsigned bits) { return sample().to_u32() & (((uint32_t)1 << bits) - 1); }
  bool check_witness(unsigned bits, Fp w) {
    observe(w);
    return sample_bits(bits) == 0;
  }
  Fp grind(unsigned bits) {  // smallest witness (p3 takes any; we fix the choice to make proofs deterministic)
    for (uint32_t w = 0;; w++) {
      Challenger c = *this;
      if (c.check_witness(bits, Fp::from_u32(w))) {
        check_witness(bits, Fp::from_u32(w));
        return Fp::from_u32(w);
      }
    }
  }
};

// ---- parameters -------------------------------------------------------------------------------------
struct StarkParams {
  unsigned log_blowup = 1, num_queries = 100, pow_bits = 16;
};

// ---- per-shard chip data ------------------------------------------------------------------------------
struct Machine {
  std::vector<Chip> chips;
  uint32_t num_pv = 0;
  const Toplevel* top = nullptr;
  uint32_t entry = 0;
  StarkParams params;
  // setup (preprocessed): only the Bytes chip has one (lair_chip.rs:141-153)
  Mat prep_trace, prep_lde;
  MerkleTree prep_tree;
  int prep_chip = -1;
  Digest prep_commit;
  // TEST HOOK (negative tests of the verifiers only): a dishonest prover that leaves out every chip whose kind bit
  // is set here, whatever MachineAir::included says
  uint32_t forge_omit_kinds = 0;
};

static inline void machine_setup(Machine& m) {
  for (size_t i = 0; i < m.chips.size(); i++)
    if (m.chips[i].prep_width) {
      if (m.prep_chip >= 0) throw std::runtime_error("more than one preprocessed chip");
      m.prep_chip = (int)i;
    }
  if (m.prep_chip >= 0) {
    m.prep_trace = from_row_major(bytes_preprocessed(), 1 << 16, 6);
    m.prep_lde = coset_lde(m.prep_trace);
    m.prep_tree = merkle_commit({&m.prep_lde});
    m.prep_commit = m.prep_tree.root();
  } else {
    for (auto& x : m.prep_commit.d) x = Fp::zero();
  }
}

struct ChipOpening {
  uint32_t chip_id, log_degree;
  Ef cumulative_sum;
  std::vector<Ef> prep_local, prep_next, main_local, main_next, perm_local, perm_next;  // perm: 4*e base cols
  std::vector<Ef> quotient[2];  // 4 base columns per chunk
};
struct QueryProof {
  std::vector<BatchOpening> inputs;  // per round
  struct Step { Ef sibling; std::vector<Digest> path; };
  std::vector<Step> steps;
};
struct ShardProof {
  uint32_t shard_index;
  Digest main_commit, perm_commit, quot_commit;
  std::vector<ChipOpening> chips;  // in chip_ordering order (height desc)
  std::vector<Fp> public_values;
  std::vector<Digest> fri_commits;
  Ef final_poly;
  Fp pow_witness;
  std::vector<QueryProof> queries;
};

// ---- serialisation (canonical u32 words) --------------------------------------------------------------
struct Writer {
  std::vector<uint32_t> w;
  void u(uint32_t x) { w.push_back(x); }
  void f(Fp x) { w.push_back(x.to_u32()); }
  void e(const Ef& x) { for (Fp c : x.c) f(c); }
  void d(const Digest& x) { for (Fp c : x.d) f(c); }
  void ev(const std::vector<Ef>& v) { u((uint32_t)v.size()); for (auto& x : v) e(x); }
};
struct Reader {
  const uint32_t* w;
  size_t n, p = 0;
  uint32_t u() { if (p >= n) throw std::runtime_error("proof truncated"); return w[p++]; }
  Fp f() { uint32_t x = u(); if (x >= P) throw std::runtime_error("non-canonical field element"); return Fp::from_u32(x); }
  Ef e() { Ef x; for (auto& c : x.c) c = f(); return x; }
  Digest d() { Digest x; for (auto& c : x.d) c = f(); return x; }
  std::vector<Ef> ev(size_t cap = 1 << 16) {
    uint32_t k = u();
    if (k > cap) throw std::runtime_error("vector too long");
    std::vector<Ef> v(k);
    for (auto& x : v) x = e();
    return v;
  }
};
static const uint32_t PROOF_MAGIC = 0x4c505246u;  // 'LPRF'

static inline void write_shard_proof(Writer& o, const ShardProof& sp) {
  o.u(sp.shard_index);
  o.d(sp.main_commit); o.d(sp.perm_commit); o.d(sp.quot_commit);
  o.u((uint32_t)sp.chips.size());
  for (auto& c : sp.chips) {
    o.u(c.chip_id); o.u(c.log_degree); o.e(c.cumulative_sum);
    o.ev(c.prep_local); o.ev(c.prep_next); o.ev(c.main_local); o.ev(c.main_next);
    o.ev(c.perm_local); o.ev(c.perm_next); o.ev(c.quotient[0]); o.ev(c.quotient[1]);
  }
  o.u((uint32_t)sp.public_values.size());
  for (Fp x : sp.public_values) o.f(x);
  o.u((uint32_t)sp.fri_commits.size());
  for (auto& d : sp.fri_commits) o.d(d);
  o.e(sp.final_poly);
  o.f(sp.pow_witness);
  o.u((uint32_t)sp.queries.size());
  for (auto& q : sp.queries) {
    o.u((uint32_t)q.inputs.size());
    for (auto& bo : q.inputs) {
      o.u((uint32_t)bo.rows.size());
      for (auto& r : bo.rows) { o.u((uint32_t)r.size()); for (Fp x : r) o.f(x); }
      o.u((uint32_t)bo.path.size());
      for (auto& d : bo.path) o.d(d);
    }
    o.u((uint32_t)q.steps.size());
    for (auto& s : q.steps) {
      o.e(s.sibling);
      o.u((uint32_t)s.path.size());
      for (auto& d : s.path) o.d(d);
    }
  }
}
static inline ShardProof read_shard_proof(Reader& r) {
  ShardProof sp;
  sp.shard_index = r.u();
  sp.main_commit = r.d(); sp.perm_commit = r.d(); sp.quot_commit = r.d();
  uint32_t nc = r.u();
  if (nc > 4096) throw std::runtime_error("too many chips");
  sp.chips.resize(nc);
  for (auto& c : sp.chips) {
    c.chip_id = r.u(); c.log_degree = r.u(); c.cumulative_sum = r.e();
    c.prep_local = r.ev(); c.prep_next = r.ev(); c.main_local = r.ev(); c.main_next = r.ev();
    c.perm_local = r.ev(); c.perm_next = r.ev(); c.quotient[0] = r.ev(); c.quotient[1] = r.ev();
  }
  uint32_t npv = r.u();
  if (npv > 4096) throw std::runtime_error("too many public values");
  for (uint32_t i = 0; i < npv; i++) sp.public_values.push_back(r.f());
  uint32_t nf = r.u();
  if (nf > 64) throw std::runtime_error("too many FRI layers");
  for (uint32_t i = 0; i < nf; i++) sp.fri_commits.push_back(r.d());
  sp.final_poly = r.e();
  sp.pow_witness = r.f();
  uint32_t nq = r.u();
  if (nq > 1024) throw std::runtime_error("too many queries");
  sp.queries.resize(nq);
  for (auto& q : sp.queries) {
    uint32_t nr = r.u();
    if (nr > 8) throw std::runtime_error("too many rounds");
    q.inputs.resize(nr);
    for (auto& bo : q.inputs) {
      uint32_t nm = r.u();
      if (nm > 4096) throw std::runtime_error("too many matrices");
      bo.rows.resize(nm);
      for (auto& row : bo.rows) {
        uint32_t k = r.u();
        if (k > (1 << 16)) throw std::runtime_error("row too wide");
        for (uint32_t i = 0; i < k; i++) row.push_back(r.f());
      }
      uint32_t pl = r.u();
      if (pl > 64) throw std::runtime_error("path too long");
      for (uint32_t i = 0; i < pl; i++) bo.path.push_back(r.d());
    }
    uint32_t ns = r.u();
    if (ns > 64) throw std::runtime_error("too many steps");
    q.steps.resize(ns);
    for (auto& s : q.steps) {
      s.sibling = r.e();
      uint32_t pl = r.u();
      if (pl > 64) throw std::runtime_error("path too long");
      for (uint32_t i = 0; i < pl; i++) s.path.push_back(r.d());
    }
  }
  return sp;
}

}  // namespace orc
