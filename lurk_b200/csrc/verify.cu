// A14: the verifier -- `machine.verify(&vk, &proof, &mut challenger)` (src/core/cli/repl.rs:186,198,
// src/core/cli/meta.rs:928-932, src/lair/lair_chip.rs:273-275, tests/fib.rs:96).
//
// In the reference this is third-party code (sphinx-core StarkMachine::verify -> Verifier::verify_shard
// -> TwoAdicFriPcs::verify) and runs on the CPU in milliseconds; it stays on the CPU
// here too (host C++ only -- this file holds no kernel and needs no device, so proofs can be checked
// on a machine without a GPU).  Written against the proof layout of DESIGN.md section 4:
//   1. transcript: preprocessed commitment, pc_start, then per shard main commitment + public values;
//   2. per shard: LogUp challenges, permutation / quotient commitments, zeta, opened values,
//      FRI batching challenge, FRI layer commitments + betas, final value, proof-of-work, queries;
//   3. per query: Merkle openings of every committed matrix at the query index (mixed heights),
//      reduced openings per height, arity-2 folds against the FRI layer openings;
//   4. per chip: the captured AIR (constraint program + LogUp constraints) evaluated on the opened
//      values at zeta must equal quotient(zeta) * Z_H(zeta);
//   5. the LogUp cumulative sums of all chips of all shards cancel.
#include <algorithm>
#include <cstring>
#include <memory>

#include "challenger.h"
#include "verify.h"

namespace lsk {
namespace {

struct VFail {
  std::string msg;
};
#define VCHECK(cond, ...)                     \
  do {                                        \
    if (!(cond)) {                            \
      char b_[256];                           \
      snprintf(b_, sizeof b_, __VA_ARGS__);   \
      throw VFail{b_};                        \
    }                                         \
  } while (0)

struct Words {  // bounds-checked cursor over the proof
  const uint32_t* w;
  size_t n, p = 0;
  uint32_t u() {
    VCHECK(p < n, "proof truncated");
    return w[p++];
  }
  uint32_t felt() {  // canonical residue -> Montgomery
    uint32_t x = u();
    VCHECK(x < P, "non-canonical field element in proof");
    return to_mont(x);
  }
  Ext ext() {
    Ext e;
    for (int k = 0; k < 4; k++) e.c[k] = felt();
    return e;
  }
  void digest(uint32_t* d) {
    for (int k = 0; k < 8; k++) d[k] = felt();
  }
};

using Digest = std::array<uint32_t, 8>;

struct ChipOpening {
  uint32_t chip = 0, logn = 0;
  Ext cum;
  std::vector<Ext> prep[2], main[2], perm[2], quot[2];  // [local, next] / the two quotient chunks
};
struct MatOpening {
  std::vector<uint32_t> row;
};
struct RoundOpening {
  std::vector<MatOpening> mats;
  std::vector<Digest> path;
};
struct FriStep {
  Ext sibling;
  std::vector<Digest> path;
};
struct Query {
  std::vector<RoundOpening> rounds;
  std::vector<FriStep> steps;
};
struct ShardProofV {
  uint32_t index = 0;
  Digest main_root, perm_root, quot_root;
  std::vector<ChipOpening> chips;
  std::vector<uint32_t> pv;
  std::vector<Digest> fri_roots;
  Ext final_poly;
  uint32_t pow_witness = 0;
  std::vector<Query> queries;
};

std::vector<Ext> read_opened(Words& r, size_t limit) {
  uint32_t n = r.u();
  VCHECK(n <= limit, "opened-value vector too long");
  std::vector<Ext> v(n);
  for (auto& e : v) e = r.ext();
  return v;
}
std::vector<Digest> read_path(Words& r) {
  uint32_t n = r.u();
  VCHECK(n <= 32, "Merkle path too long");
  std::vector<Digest> p(n);
  for (auto& d : p) r.digest(d.data());
  return p;
}

ShardProofV read_shard(Words& r, const VerifyingKey& vk) {
  ShardProofV sp;
  sp.index = r.u();
  r.digest(sp.main_root.data());
  r.digest(sp.perm_root.data());
  r.digest(sp.quot_root.data());
  uint32_t nchips = r.u();
  VCHECK(nchips >= 1 && nchips <= vk.chips.size(), "bad chip count");
  sp.chips.resize(nchips);
  for (auto& co : sp.chips) {
    co.chip = r.u();
    co.logn = r.u();
    co.cum = r.ext();
    for (int k = 0; k < 2; k++) co.prep[k] = read_opened(r, 1 << 16);
    for (int k = 0; k < 2; k++) co.main[k] = read_opened(r, 1 << 16);
    for (int k = 0; k < 2; k++) co.perm[k] = read_opened(r, 1 << 16);
    for (int k = 0; k < 2; k++) co.quot[k] = read_opened(r, 4);
  }
  uint32_t npv = r.u();
  VCHECK(npv == vk.num_pv, "public value count");
  sp.pv.resize(npv);
  for (auto& x : sp.pv) x = r.felt();
  uint32_t nfri = r.u();
  VCHECK(nfri <= 32, "too many FRI layers");
  sp.fri_roots.resize(nfri);
  for (auto& d : sp.fri_roots) r.digest(d.data());
  sp.final_poly = r.ext();
  sp.pow_witness = r.felt();
  uint32_t nq = r.u();
  VCHECK(nq == vk.num_queries, "query count");
  sp.queries.resize(nq);
  for (auto& q : sp.queries) {
    uint32_t nr = r.u();
    VCHECK(nr >= 3 && nr <= 4, "round count");
    q.rounds.resize(nr);
    for (auto& ro : q.rounds) {
      uint32_t nm = r.u();
      VCHECK(nm <= 2 * vk.chips.size(), "opened matrix count");
      ro.mats.resize(nm);
      for (auto& mo : ro.mats) {
        uint32_t w = r.u();
        VCHECK(w <= (1u << 18), "opened row too wide");
        mo.row.resize(w);
        for (auto& x : mo.row) x = r.felt();
      }
      ro.path = read_path(r);
    }
    uint32_t ns = r.u();
    VCHECK(ns == nfri, "FRI step count");
    q.steps.resize(ns);
    for (auto& s : q.steps) {
      s.sibling = r.ext();
      s.path = read_path(r);
    }
  }
  return sp;
}

// ---- hashing (PaddingFreeSponge<Perm16,16,8,8>, TruncatedPermutation<Perm16,2,8,16>) -------------------
struct Sponge {
  const HostPerm16& perm;
  uint32_t st[16];
  int fill = 0;
  explicit Sponge(const HostPerm16& p) : perm(p) { memset(st, 0, sizeof st); }
  void absorb(uint32_t x) {
    st[fill++] = x;
    if (fill == 8) perm.permute(st), fill = 0;
  }
  Digest finish() {
    if (fill) perm.permute(st), fill = 0;
    Digest d;
    std::copy(st, st + 8, d.begin());
    return d;
  }
};
Digest compress(const HostPerm16& perm, const Digest& l, const Digest& r) {
  uint32_t st[16];
  std::copy(l.begin(), l.end(), st);
  std::copy(r.begin(), r.end(), st + 8);
  perm.permute(st);
  Digest d;
  std::copy(st, st + 8, d.begin());
  return d;
}

// FieldMerkleTreeMmcs::verify_batch: matrices grouped by height (tallest first, stable); the rows of a
// shorter group are hashed and compressed into the running node when the walk reaches their height
bool merkle_verify(const HostPerm16& perm, const Digest& root, const std::vector<unsigned>& log_heights, size_t index,
                   const std::vector<const std::vector<uint32_t>*>& rows, const std::vector<Digest>& path) {
  std::vector<size_t> order(log_heights.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return log_heights[a] > log_heights[b]; });
  if (order.empty()) return false;
  size_t pos = 0;
  unsigned cur = log_heights[order[0]];
  if (path.size() != cur) return false;
  auto hash_group = [&](unsigned lh) {
    Sponge sp(perm);
    while (pos < order.size() && log_heights[order[pos]] == lh) {
      for (uint32_t x : *rows[order[pos]]) sp.absorb(x);
      pos++;
    }
    return sp.finish();
  };
  Digest node = hash_group(cur);
  for (const Digest& sib : path) {
    node = (index & 1) ? compress(perm, sib, node) : compress(perm, node, sib);
    index >>= 1;
    cur--;
    if (pos < order.size() && log_heights[order[pos]] == cur) node = compress(perm, node, hash_group(cur));
  }
  return pos == order.size() && node == root;
}

// ---- AIR at zeta ------------------------------------------------------------------------------------
Ext unflatten(const Ext* four) {  // sum_k x^k * four[k]: an extension column was committed as 4 base columns
  Ext r = four[0];
  for (int k = 1; k < 4; k++) {
    // x^k * e: rotate the coefficients, multiplying the ones that wrap by 11
    const Ext& e = four[k];
    Ext s;
    for (int j = 0; j < 4; j++) {
      int src = (j - k + 4) & 3;
      s.c[j] = j < k ? fmul11(e.c[src]) : e.c[src];
    }
    r = eadd(r, s);
  }
  return r;
}

struct PointView {
  const ChipH* chip;
  const ChipOpening* co;
  const std::vector<uint32_t>* pv;
  Ext sel[3];  // is_first, is_last, is_transition
};
Ext eval_affine(const AffineH& a, const PointView& v) {
  Ext acc = ext_from_base(to_mont(a.cst % P));
  for (auto& t : a.terms) {
    const Ext& cell = t[0] ? v.co->prep[0][t[1]] : v.co->main[0][t[1]];
    acc = eadd(acc, emul_base(cell, to_mont(t[2] % P)));
  }
  return acc;
}

Ext constraints_at_zeta(const ChipH& c, const PointView& v, const Ext& alpha, const Ext& chal_alpha, const Ext& chal_beta) {
  std::vector<Ext> regs(std::max<uint32_t>(c.n_regs, 1), ext_zero());
  auto operand = [&](uint32_t w) -> Ext {
    uint32_t idx = w & 0x0fffffffu;
    switch (w >> 28) {
      case 0: VCHECK(idx < regs.size(), "bad register"); return regs[idx];
      case 1: VCHECK(idx < c.consts.size(), "bad constant"); return ext_from_base(to_mont(c.consts[idx] % P));
      case 2: VCHECK(idx < c.width, "bad column"); return v.co->main[0][idx];
      case 3: VCHECK(idx < c.width, "bad column"); return v.co->main[1][idx];
      case 4: VCHECK(idx < c.prep_width, "bad column"); return v.co->prep[0][idx];
      case 5: VCHECK(idx < c.prep_width, "bad column"); return v.co->prep[1][idx];
      case 6: VCHECK(idx < v.pv->size(), "bad public value"); return ext_from_base((*v.pv)[idx]);
      case 7: VCHECK(idx < 3, "bad selector"); return v.sel[idx];
      default: throw VFail{"bad operand"};
    }
  };
  // constraints are folded Horner-style: acc = acc * alpha + c_i, i.e. the first one gets the highest power
  Ext acc = ext_zero();
  for (size_t i = 0; i + 3 < c.code.size(); i += 4) {
    uint32_t op = c.code[i], dst = c.code[i + 1];
    if (op != COP_ASSERT) VCHECK(dst < regs.size(), "bad destination");
    switch (op) {
      case COP_ADD: regs[dst] = eadd(operand(c.code[i + 2]), operand(c.code[i + 3])); break;
      case COP_SUB: regs[dst] = esub(operand(c.code[i + 2]), operand(c.code[i + 3])); break;
      case COP_MUL: regs[dst] = emul(operand(c.code[i + 2]), operand(c.code[i + 3])); break;
      case COP_NEG: regs[dst] = eneg(operand(c.code[i + 2])); break;
      case COP_ASSERT: acc = eadd(emul(acc, alpha), operand(c.code[i + 2])); break;
      default: throw VFail{"bad constraint opcode"};
    }
  }
  // LogUp (sphinx eval_permutation_constraints): per batch of two interactions
  //   perm_b * rlc_0 * rlc_1 = m_0 * rlc_1 + m_1 * rlc_0, then the running-sum column phi:
  //   transition: phi' - phi = sum_b perm'_b;  first row: phi = sum_b perm_b;  last row: phi = cumulative sum
  const size_t pw = c.perm_width();
  std::vector<Ext> pl(pw), pn(pw);
  for (size_t j = 0; j < pw; j++) pl[j] = unflatten(&v.co->perm[0][4 * j]), pn[j] = unflatten(&v.co->perm[1][4 * j]);
  std::vector<Ext> beta_pows(std::max<size_t>(c.max_values(), 1));
  {
    Ext p = chal_beta;
    for (auto& b : beta_pows) b = p, p = emul(p, chal_beta);
  }
  auto interaction = [&](size_t k, Ext& rlc, Ext& m) {
    bool recv = k >= c.sends.size();
    const InteractionH& it = recv ? c.receives[k - c.sends.size()] : c.sends[k];
    m = eval_affine(it.mult, v);
    if (recv) m = eneg(m);
    rlc = eadd_base(chal_alpha, to_mont(it.kind % P));
    for (size_t j = 0; j < it.values.size(); j++) rlc = eadd(rlc, emul(beta_pows[j], eval_affine(it.values[j], v)));
  };
  const size_t nint = c.n_interactions();
  Ext sum_local = ext_zero(), sum_next = ext_zero();
  for (size_t b = 0; 2 * b < nint; b++) {
    Ext r0, m0, lhs, rhs;
    interaction(2 * b, r0, m0);
    if (2 * b + 1 < nint) {
      Ext r1, m1;
      interaction(2 * b + 1, r1, m1);
      lhs = emul(emul(pl[b], r0), r1);
      rhs = eadd(emul(r1, m0), emul(r0, m1));
    } else {
      lhs = emul(pl[b], r0);
      rhs = m0;
    }
    acc = eadd(emul(acc, alpha), esub(lhs, rhs));
    sum_local = eadd(sum_local, pl[b]);
    sum_next = eadd(sum_next, pn[b]);
  }
  const Ext &phi_l = pl[pw - 1], &phi_n = pn[pw - 1];
  acc = eadd(emul(acc, alpha), emul(esub(esub(phi_n, phi_l), sum_next), v.sel[2]));
  acc = eadd(emul(acc, alpha), emul(esub(phi_l, sum_local), v.sel[0]));
  acc = eadd(emul(acc, alpha), emul(esub(phi_l, v.co->cum), v.sel[1]));
  return acc;
}

void verify_shard(const VerifyingKey& vk, const ShardProofV& sp, Challenger ch) {
  const HostPerm16& perm = vk.perm;
  // shapes and chip ordering (height descending)
  std::vector<bool> seen(vk.chips.size(), false);
  uint32_t prev = 64;
  for (auto& co : sp.chips) {
    VCHECK(co.chip < vk.chips.size() && !seen[co.chip], "bad or repeated chip id");
    seen[co.chip] = true;
    const ChipH& c = vk.chips[co.chip];
    VCHECK(co.logn <= 26 && co.logn <= prev, "chips are not ordered by height");
    prev = co.logn;
    for (int k = 0; k < 2; k++) {
      VCHECK(co.main[k].size() == c.width, "main opening width of %s", c.name.c_str());
      VCHECK(co.prep[k].size() == c.prep_width, "preprocessed opening width of %s", c.name.c_str());
      VCHECK(co.perm[k].size() == 4 * c.perm_width(), "permutation opening width of %s", c.name.c_str());
      VCHECK(co.quot[k].size() == 4, "quotient opening width");
    }
    if (c.kind == CHIP_BYTES) VCHECK(co.logn == 16, "byte table height");
    if (c.kind == CHIP_ENTRY) VCHECK(co.logn == 0 && sp.index == 0, "entrypoint shape");
    if (c.kind == CHIP_MEM) VCHECK(sp.index == 0, "memory chip outside shard 0");
  }
  // The chip set of a shard is not the prover's choice (MachineAir::included, src/lair/lair_chip.rs:124-139): the
  // Entrypoint -- the only chip that ties the public values to the computation (lair_chip.rs:166-191) -- and every
  // memory table belong to shard 0 and to no other shard, the byte table to every shard (sphinx likewise insists on
  // its "CPU" chip, which is why the byte chip carries that name: lair_chip.rs:90-92).  Func chips are present when
  // they have rows; leaving one out only removes provides, which the cumulative-sum check then misses.
  for (size_t i = 0; i < vk.chips.size(); i++) {
    const ChipH& c = vk.chips[i];
    const bool must = c.kind == CHIP_BYTES || (sp.index == 0 && (c.kind == CHIP_ENTRY || c.kind == CHIP_MEM));
    VCHECK(!must || seen[i], "shard %u lacks the %s chip", sp.index, c.name.c_str());
  }
  Ext chal_alpha = ch.sample_ext(), chal_beta = ch.sample_ext();
  ch.observe_n(sp.perm_root.data(), 8);
  Ext alpha = ch.sample_ext();
  ch.observe_n(sp.quot_root.data(), 8);
  Ext zeta = ch.sample_ext();

  // rounds in PCS order: preprocessed (if any chip has one), main, permutation, quotient
  struct VMat {
    unsigned lh;                       // log2 of the LDE height
    std::vector<Ext> points;
    std::vector<const std::vector<Ext>*> vals;
  };
  struct VRound {
    Digest commit;
    std::vector<VMat> mats;
  };
  VRound r_prep{vk.prep_commit, {}}, r_main{sp.main_root, {}}, r_perm{sp.perm_root, {}}, r_quot{sp.quot_root, {}};
  for (auto& co : sp.chips) {
    const ChipH& c = vk.chips[co.chip];
    unsigned lh = co.logn + 1;
    Ext zn = emul_base(zeta, two_adic_generator(co.logn));
    if (c.prep_width) r_prep.mats.push_back({lh, {zeta, zn}, {&co.prep[0], &co.prep[1]}});
    r_main.mats.push_back({lh, {zeta, zn}, {&co.main[0], &co.main[1]}});
    r_perm.mats.push_back({lh, {zeta, zn}, {&co.perm[0], &co.perm[1]}});
    for (int j = 0; j < 2; j++) r_quot.mats.push_back({lh, {zeta}, {&co.quot[j]}});
  }
  std::vector<VRound> rounds;
  if (!r_prep.mats.empty()) {
    VCHECK(vk.has_prep, "proof opens a preprocessed trace the key does not have");
    rounds.push_back(std::move(r_prep));
  }
  rounds.push_back(std::move(r_main));
  rounds.push_back(std::move(r_perm));
  rounds.push_back(std::move(r_quot));
  for (auto& rd : rounds)
    for (auto& mt : rd.mats)
      for (auto* vals : mt.vals)
        for (auto& x : *vals) ch.observe(x);
  Ext fri_alpha = ch.sample_ext();

  unsigned log_max = 0;
  for (auto& rd : rounds)
    for (auto& mt : rd.mats) log_max = std::max(log_max, mt.lh);
  VCHECK(log_max > 1, "degenerate shard");
  VCHECK(sp.fri_roots.size() == log_max - 1, "FRI layer count");
  std::vector<Ext> betas;
  for (auto& root : sp.fri_roots) {
    ch.observe_n(root.data(), 8);
    betas.push_back(ch.sample_ext());
  }
  ch.observe(sp.final_poly);
  VCHECK(ch.check_witness(vk.pow_bits, sp.pow_witness), "invalid proof-of-work witness");

  const uint32_t half = finv(to_mont(2)), gen = to_mont(GEN);
  for (auto& q : sp.queries) {
    size_t index = ch.sample_bits(log_max);
    VCHECK(q.rounds.size() == rounds.size(), "round count");
    std::vector<Ext> ro(log_max + 1, ext_zero()), apow(log_max + 1, ext_one());
    for (size_t ri = 0; ri < rounds.size(); ri++) {
      const VRound& rd = rounds[ri];
      const RoundOpening& op = q.rounds[ri];
      VCHECK(op.mats.size() == rd.mats.size(), "opened matrix count");
      unsigned lt = 0;
      std::vector<unsigned> lhs;
      std::vector<const std::vector<uint32_t>*> rows;
      for (size_t mi = 0; mi < rd.mats.size(); mi++) {
        VCHECK(op.mats[mi].row.size() == rd.mats[mi].vals[0]->size(), "opened row width");
        lt = std::max(lt, rd.mats[mi].lh);
        lhs.push_back(rd.mats[mi].lh);
        rows.push_back(&op.mats[mi].row);
      }
      VCHECK(merkle_verify(perm, rd.commit, lhs, index >> (log_max - lt), rows, op.path), "input Merkle path (round %zu)", ri);
      for (size_t mi = 0; mi < rd.mats.size(); mi++) {
        const VMat& mt = rd.mats[mi];
        size_t r = index >> (log_max - mt.lh);
        uint32_t x = fmul(gen, fpow(two_adic_generator(mt.lh), bitrev32((uint32_t)r, mt.lh)));
        for (size_t pi = 0; pi < mt.points.size(); pi++) {
          Ext inv = einv(esub_base(mt.points[pi], x));
          const std::vector<Ext>& pz = *mt.vals[pi];
          for (size_t c = 0; c < pz.size(); c++) {
            ro[mt.lh] = eadd(ro[mt.lh], emul(emul(apow[mt.lh], esub_base(pz[c], op.mats[mi].row[c])), inv));
            apow[mt.lh] = emul(apow[mt.lh], fri_alpha);
          }
        }
      }
    }
    Ext folded = ext_zero();
    for (size_t li = 0; li < q.steps.size(); li++) {
      unsigned lfh = log_max - 1 - (unsigned)li;
      size_t idx = index >> li, pair = idx >> 1;
      folded = eadd(folded, ro[lfh + 1]);
      Ext ev[2];
      ev[idx & 1] = folded;
      ev[(idx & 1) ^ 1] = q.steps[li].sibling;
      std::vector<uint32_t> row(8);
      for (int k = 0; k < 4; k++) row[k] = ev[0].c[k], row[4 + k] = ev[1].c[k];
      VCHECK(merkle_verify(perm, sp.fri_roots[li], {lfh}, pair, {&row}, q.steps[li].path), "FRI layer %zu Merkle path", li);
      uint32_t xinv = finv(fpow(two_adic_generator(lfh + 1), bitrev32((uint32_t)pair, lfh)));
      Ext t = emul_base(betas[li], fmul(half, xinv));
      folded = eadd(emul(eadd_base(t, half), ev[0]), emul(eadd_base(eneg(t), half), ev[1]));
    }
    folded = eadd(folded, ro[1]);
    VCHECK(ext_eq(ro[0], ext_zero()), "matrix below the blow-up height");
    VCHECK(ext_eq(folded, sp.final_poly), "FRI final value mismatch");
  }

  for (auto& co : sp.chips) {
    const ChipH& c = vk.chips[co.chip];
    const unsigned logn = co.logn;
    const uint64_t n = (uint64_t)1 << logn;
    uint32_t wn_inv = finv(two_adic_generator(logn));
    Ext zh = esub_base(epow(zeta, n), R1);
    PointView v{&c, &co, &sp.pv, {}};
    v.sel[0] = emul(zh, einv(esub_base(zeta, R1)));
    v.sel[1] = emul(zh, einv(esub_base(zeta, wn_inv)));
    v.sel[2] = esub_base(zeta, wn_inv);
    Ext folded = constraints_at_zeta(c, v, alpha, chal_alpha, chal_beta);
    // quotient(zeta) from its two chunks over the cosets g*H and g*w_2N*H
    uint32_t w2n = two_adic_generator(logn + 1);
    uint32_t shifts[2] = {gen, fmul(gen, w2n)};
    Ext quotient = ext_zero();
    for (int j = 0; j < 2; j++) {
      int o = 1 - j;
      Ext num = esub_base(epow(emul_base(zeta, finv(shifts[o])), n), R1);
      uint32_t den = fsub(fpow(fmul(shifts[j], finv(shifts[o])), n), R1);
      quotient = eadd(quotient, emul(emul_base(num, finv(den)), unflatten(co.quot[j].data())));
    }
    VCHECK(ext_eq(folded, emul(quotient, zh)), "constraints do not match the quotient for chip %s", c.name.c_str());
  }
}

}  // namespace

void verify_proof(const VerifyingKey& vk, const uint32_t* words, size_t n) {
  try {
    Words r{words, n};
    VCHECK(r.u() == 0x4c505246u, "bad proof magic");
    uint32_t ns = r.u();
    VCHECK(ns >= 1 && ns <= 4096, "bad shard count");
    std::vector<ShardProofV> sps;
    for (uint32_t i = 0; i < ns; i++) sps.push_back(read_shard(r, vk));
    VCHECK(r.p == n, "trailing data after the proof");
    Challenger ch(&vk.perm);
    ch.observe_n(vk.prep_commit.data(), 8);
    ch.observe(0);  // pc_start, src/lair/lair_chip.rs:71-77
    for (uint32_t i = 0; i < ns; i++) {
      VCHECK(sps[i].index == i, "shard index");
      VCHECK(sps[i].pv == sps[0].pv, "public values differ between shards");
      ch.observe_n(sps[i].main_root.data(), 8);
      ch.observe_n(sps[i].pv.data(), sps[i].pv.size());
    }
    Ext total = ext_zero();
    for (auto& sp : sps) {
      verify_shard(vk, sp, ch);
      for (auto& co : sp.chips) total = eadd(total, co.cum);
    }
    VCHECK(ext_eq(total, ext_zero()), "lookup argument: cumulative sums do not cancel");
  } catch (VFail& f) {
    fail(LSK_ERR_VERIFY, "%s", f.msg.c_str());
  }
}

}  // namespace lsk

using namespace lsk;

struct lsk_vk {
  VerifyingKey key;
};

extern "C" {

int lsk_vk_new(const uint32_t* blob, size_t n, const uint32_t prep_commit[8], int has_prep, uint32_t num_queries,
               uint32_t pow_bits, lsk_vk** out) {
  LSK_API_BEGIN
  if (!blob || !out || !prep_commit) fail(LSK_ERR_ARG, "null argument");
  auto vk = std::make_unique<lsk_vk>();
  if (num_queries == 0 || num_queries > 1024 || pow_bits > 30) fail(LSK_ERR_ARG, "bad FRI parameters");
  vk->key.chips = parse_machine_blob(blob, n, &vk->key.num_pv);
  for (int k = 0; k < 8; k++) {
    if (prep_commit[k] >= P) fail(LSK_ERR_ARG, "non-canonical commitment");
    vk->key.prep_commit[k] = to_mont(prep_commit[k]);
  }
  vk->key.has_prep = has_prep != 0;
  vk->key.num_queries = num_queries;
  vk->key.pow_bits = pow_bits;
  *out = vk.release();
  return LSK_OK;
  LSK_API_END
}
void lsk_vk_free(lsk_vk* vk) { delete vk; }
int lsk_vk_verify(const lsk_vk* vk, const uint32_t* words, size_t len) {
  LSK_API_BEGIN
  if (!vk || !words) fail(LSK_ERR_ARG, "null argument");
  verify_proof(vk->key, words, len);
  return LSK_OK;
  LSK_API_END
}

}  // extern "C"
