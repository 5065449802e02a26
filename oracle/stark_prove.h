// ORACLE (test infrastructure, NOT product code).  CPU prover and verifier on top of stark.h.
// Same provenance and the same "PARITY UNPINNED" caveat as stark.h: the protocol is restated from
// the SP1-v1 / Plonky3-2024 lineage that sphinx-core @ 8a39b95 derives from (third party, absent
// from /root/reference); the reference's call sites are `machine.prove::<LocalProver<_, _>>` and
// `machine.verify` (src/core/cli/repl.rs:186-199, src/lair/lair_chip.rs:258-275).
#pragma once
#include "stark.h"

namespace orc {

static inline Ef ext_x_pow(int k) {  // the basis element x^k of F[x]/(x^4-11)
  Ef e = Ef::zero();
  e.c[k] = Fp::one();
  return e;
}
static inline Ef unflatten(const Ef* four) {  // sum_k x^k * four[k]
  Ef acc = Ef::zero();
  for (int k = 0; k < 4; k++) acc += four[k] * ext_x_pow(k);
  return acc;
}
static inline void batch_inverse(std::vector<Ef>& v) {
  size_t n = v.size();
  if (!n) return;
  std::vector<Ef> pre(n);
  Ef acc = Ef::one();
  for (size_t i = 0; i < n; i++) pre[i] = acc, acc = acc * v[i];
  Ef inv = ef_inv(acc);
  for (size_t i = n; i-- > 0;) {
    Ef t = inv * pre[i];
    inv = inv * v[i];
    v[i] = t;
  }
}

// ---- shard traces ---------------------------------------------------------------------------------
struct ShardData {
  uint32_t shard;
  std::vector<uint32_t> chip_ids;  // sorted by height desc (stable) == chip_ordering
  std::vector<Mat> main;           // natural row order, on H_N
  std::vector<Mat> main_lde;
  MerkleTree main_tree;
};

static inline bool chip_included(const Machine& m, const Chip& c, const QueryRecord& q, uint32_t shard, uint32_t ss) {
  if ((m.forge_omit_kinds >> c.kind) & 1u) return false;
  switch (c.kind) {  // MachineAir::included, src/lair/lair_chip.rs:124-139
    case KIND_FUNC: return shard_func_range(q, c.index, shard, ss).len() > 0;
    case KIND_MEM: return shard == 0;
    case KIND_ENTRY: return shard == 0;
    default: return true;
  }
}
static inline Mat chip_trace(const Machine& m, const Chip& c, const QueryRecord& q, uint32_t shard, uint32_t ss) {
  size_t h = 0;
  switch (c.kind) {  // MachineAir::generate_trace, src/lair/lair_chip.rs:96-120
    case KIND_FUNC: {
      auto rows = func_trace(*m.top, q, c.index, shard, ss, &h);
      return from_row_major(rows, h, c.width);
    }
    case KIND_MEM: {
      auto rows = mem_trace(q, (int)c.index, &h);
      return from_row_major(rows, h, c.width);
    }
    case KIND_BYTES: return from_row_major(bytes_trace(q, shard), 1 << 16, 13);
    default: {
      if (q.public_values.size() != c.width) throw std::runtime_error("public value count mismatch");
      return from_row_major(q.public_values, 1, c.width);
    }
  }
}
static inline ShardData generate_shard(const Machine& m, const QueryRecord& q, uint32_t shard, uint32_t ss) {
  ShardData sd;
  sd.shard = shard;
  std::vector<uint32_t> ids;
  std::vector<Mat> traces;
  for (uint32_t i = 0; i < m.chips.size(); i++)
    if (chip_included(m, m.chips[i], q, shard, ss)) {
      ids.push_back(i);
      traces.push_back(chip_trace(m, m.chips[i], q, shard, ss));
    }
  std::vector<size_t> order(ids.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return traces[a].h > traces[b].h; });
  for (size_t o : order) {
    sd.chip_ids.push_back(ids[o]);
    sd.main.push_back(std::move(traces[o]));
  }
  for (auto& t : sd.main) sd.main_lde.push_back(coset_lde(t));
  std::vector<const Mat*> ptrs;
  for (auto& l : sd.main_lde) ptrs.push_back(&l);
  sd.main_tree = merkle_commit(ptrs);
  return sd;
}

// ---- LogUp permutation trace (sphinx generate_permutation_trace) -----------------------------------------
// returns N x (4 * perm_width) base matrix (extension columns flattened), natural rows
static inline Mat permutation_trace(const Chip& c, const Mat& main, const Mat* prep, Ef chal_alpha, Ef chal_beta,
                                    Ef* cumulative_sum) {
  size_t n = main.h, pw = c.perm_width(), batch = (size_t)1 << c.lqd, nint = c.n_interactions();
  std::vector<Ef> denoms(n * nint), mults(n * nint);
  parallel_for(n, [&](size_t r) {
    auto gm = [&](uint32_t col) { return main.at(r, col); };
    auto gp = [&](uint32_t col) { return prep->at(r, col); };
    for (size_t k = 0; k < nint; k++) {
      bool is_send = k < c.sends.size();
      const Interaction& it = is_send ? c.sends[k] : c.receives[k - c.sends.size()];
      Ef d = chal_alpha + Fp::from_u32(it.kind);
      Ef bp = chal_beta;
      for (auto& v : it.values) {
        d = d + bp * v.eval<Fp>(gm, gp);
        bp = bp * chal_beta;
      }
      Fp mlt = it.mult.eval<Fp>(gm, gp);
      denoms[r * nint + k] = d;
      mults[r * nint + k] = Ef::from_base(is_send ? mlt : -mlt);
    }
  }, 1024);
  // batched inversion in independent chunks
  size_t chunk = 4096;
  parallel_for((denoms.size() + chunk - 1) / chunk, [&](size_t ci) {
    size_t b = ci * chunk, e = std::min(denoms.size(), b + chunk);
    std::vector<Ef> tmp(denoms.begin() + b, denoms.begin() + e);
    batch_inverse(tmp);
    std::copy(tmp.begin(), tmp.end(), denoms.begin() + b);
  }, 1);
  std::vector<Ef> perm(n * pw, Ef::zero());
  parallel_for(n, [&](size_t r) {
    Ef row_sum = Ef::zero();
    for (size_t b = 0; b * batch < nint; b++) {
      Ef s = Ef::zero();
      for (size_t k = b * batch; k < std::min(nint, (b + 1) * batch); k++) s += mults[r * nint + k] * denoms[r * nint + k];
      perm[r * pw + b] = s;
      row_sum += s;
    }
    perm[r * pw + pw - 1] = row_sum;  // turned into the running sum below
  }, 1024);
  Ef run = Ef::zero();
  for (size_t r = 0; r < n; r++) {
    run += perm[r * pw + pw - 1];
    perm[r * pw + pw - 1] = run;
  }
  *cumulative_sum = run;
  Mat out(n, 4 * pw);
  for (size_t r = 0; r < n; r++)
    for (size_t j = 0; j < pw; j++)
      for (int k = 0; k < 4; k++) out.at(r, 4 * j + k) = perm[r * pw + j].c[k];
  return out;
}

// ---- quotient ------------------------------------------------------------------------------------------------
// main_lde / perm_lde / prep_lde: evaluations over g*H_2N, bit-reversed rows.  Returns the quotient
// values (extension) in the same bit-reversed order.
static inline std::vector<Ef> quotient_values(const Chip& c, unsigned logn, const Mat& main_lde, const Mat& perm_lde,
                                              const Mat* prep_lde, const std::vector<Fp>& pv, Ef cumulative_sum,
                                              Ef chal_alpha, Ef chal_beta, Ef alpha) {
  size_t n2 = (size_t)2 << logn;
  unsigned l2 = logn + 1;
  std::vector<Ef> out(n2);
  Fp g = fp_generator(), w2n = two_adic_generator(l2);
  Fp wn_inv = fp_inv(two_adic_generator(logn));
  Fp gn = fp_pow(g, (uint64_t)1 << logn);
  size_t pw = c.perm_width();
  parallel_chunks(n2, [&](size_t rb, size_t re) {
    std::vector<Fp> ml(c.width), mn(c.width), pl(c.prep_width), pn(c.prep_width), regs;
    std::vector<Ef> permL(pw), permN(pw);
    for (size_t r = rb; r < re; r++) {
      size_t i = bitrev(r, l2);
      size_t rn = bitrev((i + 2) & (n2 - 1), l2);
      Fp x = g * fp_pow(w2n, i);
      Fp zh = ((i & 1) ? -gn : gn) - Fp::one();  // x^N - 1
      Fp inv_zh = fp_inv(zh);
      RowView<Fp> row;
      row.is_first = zh * fp_inv(x - Fp::one());
      row.is_last = zh * fp_inv(x - wn_inv);
      row.is_transition = x - wn_inv;
      for (size_t k = 0; k < c.width; k++) ml[k] = main_lde.at(r, k), mn[k] = main_lde.at(rn, k);
      for (size_t k = 0; k < c.prep_width; k++) pl[k] = prep_lde->at(r, k), pn[k] = prep_lde->at(rn, k);
      for (size_t j = 0; j < pw; j++)
        for (int k = 0; k < 4; k++) permL[j].c[k] = perm_lde.at(r, 4 * j + k), permN[j].c[k] = perm_lde.at(rn, 4 * j + k);
      row.main_local = ml.data(); row.main_next = mn.data();
      row.prep_local = pl.data(); row.prep_next = pn.data();
      row.pv = pv.data();
      Ef acc = fold_air_constraints<Fp>(c, row, alpha, Ef::zero(), regs);
      acc = fold_perm_constraints<Fp>(c, row, permL.data(), permN.data(), cumulative_sum, chal_alpha, chal_beta, alpha, acc);
      out[r] = acc * inv_zh;
    }
  }, 256);
  return out;
}

// ---- openings ---------------------------------------------------------------------------------------------------
// Barycentric weights for evaluating, at z, a polynomial of degree < N given on the coset g*H_N in
// bit-reversed order (the first half of an LDE): p(z) = sum_r w_r * p_r.
static inline std::vector<Ef> bary_weights(unsigned logn, Ef z) {
  size_t n = (size_t)1 << logn;
  Fp g = fp_generator(), wn = two_adic_generator(logn);
  std::vector<Fp> xs(n);
  Fp p = g;
  std::vector<Fp> nat(n);
  for (size_t k = 0; k < n; k++) nat[k] = p, p = p * wn;
  for (size_t r = 0; r < n; r++) xs[r] = nat[bitrev(r, logn)];
  std::vector<Ef> den(n);
  for (size_t r = 0; r < n; r++) den[r] = z - xs[r];
  batch_inverse(den);
  Fp gn = fp_pow(g, n);
  Ef scale = (ef_pow(z, n) - gn) * fp_inv(gn * Fp::from_u32((uint32_t)n));
  for (size_t r = 0; r < n; r++) den[r] = den[r] * xs[r] * scale;
  return den;
}
static inline std::vector<Ef> open_matrix(const Mat& lde, const std::vector<Ef>& wts) {
  size_t n = wts.size();
  std::vector<Ef> out(lde.w);
  parallel_for(lde.w, [&](size_t c) {
    Ef acc = Ef::zero();
    const Fp* col = lde.col(c);
    for (size_t r = 0; r < n; r++) acc += wts[r] * col[r];
    out[c] = acc;
  }, 1);
  return out;
}

// ---- FRI fold -------------------------------------------------------------------------------------------------------
static inline std::vector<Ef> fri_fold(const std::vector<Ef>& cur, Ef beta) {
  size_t half = cur.size() / 2;
  unsigned logh = log2_strict(half);
  Fp ginv = fp_inv(two_adic_generator(logh + 1));
  Fp one_half = fp_inv(Fp::from_u32(2));
  std::vector<Fp> pw(half);
  Fp p = Fp::one();
  for (size_t k = 0; k < half; k++) pw[k] = p, p = p * ginv;
  std::vector<Ef> out(half);
  Ef half_beta = beta * one_half;
  parallel_for(half, [&](size_t i) {
    Ef t = half_beta * pw[bitrev(i, logh)];
    out[i] = (t + one_half) * cur[2 * i] + (-t + one_half) * cur[2 * i + 1];
  }, 4096);
  return out;
}
static inline Mat fri_layer_matrix(const std::vector<Ef>& cur) {  // (len/2) x 8 base
  size_t half = cur.size() / 2;
  Mat m(half, 8);
  for (size_t i = 0; i < half; i++)
    for (int k = 0; k < 4; k++) m.at(i, k) = cur[2 * i].c[k], m.at(i, 4 + k) = cur[2 * i + 1].c[k];
  return m;
}

// ---- prove ---------------------------------------------------------------------------------------------------------------
struct Round {
  const MerkleTree* tree;
  std::vector<const Mat*> ldes;
  std::vector<std::vector<Ef>> points;            // per matrix
  std::vector<std::vector<std::vector<Ef>>> vals;  // per matrix, per point, per column
};

static inline void observe_machine(const Machine& m, Challenger& ch) {
  ch.observe(m.prep_commit);
  ch.observe(Fp::zero());  // pc_start (LairMachineProgram, src/lair/lair_chip.rs:71-77)
}

static inline ShardProof prove_shard(const Machine& m, const QueryRecord& q, ShardData& sd, Challenger ch) {
  ShardProof sp;
  sp.shard_index = sd.shard;
  sp.public_values = q.public_values;
  sp.main_commit = sd.main_tree.root();
  size_t nchips = sd.chip_ids.size();
  Ef chal_alpha = ch.sample_ext(), chal_beta = ch.sample_ext();
  // permutation traces
  std::vector<Mat> perm(nchips), perm_lde(nchips);
  std::vector<Ef> cums(nchips);
  for (size_t i = 0; i < nchips; i++) {
    const Chip& c = m.chips[sd.chip_ids[i]];
    perm[i] = permutation_trace(c, sd.main[i], c.prep_width ? &m.prep_trace : nullptr, chal_alpha, chal_beta, &cums[i]);
    perm_lde[i] = coset_lde(perm[i]);
  }
  std::vector<const Mat*> ptrs;
  for (auto& l : perm_lde) ptrs.push_back(&l);
  MerkleTree perm_tree = merkle_commit(ptrs);
  sp.perm_commit = perm_tree.root();
  ch.observe(sp.perm_commit);
  Ef alpha = ch.sample_ext();
  // quotient
  std::vector<Mat> quot_lde;
  for (size_t i = 0; i < nchips; i++) {
    const Chip& c = m.chips[sd.chip_ids[i]];
    unsigned logn = log2_strict(sd.main[i].h);
    size_t n = sd.main[i].h;
    auto qv = quotient_values(c, logn, sd.main_lde[i], perm_lde[i], c.prep_width ? &m.prep_lde : nullptr,
                              q.public_values, cums[i], chal_alpha, chal_beta, alpha);
    Fp w2n = two_adic_generator(logn + 1);
    for (int j = 0; j < 2; j++) {
      Mat chunk(n, 4);
      for (size_t k = 0; k < n; k++)
        for (int e = 0; e < 4; e++) chunk.at(k, e) = qv[(size_t)j * n + bitrev(k, logn)].c[e];
      Fp shift = j ? fp_generator() * w2n : fp_generator();
      quot_lde.push_back(coset_lde(chunk, shift));
    }
  }
  ptrs.clear();
  for (auto& l : quot_lde) ptrs.push_back(&l);
  MerkleTree quot_tree = merkle_commit(ptrs);
  sp.quot_commit = quot_tree.root();
  ch.observe(sp.quot_commit);
  Ef zeta = ch.sample_ext();

  // openings
  std::vector<Round> rounds;
  bool has_prep = false;
  for (size_t i = 0; i < nchips; i++) has_prep |= m.chips[sd.chip_ids[i]].prep_width != 0;
  std::map<unsigned, std::pair<std::vector<Ef>, std::vector<Ef>>> wcache;  // logn -> weights at zeta, zeta*w_N
  auto weights = [&](unsigned logn) -> std::pair<std::vector<Ef>, std::vector<Ef>>& {
    auto it = wcache.find(logn);
    if (it == wcache.end()) {
      Ef zn = zeta * two_adic_generator(logn);
      it = wcache.emplace(logn, std::make_pair(bary_weights(logn, zeta), bary_weights(logn, zn))).first;
    }
    return it->second;
  };
  sp.chips.resize(nchips);
  Round r_prep{&m.prep_tree, {}, {}, {}}, r_main{&sd.main_tree, {}, {}, {}}, r_perm{&perm_tree, {}, {}, {}},
      r_quot{&quot_tree, {}, {}, {}};
  for (size_t i = 0; i < nchips; i++) {
    const Chip& c = m.chips[sd.chip_ids[i]];
    unsigned logn = log2_strict(sd.main[i].h);
    auto& w = weights(logn);
    Ef zn = zeta * two_adic_generator(logn);
    ChipOpening& co = sp.chips[i];
    co.chip_id = sd.chip_ids[i];
    co.log_degree = logn;
    co.cumulative_sum = cums[i];
    if (c.prep_width) {
      co.prep_local = open_matrix(m.prep_lde, w.first);
      co.prep_next = open_matrix(m.prep_lde, w.second);
      r_prep.ldes.push_back(&m.prep_lde);
      r_prep.points.push_back({zeta, zn});
      r_prep.vals.push_back({co.prep_local, co.prep_next});
    }
    co.main_local = open_matrix(sd.main_lde[i], w.first);
    co.main_next = open_matrix(sd.main_lde[i], w.second);
    r_main.ldes.push_back(&sd.main_lde[i]);
    r_main.points.push_back({zeta, zn});
    r_main.vals.push_back({co.main_local, co.main_next});
    co.perm_local = open_matrix(perm_lde[i], w.first);
    co.perm_next = open_matrix(perm_lde[i], w.second);
    r_perm.ldes.push_back(&perm_lde[i]);
    r_perm.points.push_back({zeta, zn});
    r_perm.vals.push_back({co.perm_local, co.perm_next});
    for (int j = 0; j < 2; j++) {
      co.quotient[j] = open_matrix(quot_lde[2 * i + j], w.first);
      r_quot.ldes.push_back(&quot_lde[2 * i + j]);
      r_quot.points.push_back({zeta});
      r_quot.vals.push_back({co.quotient[j]});
    }
  }
  if (has_prep) rounds.push_back(r_prep);
  rounds.push_back(r_main);
  rounds.push_back(r_perm);
  rounds.push_back(r_quot);
  for (auto& rd : rounds)
    for (auto& mv : rd.vals)
      for (auto& pv : mv)
        for (auto& x : pv) ch.observe(x);
  Ef fri_alpha = ch.sample_ext();

  // reduced openings per LDE log-height
  unsigned log_max = 0;
  for (auto& rd : rounds)
    for (auto* l : rd.ldes) log_max = std::max(log_max, log2_strict(l->h));
  std::vector<std::vector<Ef>> ro(log_max + 1);
  std::vector<Ef> alpha_off(log_max + 1, Ef::one());
  std::map<std::pair<unsigned, int>, std::vector<Ef>> inv_den;  // (lh, point idx) -> 1/(z - x_r)
  auto inv_denoms = [&](unsigned lh, int which) -> const std::vector<Ef>& {
    auto key = std::make_pair(lh, which);
    auto it = inv_den.find(key);
    if (it != inv_den.end()) return it->second;
    size_t n2 = (size_t)1 << lh;
    Ef z = which == 0 ? zeta : zeta * two_adic_generator(lh - 1);
    std::vector<Ef> d(n2);
    Fp w = two_adic_generator(lh), g = fp_generator();
    std::vector<Fp> nat(n2);
    Fp p = g;
    for (size_t k = 0; k < n2; k++) nat[k] = p, p = p * w;
    for (size_t r = 0; r < n2; r++) d[r] = z - nat[bitrev(r, lh)];
    size_t chunk = 4096;
    parallel_for((n2 + chunk - 1) / chunk, [&](size_t ci) {
      size_t b = ci * chunk, e = std::min(n2, b + chunk);
      std::vector<Ef> tmp(d.begin() + b, d.begin() + e);
      batch_inverse(tmp);
      std::copy(tmp.begin(), tmp.end(), d.begin() + b);
    }, 1);
    return inv_den.emplace(key, std::move(d)).first->second;
  };
  for (auto& rd : rounds)
    for (size_t mi = 0; mi < rd.ldes.size(); mi++) {
      const Mat& l = *rd.ldes[mi];
      unsigned lh = log2_strict(l.h);
      if (ro[lh].empty()) ro[lh].assign(l.h, Ef::zero());
      std::vector<Ef> apow(l.w);
      Ef a = Ef::one();
      for (size_t c = 0; c < l.w; c++) apow[c] = a, a = a * fri_alpha;
      Ef a_width = a;
      std::vector<Ef> S(l.h);
      parallel_for(l.h, [&](size_t r) {
        Ef s = Ef::zero();
        for (size_t c = 0; c < l.w; c++) s += apow[c] * l.at(r, c);
        S[r] = s;
      }, 1024);
      for (size_t pi = 0; pi < rd.points[mi].size(); pi++) {
        Ef Y = Ef::zero();
        for (size_t c = 0; c < l.w; c++) Y += apow[c] * rd.vals[mi][pi][c];
        const auto& inv = inv_denoms(lh, (int)pi);
        Ef off = alpha_off[lh];
        auto& dst = ro[lh];
        parallel_for(l.h, [&](size_t r) { dst[r] += off * (Y - S[r]) * inv[r]; }, 4096);
        alpha_off[lh] = alpha_off[lh] * a_width;
      }
    }

  // FRI commit phase
  const StarkParams& pr = m.params;
  std::vector<Ef> cur = ro[log_max];
  std::vector<Mat> layer_mats;
  std::vector<MerkleTree> layer_trees;
  layer_mats.reserve(log_max);
  layer_trees.reserve(log_max);
  for (unsigned lfh = log_max - 1;; lfh--) {
    layer_mats.push_back(fri_layer_matrix(cur));
    layer_trees.push_back(merkle_commit({&layer_mats.back()}));
    Digest root = layer_trees.back().root();
    sp.fri_commits.push_back(root);
    ch.observe(root);
    Ef beta = ch.sample_ext();
    cur = fri_fold(cur, beta);
    if (!ro[lfh].empty())
      for (size_t i = 0; i < cur.size(); i++) cur[i] += ro[lfh][i];
    if (lfh == pr.log_blowup) break;
  }
  for (size_t i = 1; i < cur.size(); i++)
    if (cur[i] != cur[0]) throw std::runtime_error("FRI final polynomial is not constant (prover bug)");
  sp.final_poly = cur[0];
  ch.observe(sp.final_poly);
  sp.pow_witness = ch.grind(pr.pow_bits);
  // query phase
  for (unsigned qi = 0; qi < pr.num_queries; qi++) {
    size_t index = ch.sample_bits(log_max);
    QueryProof qp;
    for (auto& rd : rounds) {
      unsigned lt = log2_strict(rd.tree->max_height());
      qp.inputs.push_back(merkle_open(*rd.tree, index >> (log_max - lt)));
    }
    for (size_t li = 0; li < layer_trees.size(); li++) {
      size_t idx = index >> li, pair = idx >> 1;
      BatchOpening bo = merkle_open(layer_trees[li], pair);
      QueryProof::Step st;
      int sib = (int)((idx ^ 1) & 1);
      for (int k = 0; k < 4; k++) st.sibling.c[k] = bo.rows[0][4 * sib + k];
      st.path = bo.path;
      qp.steps.push_back(std::move(st));
    }
    sp.queries.push_back(std::move(qp));
  }
  return sp;
}

static inline std::vector<uint32_t> prove(const Machine& m, const QueryRecord& q, uint32_t shard_size) {
  uint32_t ns = num_shards(q, shard_size);
  if (!ns && m.forge_omit_kinds) ns = 1;   // the forger proves a record in which nothing was executed
  if (!ns) throw std::runtime_error("empty record");
  Challenger ch;
  observe_machine(m, ch);
  std::vector<ShardData> shards;
  for (uint32_t s = 0; s < ns; s++) shards.push_back(generate_shard(m, q, s, shard_size));
  for (auto& sd : shards) {
    ch.observe(sd.main_tree.root());
    for (uint32_t i = 0; i < m.num_pv; i++) ch.observe(q.public_values[i]);
  }
  Writer w;
  w.u(PROOF_MAGIC);
  w.u(ns);
  for (auto& sd : shards) {
    // the stored pointers inside main_tree refer to sd.main_lde, which lives as long as `shards`
    ShardProof sp = prove_shard(m, q, sd, ch);
    write_shard_proof(w, sp);
  }
  return w.w;
}

// ---- verify ------------------------------------------------------------------------------------------------------------------
struct VerifyError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define VCHECK(cond, msg) do { if (!(cond)) throw VerifyError(msg); } while (0)

static inline void verify_shard(const Machine& m, const ShardProof& sp, Challenger ch) {
  const StarkParams& pr = m.params;
  size_t nchips = sp.chips.size();
  VCHECK(nchips > 0, "no chips");
  VCHECK(sp.public_values.size() == m.num_pv, "public value count");
  // shapes + ordering
  uint32_t prev_log = 64;
  std::vector<bool> seen(m.chips.size(), false);
  for (auto& co : sp.chips) {
    VCHECK(co.chip_id < m.chips.size() && !seen[co.chip_id], "bad chip id");
    seen[co.chip_id] = true;
    const Chip& c = m.chips[co.chip_id];
    VCHECK(co.log_degree <= 26, "log_degree too large");
    VCHECK(co.log_degree <= prev_log, "chips not ordered by height");
    prev_log = co.log_degree;
    VCHECK(co.main_local.size() == c.width && co.main_next.size() == c.width, "main opening width");
    VCHECK(co.prep_local.size() == c.prep_width && co.prep_next.size() == c.prep_width, "prep opening width");
    VCHECK(co.perm_local.size() == 4 * c.perm_width() && co.perm_next.size() == 4 * c.perm_width(), "perm opening width");
    VCHECK(co.quotient[0].size() == 4 && co.quotient[1].size() == 4, "quotient opening width");
    if (c.kind == KIND_BYTES) VCHECK(co.log_degree == 16, "byte table height");
    if (c.kind == KIND_ENTRY) VCHECK(co.log_degree == 0 && sp.shard_index == 0, "entrypoint shape");
    if (c.kind == KIND_MEM) VCHECK(sp.shard_index == 0, "memory chip outside shard 0");
  }
  // MachineAir::included (src/lair/lair_chip.rs:124-139) is binding on the prover: Entrypoint and the memory tables
  // in shard 0 only, the byte table ("CPU", lair_chip.rs:90-92) in every shard.  Without the Entrypoint nothing ties
  // the public values to the computation.
  for (size_t i = 0; i < m.chips.size(); i++) {
    const Chip& c = m.chips[i];
    bool must = c.kind == KIND_BYTES || (sp.shard_index == 0 && (c.kind == KIND_ENTRY || c.kind == KIND_MEM));
    VCHECK(!must || seen[i], "a chip MachineAir::included demands is missing from the shard");
  }
  Ef chal_alpha = ch.sample_ext(), chal_beta = ch.sample_ext();
  ch.observe(sp.perm_commit);
  Ef alpha = ch.sample_ext();
  ch.observe(sp.quot_commit);
  Ef zeta = ch.sample_ext();

  // rounds as the verifier sees them: (commit, per matrix: lde height, points, values)
  struct VRound {
    Digest commit;
    std::vector<size_t> heights;
    std::vector<std::vector<Ef>> points;
    std::vector<std::vector<const std::vector<Ef>*>> vals;
  };
  VRound r_prep{m.prep_commit, {}, {}, {}}, r_main{sp.main_commit, {}, {}, {}}, r_perm{sp.perm_commit, {}, {}, {}},
      r_quot{sp.quot_commit, {}, {}, {}};
  bool has_prep = false;
  for (auto& co : sp.chips) {
    const Chip& c = m.chips[co.chip_id];
    size_t h2 = (size_t)2 << co.log_degree;
    Ef zn = zeta * two_adic_generator(co.log_degree);
    if (c.prep_width) {
      has_prep = true;
      r_prep.heights.push_back(h2); r_prep.points.push_back({zeta, zn}); r_prep.vals.push_back({&co.prep_local, &co.prep_next});
    }
    r_main.heights.push_back(h2); r_main.points.push_back({zeta, zn}); r_main.vals.push_back({&co.main_local, &co.main_next});
    r_perm.heights.push_back(h2); r_perm.points.push_back({zeta, zn}); r_perm.vals.push_back({&co.perm_local, &co.perm_next});
    for (int j = 0; j < 2; j++) {
      r_quot.heights.push_back(h2); r_quot.points.push_back({zeta}); r_quot.vals.push_back({&co.quotient[j]});
    }
  }
  std::vector<VRound> rounds;
  if (has_prep) rounds.push_back(r_prep);
  rounds.push_back(r_main); rounds.push_back(r_perm); rounds.push_back(r_quot);
  for (auto& rd : rounds)
    for (auto& mv : rd.vals)
      for (auto* pv : mv)
        for (auto& x : *pv) ch.observe(x);
  Ef fri_alpha = ch.sample_ext();

  unsigned log_max = 0;
  for (auto& rd : rounds)
    for (size_t h : rd.heights) log_max = std::max(log_max, log2_strict(h));
  VCHECK(log_max > pr.log_blowup, "degenerate shard");
  VCHECK(sp.fri_commits.size() == log_max - pr.log_blowup, "FRI layer count");
  std::vector<Ef> betas;
  for (auto& cm : sp.fri_commits) {
    ch.observe(cm);
    betas.push_back(ch.sample_ext());
  }
  ch.observe(sp.final_poly);
  VCHECK(ch.check_witness(pr.pow_bits, sp.pow_witness), "invalid proof-of-work witness");
  VCHECK(sp.queries.size() == pr.num_queries, "query count");
  Fp one_half = fp_inv(Fp::from_u32(2));
  for (auto& qp : sp.queries) {
    size_t index = ch.sample_bits(log_max);
    VCHECK(qp.inputs.size() == rounds.size(), "round count");
    std::vector<Ef> ro(log_max + 1, Ef::zero());
    std::vector<Ef> apow(log_max + 1, Ef::one());
    for (size_t ri = 0; ri < rounds.size(); ri++) {
      const VRound& rd = rounds[ri];
      const BatchOpening& bo = qp.inputs[ri];
      VCHECK(bo.rows.size() == rd.heights.size(), "opened matrix count");
      size_t maxh = *std::max_element(rd.heights.begin(), rd.heights.end());
      unsigned lt = log2_strict(maxh);
      for (size_t mi = 0; mi < bo.rows.size(); mi++) VCHECK(bo.rows[mi].size() == rd.vals[mi][0]->size(), "opened row width");
      VCHECK(merkle_verify(rd.commit, rd.heights, index >> (log_max - lt), bo.rows, bo.path), "input Merkle path");
      for (size_t mi = 0; mi < bo.rows.size(); mi++) {
        unsigned lh = log2_strict(rd.heights[mi]);
        size_t r = index >> (log_max - lh);
        Fp x = fp_generator() * fp_pow(two_adic_generator(lh), bitrev(r, lh));
        for (size_t pi = 0; pi < rd.points[mi].size(); pi++) {
          Ef inv = ef_inv(rd.points[mi][pi] - x);
          const std::vector<Ef>& pz = *rd.vals[mi][pi];
          for (size_t c = 0; c < pz.size(); c++) {
            ro[lh] += apow[lh] * (pz[c] - bo.rows[mi][c]) * inv;
            apow[lh] = apow[lh] * fri_alpha;
          }
        }
      }
    }
    VCHECK(qp.steps.size() == sp.fri_commits.size(), "step count");
    Ef folded = Ef::zero();
    for (size_t li = 0; li < qp.steps.size(); li++) {
      unsigned lfh = log_max - 1 - (unsigned)li;  // log height of the folded layer
      size_t idx = index >> li, pair = idx >> 1;
      folded += ro[lfh + 1];
      Ef ev[2];
      ev[idx & 1] = folded;
      ev[(idx & 1) ^ 1] = qp.steps[li].sibling;
      std::vector<Fp> row(8);
      for (int k = 0; k < 4; k++) row[k] = ev[0].c[k], row[4 + k] = ev[1].c[k];
      VCHECK(merkle_verify(sp.fri_commits[li], {(size_t)1 << lfh}, pair, {row}, qp.steps[li].path), "FRI layer Merkle path");
      Fp xinv = fp_inv(fp_pow(two_adic_generator(lfh + 1), bitrev(pair, lfh)));
      Ef t = betas[li] * (one_half * xinv);
      folded = (t + one_half) * ev[0] + (-t + one_half) * ev[1];
    }
    folded += ro[pr.log_blowup];
    for (unsigned lh = 0; lh < pr.log_blowup; lh++) VCHECK(ro[lh].is_zero(), "matrix below the blowup height");
    VCHECK(folded == sp.final_poly, "FRI final value mismatch");
  }

  // constraints at zeta
  for (auto& co : sp.chips) {
    const Chip& c = m.chips[co.chip_id];
    unsigned logn = co.log_degree;
    size_t pw = c.perm_width();
    Fp wn_inv = fp_inv(two_adic_generator(logn));
    Ef zh = ef_pow(zeta, (uint64_t)1 << logn) - Fp::one();
    RowView<Ef> row;
    row.is_first = zh * ef_inv(zeta - Fp::one());
    row.is_last = zh * ef_inv(zeta - wn_inv);
    row.is_transition = zeta - wn_inv;
    row.main_local = co.main_local.data(); row.main_next = co.main_next.data();
    row.prep_local = co.prep_local.data(); row.prep_next = co.prep_next.data();
    row.pv = sp.public_values.data();
    std::vector<Ef> permL(pw), permN(pw), regs;
    for (size_t j = 0; j < pw; j++) permL[j] = unflatten(&co.perm_local[4 * j]), permN[j] = unflatten(&co.perm_next[4 * j]);
    Ef acc = fold_air_constraints<Ef>(c, row, alpha, Ef::zero(), regs);
    acc = fold_perm_constraints<Ef>(c, row, permL.data(), permN.data(), co.cumulative_sum, chal_alpha, chal_beta, alpha, acc);
    // quotient(zeta) from the two chunks
    Fp g = fp_generator(), w2n = two_adic_generator(logn + 1);
    Fp shifts[2] = {g, g * w2n};
    Ef quotient = Ef::zero();
    for (int j = 0; j < 2; j++) {
      int o = 1 - j;
      uint64_t n = (uint64_t)1 << logn;
      Ef num = ef_pow(zeta * fp_inv(shifts[o]), n) - Fp::one();
      Fp den = fp_pow(shifts[j] * fp_inv(shifts[o]), n) - Fp::one();
      Ef zp = num * fp_inv(den);
      quotient += zp * unflatten(co.quotient[j].data());
    }
    VCHECK(acc * ef_inv(zh) == quotient, ("constraints do not match the quotient for chip " + c.name).c_str());
  }
}

static inline void verify(const Machine& m, const uint32_t* words, size_t n) {
  Reader r{words, n};
  VCHECK(r.u() == PROOF_MAGIC, "bad proof magic");
  uint32_t ns = r.u();
  VCHECK(ns >= 1 && ns <= 1024, "bad shard count");
  std::vector<ShardProof> sps;
  for (uint32_t i = 0; i < ns; i++) sps.push_back(read_shard_proof(r));
  VCHECK(r.p == n, "trailing data");
  Challenger ch;
  observe_machine(m, ch);
  for (uint32_t i = 0; i < ns; i++) {
    VCHECK(sps[i].shard_index == i, "shard index");
    VCHECK(sps[i].public_values.size() == m.num_pv, "public values");
    for (size_t k = 0; k < m.num_pv; k++) VCHECK(sps[i].public_values[k] == sps[0].public_values[k], "public values differ");
    ch.observe(sps[i].main_commit);
    for (Fp x : sps[i].public_values) ch.observe(x);
  }
  Ef total = Ef::zero();
  for (auto& sp : sps) {
    verify_shard(m, sp, ch);
    for (auto& co : sp.chips) total += co.cumulative_sum;
  }
  VCHECK(total.is_zero(), "lookup argument: cumulative sums do not cancel");
}

}  // namespace orc
