// ORACLE (test infrastructure, NOT product code): C entry points over stark.h / stark_prove.h.
#include "stark_prove.h"
using namespace orc;

static thread_local std::string g_serr;
extern "C" {
const char* orc_stark_last_error() { return g_serr.c_str(); }

void* orc_machine_new(void* toplevel, uint32_t entry_func, const uint32_t* blob, size_t n, uint32_t num_queries,
                      uint32_t pow_bits) {
  try {
    auto* m = new Machine;
    m->top = (Toplevel*)toplevel;
    m->entry = entry_func;
    m->chips = parse_machine(blob, n, &m->num_pv);
    m->params.num_queries = num_queries;
    m->params.pow_bits = pow_bits;
    machine_setup(*m);
    return m;
  } catch (std::exception& e) {
    g_serr = e.what();
    return nullptr;
  }
}
void orc_machine_free(void* m) { delete (Machine*)m; }
void orc_machine_prep_commit(void* m, uint32_t* out8) {
  for (int i = 0; i < 8; i++) out8[i] = ((Machine*)m)->prep_commit.d[i].to_u32();
}

// negative tests: a prover that omits chips / claims public values it did not compute
void orc_machine_forge_omit_kinds(void* m, uint32_t kind_mask) { ((Machine*)m)->forge_omit_kinds = kind_mask; }
void orc_record_set_public_values(void* record, const uint32_t* pv, size_t n) {
  auto& v = ((QueryRecord*)record)->public_values;
  v.clear();
  for (size_t i = 0; i < n; i++) v.push_back(Fp::from_u32(pv[i]));
}

// returns a malloc'd buffer of u32 words (free with orc_free)
int orc_prove(void* m, void* record, uint32_t shard_size, uint32_t** out, size_t* out_len) {
  try {
    auto w = prove(*(Machine*)m, *(QueryRecord*)record, shard_size);
    *out = (uint32_t*)malloc(w.size() * 4);
    memcpy(*out, w.data(), w.size() * 4);
    *out_len = w.size();
    return 0;
  } catch (std::exception& e) {
    g_serr = e.what();
    return -1;
  }
}
void orc_free(void* p) { free(p); }

// 0 = accepted, -4 = rejected (message in orc_stark_last_error), -1 = malformed
int orc_verify(void* m, const uint32_t* proof, size_t n) {
  try {
    verify(*(Machine*)m, proof, n);
    return 0;
  } catch (VerifyError& e) {
    g_serr = e.what();
    return -4;
  } catch (std::exception& e) {
    g_serr = e.what();
    return -1;
  }
}

// ---- stage-level helpers for parity tests of the CUDA stages -------------------------------------
// row-major canonical in (h x w on H_N, natural rows) -> row-major canonical out (2h x w, bit-reversed rows)
int orc_coset_lde(const uint32_t* rows, size_t h, size_t w, uint32_t in_shift, uint32_t* out) {
  try {
    std::vector<Fp> v(h * w);
    for (size_t i = 0; i < h * w; i++) v[i] = Fp::from_u32(rows[i]);
    Mat m = from_row_major(v, h, w);
    Mat l = coset_lde(m, Fp::from_u32(in_shift));
    for (size_t r = 0; r < l.h; r++)
      for (size_t c = 0; c < w; c++) out[r * w + c] = l.at(r, c).to_u32();
    return 0;
  } catch (std::exception& e) {
    g_serr = e.what();
    return -1;
  }
}
// Merkle root over `nmats` row-major canonical matrices (heights powers of two)
int orc_merkle_root(const uint32_t* const* mats, const size_t* hs, const size_t* ws, size_t nmats, uint32_t* out8) {
  try {
    std::vector<Mat> ms;
    for (size_t i = 0; i < nmats; i++) {
      std::vector<Fp> v(hs[i] * ws[i]);
      for (size_t k = 0; k < v.size(); k++) v[k] = Fp::from_u32(mats[i][k]);
      ms.push_back(from_row_major(v, hs[i], ws[i]));
    }
    std::vector<const Mat*> ptrs;
    for (auto& m : ms) ptrs.push_back(&m);
    MerkleTree t = merkle_commit(ptrs);
    for (int i = 0; i < 8; i++) out8[i] = t.root().d[i].to_u32();
    return 0;
  } catch (std::exception& e) {
    g_serr = e.what();
    return -1;
  }
}
// ---- A8 / A10 / A12 one stage at a time (counterparts of lsk_perm_trace / lsk_quotient / lsk_fri_prove) ----------
static Ef ef_from(const uint32_t* w) {
  Ef e;
  for (int k = 0; k < 4; k++) e.c[k] = Fp::from_u32(w[k]);
  return e;
}
int orc_chip_perm_width(void* m, uint32_t chip) { return (int)((Machine*)m)->chips.at(chip).perm_width(); }
// main: h x w row-major canonical, natural rows -> perm: h x 4*pw, cum[4]
int orc_perm_trace(void* mp, uint32_t chip, const uint32_t* main, size_t h, const uint32_t* alpha, const uint32_t* beta,
                   uint32_t* perm_out, uint32_t* cum_out) {
  try {
    Machine& m = *(Machine*)mp;
    const Chip& c = m.chips.at(chip);
    std::vector<Fp> v(h * c.width);
    for (size_t i = 0; i < v.size(); i++) v[i] = Fp::from_u32(main[i]);
    Mat mm = from_row_major(v, h, c.width);
    Ef cum;
    Mat perm = permutation_trace(c, mm, c.prep_width ? &m.prep_trace : nullptr, ef_from(alpha), ef_from(beta), &cum);
    for (size_t r = 0; r < h; r++)
      for (size_t k = 0; k < perm.w; k++) perm_out[r * perm.w + k] = perm.at(r, k).to_u32();
    for (int k = 0; k < 4; k++) cum_out[k] = cum.c[k].to_u32();
    return 0;
  } catch (std::exception& e) {
    g_serr = e.what();
    return -1;
  }
}
// quotient values on g*H_2h in commitment (bit-reversed) order: 2h x 4
int orc_quotient(void* mp, uint32_t chip, const uint32_t* main, const uint32_t* perm, size_t h, const uint32_t* pv,
                 const uint32_t* cum, const uint32_t* chal_alpha, const uint32_t* chal_beta, const uint32_t* alpha,
                 uint32_t* out) {
  try {
    Machine& m = *(Machine*)mp;
    const Chip& c = m.chips.at(chip);
    size_t pcols = 4 * c.perm_width();
    std::vector<Fp> a(h * c.width), b(h * pcols), pvs(m.num_pv);
    for (size_t i = 0; i < a.size(); i++) a[i] = Fp::from_u32(main[i]);
    for (size_t i = 0; i < b.size(); i++) b[i] = Fp::from_u32(perm[i]);
    for (size_t i = 0; i < pvs.size(); i++) pvs[i] = Fp::from_u32(pv[i]);
    Mat main_lde = coset_lde(from_row_major(a, h, c.width)), perm_lde = coset_lde(from_row_major(b, h, pcols));
    std::vector<Ef> q = quotient_values(c, log2_strict(h), main_lde, perm_lde, c.prep_width ? &m.prep_lde : nullptr, pvs,
                                        ef_from(cum), ef_from(chal_alpha), ef_from(chal_beta), ef_from(alpha));
    for (size_t r = 0; r < q.size(); r++)
      for (int k = 0; k < 4; k++) out[4 * r + k] = q[r].c[k].to_u32();
    return 0;
  } catch (std::exception& e) {
    g_serr = e.what();
    return -1;
  }
}
// FRI on one vector (2^log_n extension values, commitment order) from a given DuplexChallenger sponge state; the word
// image of lsk_fri_prove (include/lurk_stark.h)
int orc_fri_prove(const uint32_t* vals, uint32_t log_n, const uint32_t* state, uint32_t num_queries, uint32_t pow_bits,
                  uint32_t** out, size_t* out_len) {
  try {
    size_t n = (size_t)1 << log_n;
    std::vector<Ef> cur(n);
    for (size_t i = 0; i < n; i++) cur[i] = ef_from(vals + 4 * i);
    Challenger ch;
    for (int k = 0; k < 16; k++) ch.state[k] = Fp::from_u32(state[k]);
    std::vector<std::vector<Ef>> layer_vals;
    std::vector<Mat> layer_mats;
    std::vector<MerkleTree> trees;
    layer_mats.reserve(log_n);
    trees.reserve(log_n);
    std::vector<uint32_t> W;
    W.push_back(log_n - 1);
    for (unsigned lfh = log_n - 1;; lfh--) {
      layer_vals.push_back(cur);
      layer_mats.push_back(fri_layer_matrix(cur));
      trees.push_back(merkle_commit({&layer_mats.back()}));
      Digest root = trees.back().root();
      for (Fp x : root.d) W.push_back(x.to_u32());
      ch.observe(root);
      cur = fri_fold(cur, ch.sample_ext());
      if (lfh == 1) break;
    }
    for (size_t i = 1; i < cur.size(); i++)
      if (cur[i] != cur[0]) throw std::runtime_error("the folded vector is not constant");
    for (Fp x : cur[0].c) W.push_back(x.to_u32());
    ch.observe(cur[0]);
    W.push_back(ch.grind(pow_bits).to_u32());
    W.push_back(num_queries);
    for (uint32_t q = 0; q < num_queries; q++) {
      size_t index = ch.sample_bits(log_n);
      W.push_back((uint32_t)index);
      for (size_t li = 0; li < trees.size(); li++) {
        size_t pos = index >> li;
        const Ef& sib = layer_vals[li][pos ^ 1];
        for (Fp x : sib.c) W.push_back(x.to_u32());
        BatchOpening op = merkle_open(trees[li], pos >> 1);
        W.push_back((uint32_t)op.path.size());
        for (auto& d : op.path)
          for (Fp x : d.d) W.push_back(x.to_u32());
      }
    }
    *out = (uint32_t*)malloc(W.size() * 4 + 4);
    memcpy(*out, W.data(), W.size() * 4);
    *out_len = W.size();
    return 0;
  } catch (std::exception& e) {
    g_serr = e.what();
    return -1;
  }
}
}  // extern "C"
