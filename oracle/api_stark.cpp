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
}  // extern "C"
