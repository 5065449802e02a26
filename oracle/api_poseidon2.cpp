// ORACLE (test infrastructure, NOT product code): C entry points over poseidon2.h for ctypes.
#include "poseidon2.h"
#include "par.h"
#include <vector>
using namespace orc;

extern "C" {

// Full permutation of n states of `width` canonical u32 words (in place semantics: out may alias in).
int orc_poseidon2_permute(int width, const uint32_t* in, size_t n, uint32_t* out) {
  if (!p2_params(width).t) return -1;
  parallel_for(n, [&](size_t i) {
    Fp s[40];
    for (int j = 0; j < width; j++) s[j] = Fp::from_u32(in[i * width + j]);
    p2_permute(s, width);
    for (int j = 0; j < width; j++) out[i * width + j] = s[j].to_u32();
  });
  return 0;
}

// PoseidonChipset::hash (src/core/poseidon.rs:29-39): digest = first 8 words of the permuted state.
int orc_poseidon2_hash(int width, const uint32_t* in, size_t n, uint32_t* out8) {
  if (!p2_params(width).t) return -1;
  parallel_for(n, [&](size_t i) {
    Fp s[40];
    for (int j = 0; j < width; j++) s[j] = Fp::from_u32(in[i * width + j]);
    p2_permute(s, width);
    for (int j = 0; j < 8; j++) out8[i * 8 + j] = s[j].to_u32();
  });
  return 0;
}

int orc_poseidon2_witness_cols(int width) { return p2_params(width).t ? 8 + p2_witness_cols(width) : -1; }

// PoseidonChipset::populate_witness (src/core/poseidon.rs:65-72): witness = out[8] ++ Poseidon2Cols.
int orc_poseidon2_witness(int width, const uint32_t* in, size_t n, uint32_t* wit) {
  if (!p2_params(width).t) return -1;
  int wc = p2_witness_cols(width);
  parallel_for(n, [&](size_t i) {
    Fp s[40];
    std::vector<Fp> w(wc);
    for (int j = 0; j < width; j++) s[j] = Fp::from_u32(in[i * width + j]);
    p2_permute(s, width, w.data());
    uint32_t* o = wit + i * (size_t)(8 + wc);
    for (int j = 0; j < 8; j++) o[j] = s[j].to_u32();
    for (int j = 0; j < wc; j++) o[8 + j] = w[j].to_u32();
  });
  return 0;
}

// PaddingFreeSponge<16,8,8> over rows of `row_len` words; out8 per row.
int orc_sponge16_hash_rows(const uint32_t* rows, size_t n, size_t row_len, uint32_t* out8) {
  parallel_for(n, [&](size_t i) {
    Sponge16 sp;
    for (size_t j = 0; j < row_len; j++) sp.absorb(Fp::from_u32(rows[i * row_len + j]));
    Fp d[8];
    sp.finish(d);
    for (int j = 0; j < 8; j++) out8[i * 8 + j] = d[j].to_u32();
  });
  return 0;
}

}  // extern "C"
