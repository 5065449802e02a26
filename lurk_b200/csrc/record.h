// Flattened QueryRecord: the GPU-ready form of src/lair/execute.rs:26-33,67-75.
//
// The reference keeps per-func `IndexMap<args -> QueryResult>` with boxed vectors and re-hashes the
// arguments of every Call/Store at trace-generation time (src/lair/trace.rs:325-373).  Here the
// executor emits, per func, structure-of-arrays tables indexed by nonce (= row), and for every
// Call / PreImg / Store executed by a row the *row index* of the callee query or memory slot
// ("hints"), so that the trace kernel only gathers.  All values canonical u32.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "program.h"

namespace lsk {

// open-addressing index: key words live in the owning table; slot stores row + 1
struct HashIndex {
  std::vector<uint32_t> slots;
  size_t used = 0;
  static uint64_t hash(const uint32_t* k, uint32_t n) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ n;
    for (uint32_t i = 0; i < n; i++) {
      h ^= k[i];
      h *= 0xff51afd7ed558ccdull;
      h ^= h >> 32;
    }
    return h;
  }
  // keys: base pointer of row-major key storage with `n` words per row
  long find(const uint32_t* keys, uint32_t n, const uint32_t* k) const {
    if (slots.empty()) return -1;
    size_t mask = slots.size() - 1;
    for (size_t i = hash(k, n) & mask;; i = (i + 1) & mask) {
      uint32_t s = slots[i];
      if (!s) return -1;
      const uint32_t* c = keys + (size_t)(s - 1) * n;
      bool eq = true;
      for (uint32_t j = 0; j < n && eq; j++) eq = c[j] == k[j];
      if (eq) return (long)s - 1;
    }
  }
  void insert_new(const uint32_t* keys, uint32_t n, uint32_t row) {  // key must be absent
    if ((used + 1) * 2 > slots.size()) grow(keys, n);
    place(keys + (size_t)row * n, n, row);
    used++;
  }
  void set(const uint32_t* keys, uint32_t n, const uint32_t* k, uint32_t row) {  // insert or overwrite
    if (!slots.empty()) {
      size_t mask = slots.size() - 1;
      for (size_t i = hash(k, n) & mask;; i = (i + 1) & mask) {
        uint32_t s = slots[i];
        if (!s) break;
        const uint32_t* c = keys + (size_t)(s - 1) * n;
        bool eq = true;
        for (uint32_t j = 0; j < n && eq; j++) eq = c[j] == k[j];
        if (eq) {
          slots[i] = row + 1;
          return;
        }
      }
    }
    insert_new(keys, n, row);
  }

 private:
  void place(const uint32_t* k, uint32_t n, uint32_t row) {
    size_t mask = slots.size() - 1;
    size_t i = hash(k, n) & mask;
    while (slots[i]) i = (i + 1) & mask;
    slots[i] = row + 1;
  }
  void grow(const uint32_t* keys, uint32_t n) {
    std::vector<uint32_t> old;
    old.swap(slots);
    slots.assign(old.empty() ? 1024 : old.size() * 2, 0);
    for (uint32_t s : old)
      if (s) place(keys + (size_t)(s - 1) * n, n, s - 1);
  }
};

struct FuncTable {
  uint32_t in = 0, out = 0;
  bool invertible = false;
  std::vector<uint32_t> args;      // [n][in]
  std::vector<uint32_t> outs;      // [n][out]
  std::vector<uint8_t> done;       // output computed
  std::vector<uint32_t> provide;   // [n][2] (last_nonce, last_count)
  std::vector<uint32_t> req_off;   // [n] index (in pairs) of the row's first require record
  std::vector<uint32_t> reqs;      // pairs (prev_nonce, prev_count) in op order
  std::vector<uint32_t> hint_off;  // [n]
  std::vector<uint32_t> hints;     // callee / memory row per Call, PreImg, Store in op order
  // partial funcs only (depth provenance, src/lair/provenance.rs; QueryResult::{depth, depth_requires}):
  bool partial = false;
  std::vector<uint32_t> depth;     // [n]
  std::vector<uint32_t> dreq_off;  // [n] index (in pairs) of the row's first depth require
  std::vector<uint32_t> dreqs;     // pairs: 2 range checks of the depth bytes, then one less-than per partial callee
  // The executor appends a row's require / hint / depth-require segments when the row RETURNS, i.e. in
  // completion order; *_len keeps each row's segment length so that lsk_record_finalize can re-pack the
  // three CSR arrays in row order (host-only bookkeeping, never uploaded).  row_ordered: offsets are
  // monotone in the row index, so a shard's rows own one contiguous range of each array.
  std::vector<uint32_t> req_len, hint_len, dreq_len;
  bool row_ordered = true;
  size_t completed = 0;            // rows that have returned so far
  HashIndex index;                 // args -> row
  // inverse queries (PreImg): out -> inp, both produced at Return and injected from outside
  std::vector<uint32_t> inv_outs, inv_inps;
  HashIndex inv_index;
  size_t rows() const { return done.size(); }
};

struct MemTable {
  uint32_t len = 0;
  std::vector<uint32_t> vals;     // [n][len]
  std::vector<uint32_t> provide;  // [n][2]
  HashIndex index;
  size_t rows() const { return provide.size() / 2; }
};

}  // namespace lsk

struct lsk_program {
  std::vector<uint32_t> words;
  uint32_t nfuncs = 0, nchips = 0;
  const uint32_t* chip(uint32_t i) const { return &words[3 + i * lsk::CHIP_WORDS]; }
  uint32_t func_offset(uint32_t f) const { return words[3 + nchips * lsk::CHIP_WORDS + f]; }
  lsk::FuncHeader header(uint32_t f) const {
    const uint32_t* w = &words[func_offset(f)];
    return {w[0], w[1], w[2], w[3], w[4], w[5], w[6]};
  }
  uint32_t body(uint32_t f) const { return func_offset(f) + lsk::FUNC_HEADER_WORDS; }
  uint32_t width(uint32_t f) const {
    auto h = header(f);
    return 1 + h.in + h.out + h.aux + h.sel;
  }
};

struct lsk_record {
  std::vector<lsk::FuncTable> funcs;
  lsk::MemTable mems[lsk::NUM_MEM_TABLES];
  std::vector<uint32_t> bytes;   // [65536][6][2] provide records (nonce, count); BytesRecord
  bool bytes_used = false;
  std::vector<uint32_t> public_values;
  std::vector<void*> pinned;     // host ranges page-locked by lsk_record_pin (released by lsk_record_unpin / free)
};
namespace lsk {
void record_unpin(lsk_record* r);  // prover.cu (needs the CUDA runtime)
bool record_row_ordered(const lsk_record& r);
}
