// A4/A5: witness-trace generation on the GPU.
//
// Replaces FuncChip::generate_trace + Func::populate_row (src/lair/trace.rs:72-135,145-418),
// MemChip::generate_trace (src/lair/memory.rs:30-68), BytesChip::generate_trace /
// preprocessed_trace (src/gadgets/bytes/trace.rs:49-101).
//
// One thread per trace row interprets the func's flat bytecode (program.h).  Everything the
// reference looks up through hash maps at this point (callee outputs, memory slots) arrives as
// row indices in the flattened record, so the kernel only gathers.  Field inversions (AssertNe /
// Inv / Not witnesses and the 1/(count+1) of every require triple) are done in-thread with a
// 62-product Fermat chain.  Output is written either row-major canonical (C-ABI parity view) or
// column-major with bit-reversed rows in Montgomery form (the layout the LDE/commit kernels eat:
// thread i writes row brev(i) of every column, the inverse NTT then needs no reordering pass).
// HBM-bound by design: 4*N*w bytes written + the record tables read once.
#include "poseidon2.cuh"
#include "comm.cuh"
#include "trace.cuh"

namespace lsk {

struct StoreRowMajor {
  uint32_t* out;
  uint32_t width, row;
  LSK_D void operator()(uint32_t col, uint32_t v_mont) const { out[(size_t)row * width + col] = from_mont(v_mont); }
};
struct StoreColBitrev {
  uint32_t* out;
  uint32_t height, brow;  // brow = bitrev(row)
  LSK_D void operator()(uint32_t col, uint32_t v_mont) const { out[(size_t)col * height + brow] = v_mont; }
};

template <class Store>
struct AuxWitness {  // sink for the wide Poseidon2 witness inside a func row
  const Store& st;
  uint32_t base;  // column of the first Poseidon2Cols entry
  int t, rp;
  LSK_D void ext_state(int r, int i, uint32_t v) { st(base + r * t + i, v); }
  LSK_D void ext_sbox(int r, int i, uint32_t v) { st(base + 8 * t + r * t + i, v); }
  LSK_D void int_init(int i, uint32_t v) { st(base + 16 * t + i, v); }
  LSK_D void int_state0(int r, uint32_t v) { st(base + 17 * t + r, v); }
  LSK_D void int_sbox(int r, uint32_t v) { st(base + 17 * t + (rp - 1) + r, v); }
};

template <int T, class Store>
__device__ __noinline__ void extern_poseidon(const uint32_t* op, uint32_t* map, uint32_t& map_len, uint32_t& aux,
                                              const Store& st) {
  uint32_t s[T];
#pragma unroll
  for (int i = 0; i < T; i++) s[i] = map[op[3 + i]];
  AuxWitness<Store> wit{st, aux + 8, T, P2Rounds<T>::RP};
  p2_permute<T, AuxWitness<Store>>(s, wit);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    map[map_len++] = s[i];
    st(aux + i, s[i]);
  }
  aux += 8 + 17 * T + 2 * P2Rounds<T>::RP - 1;
}

template <int MAP_CAP, class Store>
__device__ void populate_row(const uint32_t* __restrict__ w, uint32_t body, FuncHeader hdr, const uint32_t* chips,
                             const DevFuncTable* __restrict__ funcs, const DevMemTable* __restrict__ mems,
                             const DevFuncTable& self, uint32_t q, const Store& st) {
  uint32_t map[MAP_CAP];
  uint32_t map_len = 0;
  const uint32_t col_in = 1, col_out = 1 + hdr.in, col_aux = 1 + hdr.in + hdr.out;
  const uint32_t col_sel = col_aux + hdr.aux;
  for (uint32_t j = 0; j < hdr.in; j++) {
    uint32_t v = to_mont(self.args[(size_t)q * hdr.in + j]);
    map[map_len++] = v;
    st(col_in + j, v);
  }
  for (uint32_t j = 0; j < hdr.out; j++) st(col_out + j, to_mont(self.outs[(size_t)q * hdr.out + j]));
  uint32_t aux = col_aux;
  st(aux++, to_mont(self.provide[(size_t)q * 2]));
  st(aux++, to_mont(self.provide[(size_t)q * 2 + 1]));
  const uint32_t* reqs = self.reqs + (size_t)self.req_off[q] * 2;
  const uint32_t* hints = self.hints + self.hint_off[q];
  auto push_record = [&](const uint32_t*& r) {  // Record::into_require, src/air/builder.rs:159-168
    uint32_t pn = r[0], pc = r[1];
    r += 2;
    st(aux++, to_mont(pn));
    st(aux++, to_mont(pc));
    st(aux++, finv(to_mont(pc + 1)));
  };
  auto push_require = [&]() { push_record(reqs); };
  // depth provenance (src/lair/trace.rs:110-121,235-254): the row's own depth as 4 range-checked bytes ...
  const uint32_t* dreqs = nullptr;
  uint32_t depth = 0;
  if (hdr.partial) {
    dreqs = self.dreqs + (size_t)self.dreq_off[q] * 2;
    depth = self.depth[q];
    for (int b = 0; b < 4; b++) st(aux++, to_mont((depth >> (8 * b)) & 255u));
    push_record(dreqs);
    push_record(dreqs);
  }
  // ... and, per partial callee, its depth bytes, the LessThanWitness (is_comp[4], lhs limb, rhs limb:
  // src/gadgets/unsigned/less_than.rs:12-41) proving dep < depth, and the byte lookup's require
  auto push_depth = [&](uint32_t dep) {
    for (int b = 0; b < 4; b++) st(aux++, to_mont((dep >> (8 * b)) & 255u));
    int cb = 3;
    while (cb > 0 && ((dep >> (8 * cb)) & 255u) == ((depth >> (8 * cb)) & 255u)) cb--;
    st(aux + cb, R1);
    st(aux + 4, to_mont((dep >> (8 * cb)) & 255u));
    st(aux + 5, to_mont((depth >> (8 * cb)) & 255u));
    aux += 6;
    push_record(dreqs);
  };
  uint32_t pc = body;
  for (;;) {
    uint32_t nops = w[pc + 1];
    pc += 2;
    for (uint32_t k = 0; k < nops; k++) {
      const uint32_t* op = w + pc;
      pc += op_words(op);
      switch (op[0]) {
        case OP_CONST: map[map_len++] = to_mont(op[1]); break;
        case OP_ADD: map[map_len] = fadd(map[op[1]], map[op[2]]); map_len++; break;
        case OP_SUB: map[map_len] = fsub(map[op[1]], map[op[2]]); map_len++; break;
        case OP_MUL: {
          uint32_t v = fmul(map[op[1]], map[op[2]]);
          map[map_len++] = v;
          if (op[3]) st(aux++, v);
        } break;
        case OP_INV: {
          uint32_t v = finv(map[op[1]]);
          map[map_len++] = v;
          if (op[2]) st(aux++, v);
        } break;
        case OP_NOT: {
          uint32_t a = map[op[1]];
          uint32_t x = a == 0 ? R1 : 0u;
          map[map_len++] = x;
          if (op[2]) {
            st(aux++, finv(a));  // finv(0) == 0
            st(aux++, x);
          }
        } break;
        case OP_ASSERT_EQ: break;
        case OP_ASSERT_NE: {  // push_inequality_witness, trace.rs:218-233
          uint32_t n = op[1];
          bool found = false;
          for (uint32_t i = 0; i < n; i++) {
            uint32_t d = fsub(map[op[2 + i]], map[op[2 + n + i]]);
            uint32_t c = 0;
            if (!found && d != 0) {
              c = finv(d);
              found = true;
            }
            st(aux++, c);
          }
        } break;
        case OP_CONTAINS: {
          uint32_t n = op[1], b = map[op[2 + n]];
          uint32_t acc = fsub(map[op[2]], b);
          for (uint32_t i = 1; i < n; i++) {
            acc = fmul(acc, fsub(map[op[2 + i]], b));
            st(aux++, acc);
          }
        } break;
        case OP_CALL: {
          const DevFuncTable& callee = funcs[op[1]];
          uint32_t row = *hints++;
          for (uint32_t j = 0; j < callee.out; j++) {
            uint32_t v = to_mont(callee.outs[(size_t)row * callee.out + j]);
            map[map_len++] = v;
            st(aux++, v);
          }
          push_require();
          if (op[4 + op[2]]) push_depth(callee.depth[row]);
        } break;
        case OP_PREIMG: {
          const DevFuncTable& callee = funcs[op[1]];
          uint32_t row = *hints++;
          for (uint32_t j = 0; j < callee.in; j++) {
            uint32_t v = to_mont(callee.args[(size_t)row * callee.in + j]);
            map[map_len++] = v;
            st(aux++, v);
          }
          push_require();
          if (op[4 + op[2]]) push_depth(callee.depth[row]);
        } break;
        case OP_STORE: {
          uint32_t row = *hints++;
          uint32_t v = to_mont(row + 1);
          map[map_len++] = v;
          st(aux++, v);
          push_require();
        } break;
        case OP_LOAD: {
          uint32_t len = op[1];
          const DevMemTable& mt = mems[mem_index_from_len((int)len)];
          uint32_t ptr = from_mont(map[op[2]]);
          for (uint32_t j = 0; j < len; j++) {
            uint32_t v = to_mont(mt.vals[(size_t)(ptr - 1) * len + j]);
            map[map_len++] = v;
            st(aux++, v);
          }
          push_require();
        } break;
        case OP_EXTERN: {
          const uint32_t* chip = chips + op[1] * CHIP_WORDS;
          const uint32_t kind = chip[0];
          if (kind == 1) extern_poseidon<24, Store>(op, map, map_len, aux, st);
          else if (kind == 2) extern_poseidon<32, Store>(op, map, map_len, aux, st);
          else if (kind == 3) extern_poseidon<40, Store>(op, map, map_len, aux, st);
          else if (kind == 10) {  // BigNumCompareWitness (src/gadgets/big_num/cmp.rs:12-45), 28 columns
            uint32_t l = 0, r = 0;
            for (int i = 7; i >= 0; i--) {
              uint32_t a = map[op[3 + i]], b = map[op[11 + i]];
              if (a != b) {
                st(aux + i, R1);      // is_comp
                st(aux + 8, a);       // lhs_comp_limb
                st(aux + 9, b);       // rhs_comp_limb
                l = from_mont(a);
                r = from_mont(b);
                break;
              }
            }
            const uint32_t vv[2] = {l, r};
            for (int k = 0; k < 2; k++) {  // FieldToWord32: is_msb_less_than, the 4 bytes
              st(aux + 10 + 5 * k, (vv[k] >> 24) < 0x78 ? R1 : 0u);
              for (int j = 0; j < 4; j++) st(aux + 11 + 5 * k + j, to_mont((vv[k] >> (8 * j)) & 255u));
            }
            int i = 3;  // CompareWitness<4> on the two words
            while (i >= 0 && ((l >> (8 * i)) & 255u) == ((r >> (8 * i)) & 255u)) i--;
            uint32_t lt = 0;
            if (i >= 0) {
              uint32_t a = (l >> (8 * i)) & 255u, b = (r >> (8 * i)) & 255u;
              uint32_t am = to_mont(a), bm = to_mont(b);
              st(aux + 20 + i, R1);
              st(aux + 24, am);
              st(aux + 25, bm);
              st(aux + 26, finv(fsub(am, bm)));
              lt = a < b ? R1 : 0u;
              st(aux + 27, lt);
            }
            map[map_len++] = lt;
            aux += 28;
          }
          else if (kind >= 4 && kind <= 9) {  // U64::populate_witness, src/core/u64.rs:130-168
            const uint32_t nin = kind == 7 ? 8u : 16u;
            uint64_t x = 0, y = 0;
            for (uint32_t i = 0; i < nin; i++) {
              uint64_t v = from_mont(map[op[3 + i]]);
              if (i < 8) x |= v << (8 * i);
              else y |= v << (8 * (i - 8));
            }
            auto byte = [](uint64_t v, int i) { return (uint32_t)(v >> (8 * i)) & 255u; };
            auto aux_word = [&](uint64_t v) {
              for (int i = 0; i < 8; i++) st(aux++, to_mont(byte(v, i)));
            };
            auto map_word = [&](uint64_t v) {
              for (int i = 0; i < 8; i++) map[map_len++] = to_mont(byte(v, i));
            };
            auto non_zero_witness = [&](uint64_t v) {  // IsZeroWitness::populate_non_zero (is_zero.rs:17-28); zeros if v == 0
              bool found = false;
              for (int i = 0; i < 8; i++) {
                uint32_t limb = byte(v, i);
                if (!found && limb) {
                  st(aux + i, finv(to_mont(limb)));
                  found = true;
                }
              }
              aux += 8;
            };
            // CompareWitness (cmp.rs:12-46): is_comp[8], lhs limb, rhs limb, 1/(lhs - rhs), is_less_than;
            // LessThanWitness (less_than.rs:12-41) is its first 10 columns
            auto compare_witness = [&](uint64_t l, uint64_t r, bool full) -> uint32_t {
              int i = 7;
              while (i >= 0 && byte(l, i) == byte(r, i)) i--;
              uint32_t lt = 0;
              if (i >= 0) {
                uint32_t lm = to_mont(byte(l, i)), rm = to_mont(byte(r, i));
                st(aux + i, R1);
                st(aux + 8, lm);
                st(aux + 9, rm);
                lt = byte(l, i) < byte(r, i) ? R1 : 0u;
                if (full) {
                  st(aux + 10, finv(fsub(lm, rm)));
                  st(aux + 11, lt);
                }
              }
              aux += full ? 12 : 10;
              return lt;
            };
            auto product_witness = [&](uint64_t l, uint64_t r) -> uint64_t {  // MulWitness carries (mul.rs:27-63), then the result
              uint32_t prod[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              for (int i = 0; i < 8; i++)
                for (int j = 0; i + j < 8; j++) prod[i + j] += byte(l, i) * byte(r, j);
              uint32_t carry = 0;
              for (int k = 0; k < 8; k++) {
                carry = (prod[k] + carry) >> 8;
                st(aux++, to_mont(carry));
              }
              aux_word(l * r);
              return l * r;
            };
            switch (kind) {
              case 4: case 5: {  // Sum / Diff (add.rs:63-175): the wrapping result's bytes
                const uint64_t z = kind == 4 ? x + y : x - y;
                map_word(z);
                aux_word(z);
              } break;
              case 6: map[map_len++] = compare_witness(x, y, true); break;
              case 7: {  // IsZeroOrEqual::populate_is_zero (is_zero.rs:113-125): inverses[8], result
                non_zero_witness(x);
                const uint32_t z = x == 0 ? R1 : 0u;
                st(aux++, z);
                map[map_len++] = z;
              } break;
              case 8: map_word(product_witness(x, y)); break;
              default: {  // DivRem (div_rem.rs:17-59): b_non_zero, q, qb, r, r_lt_b, qb_cmp_a
                const uint64_t q = y ? x / y : 0;
                non_zero_witness(y);
                aux_word(q);
                const uint64_t qb = product_witness(q, y);
                const uint64_t r = x - qb;
                aux_word(r);
                compare_witness(r, y, false);
                compare_witness(qb, x, true);
                map_word(q);
                map_word(r);
              }
            }
          }
          for (uint32_t i = 0; i < chip[4]; i++) push_require();  // order: witness, requires (trace.rs:398-406)
        } break;
        case OP_RANGE_U8: {
          uint32_t n = (op[1] + 1) / 2;
          for (uint32_t i = 0; i < n; i++) push_require();
        } break;
        default: break;
      }
    }
    const uint32_t* c = w + pc;
    if (c[0] == CTRL_RETURN) {
      st(col_sel + c[1], R1);
      return;
    }
    // CHOOSE / CHOOSE_MANY
    uint32_t nv = c[1];
    const uint32_t* vars = c + 2;
    uint32_t nb = c[2 + nv], has_d = c[3 + nv];
    uint32_t p = pc + 4 + nv;
    uint32_t chosen = 0;
    for (uint32_t b = 0; b < nb; b++) {
      uint32_t nk = w[p++];
      bool hit = false;
      if (c[0] == CTRL_CHOOSE) {
        uint32_t v = from_mont(map[vars[0]]);
        for (uint32_t k = 0; k < nk; k++) hit |= w[p + k] == v;
        p += nk;
      } else {
        hit = true;
        for (uint32_t k = 0; k < nv; k++) hit &= w[p + k] == from_mont(map[vars[k]]);
        p += nv * nk;
      }
      if (hit && !chosen) chosen = p;
      p += w[p];
    }
    if (!chosen && has_d) chosen = p;
    if (!chosen) return;  // unreachable for records produced by the executor
    pc = chosen;
  }
}

template <int MAP_CAP, bool COLMAJOR>
__global__ void __launch_bounds__(128) func_trace_kernel(const uint32_t* __restrict__ w, uint32_t body, FuncHeader hdr,
                                                          const uint32_t* chips, uint32_t func,
                                                          const DevFuncTable* __restrict__ funcs,
                                                          const DevMemTable* __restrict__ mems, uint32_t row_start,
                                                          uint32_t real_rows, uint32_t height, uint32_t log_h,
                                                          uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= height) return;
  if (COLMAJOR && log_h >= 5) {
    // Row i is stored at position bitrev(i) of every column.  With thread = row, a warp's store to one
    // column would hit 32 different sectors (4 useful bytes each; beyond L2 capacity that is a DRAM
    // read-modify-write per cell).  Instead a warp takes the 32 rows whose top three and bottom two index
    // bits vary: the 8 rows sharing the bottom bits land on 8 consecutive positions -- one full 32 B
    // sector per store -- while the record tables are still read in runs of 4 consecutive rows.
    const uint32_t lane = i & 31, m = i >> 5;
    i = ((lane & 7u) << (log_h - 3)) | (m << 2) | (lane >> 3);
  }
  uint32_t width = 1 + hdr.in + hdr.out + hdr.aux + hdr.sel;
  const DevFuncTable self = funcs[func];
  if (COLMAJOR) {
    StoreColBitrev st{out, height, bitrev32(i, log_h)};
    st(0, to_mont(row_start + i));  // nonces are unique even on padding rows (trace.rs:82-84)
    if (i < real_rows) populate_row<MAP_CAP>(w, body, hdr, chips, funcs, mems, self, row_start + i, st);
  } else {
    StoreRowMajor st{out, width, i};
    st(0, to_mont(row_start + i));
    if (i < real_rows) populate_row<MAP_CAP>(w, body, hdr, chips, funcs, mems, self, row_start + i, st);
  }
}

// MemChip: [is_real, ptr = i + 1, last_nonce, last_count, values...]
template <bool COLMAJOR>
__global__ void mem_trace_kernel(DevMemTable mt, uint32_t height, uint32_t log_h, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mt.rows) return;
  uint32_t width = 4 + mt.len;
  auto put = [&](uint32_t col, uint32_t canonical) {
    if (COLMAJOR) out[(size_t)col * height + bitrev32(i, log_h)] = to_mont(canonical);
    else out[(size_t)i * width + col] = canonical;
  };
  put(0, 1);
  put(1, i + 1);
  put(2, mt.provide[(size_t)i * 2]);
  put(3, mt.provide[(size_t)i * 2 + 1]);
  for (uint32_t j = 0; j < mt.len; j++) put(4 + j, mt.vals[(size_t)i * mt.len + j]);
}

// BytesChip main: [is_real, 6 x (last_nonce, last_count)] for the 2^16 byte pairs
template <bool COLMAJOR>
__global__ void bytes_trace_kernel(const uint32_t* __restrict__ rec, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 65536) return;
  auto put = [&](uint32_t col, uint32_t canonical) {
    if (COLMAJOR) out[(size_t)col * 65536 + bitrev32(i, 16)] = to_mont(canonical);
    else out[(size_t)i * 13 + col] = canonical;
  };
  put(0, 1);
  for (int k = 0; k < 12; k++) put(1 + k, rec[(size_t)i * 12 + k]);
}
// BytesChip preprocessed: (b1, b2, b1 < b2, and, xor, or), index = b1 | b2 << 8
template <bool COLMAJOR>
__global__ void bytes_prep_kernel(uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 65536) return;
  uint32_t b1 = i & 255, b2 = i >> 8;
  uint32_t v[6] = {b1, b2, b1 < b2 ? 1u : 0u, b1 & b2, b1 ^ b2, b1 | b2};
  for (int k = 0; k < 6; k++) {
    if (COLMAJOR) out[(size_t)k * 65536 + bitrev32(i, 16)] = to_mont(v[k]);
    else out[(size_t)i * 6 + k] = v[k];
  }
}

void init_trace_kernels() { LSK_CUDA(p2_upload_constants()); }

void DeviceRecord::upload(lsk_ctx* ctx, const lsk_program& prog, const lsk_record& rec, uint32_t shard,
                          uint32_t shard_size) {
  bufs.clear();
  h2d_bytes = 0;
  const bool sliced = shard_size != 0 && record_row_ordered(rec);
  // words [first, first + count) of v; the returned pointer is biased by -first so that v's own indices work
  // In a device group (comm.cuh) every rank holds the same host record.  The large tables are laid out in ONE device
  // arena; rank r copies only the r-th 1/G of the arena over its own PCIe link and a single all-gather over NVLink fills
  // in the rest -- G links share the upload instead of every rank (or a leader) pushing the whole record.
  const bool grouped = comm_active(ctx);
  struct Piece { size_t off; const uint32_t* src; size_t count; };
  std::vector<Piece> pieces;
  uint32_t* arena = nullptr;
  size_t arena_off = 0, arena_per = 0;
  if (grouped) {
    size_t total = 0;
    auto add = [&](const std::vector<uint32_t>& v) { total += v.size() + 4; };
    for (const FuncTable& t : rec.funcs) {
      add(t.args); add(t.outs); add(t.provide); add(t.req_off); add(t.reqs); add(t.hint_off); add(t.hints);
      add(t.depth); add(t.dreq_off); add(t.dreqs);
    }
    for (const MemTable& m : rec.mems) { add(m.vals); add(m.provide); }
    const size_t g = (size_t)ctx->comm->size;
    arena_per = ((total + g - 1) / g + 3) & ~(size_t)3;
    bufs.emplace_back(ctx, arena_per * g);
    arena = bufs.back().p;
  }
  // copies are collected first (tag: -1 = read by any trace kernel, f = read by func f's own trace only) and issued in
  // waves below
  struct Copy { uint32_t* dst; const uint32_t* src; size_t count; int tag; };
  std::vector<Copy> copies;
  int tag = -1;
  auto up_range = [&](const std::vector<uint32_t>& v, size_t first, size_t count) -> const uint32_t* {
    if (!count) return nullptr;
    if (grouped && count * 4 >= (64u << 10)) {
      const size_t off = arena_off;
      arena_off += (count + 3) & ~(size_t)3;
      pieces.push_back({off, v.data() + first, count});
      return arena + off - first;
    }
    bufs.emplace_back(ctx, count);
    copies.push_back({bufs.back().p, v.data() + first, count, tag});
    h2d_bytes += count * 4;
    return bufs.back().p - first;
  };
  auto up = [&](const std::vector<uint32_t>& v) { return up_range(v, 0, v.size()); };
  h_funcs.clear();
  for (const FuncTable& t : rec.funcs) {
    DevFuncTable d{};
    const size_t n = t.rows();
    size_t s = 0, e = n;
    if (sliced) s = std::min(n, (size_t)shard * shard_size), e = std::min(n, (size_t)(shard + 1) * shard_size);
    const size_t cnt = e - s;
    const int self = (int)h_funcs.size();
    tag = -1;
    d.outs = up(t.outs);
    d.depth = up(t.depth);
    tag = t.invertible ? -1 : self;   // arguments are gathered by other funcs' rows only through `preimg` of an invertible func
    d.args = t.invertible || !sliced ? up(t.args) : up_range(t.args, s * t.in, cnt * t.in);
    tag = self;
    d.provide = up_range(t.provide, 2 * s, 2 * cnt);
    d.req_off = up_range(t.req_off, s, cnt);
    d.hint_off = up_range(t.hint_off, s, cnt);
    d.dreq_off = t.partial ? up_range(t.dreq_off, s, cnt) : nullptr;
    if (cnt) {
      // row-ordered CSR: the rows [s, e) own [off[s], off[e]) of each array
      size_t r0 = t.req_off[s], r1 = e < n ? t.req_off[e] : t.reqs.size() / 2;
      size_t h0 = t.hint_off[s], h1 = e < n ? t.hint_off[e] : t.hints.size();
      if (!sliced) r0 = 0, r1 = t.reqs.size() / 2, h0 = 0, h1 = t.hints.size();
      d.reqs = up_range(t.reqs, 2 * r0, 2 * (r1 - r0));
      d.hints = up_range(t.hints, h0, h1 - h0);
      if (t.partial) {
        size_t q0 = t.dreq_off[s], q1 = e < n ? t.dreq_off[e] : t.dreqs.size() / 2;
        if (!sliced) q0 = 0, q1 = t.dreqs.size() / 2;
        d.dreqs = up_range(t.dreqs, 2 * q0, 2 * (q1 - q0));
      }
    }
    d.in = t.in; d.out = t.out; d.rows = (uint32_t)n;
    h_funcs.push_back(d);
  }
  h_mems.clear();
  tag = -1;
  for (const MemTable& m : rec.mems) {
    DevMemTable d{};
    d.vals = up(m.vals);
    d.provide = !sliced || shard == 0 ? up(m.provide) : nullptr;  // the Mem chips live in shard 0
    d.len = m.len; d.rows = (uint32_t)m.rows();
    h_mems.push_back(d);
  }
  if (grouped && arena_off) {
    const size_t b = (size_t)ctx->comm->rank * arena_per, e = b + arena_per;   // this rank's words of the arena
    for (const Piece& pc : pieces) {
      const size_t lo = std::max(b, pc.off), hi = std::min(e, pc.off + pc.count);
      if (lo >= hi) continue;
      LSK_CUDA(cudaMemcpyAsync(arena + lo, pc.src + (lo - pc.off), (hi - lo) * 4, cudaMemcpyHostToDevice, ctx->stream));
      h2d_bytes += (hi - lo) * 4;
    }
    comm_allgather(ctx, arena, arena_per * 4);
  }
  // the copy stream: the side stream (behind everything the main stream has queued: the pool may hand out blocks that
  // earlier kernels still use), or the main stream itself inside a device group
  const bool async_up = !grouped && ctx->side && !getenv("LSK_NO_SIDE_STREAM");
  cudaStream_t cs = async_up ? ctx->side : ctx->stream;
  if (async_up) {
    LSK_CUDA(cudaEventRecord(ctx->side_fork, ctx->stream));
    LSK_CUDA(cudaStreamWaitEvent(cs, ctx->side_fork, 0));
  }
  auto h2d = [&](void* dst, const void* src, size_t bytes) {
    LSK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cs));
  };
  funcs.alloc(ctx, h_funcs.size());
  if (!h_funcs.empty()) h2d(funcs.p, h_funcs.data(), h_funcs.size() * sizeof(DevFuncTable));
  mems.alloc(ctx, h_mems.size());
  h2d(mems.p, h_mems.data(), h_mems.size() * sizeof(DevMemTable));
  program.alloc(ctx, prog.words.size());
  h2d(program.p, prog.words.data(), prog.words.size() * 4);
  if (rec.bytes_used && (!sliced || shard == 0)) {  // the byte table's main trace is non-empty in shard 0 only
    bytes.alloc(ctx, rec.bytes.size());
    h2d(bytes.p, rec.bytes.data(), rec.bytes.size() * 4);
    h2d_bytes += rec.bytes.size() * 4;
  } else {
    bytes.release();
  }
  auto event = [&](cudaEvent_t& e) {
    if (!e) LSK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    LSK_CUDA(cudaEventRecord(e, cs));
  };
  for (const Copy& c : copies)
    if (c.tag < 0) h2d(c.dst, c.src, c.count * 4);
  event(common_ready);
  func_ready.resize(h_funcs.size(), nullptr);
  size_t at = 0;   // copies are in func order
  for (size_t f = 0; f < h_funcs.size(); f++) {
    bool any = false;
    for (; at < copies.size() && copies[at].tag <= (int)f; at++)
      if (copies[at].tag == (int)f) h2d(copies[at].dst, copies[at].src, copies[at].count * 4), any = true;
    if (any) event(func_ready[f]);
    else if (func_ready[f]) cudaEventDestroy(func_ready[f]), func_ready[f] = nullptr;
  }
  event(all_ready);
}

void func_trace_dev(lsk_ctx* ctx, const lsk_program& prog, const DeviceRecord& drec, uint32_t func, uint32_t row_start,
                    uint32_t real_rows, uint32_t height, TraceLayout layout, uint32_t* d_out) {
  FuncHeader hdr = prog.header(func);
  uint32_t width = prog.width(func);
  LSK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)height * width * 4, ctx->stream));
  if (!height) return;
  drec.wait_func(ctx, func);   // the tables this trace reads may still be on their way (DeviceRecord::upload)
  uint32_t log_h = log2_exact(height);
  unsigned grid = div_up(height, 128);
  const uint32_t* w = drec.program.p;
  const uint32_t* chips = w + 3;
  uint32_t body = prog.body(func);
  bool col = layout == TraceLayout::ColMajorBitrevMont;
  ProfScope ps(ctx, PT_TRACE, 1, 4.0 * height * width + 4.0 * real_rows * (hdr.in + hdr.out + 6));
#define LAUNCH(CAP)                                                                                                  \
  do {                                                                                                               \
    if (col)                                                                                                         \
      func_trace_kernel<CAP, true><<<grid, 128, 0, ctx->stream>>>(w, body, hdr, chips, func, drec.funcs.p,          \
                                                                  drec.mems.p, row_start, real_rows, height, log_h, \
                                                                  d_out);                                            \
    else                                                                                                             \
      func_trace_kernel<CAP, false><<<grid, 128, 0, ctx->stream>>>(w, body, hdr, chips, func, drec.funcs.p,         \
                                                                   drec.mems.p, row_start, real_rows, height, log_h, \
                                                                   d_out);                                           \
  } while (0)
  if (hdr.max_map <= 32) LAUNCH(32);
  else if (hdr.max_map <= 128) LAUNCH(128);
  else if (hdr.max_map <= 512) LAUNCH(512);
  else fail(LSK_ERR_ARG, "func %u needs a value stack of %u entries (> 512)", func, hdr.max_map);
#undef LAUNCH
  LSK_CUDA(cudaGetLastError());
}

void mem_trace_dev(lsk_ctx* ctx, const DeviceRecord& drec, uint32_t mem_idx, uint32_t height, TraceLayout layout,
                   uint32_t* d_out) {
  const DevMemTable& mt = drec.h_mems[mem_idx];
  LSK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)height * (4 + mt.len) * 4, ctx->stream));
  if (!mt.rows) return;
  drec.wait_common(ctx);
  uint32_t log_h = log2_exact(height);
  ProfScope ps(ctx, PT_TRACE, 1, 8.0 * mt.rows * (4 + mt.len));
  if (layout == TraceLayout::ColMajorBitrevMont)
    mem_trace_kernel<true><<<div_up(mt.rows, 128), 128, 0, ctx->stream>>>(mt, height, log_h, d_out);
  else
    mem_trace_kernel<false><<<div_up(mt.rows, 128), 128, 0, ctx->stream>>>(mt, height, log_h, d_out);
  LSK_CUDA(cudaGetLastError());
}

void bytes_trace_dev(lsk_ctx* ctx, const DeviceRecord& drec, bool real, TraceLayout layout, uint32_t* d_out) {
  LSK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)65536 * 13 * 4, ctx->stream));
  if (!real || !drec.bytes.p) return;
  drec.wait_common(ctx);
  ProfScope ps(ctx, PT_TRACE, 1, 65536.0 * 25 * 4);
  if (layout == TraceLayout::ColMajorBitrevMont)
    bytes_trace_kernel<true><<<65536 / 256, 256, 0, ctx->stream>>>(drec.bytes.p, d_out);
  else
    bytes_trace_kernel<false><<<65536 / 256, 256, 0, ctx->stream>>>(drec.bytes.p, d_out);
  LSK_CUDA(cudaGetLastError());
}

void bytes_preprocessed_dev(lsk_ctx* ctx, TraceLayout layout, uint32_t* d_out) {
  ProfScope ps(ctx, PT_TRACE, 1, 65536.0 * 6 * 4);
  if (layout == TraceLayout::ColMajorBitrevMont)
    bytes_prep_kernel<true><<<65536 / 256, 256, 0, ctx->stream>>>(d_out);
  else
    bytes_prep_kernel<false><<<65536 / 256, 256, 0, ctx->stream>>>(d_out);
  LSK_CUDA(cudaGetLastError());
}

}  // namespace lsk
