// Host-side Lair executor that emits the flattened, GPU-ready QueryRecord (record.h).
//
// Mirrors the behaviour of Toplevel::execute / Func::execute (src/lair/execute.rs:375-783):
// memoised evaluation where every distinct (func, args) becomes one row whose index is its
// insertion order (the nonce), `provide`/`require` bookkeeping through Record::new_lookup
// (src/air/builder.rs:141-148), the pre-seeded `provide.count = 1` of the entry query
// (execute.rs:444-447), memory tables by length class (:641-664) and byte range checks (:673-682).
// Re-designed rather than transliterated: it walks the flat program image with a program counter
// and three shared stacks (values, require records, hints) instead of cloning per-call vectors,
// and it records callee / slot row indices so no hashing is needed at trace time.
//
// This is the producer of the hot path's input; it is sequential by construction (memo table +
// nonce order) and stays on the CPU exactly like the reference's.
#include <pthread.h>
#include <sys/mman.h>
#include <exception>
#include <cstring>
#include <memory>

#include "common.h"
#include "poseidon2_constants.h"
#include "record.h"

namespace lsk {

static const uint64_t Pq = 0x78000001ull;
static inline uint32_t hadd(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= Pq ? s - Pq : s; }
static inline uint32_t hsub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + (uint32_t)Pq - b; }
static inline uint32_t hmul(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b % Pq); }
static inline uint32_t hpow(uint32_t b, uint64_t e) {
  uint32_t r = 1;
  while (e) { if (e & 1) r = hmul(r, b); b = hmul(b, b); e >>= 1; }
  return r;
}
static inline uint32_t hinv(uint32_t a) { return hpow(a, Pq - 2); }

// Host Poseidon2 (canonical arithmetic) for ExternCall hashers: PoseidonChipset::execute_simple,
// src/core/poseidon.rs:61-63.  Only used by the executor; the proving path hashes on the GPU.
static void host_external_linear(uint32_t* s, int t) {
  for (int i = 0; i < t; i += 4) {
    uint64_t x0 = s[i], x1 = s[i + 1], x2 = s[i + 2], x3 = s[i + 3];
    s[i] = (uint32_t)((2 * x0 + 3 * x1 + x2 + x3) % Pq);
    s[i + 1] = (uint32_t)((x0 + 2 * x1 + 3 * x2 + x3) % Pq);
    s[i + 2] = (uint32_t)((x0 + x1 + 2 * x2 + 3 * x3) % Pq);
    s[i + 3] = (uint32_t)((3 * x0 + x1 + x2 + 2 * x3) % Pq);
  }
  uint64_t sums[4] = {0, 0, 0, 0};
  for (int i = 0; i < t; i++) sums[i & 3] += s[i];
  for (int i = 0; i < t; i++) s[i] = (uint32_t)((s[i] + sums[i & 3]) % Pq);
}
static inline uint32_t host_sbox(uint32_t x) {
  uint32_t x3 = hmul(hmul(x, x), x);
  return hmul(x, hmul(x3, x3));
}
static void host_poseidon2(uint32_t* s, int t) {
  const uint32_t *diag, *full, *part;
  int rp;
  switch (t) {
    case 24: diag = P2_DIAG_24; full = P2_FULL_RC_24; part = P2_PART_RC_24; rp = P2_RP_24; break;
    case 32: diag = P2_DIAG_32; full = P2_FULL_RC_32; part = P2_PART_RC_32; rp = P2_RP_32; break;
    case 40: diag = P2_DIAG_40; full = P2_FULL_RC_40; part = P2_PART_RC_40; rp = P2_RP_40; break;
    default: fail(LSK_ERR_EXEC, "unknown extern chip width %d", t);
  }
  host_external_linear(s, t);
  auto full_round = [&](int r) {
    for (int i = 0; i < t; i++) s[i] = host_sbox(hadd(s[i], full[r * t + i]));
    host_external_linear(s, t);
  };
  for (int r = 0; r < 4; r++) full_round(r);
  for (int r = 0; r < rp; r++) {
    s[0] = host_sbox(hadd(s[0], part[r]));
    uint64_t sum = 0;
    for (int i = 0; i < t; i++) sum += s[i];
    sum %= Pq;
    for (int i = 0; i < t; i++) s[i] = (uint32_t)(((uint64_t)s[i] * diag[i] + sum) % Pq);
  }
  for (int r = 4; r < 8; r++) full_round(r);
}
static inline int chip_width(uint32_t kind) { return kind == 1 ? 24 : kind == 2 ? 32 : kind == 3 ? 40 : 0; }

struct Frame {
  uint32_t func, nonce, pc, ops_left;
  size_t map_base, req_base, hint_base, depth_base;
  bool preimg;  // this frame was entered through PreImg: the caller receives the inputs
};

struct Executor {
  const lsk_program& prog;
  lsk_record& rec;
  const uint32_t* w;
  std::vector<uint32_t> vals, req_stack, hint_stack, depth_stack;
  std::vector<Frame> frames;
  // replay: the record is complete (imported from the caller's own executor, lsk_record_import_*); rows are re-run one
  // by one only to find which callee / memory row each Call, PreImg and Store refers to (the hints); nothing else changes
  bool replay = false;
  uint32_t replay_sink[2] = {0, 0};
  Executor(const lsk_program& p, lsk_record& r) : prog(p), rec(r), w(p.words.data()) {}

  void new_lookup(uint32_t* provide, uint32_t nonce, std::vector<uint32_t>& reqs) {
    if (replay) return;
    reqs.push_back(provide[0]);
    reqs.push_back(provide[1]);
    provide[0] = nonce;
    provide[1] += 1;
  }
  uint32_t insert_query(uint32_t f, const uint32_t* args, uint32_t seed_count) {
    FuncTable& t = rec.funcs[f];
    uint32_t row = (uint32_t)t.rows();
    t.args.insert(t.args.end(), args, args + t.in);
    t.outs.resize(t.outs.size() + t.out, 0);
    t.done.push_back(0);
    t.provide.push_back(0);
    t.provide.push_back(seed_count);
    t.req_off.push_back(0);
    t.hint_off.push_back(0);
    t.req_len.push_back(0);
    t.hint_len.push_back(0);
    if (t.partial) {
      t.depth.push_back(0);
      t.dreq_off.push_back(0);
      t.dreq_len.push_back(0);
    }
    t.index.insert_new(t.args.data(), t.in, row);
    return row;
  }
  void enter_block(Frame& f, uint32_t block) {
    f.ops_left = w[block + 1];
    f.pc = block + 2;
  }
  void push_frame(uint32_t func, uint32_t nonce, const uint32_t* inp, uint32_t n, bool preimg) {
    Frame f{func, nonce, 0, 0, vals.size(), req_stack.size(), hint_stack.size(), depth_stack.size(), preimg};
    // NB `inp` may point into `vals`; copy before growing
    std::vector<uint32_t> tmp(inp, inp + n);
    vals.insert(vals.end(), tmp.begin(), tmp.end());
    enter_block(f, prog.body(func));
    frames.push_back(f);
  }
  // a memoised or just-finished callee row is consumed by the frame on top of the stack
  void consume(uint32_t callee, uint32_t row, bool preimg) {
    Frame& f = frames.back();
    FuncTable& t = rec.funcs[callee];
    const uint32_t* src = preimg ? &t.args[(size_t)row * t.in] : &t.outs[(size_t)row * t.out];
    uint32_t n = preimg ? t.in : t.out;
    std::vector<uint32_t> tmp(src, src + n);
    vals.insert(vals.end(), tmp.begin(), tmp.end());
    new_lookup(&t.provide[(size_t)row * 2], f.nonce, req_stack);
    hint_stack.push_back(row);
    if (t.partial && rec.funcs[f.func].partial) depth_stack.push_back(t.depth[row]);  // execute.rs:510-512,751-753
  }
  uint32_t* byte_record(uint32_t i1, uint32_t i2, int slot) {  // BytesInputRecord field of ByteInput (i1, i2)
    if (replay) return replay_sink;
    if (rec.bytes.empty()) rec.bytes.assign((size_t)65536 * 12, 0);
    rec.bytes_used = true;
    return &rec.bytes[(size_t)(i1 | (i2 << 8)) * 12 + 2 * slot];
  }

  void run(uint32_t entry, const uint32_t* args, uint32_t nargs, std::vector<uint32_t>& out) {
    FuncTable& et = rec.funcs[entry];
    if (nargs != et.in) fail(LSK_ERR_EXEC, "Argument mismatch");
    long existing = et.index.find(et.args.data(), et.in, args);
    uint32_t row;
    if (existing >= 0) {
      // IndexMap::insert_full overwrites the value of an existing key (execute.rs:446-447)
      row = (uint32_t)existing;
      et.done[row] = 0;
      et.provide[(size_t)row * 2] = 0;
      et.provide[(size_t)row * 2 + 1] = 1;
    } else {
      row = insert_query(entry, args, 1);
    }
    push_frame(entry, row, args, nargs, false);
    while (!frames.empty()) step(out);
    rec.public_values.assign(args, args + nargs);
    rec.public_values.insert(rec.public_values.end(), out.begin(), out.end());
    if (et.partial)  // execute.rs:387-389: the entry's depth as little-endian bytes
      for (int b = 0; b < 4; b++) rec.public_values.push_back((et.depth[row] >> (8 * b)) & 255u);
  }

  void step(std::vector<uint32_t>& final_out) {
    Frame& f = frames.back();
    uint32_t* m = vals.data() + f.map_base;
    if (f.ops_left) {
      const uint32_t* op = w + f.pc;
      f.pc += op_words(op);
      f.ops_left--;
      switch (op[0]) {
        case OP_CONST: vals.push_back(op[1]); break;
        case OP_ADD: vals.push_back(hadd(m[op[1]], m[op[2]])); break;
        case OP_SUB: vals.push_back(hsub(m[op[1]], m[op[2]])); break;
        case OP_MUL: vals.push_back(hmul(m[op[1]], m[op[2]])); break;
        case OP_INV:
          if (!m[op[1]]) fail(LSK_ERR_EXEC, "inverse of zero");
          vals.push_back(hinv(m[op[1]]));
          break;
        case OP_NOT: vals.push_back(m[op[1]] == 0 ? 1u : 0u); break;
        case OP_ASSERT_EQ: {
          uint32_t n = op[1];
          for (uint32_t i = 0; i < n; i++)
            if (m[op[2 + i]] != m[op[2 + n + i]]) fail(LSK_ERR_EXEC, "AssertEq failed");
        } break;
        case OP_ASSERT_NE: {
          uint32_t n = op[1];
          bool ne = false;
          for (uint32_t i = 0; i < n; i++) ne |= m[op[2 + i]] != m[op[2 + n + i]];
          if (!ne) fail(LSK_ERR_EXEC, "AssertNe failed");
        } break;
        case OP_CONTAINS: {
          uint32_t n = op[1], b = m[op[2 + n]];
          bool found = false;
          for (uint32_t i = 0; i < n; i++) found |= m[op[2 + i]] == b;
          if (!found) fail(LSK_ERR_EXEC, "Contains failed");
        } break;
        case OP_CALL: {
          uint32_t callee = op[1], n = op[2];
          FuncTable& t = rec.funcs[callee];
          uint32_t inp[64];
          if (n > 64) fail(LSK_ERR_EXEC, "call arity too large");
          for (uint32_t i = 0; i < n; i++) inp[i] = m[op[3 + i]];
          long row = t.index.find(t.args.data(), t.in, inp);
          if (row >= 0) {
            if (!t.done[row]) fail(LSK_ERR_EXEC, "Loop detected");
            consume(callee, (uint32_t)row, false);
          } else {
            if (replay) fail(LSK_ERR_EXEC, "imported record lacks a query of func %u that a row calls", callee);
            uint32_t r = insert_query(callee, inp, 0);
            push_frame(callee, r, inp, n, false);
          }
        } break;
        case OP_PREIMG: {
          uint32_t callee = op[1], n = op[2];
          FuncTable& t = rec.funcs[callee];
          if (!t.invertible) fail(LSK_ERR_EXEC, "Missing inverse map");
          uint32_t outv[64];
          if (n > 64) fail(LSK_ERR_EXEC, "preimg arity too large");
          for (uint32_t i = 0; i < n; i++) outv[i] = m[op[3 + i]];
          long iv = t.inv_index.find(t.inv_outs.data(), t.out, outv);
          if (iv < 0) fail(LSK_ERR_EXEC, "Preimg not found");
          std::vector<uint32_t> inp(t.inv_inps.begin() + (size_t)iv * t.in, t.inv_inps.begin() + (size_t)(iv + 1) * t.in);
          long row = t.index.find(t.args.data(), t.in, inp.data());
          if (row >= 0) {
            if (!t.done[row]) fail(LSK_ERR_EXEC, "Loop detected");
            if (memcmp(&t.outs[(size_t)row * t.out], outv, t.out * 4)) fail(LSK_ERR_EXEC, "preimage output mismatch");
            consume(callee, (uint32_t)row, true);
          } else {
            if (replay) fail(LSK_ERR_EXEC, "imported record lacks a query of func %u that a row takes the preimage of", callee);
            uint32_t r = insert_query(callee, inp.data(), 0);
            push_frame(callee, r, inp.data(), t.in, true);
          }
        } break;
        case OP_STORE: {
          uint32_t n = op[1];
          int mi = mem_index_from_len((int)n);
          if (mi < 0) fail(LSK_ERR_EXEC, "There are no mem tables of size %u", n);
          MemTable& mt = rec.mems[mi];
          uint32_t v[8];
          for (uint32_t i = 0; i < n; i++) v[i] = m[op[2 + i]];
          long row = mt.index.find(mt.vals.data(), n, v);
          if (row < 0) {
            if (replay) fail(LSK_ERR_EXEC, "imported record lacks a memory slot that a row stores");
            row = (long)mt.rows();
            mt.vals.insert(mt.vals.end(), v, v + n);
            mt.provide.push_back(0);
            mt.provide.push_back(0);
            mt.index.insert_new(mt.vals.data(), n, (uint32_t)row);
          }
          vals.push_back((uint32_t)row + 1);
          new_lookup(&mt.provide[(size_t)row * 2], f.nonce, req_stack);
          hint_stack.push_back((uint32_t)row);
        } break;
        case OP_LOAD: {
          uint32_t len = op[1], ptr = m[op[2]];
          int mi = mem_index_from_len((int)len);
          if (mi < 0) fail(LSK_ERR_EXEC, "There are no mem tables of size %u", len);
          MemTable& mt = rec.mems[mi];
          if (ptr == 0 || ptr > mt.rows()) fail(LSK_ERR_EXEC, "Unbound pointer");
          uint32_t v[8];
          memcpy(v, &mt.vals[(size_t)(ptr - 1) * len], len * 4);
          vals.insert(vals.end(), v, v + len);
          new_lookup(&mt.provide[(size_t)(ptr - 1) * 2], f.nonce, req_stack);
        } break;
        case OP_EXTERN: {
          const uint32_t* chip = prog.chip(op[1]);
          if (chip[0] == 10) {  // BigNum::LessThan::execute (src/core/big_num.rs:42-60) = BigNumCompareWitness::populate
            if (op[2] != 16) fail(LSK_ERR_EXEC, "bad extern chip call");
            uint32_t l = 0, r = 0;  // most significant differing field limbs (big_num/cmp.rs:31-45); (0, 0) when equal
            for (int i = 7; i >= 0; i--)
              if (m[op[3 + i]] != m[op[11 + i]]) {
                l = m[op[3 + i]];
                r = m[op[11 + i]];
                break;
              }
            for (uint32_t v : {l, r}) {  // FieldToWord32::populate: msb < 0x78 lookup, then the 4 bytes in pairs
              new_lookup(byte_record(v >> 24, 0x78, 2), f.nonce, req_stack);
              new_lookup(byte_record(v & 255u, (v >> 8) & 255u, 0), f.nonce, req_stack);
              new_lookup(byte_record((v >> 16) & 255u, v >> 24, 0), f.nonce, req_stack);
            }
            int i = 3;  // CompareWitness<4>::populate on the two words
            while (i >= 0 && ((l >> (8 * i)) & 255u) == ((r >> (8 * i)) & 255u)) i--;
            new_lookup(byte_record(i >= 0 ? (l >> (8 * i)) & 255u : 0, i >= 0 ? (r >> (8 * i)) & 255u : 0, 2), f.nonce, req_stack);
            vals.push_back(l < r ? 1u : 0u);
            break;
          }
          if (chip[0] >= 4 && chip[0] <= 9) {  // U64::execute, src/core/u64.rs:76-128
            const uint32_t nin = chip[0] == 7 ? 8u : 16u;
            if (op[2] != nin) fail(LSK_ERR_EXEC, "bad extern chip call");
            uint64_t x = 0, y = 0;
            for (uint32_t i = 0; i < nin; i++) {
              uint64_t v = m[op[3 + i]];
              if (v > 255) fail(LSK_ERR_EXEC, "u64 chip input is not a byte");  // into_u64: try_into().unwrap()
              if (i < 8) x |= v << (8 * i);
              else y |= v << (8 * (i - 8));
            }
            auto byte = [](uint64_t v, int i) { return (uint32_t)(v >> (8 * i)) & 255u; };
            auto range_pairs = [&](uint64_t v) {  // UncheckedWord::assign_bytes -> range_check_u8_iter
              for (int i = 0; i < 8; i += 2) new_lookup(byte_record(byte(v, i), byte(v, i + 1), 0), f.nonce, req_stack);
            };
            auto push_word = [&](uint64_t v) {
              for (int i = 0; i < 8; i++) vals.push_back(byte(v, i));
            };
            auto compare = [&](uint64_t l, uint64_t r) {  // CompareWitness::populate, cmp.rs:23-46
              int i = 7;
              while (i >= 0 && byte(l, i) == byte(r, i)) i--;
              uint32_t a = i >= 0 ? byte(l, i) : 0, b = i >= 0 ? byte(r, i) : 0;
              new_lookup(byte_record(a, b, 2), f.nonce, req_stack);
            };
            auto product = [&](uint64_t l, uint64_t r) {  // MulWitness::populate + result range checks, mul.rs:27-63,148-155
              uint32_t prod[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              for (int i = 0; i < 8; i++)
                for (int j = 0; i + j < 8; j++) prod[i + j] += byte(l, i) * byte(r, j);
              uint32_t carry = 0;
              for (int k = 0; k < 8; k++) {
                carry = (prod[k] + carry) >> 8;
                new_lookup(byte_record(carry & 255u, carry >> 8, 1), f.nonce, req_stack);  // range_check_u16(carry)
              }
              range_pairs(l * r);
              return l * r;
            };
            switch (chip[0]) {
              case 4: case 5: {  // Sum / Diff::populate: wrapping result
                uint64_t z = chip[0] == 4 ? x + y : x - y;
                range_pairs(z);
                push_word(z);
              } break;
              case 6:
                compare(x, y);
                vals.push_back(x < y ? 1u : 0u);
                break;
              case 7: vals.push_back(x == 0 ? 1u : 0u); break;  // IsZero: no lookups
              case 8: push_word(product(x, y)); break;
              default: {  // DivRem::populate, div_rem.rs:35-59
                if (y == 0) fail(LSK_ERR_EXEC, "expected input to be non-zero");
                uint64_t q = x / y;
                range_pairs(q);
                uint64_t qb = product(q, y);
                uint64_t r = x - qb;
                range_pairs(r);
                {  // LessThanWitness::populate(r, b)
                  int i = 7;
                  while (byte(r, i) == byte(y, i)) i--;
                  new_lookup(byte_record(byte(r, i), byte(y, i), 2), f.nonce, req_stack);
                }
                compare(qb, x);
                push_word(q);
                push_word(r);
              }
            }
            break;
          }
          int t = chip_width(chip[0]);
          if (!t || op[2] != (uint32_t)t) fail(LSK_ERR_EXEC, "bad extern chip call");
          uint32_t st[40];
          for (int i = 0; i < t; i++) st[i] = m[op[3 + i]];
          host_poseidon2(st, t);
          vals.insert(vals.end(), st, st + 8);
        } break;
        case OP_RANGE_U8: {
          uint32_t n = op[1];
          for (uint32_t i = 0; i < n; i += 2) {
            uint32_t i1 = m[op[2 + i]], i2 = i + 1 < n ? m[op[3 + i]] : 0;
            if (i1 > 255 || i2 > 255) fail(LSK_ERR_EXEC, "Variable not in u8 range");
            new_lookup(byte_record(i1, i2, 0), f.nonce, req_stack);  // slot 0 = range_u8 pair
          }
        } break;
        default: fail(LSK_ERR_EXEC, "bad opcode %u", op[0]);
      }
      return;
    }
    // control node
    const uint32_t* c = w + f.pc;
    if (c[0] == CTRL_RETURN) {
      uint32_t n = c[2];
      FuncTable& t = rec.funcs[f.func];
      if (n != t.out) fail(LSK_ERR_EXEC, "return arity mismatch");
      uint32_t outv[64];
      for (uint32_t i = 0; i < n; i++) outv[i] = m[c[3 + i]];
      if (replay) {
        if (memcmp(&t.outs[(size_t)f.nonce * t.out], outv, n * 4))
          fail(LSK_ERR_EXEC, "imported record: row %u of func %u does not return the recorded output", f.nonce, f.func);
        t.hint_off[f.nonce] = (uint32_t)t.hints.size();
        t.hints.insert(t.hints.end(), hint_stack.begin() + f.hint_base, hint_stack.end());
        t.hint_len[f.nonce] = (uint32_t)(hint_stack.size() - f.hint_base);
        hint_stack.resize(f.hint_base);
        depth_stack.resize(f.depth_base);
        vals.resize(f.map_base);
        frames.pop_back();
        if (!frames.empty()) fail(LSK_ERR_EXEC, "replay frame stack not empty");
        return;
      }
      if (t.done[f.nonce]) fail(LSK_ERR_EXEC, "output already set");
      memcpy(&t.outs[(size_t)f.nonce * t.out], outv, n * 4);
      t.done[f.nonce] = 1;
      if (t.invertible) {  // execute.rs:690-692: inv_map.insert(out, inp) (later insert wins)
        long iv = t.inv_index.find(t.inv_outs.data(), t.out, outv);
        if (iv >= 0) {
          memcpy(&t.inv_inps[(size_t)iv * t.in], &t.args[(size_t)f.nonce * t.in], t.in * 4);
        } else {
          uint32_t slot = (uint32_t)t.inv_index.used;
          t.inv_outs.insert(t.inv_outs.end(), outv, outv + t.out);
          t.inv_inps.insert(t.inv_inps.end(), t.args.begin() + (size_t)f.nonce * t.in,
                            t.args.begin() + (size_t)(f.nonce + 1) * t.in);
          t.inv_index.insert_new(t.inv_outs.data(), t.out, slot);
        }
      }
      if (t.partial) {  // execute.rs:702-711: depth = 1 + max over the partial callees, range check + comparisons
        uint32_t depth = 0;
        for (size_t i = f.depth_base; i < depth_stack.size(); i++) depth = std::max(depth, depth_stack[i] + 1);
        t.dreq_off[f.nonce] = (uint32_t)(t.dreqs.size() / 2);
        for (int b = 0; b < 4; b += 2)
          new_lookup(byte_record((depth >> (8 * b)) & 255u, (depth >> (8 * b + 8)) & 255u, 0), f.nonce, t.dreqs);
        for (size_t i = f.depth_base; i < depth_stack.size(); i++) {
          uint32_t dep = depth_stack[i];  // dep < depth; most significant differing byte (LessThanWitness::populate)
          int b = 3;
          while (((dep >> (8 * b)) & 255u) == ((depth >> (8 * b)) & 255u)) b--;
          new_lookup(byte_record((dep >> (8 * b)) & 255u, (depth >> (8 * b)) & 255u, 2), f.nonce, t.dreqs);
        }
        t.depth[f.nonce] = depth;
        t.dreq_len[f.nonce] = (uint32_t)(t.dreqs.size() / 2) - t.dreq_off[f.nonce];
        depth_stack.resize(f.depth_base);
      }
      // segments are appended in completion order; a row finishing after a higher-numbered one breaks row order
      if (f.nonce != t.completed++) t.row_ordered = false;
      t.req_off[f.nonce] = (uint32_t)(t.reqs.size() / 2);
      t.reqs.insert(t.reqs.end(), req_stack.begin() + f.req_base, req_stack.end());
      t.req_len[f.nonce] = (uint32_t)(req_stack.size() - f.req_base) / 2;
      t.hint_off[f.nonce] = (uint32_t)t.hints.size();
      t.hints.insert(t.hints.end(), hint_stack.begin() + f.hint_base, hint_stack.end());
      t.hint_len[f.nonce] = (uint32_t)(hint_stack.size() - f.hint_base);
      req_stack.resize(f.req_base);
      hint_stack.resize(f.hint_base);
      vals.resize(f.map_base);
      Frame done = f;
      frames.pop_back();
      if (frames.empty()) {
        final_out.assign(outv, outv + n);
      } else {
        consume(done.func, done.nonce, done.preimg);
      }
      return;
    }
    if (c[0] == CTRL_CHOOSE || c[0] == CTRL_CHOOSE_MANY) {
      uint32_t nv = c[1];
      const uint32_t* vars = c + 2;
      uint32_t nb = c[2 + nv], has_d = c[3 + nv];
      uint32_t p = f.pc + 4 + nv;
      long chosen = -1;
      for (uint32_t b = 0; b < nb; b++) {
        uint32_t nk = w[p++];
        bool hit = false;
        if (c[0] == CTRL_CHOOSE) {
          for (uint32_t k = 0; k < nk; k++) hit |= w[p + k] == m[vars[0]];
          p += nk;
        } else {
          hit = true;
          for (uint32_t k = 0; k < nv; k++) hit &= w[p + k] == m[vars[k]];
          p += nv * nk;
        }
        if (hit && chosen < 0) chosen = p;
        p += w[p];  // skip the block
      }
      if (chosen < 0 && has_d) chosen = p;
      if (chosen < 0) fail(LSK_ERR_EXEC, "No match");
      enter_block(f, (uint32_t)chosen);
      return;
    }
    fail(LSK_ERR_EXEC, "bad control word %u", c[0]);
  }
};

// Re-pack the CSR arrays of every func table in row order (see FuncTable::row_ordered).
static void repack(std::vector<uint32_t>& off, const std::vector<uint32_t>& len, std::vector<uint32_t>& data, uint32_t unit) {
  std::vector<uint32_t> out;
  out.reserve(data.size());
  for (size_t i = 0; i < off.size(); i++) {
    uint32_t start = (uint32_t)(out.size() / unit);
    out.insert(out.end(), data.begin() + (size_t)off[i] * unit, data.begin() + ((size_t)off[i] + len[i]) * unit);
    off[i] = start;
  }
  data.swap(out);
}
bool record_row_ordered(const lsk_record& r) {
  for (auto& t : r.funcs)
    if (!t.row_ordered) return false;
  return true;
}

}  // namespace lsk

using namespace lsk;

// The executor and the trace kernel walk the program image without further checks, so lsk_program_load accepts
// only images every walk of which stays inside the image and inside the per-row value stack: block lengths and
// nesting, operand indices against the stack height at that point (and the stack against the declared
// max_map), func / chip / memory-table references and arities, the shape of every control node.  It also keeps
// the reference's load-time rule that a total func may not call a partial one ("assertion failed: ctx.partial",
// src/lair/execute.rs:1005-1023).  The column budget (aux, sel) is the compiler's statement and is trusted.
namespace {
struct ImageChecker {
  const lsk_program& p;
  const uint32_t* w;
  size_t n;
  uint32_t func = 0, max_map = 0, out_size = 0;
  bool total = false;
  [[noreturn]] void bad(const char* what) const { fail(LSK_ERR_ARG, "bad program image: %s (func %u)", what, func); }
  uint32_t at(size_t i, size_t end) const {
    if (i >= end) bad("truncated block");
    return w[i];
  }
  void operands(size_t first, uint32_t count, size_t end, uint32_t height) const {
    if (count > end || first + count > end) bad("operand list runs past its block");
    for (uint32_t k = 0; k < count; k++)
      if (w[first + k] >= height) bad("operand above the value stack");
  }
  uint32_t push(uint32_t height, uint32_t count) const {
    if (count > max_map || height + count > max_map) bad("value stack exceeds max_map");
    return height + count;
  }
  // returns the end of the block starting at `b`; `limit` = end of the enclosing region
  size_t block(size_t b, size_t limit, uint32_t height, unsigned depth) const {
    if (depth > 4096) bad("blocks nested too deeply");
    if (b + 2 > limit) bad("truncated block");
    const size_t len = w[b];
    if (len < 2 || len > limit - b) bad("block length");
    const size_t end = b + len;
    const uint32_t nops = w[b + 1];
    size_t pc = b + 2;
    for (uint32_t k = 0; k < nops; k++) {
      const uint32_t opc = at(pc, end);
      switch (opc) {
        case OP_CONST: at(pc + 1, end); height = push(height, 1); pc += 2; break;
        case OP_ADD: case OP_SUB: operands(pc + 1, 2, end, height); height = push(height, 1); pc += 3; break;
        case OP_MUL: operands(pc + 1, 2, end, height); at(pc + 3, end); height = push(height, 1); pc += 4; break;
        case OP_INV: case OP_NOT: operands(pc + 1, 1, end, height); at(pc + 2, end); height = push(height, 1); pc += 3; break;
        case OP_ASSERT_EQ: case OP_ASSERT_NE: {
          const uint32_t cnt = at(pc + 1, end);
          operands(pc + 2, cnt, end, height);
          operands(pc + 2 + cnt, cnt, end, height);
          pc += 2 + 2 * (size_t)cnt;
        } break;
        case OP_CONTAINS: {
          const uint32_t cnt = at(pc + 1, end);
          operands(pc + 2, cnt, end, height);
          operands(pc + 2 + cnt, 1, end, height);
          pc += 3 + (size_t)cnt;
        } break;
        case OP_CALL: case OP_PREIMG: {
          const uint32_t callee = at(pc + 1, end), cnt = at(pc + 2, end);
          if (callee >= p.nfuncs) bad("call of an unknown func");
          const FuncHeader h = p.header(callee);
          if (total && h.partial) fail(LSK_ERR_ARG, "assertion failed: ctx.partial (func %u is total but calls partial func %u)", func, callee);
          if (cnt > 64 || cnt != (opc == OP_CALL ? h.in : h.out)) bad("call arity");
          operands(pc + 3, cnt, end, height);
          const uint32_t produced = at(pc + 3 + cnt, end);
          at(pc + 4 + cnt, end);
          if (produced != (opc == OP_CALL ? h.out : h.in)) bad("call result size");
          if (opc == OP_PREIMG && !h.invertible) bad("preimage of a func that is not invertible");
          height = push(height, produced);
          pc += 5 + (size_t)cnt;
        } break;
        case OP_STORE: {
          const uint32_t cnt = at(pc + 1, end);
          if (mem_index_from_len((int)std::min<uint32_t>(cnt, 1000)) < 0) bad("store of an unsupported length");
          operands(pc + 2, cnt, end, height);
          height = push(height, 1);
          pc += 2 + (size_t)cnt;
        } break;
        case OP_LOAD: {
          const uint32_t len = at(pc + 1, end);
          if (mem_index_from_len((int)std::min<uint32_t>(len, 1000)) < 0) bad("load of an unsupported length");
          operands(pc + 2, 1, end, height);
          height = push(height, len);
          pc += 3;
        } break;
        case OP_EXTERN: {
          const uint32_t chip = at(pc + 1, end), cnt = at(pc + 2, end);
          if (chip >= p.nchips) bad("unknown extern chip");
          const uint32_t* c = p.chip(chip);   // kind in out witness requires
          if (cnt != c[1]) bad("extern chip arity");
          operands(pc + 3, cnt, end, height);
          height = push(height, c[2]);
          pc += 3 + (size_t)cnt;
        } break;
        case OP_RANGE_U8: {
          const uint32_t cnt = at(pc + 1, end);
          operands(pc + 2, cnt, end, height);
          pc += 2 + (size_t)cnt;
        } break;
        default: bad("unknown opcode");
      }
      if (pc > end) bad("op runs past its block");
    }
    const uint32_t ctrl = at(pc, end);
    if (ctrl == CTRL_RETURN) {
      at(pc + 1, end);
      const uint32_t cnt = at(pc + 2, end);
      if (cnt != out_size) bad("return arity");
      operands(pc + 3, cnt, end, height);
      pc += 3 + (size_t)cnt;
    } else if (ctrl == CTRL_CHOOSE || ctrl == CTRL_CHOOSE_MANY) {
      const uint32_t nv = at(pc + 1, end);
      if (nv == 0 || nv > 64 || (ctrl == CTRL_CHOOSE && nv != 1)) bad("choose arity");
      operands(pc + 2, nv, end, height);
      const uint32_t nb = at(pc + 2 + nv, end), has_d = at(pc + 3 + nv, end);
      if (has_d > 1) bad("choose default flag");
      size_t q = pc + 4 + nv;
      for (uint32_t i = 0; i < nb; i++) {
        const uint32_t nk = at(q, end);
        if (ctrl == CTRL_CHOOSE_MANY && nk != 1) bad("choose-many branch key count");
        const size_t keys = ctrl == CTRL_CHOOSE ? (size_t)nk : (size_t)nv * nk;
        if (keys > end - q) bad("branch keys run past their block");
        q = block(q + 1 + keys, end, height, depth + 1);
      }
      if (has_d) q = block(q, end, height, depth + 1);
      pc = q;
    } else {
      bad("unknown control node");
    }
    const uint32_t nret = at(pc, end);
    if (nret > end - pc - 1) bad("return ident list runs past its block");
    if (pc + 1 + nret != end) bad("block length does not match its contents");
    return end;
  }
};

void validate_program(const lsk_program& p) {
  const size_t n = p.words.size();
  const size_t table = 3 + (size_t)p.nchips * CHIP_WORDS;
  if (p.nchips > n || p.nfuncs > n || table + p.nfuncs > n) fail(LSK_ERR_ARG, "truncated program image");
  for (uint32_t c = 0; c < p.nchips; c++) {
    const uint32_t* ch = p.chip(c);
    if (ch[1] > 64 || ch[2] > 4096 || ch[3] > (1u << 16) || ch[4] > 4096) fail(LSK_ERR_ARG, "bad program image: chip %u sizes", c);
  }
  for (uint32_t f = 0; f < p.nfuncs; f++) {
    const size_t off = p.func_offset(f);
    if (off < table + p.nfuncs || off + FUNC_HEADER_WORDS + 2 > n) fail(LSK_ERR_ARG, "bad program image: func %u offset", f);
    const FuncHeader h = p.header(f);
    if (h.in > 64 || h.out > 4096 || h.partial > 1 || h.invertible > 1 || h.aux > (1u << 20) || h.sel > (1u << 20) ||
        h.max_map > (1u << 20) || h.in > h.max_map)
      fail(LSK_ERR_ARG, "bad program image: func %u header", f);
  }
  for (uint32_t f = 0; f < p.nfuncs; f++) {
    const FuncHeader h = p.header(f);
    ImageChecker ck{p, p.words.data(), n};
    ck.func = f;
    ck.max_map = h.max_map;
    ck.out_size = h.out;
    ck.total = !h.partial;
    ck.block(p.body(f), n, h.in, 0);
  }
}
}  // namespace

extern "C" {

int lsk_record_finalize(lsk_record* r) {
  LSK_API_BEGIN
  if (!r) fail(LSK_ERR_ARG, "null record");
  lsk::record_unpin(r);
  for (auto& t : r->funcs) {
    if (t.row_ordered) continue;
    for (size_t i = 0; i < t.rows(); i++)
      if (!t.done[i]) fail(LSK_ERR_ARG, "record has an unfinished query");
    repack(t.req_off, t.req_len, t.reqs, 2);
    repack(t.hint_off, t.hint_len, t.hints, 1);
    if (t.partial) repack(t.dreq_off, t.dreq_len, t.dreqs, 2);
    t.row_ordered = true;
  }
  return LSK_OK;
  LSK_API_END
}

int lsk_program_load(const uint32_t* words, size_t n, lsk_program** out) {
  LSK_API_BEGIN
  if (!words || !out || n < 3 || words[0] != PROGRAM_MAGIC) fail(LSK_ERR_ARG, "bad program image");
  auto p = std::make_unique<lsk_program>();
  p->words.assign(words, words + n);
  p->nfuncs = words[1];
  p->nchips = words[2];
  validate_program(*p);
  *out = p.release();
  return LSK_OK;
  LSK_API_END
}
void lsk_program_free(lsk_program* p) { delete p; }

int lsk_poseidon2_hash_host(int width, const uint32_t* in, size_t n, uint32_t* out8) {
  LSK_API_BEGIN
  if ((!in || !out8) && n) fail(LSK_ERR_ARG, "null argument");
  if (width != 24 && width != 32 && width != 40) fail(LSK_ERR_ARG, "unsupported Poseidon2 width %d", width);
  for (size_t i = 0; i < n; i++) {
    uint32_t st[40];
    for (int k = 0; k < width; k++) {
      st[k] = in[i * width + k];
      if (st[k] >= Pq) fail(LSK_ERR_ARG, "non-canonical field element in a hash preimage");
    }
    host_poseidon2(st, width);
    for (int k = 0; k < 8; k++) out8[i * 8 + k] = st[k];
  }
  return LSK_OK;
  LSK_API_END
}

int lsk_poseidon2_constants(int width, uint32_t* diag, uint32_t* full_rc, uint32_t* part_rc, size_t part_cap) {
  LSK_API_BEGIN
  const uint32_t *d, *f, *pr;
  int rp;
  switch (width) {
    case 16: d = P2_DIAG_16; f = P2_FULL_RC_16; pr = P2_PART_RC_16; rp = P2_RP_16; break;
    case 24: d = P2_DIAG_24; f = P2_FULL_RC_24; pr = P2_PART_RC_24; rp = P2_RP_24; break;
    case 32: d = P2_DIAG_32; f = P2_FULL_RC_32; pr = P2_PART_RC_32; rp = P2_RP_32; break;
    case 40: d = P2_DIAG_40; f = P2_FULL_RC_40; pr = P2_PART_RC_40; rp = P2_RP_40; break;
    default: fail(LSK_ERR_ARG, "unsupported Poseidon2 width %d", width);
  }
  if (diag) std::copy(d, d + width, diag);
  if (full_rc) std::copy(f, f + 8 * width, full_rc);
  if (part_rc) std::copy(pr, pr + std::min<size_t>(rp, part_cap), part_rc);
  return rp;
  LSK_API_END
}
int lsk_program_num_funcs(const lsk_program* p) { return p ? (int)p->nfuncs : LSK_ERR_ARG; }
int lsk_program_func_width(const lsk_program* p, uint32_t func) {
  return p && func < p->nfuncs ? (int)p->width(func) : LSK_ERR_ARG;
}

int lsk_record_new(const lsk_program* p, lsk_record** out) {
  LSK_API_BEGIN
  if (!p || !out) fail(LSK_ERR_ARG, "null argument");
  auto* r = new lsk_record;
  r->funcs.resize(p->nfuncs);
  for (uint32_t f = 0; f < p->nfuncs; f++) {
    auto h = p->header(f);
    r->funcs[f].in = h.in;
    r->funcs[f].out = h.out;
    r->funcs[f].invertible = h.invertible != 0;
    r->funcs[f].partial = h.partial != 0;
  }
  for (int i = 0; i < NUM_MEM_TABLES; i++) r->mems[i].len = (uint32_t)mem_table_len(i);
  *out = r;
  return LSK_OK;
  LSK_API_END
}
void lsk_record_free(lsk_record* r) {
  if (!r) return;
  lsk::record_unpin(r);
  delete r;
}

// The executor recurses as deep as the Lair program does (a Lurk `(fib 100000)` nests ~10^6 Lair calls), far beyond a
// default 8 MB thread stack: it runs on its own thread whose stack is a lazily committed reservation (LSK_EXEC_STACK_GB,
// default 64 GB of address space; only the pages actually touched are ever backed).
namespace {
struct ExecJob {
  Executor* ex;
  uint32_t func;
  const uint32_t* args;
  uint32_t nargs;
  std::vector<uint32_t>* out;
  std::exception_ptr err;
};
void* exec_thread(void* arg) {
  ExecJob* j = (ExecJob*)arg;
  try {
    j->ex->run(j->func, j->args, j->nargs, *j->out);
  } catch (...) {
    j->err = std::current_exception();
  }
  return nullptr;
}
void run_on_big_stack(ExecJob& job) {
  size_t gb = 64;
  if (const char* e = getenv("LSK_EXEC_STACK_GB")) gb = (size_t)std::max(1L, atol(e));
  void* stack = MAP_FAILED;
  size_t size = 0;
  for (; gb >= 1 && stack == MAP_FAILED; gb /= 2) {
    size = gb << 30;
    stack = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
  }
  pthread_attr_t attr;
  pthread_t th;
  bool started = false;
  if (stack != MAP_FAILED && pthread_attr_init(&attr) == 0) {
    if (pthread_attr_setstack(&attr, stack, size) == 0 && pthread_create(&th, &attr, exec_thread, &job) == 0) started = true;
    pthread_attr_destroy(&attr);
  }
  if (started) pthread_join(th, nullptr);
  else exec_thread(&job);   // no reservation to be had: the caller's stack will have to do
  if (stack != MAP_FAILED) munmap(stack, size);
  if (job.err) std::rethrow_exception(job.err);
}
}  // namespace

int lsk_execute(const lsk_program* p, lsk_record* r, uint32_t func, const uint32_t* args, size_t nargs, uint32_t* out,
                size_t nout) {
  LSK_API_BEGIN
  if (!p || !r || func >= p->nfuncs) fail(LSK_ERR_ARG, "bad argument");
  for (size_t i = 0; i < nargs; i++)
    if (args[i] >= Pq) fail(LSK_ERR_ARG, "argument not canonical");
  lsk::record_unpin(r);  // the tables are about to grow
  Executor ex(*p, *r);
  std::vector<uint32_t> o;
  ExecJob job{&ex, func, args, (uint32_t)nargs, &o, nullptr};
  run_on_big_stack(job);
  if (o.size() != nout) fail(LSK_ERR_ARG, "output size mismatch");
  for (size_t i = 0; i < nout; i++) out[i] = o[i];
  return LSK_OK;
  LSK_API_END
}


// ---- importing a record produced elsewhere (the reference's own executor, src/lair/execute.rs:375-783) -----------------
// QueryRecord::func_queries[func] in insertion (= nonce) order: args, output, provide, requires; depth and
// depth_requires of partial funcs.  CSR offsets have n + 1 entries, in PAIRS (prev_nonce, prev_count).
int lsk_record_import_func(const lsk_program* p, lsk_record* r, uint32_t func, size_t n, const uint32_t* args,
                           const uint32_t* outs, const uint32_t* provide, const uint32_t* req_off, const uint32_t* reqs,
                           const uint32_t* depth, const uint32_t* dreq_off, const uint32_t* dreqs) {
  LSK_API_BEGIN
  if (!p || !r || func >= p->nfuncs || (n && (!provide || !req_off))) fail(LSK_ERR_ARG, "bad argument");
  FuncTable& t = r->funcs[func];
  if (t.rows()) fail(LSK_ERR_ARG, "func %u already has queries", func);
  if (n > 0xfffffffeull) fail(LSK_ERR_ARG, "too many queries");
  if (n && ((t.in && !args) || (t.out && !outs) || (t.partial && (!depth || !dreq_off)))) fail(LSK_ERR_ARG, "missing table");
  lsk::record_unpin(r);
  auto canonical = [&](const uint32_t* v, size_t k, const char* what) {
    for (size_t i = 0; i < k; i++)
      if (v[i] >= Pq) fail(LSK_ERR_ARG, "non-canonical value in the %s of func %u", what, func);
  };
  if (n) {
    canonical(args, n * t.in, "arguments");
    canonical(outs, n * t.out, "outputs");
    for (size_t i = 0; i < n; i++)
      if (req_off[i] > req_off[i + 1]) fail(LSK_ERR_ARG, "require offsets of func %u are not monotone", func);
    if (req_off[0] != 0) fail(LSK_ERR_ARG, "require offsets must start at 0");
    if (req_off[n] && !reqs) fail(LSK_ERR_ARG, "missing require records");
  }
  t.args.assign(args, args + n * t.in);
  t.outs.assign(outs, outs + n * t.out);
  t.done.assign(n, 1);
  t.provide.assign(provide, provide + 2 * n);
  t.req_off.assign(req_off, req_off + n);
  t.req_len.resize(n);
  for (size_t i = 0; i < n; i++) t.req_len[i] = req_off[i + 1] - req_off[i];
  t.reqs.assign(reqs, reqs + (n ? 2 * (size_t)req_off[n] : 0));
  t.hint_off.assign(n, 0);
  t.hint_len.assign(n, 0);
  t.hints.clear();
  if (t.partial) {
    for (size_t i = 0; i < n; i++)
      if (dreq_off[i] > dreq_off[i + 1]) fail(LSK_ERR_ARG, "depth-require offsets of func %u are not monotone", func);
    t.depth.assign(depth, depth + n);
    t.dreq_off.assign(dreq_off, dreq_off + n);
    t.dreq_len.resize(n);
    for (size_t i = 0; i < n; i++) t.dreq_len[i] = dreq_off[i + 1] - dreq_off[i];
    t.dreqs.assign(dreqs, dreqs + (n ? 2 * (size_t)dreq_off[n] : 0));
  }
  t.row_ordered = true;
  t.completed = n;
  t.index = HashIndex();
  for (size_t i = 0; i < n; i++) {
    if (t.index.find(t.args.data(), t.in, &t.args[i * t.in]) >= 0) fail(LSK_ERR_ARG, "func %u: two queries with the same arguments", func);
    t.index.insert_new(t.args.data(), t.in, (uint32_t)i);
  }
  if (t.invertible)   // execute.rs:690-692: every finished query also feeds the inverse map (a later one wins)
    for (size_t i = 0; i < n; i++) {
      const uint32_t* o = &t.outs[i * t.out];
      long iv = t.inv_index.find(t.inv_outs.data(), t.out, o);
      if (iv >= 0) {
        memcpy(&t.inv_inps[(size_t)iv * t.in], &t.args[i * t.in], t.in * 4);
      } else {
        uint32_t slot = (uint32_t)t.inv_index.used;
        t.inv_outs.insert(t.inv_outs.end(), o, o + t.out);
        t.inv_inps.insert(t.inv_inps.end(), t.args.begin() + i * t.in, t.args.begin() + (i + 1) * t.in);
        t.inv_index.insert_new(t.inv_outs.data(), t.out, slot);
      }
    }
  return LSK_OK;
  LSK_API_END
}
// QueryRecord::mem_queries[mem_idx] in insertion order (pointer = index + 1): values and provide records
int lsk_record_import_mem(lsk_record* r, uint32_t mem_idx, size_t n, const uint32_t* vals, const uint32_t* provide) {
  LSK_API_BEGIN
  if (!r || mem_idx >= NUM_MEM_TABLES || (n && (!vals || !provide))) fail(LSK_ERR_ARG, "bad argument");
  MemTable& mt = r->mems[mem_idx];
  if (mt.rows()) fail(LSK_ERR_ARG, "memory table %u already has entries", mem_idx);
  lsk::record_unpin(r);
  for (size_t i = 0; i < n * mt.len; i++)
    if (vals[i] >= Pq) fail(LSK_ERR_ARG, "non-canonical value in memory table %u", mem_idx);
  mt.vals.assign(vals, vals + n * mt.len);
  mt.provide.assign(provide, provide + 2 * n);
  mt.index = HashIndex();
  for (size_t i = 0; i < n; i++) {
    if (mt.index.find(mt.vals.data(), mt.len, &mt.vals[i * mt.len]) >= 0) fail(LSK_ERR_ARG, "memory table %u: duplicate entry", mem_idx);
    mt.index.insert_new(mt.vals.data(), mt.len, (uint32_t)i);
  }
  return LSK_OK;
  LSK_API_END
}
// BytesRecord (src/gadgets/bytes/record.rs:14-110): for each of the 65536 ByteInputs the provide records
// (nonce, count) of range_u8, range_u16, less_than, and, xor, or -- [65536][6][2] words; NULL / all zero = no byte lookups
int lsk_record_import_bytes(lsk_record* r, const uint32_t* records) {
  LSK_API_BEGIN
  if (!r) fail(LSK_ERR_ARG, "null record");
  lsk::record_unpin(r);
  r->bytes.clear();
  r->bytes_used = false;
  if (records) {
    bool any = false;
    for (size_t i = 0; i < (size_t)65536 * 12 && !any; i++) any = records[i] != 0;
    if (any) {
      r->bytes.assign(records, records + (size_t)65536 * 12);
      r->bytes_used = true;
    }
  }
  return LSK_OK;
  LSK_API_END
}
int lsk_record_import_public_values(lsk_record* r, const uint32_t* pv, size_t n) {
  LSK_API_BEGIN
  if (!r || (n && !pv)) fail(LSK_ERR_ARG, "bad argument");
  for (size_t i = 0; i < n; i++)
    if (pv[i] >= Pq) fail(LSK_ERR_ARG, "non-canonical public value");
  r->public_values.assign(pv, pv + n);
  return LSK_OK;
  LSK_API_END
}
// The hints the trace kernels gather through (record.h): replays every imported row against the imported tables --
// the lookups the reference repeats at trace-generation time (src/lair/trace.rs:325-373) -- and checks on the way that
// each row returns its recorded output.  Must follow the imports (and lsk_inject_inv_queries, for preimages that no
// imported query produced); afterwards the record is what lsk_execute would have built.
int lsk_record_resolve(const lsk_program* p, lsk_record* r) {
  LSK_API_BEGIN
  if (!p || !r) fail(LSK_ERR_ARG, "null argument");
  lsk::record_unpin(r);
  Executor ex(*p, *r);
  ex.replay = true;
  std::vector<uint32_t> out;
  for (uint32_t f = 0; f < p->nfuncs; f++) {
    FuncTable& t = r->funcs[f];
    t.hints.clear();
    for (size_t row = 0; row < t.rows(); row++) {
      std::vector<uint32_t> a(t.args.begin() + row * t.in, t.args.begin() + (row + 1) * t.in);
      ex.push_frame(f, (uint32_t)row, a.data(), t.in, false);
      while (!ex.frames.empty()) ex.step(out);
    }
  }
  return LSK_OK;
  LSK_API_END
}
// ---- exporting (tests; the inverse of the imports) ---------------------------------------------------------------------
// sizes = {rows, require pairs, hints, depth-require pairs}
int lsk_record_func_sizes(const lsk_record* r, uint32_t func, size_t sizes[4]) {
  LSK_API_BEGIN
  if (!r || func >= r->funcs.size() || !sizes) fail(LSK_ERR_ARG, "bad argument");
  const FuncTable& t = r->funcs[func];
  sizes[0] = t.rows(); sizes[1] = t.reqs.size() / 2; sizes[2] = t.hints.size(); sizes[3] = t.dreqs.size() / 2;
  return LSK_OK;
  LSK_API_END
}
// row-ordered tables (call lsk_record_finalize first); offsets get rows + 1 entries; any pointer may be NULL
int lsk_record_export_func(const lsk_record* r, uint32_t func, uint32_t* args, uint32_t* outs, uint32_t* provide,
                           uint32_t* req_off, uint32_t* reqs, uint32_t* hint_off, uint32_t* hints, uint32_t* depth,
                           uint32_t* dreq_off, uint32_t* dreqs) {
  LSK_API_BEGIN
  if (!r || func >= r->funcs.size()) fail(LSK_ERR_ARG, "bad argument");
  const FuncTable& t = r->funcs[func];
  if (!t.row_ordered) fail(LSK_ERR_ARG, "export needs row-ordered tables: call lsk_record_finalize");
  const size_t n = t.rows();
  auto cp = [](uint32_t* dst, const std::vector<uint32_t>& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * 4); };
  cp(args, t.args); cp(outs, t.outs); cp(provide, t.provide); cp(reqs, t.reqs); cp(hints, t.hints);
  if (req_off) { cp(req_off, t.req_off); req_off[n] = (uint32_t)(t.reqs.size() / 2); }
  if (hint_off) { cp(hint_off, t.hint_off); hint_off[n] = (uint32_t)t.hints.size(); }
  if (t.partial) {
    cp(depth, t.depth); cp(dreqs, t.dreqs);
    if (dreq_off) { cp(dreq_off, t.dreq_off); dreq_off[n] = (uint32_t)(t.dreqs.size() / 2); }
  }
  return LSK_OK;
  LSK_API_END
}
int lsk_record_export_mem(const lsk_record* r, uint32_t mem_idx, uint32_t* vals, uint32_t* provide) {
  LSK_API_BEGIN
  if (!r || mem_idx >= NUM_MEM_TABLES) fail(LSK_ERR_ARG, "bad argument");
  const MemTable& mt = r->mems[mem_idx];
  if (vals && !mt.vals.empty()) memcpy(vals, mt.vals.data(), mt.vals.size() * 4);
  if (provide && !mt.provide.empty()) memcpy(provide, mt.provide.data(), mt.provide.size() * 4);
  return LSK_OK;
  LSK_API_END
}
// [65536][6][2] words (all zero when the record has no byte lookups)
int lsk_record_export_bytes(const lsk_record* r, uint32_t* records) {
  LSK_API_BEGIN
  if (!r || !records) fail(LSK_ERR_ARG, "bad argument");
  if (r->bytes.empty()) memset(records, 0, (size_t)65536 * 12 * 4);
  else memcpy(records, r->bytes.data(), (size_t)65536 * 12 * 4);
  return LSK_OK;
  LSK_API_END
}

int lsk_inject_inv_queries(const lsk_program* p, lsk_record* r, uint32_t func, const uint32_t* inps,
                           const uint32_t* outs, size_t n) {
  LSK_API_BEGIN
  if (!p || !r || func >= p->nfuncs) fail(LSK_ERR_ARG, "bad argument");
  FuncTable& t = r->funcs[func];
  if (!t.invertible) fail(LSK_ERR_ARG, "Inverse query map not found");
  for (size_t i = 0; i < n; i++) {
    const uint32_t* o = outs + i * t.out;
    long iv = t.inv_index.find(t.inv_outs.data(), t.out, o);
    if (iv >= 0) {
      memcpy(&t.inv_inps[(size_t)iv * t.in], inps + i * t.in, t.in * 4);
    } else {
      uint32_t slot = (uint32_t)t.inv_index.used;
      t.inv_outs.insert(t.inv_outs.end(), o, o + t.out);
      t.inv_inps.insert(t.inv_inps.end(), inps + i * t.in, inps + (i + 1) * t.in);
      t.inv_index.insert_new(t.inv_outs.data(), t.out, slot);
    }
  }
  return LSK_OK;
  LSK_API_END
}

size_t lsk_record_num_queries(const lsk_record* r, uint32_t func) { return func < r->funcs.size() ? r->funcs[func].rows() : 0; }
size_t lsk_record_num_mem(const lsk_record* r, uint32_t mem_idx) { return mem_idx < NUM_MEM_TABLES ? r->mems[mem_idx].rows() : 0; }
size_t lsk_record_public_values(const lsk_record* r, uint32_t* out, size_t cap) {
  for (size_t i = 0; i < r->public_values.size() && i < cap; i++) out[i] = r->public_values[i];
  return r->public_values.size();
}
uint32_t lsk_record_num_shards(const lsk_record* r, uint32_t shard_size) {
  size_t mx = 0;
  for (auto& t : r->funcs) mx = std::max(mx, t.rows());
  if (!shard_size) return 0;
  return (uint32_t)(mx / shard_size + (mx % shard_size ? 1 : 0));
}
size_t lsk_func_trace_height(const lsk_record* r, uint32_t func, uint32_t shard, uint32_t shard_size) {
  size_t n = r->funcs[func].rows();
  size_t s = (size_t)shard * shard_size, e = std::min<size_t>((size_t)(shard + 1) * shard_size, n);
  size_t real = e > s ? e - s : 0, h = 1;
  while (h < real) h <<= 1;
  return h;
}
size_t lsk_mem_trace_height(const lsk_record* r, uint32_t mem_idx) {
  size_t n = r->mems[mem_idx].rows(), h = 4;
  while (h < n) h <<= 1;
  return h;
}
size_t lsk_record_device_bytes(const lsk_record* r) {
  size_t b = 0;
  for (auto& t : r->funcs)
    b += 4 * (t.args.size() + t.outs.size() + t.provide.size() + t.req_off.size() + t.reqs.size() + t.hint_off.size() +
              t.hints.size() + t.depth.size() + t.dreq_off.size() + t.dreqs.size());
  for (auto& m : r->mems) b += 4 * (m.vals.size() + m.provide.size());
  b += 4 * r->bytes.size();
  return b;
}

}  // extern "C"
