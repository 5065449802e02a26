// ORACLE (test infrastructure, NOT product code).  CPU restatement of the Lair executor and of
// per-chip trace generation, following the reference file by file:
//   program image      decoded from the flat u32 serialisation produced by lurk_b200/lair.py
//                      (mirrors src/lair/bytecode.rs:12-146)
//   QueryRecord &c     src/lair/execute.rs:26-83, src/air/builder.rs:135-169
//   Func::execute      src/lair/execute.rs:436-783  (explicit stack, memoised, nonce = insertion index)
//   FuncChip trace     src/lair/trace.rs:72-135 (rows), :145-418 (populate_row per op)
//   MemChip trace      src/lair/memory.rs:30-68
//   BytesChip traces   src/gadgets/bytes/trace.rs:49-67 (preprocessed), :75-101 (main)
//   Entrypoint trace   src/lair/lair_chip.rs:112-118
//   sharding           src/lair/execute.rs:106-111,186-216
// Pinned by the reference's golden matrices (tests/test_oracle_lair.py cites each one).
//   depth provenance   src/lair/provenance.rs, src/lair/execute.rs:510-512,702-711,751-753,
//                      src/lair/trace.rs:110-121,235-254, src/gadgets/unsigned/less_than.rs:19-41
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "bb.h"
#include "par.h"
#include "poseidon2.h"

namespace orc {

enum OpCode : uint32_t {
  OP_CONST = 1, OP_ADD, OP_SUB, OP_MUL, OP_INV, OP_NOT, OP_ASSERT_EQ, OP_ASSERT_NE, OP_CONTAINS,
  OP_CALL, OP_PREIMG, OP_STORE, OP_LOAD, OP_EXTERN, OP_RANGE_U8,
  CTRL_RETURN = 100, CTRL_CHOOSE = 101, CTRL_CHOOSE_MANY = 102
};

struct Block;
struct OpNode {
  uint32_t code = 0;
  uint32_t a = 0, b = 0;          // scalar operands (var indices, func index, len ...)
  uint32_t has_aux = 0;           // mul / inv / not: whether a witness column is allocated
  Fp cst = Fp::zero();            // const
  std::vector<uint32_t> xs, ys;   // list operands
  uint32_t n_out = 0, callee_partial = 0;
};
struct Branch {
  std::vector<Fp> keys;  // choose: several scalar keys; choose_many: one tuple
  std::unique_ptr<Block> block;
};
struct CtrlNode {
  uint32_t kind = 0;
  uint32_t ident = 0;
  std::vector<uint32_t> vars;
  std::vector<Branch> branches;
  std::unique_ptr<Block> dflt;
};
struct Block {
  std::vector<OpNode> ops;
  CtrlNode ctrl;
  std::vector<uint32_t> return_idents;
};
struct ChipDef { uint32_t kind, input_size, output_size, witness_size, require_size; };
struct FuncDef {
  uint32_t index, input_size, output_size, partial, invertible, aux, sel, max_map;
  Block body;
  uint32_t width() const { return 1 + input_size + output_size + aux + sel; }
};
struct Toplevel {
  std::vector<ChipDef> chips;
  std::vector<FuncDef> funcs;
};

static inline void parse_block(const uint32_t* w, size_t& p, Block& b) {
  size_t start = p;
  uint32_t len = w[p++];
  uint32_t nops = w[p++];
  b.ops.resize(nops);
  for (auto& op : b.ops) {
    op.code = w[p++];
    switch (op.code) {
      case OP_CONST: op.cst = Fp::from_u32(w[p++]); break;
      case OP_ADD: case OP_SUB: op.a = w[p++]; op.b = w[p++]; break;
      case OP_MUL: op.a = w[p++]; op.b = w[p++]; op.has_aux = w[p++]; break;
      case OP_INV: case OP_NOT: op.a = w[p++]; op.has_aux = w[p++]; break;
      case OP_ASSERT_EQ: case OP_ASSERT_NE: {
        uint32_t n = w[p++];
        op.xs.assign(w + p, w + p + n); p += n;
        op.ys.assign(w + p, w + p + n); p += n;
      } break;
      case OP_CONTAINS: {
        uint32_t n = w[p++];
        op.xs.assign(w + p, w + p + n); p += n;
        op.b = w[p++];
      } break;
      case OP_CALL: case OP_PREIMG: {
        op.a = w[p++];
        uint32_t n = w[p++];
        op.xs.assign(w + p, w + p + n); p += n;
        op.n_out = w[p++]; op.callee_partial = w[p++];
      } break;
      case OP_STORE: case OP_RANGE_U8: {
        uint32_t n = w[p++];
        op.xs.assign(w + p, w + p + n); p += n;
      } break;
      case OP_LOAD: op.a = w[p++]; op.b = w[p++]; break;
      case OP_EXTERN: {
        op.a = w[p++];
        uint32_t n = w[p++];
        op.xs.assign(w + p, w + p + n); p += n;
      } break;
      default: throw std::runtime_error("bad opcode");
    }
  }
  CtrlNode& c = b.ctrl;
  c.kind = w[p++];
  if (c.kind == CTRL_RETURN) {
    c.ident = w[p++];
    uint32_t n = w[p++];
    c.vars.assign(w + p, w + p + n); p += n;
  } else if (c.kind == CTRL_CHOOSE || c.kind == CTRL_CHOOSE_MANY) {
    uint32_t nv = w[p++];
    c.vars.assign(w + p, w + p + nv); p += nv;
    uint32_t nb = w[p++], has_d = w[p++];
    c.branches.resize(nb);
    for (auto& br : c.branches) {
      uint32_t nk = w[p++];
      uint32_t words = c.kind == CTRL_CHOOSE ? nk : nk * nv;
      for (uint32_t i = 0; i < words; i++) br.keys.push_back(Fp::from_u32(w[p++]));
      br.block.reset(new Block);
      parse_block(w, p, *br.block);
    }
    if (has_d) {
      c.dflt.reset(new Block);
      parse_block(w, p, *c.dflt);
    }
  } else {
    throw std::runtime_error("bad ctrl");
  }
  uint32_t nr = w[p++];
  b.return_idents.assign(w + p, w + p + nr); p += nr;
  if (p - start != len) throw std::runtime_error("block length mismatch");
}

static inline Toplevel* parse_toplevel(const uint32_t* w, size_t n) {
  if (n < 3 || w[0] != 0x4c414952u) throw std::runtime_error("bad program magic");
  auto* t = new Toplevel;
  uint32_t nf = w[1], nc = w[2];
  size_t p = 3;
  for (uint32_t i = 0; i < nc; i++, p += 5) t->chips.push_back({w[p], w[p + 1], w[p + 2], w[p + 3], w[p + 4]});
  const uint32_t* table = w + p;
  t->funcs.resize(nf);
  for (uint32_t i = 0; i < nf; i++) {
    size_t q = table[i];
    FuncDef& f = t->funcs[i];
    f.index = i;
    f.input_size = w[q++]; f.output_size = w[q++]; f.partial = w[q++]; f.invertible = w[q++];
    f.aux = w[q++]; f.sel = w[q++]; f.max_map = w[q++];
    parse_block(w, q, f.body);
  }
  return t;
}

// ---- records ------------------------------------------------------------------------------
struct Record {  // src/air/builder.rs:135-148
  uint32_t nonce = 0, count = 0;
  Record new_lookup(uint32_t n) {
    Record r = *this;
    nonce = n;
    count += 1;
    return r;
  }
};
struct QueryResult {  // src/lair/execute.rs:26-33
  bool has_output = false;
  std::vector<Fp> output;
  Record provide;
  std::vector<Record> requires_;
  uint32_t depth = 0;
  std::vector<Record> depth_requires;
};
struct VecHash {
  size_t operator()(const std::vector<uint32_t>& v) const {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint32_t x : v) h = (h ^ x) * 0x100000001b3ull;
    return (size_t)(h ^ (h >> 29));
  }
};
// insertion-ordered map (indexmap) keyed by the Montgomery words of the argument list
struct QueryMap {
  std::vector<std::vector<Fp>> keys;
  std::vector<QueryResult> vals;
  std::unordered_map<std::vector<uint32_t>, uint32_t, VecHash> index;
  static std::vector<uint32_t> raw(const std::vector<Fp>& k) {
    std::vector<uint32_t> r(k.size());
    for (size_t i = 0; i < k.size(); i++) r[i] = k[i].v;
    return r;
  }
  long find(const std::vector<Fp>& k) const {
    auto it = index.find(raw(k));
    return it == index.end() ? -1 : (long)it->second;
  }
  uint32_t insert(const std::vector<Fp>& k, const QueryResult& r) {
    uint32_t i = (uint32_t)keys.size();
    index.emplace(raw(k), i);
    keys.push_back(k);
    vals.push_back(r);
    return i;
  }
  size_t size() const { return keys.size(); }
};
struct BytesInputRecord { Record rec[6]; };  // range_u8, range_u16, less_than, and, xor, or
struct QueryRecord {  // src/lair/execute.rs:67-75
  std::vector<Fp> public_values;
  std::vector<QueryMap> func_queries;
  std::vector<std::unordered_map<std::vector<uint32_t>, std::vector<Fp>, VecHash>> inv_func_queries;
  QueryMap mem_queries[6];
  std::map<uint16_t, BytesInputRecord> bytes;  // BTreeMap<ByteInput, ..>
  explicit QueryRecord(const Toplevel& t) : func_queries(t.funcs.size()), inv_func_queries(t.funcs.size()) {}
};

static const int MEM_TABLE_SIZES[6] = {2, 3, 4, 5, 6, 8};
static inline int mem_index_from_len(size_t len) {
  for (int i = 0; i < 6; i++)
    if ((size_t)MEM_TABLE_SIZES[i] == len) return i;
  throw std::runtime_error("There are no mem tables of that size");
}
static inline int chip_width(uint32_t kind) { return kind == 1 ? 24 : kind == 2 ? 32 : kind == 3 ? 40 : 0; }

// ---- executor (src/lair/execute.rs:436-783) -----------------------------------------------
struct ExecEntry {
  const OpNode* op;
  const CtrlNode* ctrl;
};
struct CallerState {
  bool preimg;
  uint32_t func_index, nonce;
  std::vector<Fp> map;
  std::vector<Record> requires_;
  bool partial;
  std::vector<uint32_t> depths;
};

static inline std::vector<Fp> execute(const Toplevel& top, uint32_t entry, const std::vector<Fp>& args, QueryRecord& q) {
  uint32_t func_index = entry;
  if (args.size() != top.funcs[entry].input_size) throw std::runtime_error("Argument mismatch");
  QueryResult first;
  first.provide.count = 1;  // execute.rs:444-447
  if (q.func_queries[func_index].find(args) >= 0) throw std::runtime_error("entry query already present");
  uint32_t nonce = q.func_queries[func_index].insert(args, first);
  std::vector<Fp> map = args;
  std::vector<Record> requires_;
  bool partial = top.funcs[entry].partial;
  std::vector<uint32_t> depths;
  std::vector<ExecEntry> stack;
  std::vector<CallerState> callers;
  auto push_block = [&](const Block& b) {
    stack.push_back({nullptr, &b.ctrl});
    for (size_t i = b.ops.size(); i-- > 0;) stack.push_back({&b.ops[i], nullptr});
  };
  auto enter = [&](bool preimg, uint32_t callee, std::vector<Fp> inp) {
    uint32_t callee_nonce = q.func_queries[callee].insert(inp, QueryResult());
    callers.push_back({preimg, func_index, nonce, std::move(map), std::move(requires_), partial, std::move(depths)});
    map = std::move(inp);
    requires_.clear();
    depths.clear();
    func_index = callee;
    nonce = callee_nonce;
    partial = top.funcs[callee].partial;
    push_block(top.funcs[callee].body);
  };
  push_block(top.funcs[entry].body);
  while (!stack.empty()) {
    ExecEntry e = stack.back();
    stack.pop_back();
    if (e.op) {
      const OpNode& op = *e.op;
      switch (op.code) {
        case OP_ASSERT_EQ:
          for (size_t i = 0; i < op.xs.size(); i++)
            if (map[op.xs[i]] != map[op.ys[i]]) throw std::runtime_error("AssertEq failed");
          break;
        case OP_ASSERT_NE: {
          bool ne = false;
          for (size_t i = 0; i < op.xs.size(); i++) ne |= map[op.xs[i]] != map[op.ys[i]];
          if (!ne) throw std::runtime_error("AssertNe failed");
        } break;
        case OP_CONTAINS: {
          bool found = false;
          for (uint32_t x : op.xs) found |= map[x] == map[op.b];
          if (!found) throw std::runtime_error("Contains failed");
        } break;
        case OP_CALL: {
          std::vector<Fp> inp;
          for (uint32_t x : op.xs) inp.push_back(map[x]);
          long qi = q.func_queries[op.a].find(inp);
          if (qi >= 0) {
            QueryResult& r = q.func_queries[op.a].vals[qi];
            if (!r.has_output) throw std::runtime_error("Loop detected");
            map.insert(map.end(), r.output.begin(), r.output.end());
            requires_.push_back(r.provide.new_lookup(nonce));
            if (partial && top.funcs[op.a].partial) depths.push_back(r.depth);
          } else {
            enter(false, op.a, std::move(inp));
          }
        } break;
        case OP_PREIMG: {
          std::vector<Fp> out;
          for (uint32_t x : op.xs) out.push_back(map[x]);
          auto it = q.inv_func_queries[op.a].find(QueryMap::raw(out));
          if (it == q.inv_func_queries[op.a].end()) throw std::runtime_error("Preimg not found");
          std::vector<Fp> inp = it->second;
          long qi = q.func_queries[op.a].find(inp);
          if (qi >= 0) {
            QueryResult& r = q.func_queries[op.a].vals[qi];
            if (!r.has_output) throw std::runtime_error("Loop detected");
            map.insert(map.end(), inp.begin(), inp.end());
            requires_.push_back(r.provide.new_lookup(nonce));
            if (partial && top.funcs[op.a].partial) depths.push_back(r.depth);
          } else {
            enter(true, op.a, std::move(inp));
          }
        } break;
        case OP_CONST: map.push_back(op.cst); break;
        case OP_ADD: map.push_back(map[op.a] + map[op.b]); break;
        case OP_SUB: map.push_back(map[op.a] - map[op.b]); break;
        case OP_MUL: map.push_back(map[op.a] * map[op.b]); break;
        case OP_INV:
          if (map[op.a].is_zero()) throw std::runtime_error("inverse of zero");
          map.push_back(fp_inv(map[op.a]));
          break;
        case OP_NOT: map.push_back(map[op.a].is_zero() ? Fp::one() : Fp::zero()); break;
        case OP_STORE: {
          std::vector<Fp> vals;
          for (uint32_t x : op.xs) vals.push_back(map[x]);
          QueryMap& mm = q.mem_queries[mem_index_from_len(vals.size())];
          long i = mm.find(vals);
          if (i < 0) i = mm.insert(vals, QueryResult());
          map.push_back(Fp::from_u32((uint32_t)i + 1));
          requires_.push_back(mm.vals[i].provide.new_lookup(nonce));
        } break;
        case OP_LOAD: {
          uint32_t ptr = map[op.b].to_u32();
          QueryMap& mm = q.mem_queries[mem_index_from_len(op.a)];
          if (ptr == 0 || ptr > mm.size()) throw std::runtime_error("Unbound pointer");
          const auto& vals = mm.keys[ptr - 1];
          map.insert(map.end(), vals.begin(), vals.end());
          requires_.push_back(mm.vals[ptr - 1].provide.new_lookup(nonce));
        } break;
        case OP_EXTERN: {
          const ChipDef& c = top.chips[op.a];
          if (c.kind == 10) {  // BigNum::LessThan::execute, src/core/big_num.rs:42-60 + big_num/cmp.rs:25-45
            uint32_t l = 0, r = 0;
            for (int i = 7; i >= 0; i--)
              if (map[op.xs[i]] != map[op.xs[8 + i]]) {
                l = map[op.xs[i]].to_u32();
                r = map[op.xs[8 + i]].to_u32();
                break;
              }
            auto lookup = [&](uint32_t i1, uint32_t i2, int slot) {
              requires_.push_back(q.bytes[(uint16_t)(i1 | (i2 << 8))].rec[slot].new_lookup(nonce));
            };
            for (uint32_t v : {l, r}) {  // FieldToWord32::populate, unsigned/field.rs:20-34,113-118
              lookup(v >> 24, 0x78, 2);
              lookup(v & 255u, (v >> 8) & 255u, 0);
              lookup((v >> 16) & 255u, v >> 24, 0);
            }
            int i = 3;
            while (i >= 0 && ((l >> (8 * i)) & 255u) == ((r >> (8 * i)) & 255u)) i--;
            lookup(i >= 0 ? (l >> (8 * i)) & 255u : 0, i >= 0 ? (r >> (8 * i)) & 255u : 0, 2);
            map.push_back(Fp::from_u32(l < r));
            break;
          }
          if (c.kind >= 4 && c.kind <= 9) {  // U64::execute, src/core/u64.rs:76-128
            const size_t nin = c.kind == 7 ? 8 : 16;
            uint64_t x = 0, y = 0;
            for (size_t i = 0; i < nin; i++) {
              uint64_t v = map[op.xs[i]].to_u32();
              if (v > 255) throw std::runtime_error("u64 chip input is not a byte");
              if (i < 8) x |= v << (8 * i);
              else y |= v << (8 * (i - 8));
            }
            auto byte = [](uint64_t v, int i) { return (uint32_t)(v >> (8 * i)) & 255u; };
            auto lookup = [&](uint32_t i1, uint32_t i2, int slot) {
              requires_.push_back(q.bytes[(uint16_t)(i1 | (i2 << 8))].rec[slot].new_lookup(nonce));
            };
            auto range_pairs = [&](uint64_t v) {  // UncheckedWord::assign_bytes, unsigned/mod.rs:114-120
              for (int i = 0; i < 8; i += 2) lookup(byte(v, i), byte(v, i + 1), 0);
            };
            auto push_word = [&](uint64_t v) {
              for (int i = 0; i < 8; i++) map.push_back(Fp::from_u32(byte(v, i)));
            };
            auto compare = [&](uint64_t l, uint64_t r) {  // cmp.rs:23-46
              int i = 7;
              while (i >= 0 && byte(l, i) == byte(r, i)) i--;
              lookup(i >= 0 ? byte(l, i) : 0, i >= 0 ? byte(r, i) : 0, 2);
            };
            auto product = [&](uint64_t l, uint64_t r) {  // mul.rs:27-63,148-155
              uint32_t prod[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              for (int i = 0; i < 8; i++)
                for (int j = 0; i + j < 8; j++) prod[i + j] += byte(l, i) * byte(r, j);
              uint32_t carry = 0;
              for (int k = 0; k < 8; k++) {
                carry = (prod[k] + carry) >> 8;
                lookup(carry & 255u, carry >> 8, 1);  // range_check_u16
              }
              range_pairs(l * r);
              return l * r;
            };
            switch (c.kind) {
              case 4: case 5: {
                uint64_t z = c.kind == 4 ? x + y : x - y;
                range_pairs(z);
                push_word(z);
              } break;
              case 6: compare(x, y); map.push_back(Fp::from_u32(x < y)); break;
              case 7: map.push_back(Fp::from_u32(x == 0)); break;
              case 8: push_word(product(x, y)); break;
              default: {  // div_rem.rs:35-59
                if (y == 0) throw std::runtime_error("expected input to be non-zero");
                uint64_t qq = x / y;
                range_pairs(qq);
                uint64_t qb = product(qq, y);
                uint64_t r = x - qb;
                range_pairs(r);
                int i = 7;
                while (byte(r, i) == byte(y, i)) i--;
                lookup(byte(r, i), byte(y, i), 2);
                compare(qb, x);
                push_word(qq);
                push_word(r);
              }
            }
            break;
          }
          int w = chip_width(c.kind);
          if (!w) throw std::runtime_error("unknown extern chip kind");
          Fp st[40];
          for (int i = 0; i < w; i++) st[i] = map[op.xs[i]];
          p2_permute(st, w);  // PoseidonChipset::execute_simple, src/core/poseidon.rs:61-63
          for (int i = 0; i < 8; i++) map.push_back(st[i]);
        } break;
        case OP_RANGE_U8: {  // execute.rs:673-682 + bytes/mod.rs:97-103 + bytes/record.rs:112-117
          for (size_t i = 0; i < op.xs.size(); i += 2) {
            uint32_t i1 = map[op.xs[i]].to_u32();
            uint32_t i2 = i + 1 < op.xs.size() ? map[op.xs[i + 1]].to_u32() : 0;
            if (i1 > 255 || i2 > 255) throw std::runtime_error("Variable not in u8 range");
            uint16_t key = (uint16_t)(i1 | (i2 << 8));
            requires_.push_back(q.bytes[key].rec[0].new_lookup(nonce));
          }
        } break;
        default: throw std::runtime_error("bad op");
      }
    } else {
      const CtrlNode& c = *e.ctrl;
      if (c.kind == CTRL_RETURN) {
        std::vector<Fp> out;
        for (uint32_t v : c.vars) out.push_back(map[v]);
        QueryMap& qm = q.func_queries[func_index];
        QueryResult& r = qm.vals[nonce];
        if (r.has_output) throw std::runtime_error("output already set");
        std::vector<Fp> inp = qm.keys[nonce];
        if (top.funcs[func_index].invertible) q.inv_func_queries[func_index][QueryMap::raw(out)] = inp;
        if (partial) {  // execute.rs:702-711
          uint32_t depth = 0;
          for (uint32_t d : depths) depth = std::max(depth, d + 1);
          auto byte_of = [](uint32_t x, int i) { return (x >> (8 * i)) & 255u; };
          for (int i = 0; i < 4; i += 2)  // range_check_u8_iter(depth.to_le_bytes()), bytes/mod.rs:97-103
            r.depth_requires.push_back(
                q.bytes[(uint16_t)(byte_of(depth, i) | (byte_of(depth, i + 1) << 8))].rec[0].new_lookup(nonce));
          for (uint32_t dep : depths) {  // LessThanWitness::populate, less_than.rs:19-41
            if (!(dep < depth)) throw std::runtime_error("depth comparison failed");
            for (int i = 3; i >= 0; i--)
              if (byte_of(dep, i) != byte_of(depth, i)) {
                r.depth_requires.push_back(
                    q.bytes[(uint16_t)(byte_of(dep, i) | (byte_of(depth, i) << 8))].rec[2].new_lookup(nonce));
                break;
              }
          }
          r.depth = depth;
        }
        r.has_output = true;
        r.output = out;
        r.requires_ = std::move(requires_);
        if (!callers.empty()) {
          bool callee_partial = partial;
          CallerState cs = std::move(callers.back());
          callers.pop_back();
          func_index = cs.func_index;
          nonce = cs.nonce;
          map = std::move(cs.map);
          requires_ = std::move(cs.requires_);
          partial = cs.partial;
          depths = std::move(cs.depths);
          const auto& ext = cs.preimg ? inp : out;
          map.insert(map.end(), ext.begin(), ext.end());
          requires_.push_back(r.provide.new_lookup(nonce));
          if (partial && callee_partial) depths.push_back(r.depth);
        } else {
          if (!stack.empty()) throw std::runtime_error("stack not empty at exit");
          map = out;
          break;
        }
      } else {
        const Block* chosen = nullptr;
        size_t nv = c.vars.size();
        for (const auto& br : c.branches) {
          if (c.kind == CTRL_CHOOSE) {
            for (Fp k : br.keys)
              if (k == map[c.vars[0]]) chosen = br.block.get();
          } else {
            bool all = true;
            for (size_t i = 0; i < nv; i++) all &= br.keys[i] == map[c.vars[i]];
            if (all) chosen = br.block.get();
          }
          if (chosen) break;
        }
        if (!chosen) chosen = c.dflt.get();
        if (!chosen) throw std::runtime_error("No match");
        push_block(*chosen);
      }
    }
  }
  // Toplevel::execute, execute.rs:376-392
  q.public_values = args;
  q.public_values.insert(q.public_values.end(), map.begin(), map.end());
  if (top.funcs[entry].partial) {
    uint32_t depth = 0;
    for (uint32_t d : depths) depth = std::max(depth, d + 1);
    for (int i = 0; i < 4; i++) q.public_values.push_back(Fp::from_u32((depth >> (8 * i)) & 255u));
  }
  return map;
}

// ---- trace generation ----------------------------------------------------------------------
static inline size_t next_pow2(size_t n) {
  size_t p = 1;
  while (p < n) p <<= 1;
  return p;
}
struct Range { size_t start, end; size_t len() const { return end > start ? end - start : 0; } };
static inline Range shard_func_range(const QueryRecord& q, uint32_t fi, uint32_t shard, uint32_t shard_size) {
  size_t n = q.func_queries[fi].size();
  size_t s = (size_t)shard * shard_size, e = std::min<size_t>((size_t)(shard + 1) * shard_size, n);
  return {s, e};
}
static inline uint32_t num_shards(const QueryRecord& q, uint32_t shard_size) {  // execute.rs:186-216
  size_t mx = 0;
  for (auto& m : q.func_queries) mx = std::max(mx, m.size());
  return (uint32_t)(mx / shard_size + (mx % shard_size ? 1 : 0));
}

struct RowCtx {
  Fp* row;
  size_t aux_at, width;
  const std::vector<Record>* requires_;
  size_t req_at = 0;
  const std::vector<Record>* depth_requires = nullptr;
  size_t dreq_at = 0;
  uint32_t depth = 0;
  void push_aux(Fp f) { row[aux_at++] = f; }
  void push_record(Record r) {  // Record::into_require, builder.rs:159-168
    push_aux(Fp::from_u32(r.nonce));
    push_aux(Fp::from_u32(r.count));
    push_aux(fp_inv(Fp::from_u32(r.count) + Fp::one()));
  }
  void push_require() {
    if (req_at >= requires_->size()) throw std::runtime_error("Not enough require hints");
    push_record((*requires_)[req_at++]);
  }
  void push_depth_require() {
    if (!depth_requires || dreq_at >= depth_requires->size()) throw std::runtime_error("Not enough require hints");
    push_record((*depth_requires)[dreq_at++]);
  }
  void push_depth(uint32_t dep) {  // trace.rs:235-254
    for (int i = 0; i < 4; i++) push_aux(Fp::from_u32((dep >> (8 * i)) & 255u));
    Fp wit[6] = {Fp::zero(), Fp::zero(), Fp::zero(), Fp::zero(), Fp::zero(), Fp::zero()};
    if (!(dep < depth)) throw std::runtime_error("depth comparison failed");
    for (int i = 3; i >= 0; i--) {
      uint32_t l = (dep >> (8 * i)) & 255u, r = (depth >> (8 * i)) & 255u;
      if (l != r) {
        wit[i] = Fp::one();
        wit[4] = Fp::from_u32(l);
        wit[5] = Fp::from_u32(r);
        break;
      }
    }
    for (Fp x : wit) push_aux(x);
    push_depth_require();
  }
};

static inline void populate_block(const Toplevel& top, const QueryRecord& q, const FuncDef& f, const Block& b,
                                  std::vector<Fp>& map, RowCtx& cx) {
  for (const OpNode& op : b.ops) {
    switch (op.code) {
      case OP_ASSERT_EQ: break;
      case OP_ASSERT_NE: {  // push_inequality_witness, trace.rs:218-233
        bool found = false;
        for (size_t i = 0; i < op.xs.size(); i++) {
          Fp d = map[op.xs[i]] - map[op.ys[i]];
          if (!found && !d.is_zero()) {
            cx.push_aux(fp_inv(d));
            found = true;
          } else {
            cx.push_aux(Fp::zero());
          }
        }
        if (!found) throw std::runtime_error("AssertNe witness: all equal");
      } break;
      case OP_CONTAINS: {
        Fp bv = map[op.b];
        Fp acc = map[op.xs[0]] - bv;
        for (size_t i = 1; i < op.xs.size(); i++) {
          acc = acc * (map[op.xs[i]] - bv);
          cx.push_aux(acc);
        }
      } break;
      case OP_CONST: map.push_back(op.cst); break;
      case OP_ADD: map.push_back(map[op.a] + map[op.b]); break;
      case OP_SUB: map.push_back(map[op.a] - map[op.b]); break;
      case OP_MUL: {
        Fp v = map[op.a] * map[op.b];
        map.push_back(v);
        if (op.has_aux) cx.push_aux(v);
      } break;
      case OP_INV: {
        Fp v = fp_inv(map[op.a]);
        map.push_back(v);
        if (op.has_aux) cx.push_aux(v);
      } break;
      case OP_NOT: {
        Fp a = map[op.a];
        Fp d = a.is_zero() ? Fp::zero() : fp_inv(a);
        Fp x = a.is_zero() ? Fp::one() : Fp::zero();
        map.push_back(x);
        if (op.has_aux) { cx.push_aux(d); cx.push_aux(x); }
      } break;
      case OP_CALL: {
        std::vector<Fp> inp;
        for (uint32_t x : op.xs) inp.push_back(map[x]);
        long qi = q.func_queries[op.a].find(inp);
        if (qi < 0) throw std::runtime_error("Cannot find query result");
        for (Fp o : q.func_queries[op.a].vals[qi].output) { map.push_back(o); cx.push_aux(o); }
        cx.push_require();
        if (top.funcs[op.a].partial) cx.push_depth(q.func_queries[op.a].vals[qi].depth);
      } break;
      case OP_PREIMG: {
        std::vector<Fp> out;
        for (uint32_t x : op.xs) out.push_back(map[x]);
        auto it = q.inv_func_queries[op.a].find(QueryMap::raw(out));
        if (it == q.inv_func_queries[op.a].end()) throw std::runtime_error("Cannot find preimage");
        for (Fp i : it->second) { map.push_back(i); cx.push_aux(i); }
        cx.push_require();
        if (top.funcs[op.a].partial) {
          long qi = q.func_queries[op.a].find(it->second);
          if (qi < 0) throw std::runtime_error("Cannot find query result");
          cx.push_depth(q.func_queries[op.a].vals[qi].depth);
        }
      } break;
      case OP_STORE: {
        std::vector<Fp> vals;
        for (uint32_t x : op.xs) vals.push_back(map[x]);
        long i = q.mem_queries[mem_index_from_len(vals.size())].find(vals);
        if (i < 0) throw std::runtime_error("Cannot find query result");
        Fp ptr = Fp::from_u32((uint32_t)i + 1);
        map.push_back(ptr);
        cx.push_aux(ptr);
        cx.push_require();
      } break;
      case OP_LOAD: {
        const QueryMap& mm = q.mem_queries[mem_index_from_len(op.a)];
        uint32_t ptr = map[op.b].to_u32();
        for (Fp v : mm.keys[ptr - 1]) { map.push_back(v); cx.push_aux(v); }
        cx.push_require();
      } break;
      case OP_EXTERN: {  // trace.rs:388-407 + core/poseidon.rs:65-72 / core/u64.rs:130-168
        const ChipDef& c = top.chips[op.a];
        if (c.kind == 10) {  // BigNumCompareWitness (big_num/cmp.rs:12-45): is_comp[8], limbs, 2 x FieldToWord32, CompareWitness<4>
          Fp wit[28];
          for (auto& t : wit) t = Fp::zero();
          uint32_t l = 0, r = 0;
          for (int i = 7; i >= 0; i--)
            if (map[op.xs[i]] != map[op.xs[8 + i]]) {
              wit[i] = Fp::one();
              wit[8] = map[op.xs[i]];
              wit[9] = map[op.xs[8 + i]];
              l = wit[8].to_u32();
              r = wit[9].to_u32();
              break;
            }
          uint32_t vv[2] = {l, r};
          for (int k = 0; k < 2; k++) {
            wit[10 + 5 * k] = Fp::from_u32((vv[k] >> 24) < 0x78);
            for (int j = 0; j < 4; j++) wit[11 + 5 * k + j] = Fp::from_u32((vv[k] >> (8 * j)) & 255u);
          }
          int i = 3;
          while (i >= 0 && ((l >> (8 * i)) & 255u) == ((r >> (8 * i)) & 255u)) i--;
          if (i >= 0) {
            uint32_t a = (l >> (8 * i)) & 255u, b = (r >> (8 * i)) & 255u;
            wit[20 + i] = Fp::one();
            wit[24] = Fp::from_u32(a);
            wit[25] = Fp::from_u32(b);
            wit[26] = fp_inv(wit[24] - wit[25]);
            wit[27] = Fp::from_u32(a < b);
          }
          map.push_back(wit[27]);
          for (Fp t : wit) cx.push_aux(t);
          for (uint32_t k = 0; k < c.require_size; k++) cx.push_require();
          break;
        }
        if (c.kind >= 4 && c.kind <= 9) {  // U64::populate_witness, src/core/u64.rs:130-168
          const size_t nin = c.kind == 7 ? 8 : 16;
          uint64_t x = 0, y = 0;
          for (size_t i = 0; i < nin; i++) {
            uint64_t v = map[op.xs[i]].to_u32();
            if (i < 8) x |= v << (8 * i);
            else y |= v << (8 * (i - 8));
          }
          auto byte = [](uint64_t v, int i) { return (uint32_t)(v >> (8 * i)) & 255u; };
          auto aux_word = [&](uint64_t v) {
            for (int i = 0; i < 8; i++) cx.push_aux(Fp::from_u32(byte(v, i)));
          };
          auto map_word = [&](uint64_t v) {
            for (int i = 0; i < 8; i++) map.push_back(Fp::from_u32(byte(v, i)));
          };
          auto non_zero_witness = [&](uint64_t v) {  // IsZeroWitness::populate_non_zero, is_zero.rs:17-28
            bool found = false;
            for (int i = 0; i < 8; i++) {
              if (!found && byte(v, i)) {
                cx.push_aux(fp_inv(Fp::from_u32(byte(v, i))));
                found = true;
              } else {
                cx.push_aux(Fp::zero());
              }
            }
          };
          auto compare_witness = [&](uint64_t l, uint64_t r, bool full) {  // CompareWitness / LessThanWitness
            Fp wit[12];
            for (auto& t : wit) t = Fp::zero();
            int i = 7;
            while (i >= 0 && byte(l, i) == byte(r, i)) i--;
            if (i >= 0) {
              wit[i] = Fp::one();
              wit[8] = Fp::from_u32(byte(l, i));
              wit[9] = Fp::from_u32(byte(r, i));
              wit[10] = fp_inv(wit[8] - wit[9]);
              wit[11] = Fp::from_u32(byte(l, i) < byte(r, i));
            }
            for (int k = 0; k < (full ? 12 : 10); k++) cx.push_aux(wit[k]);
          };
          auto product_witness = [&](uint64_t l, uint64_t r) {  // MulWitness carries, then the result
            uint32_t prod[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 8; i++)
              for (int j = 0; i + j < 8; j++) prod[i + j] += byte(l, i) * byte(r, j);
            uint32_t carry = 0;
            for (int k = 0; k < 8; k++) {
              carry = (prod[k] + carry) >> 8;
              cx.push_aux(Fp::from_u32(carry));
            }
            aux_word(l * r);
            return l * r;
          };
          switch (c.kind) {
            case 4: case 5: {
              uint64_t z = c.kind == 4 ? x + y : x - y;
              map_word(z);
              aux_word(z);
            } break;
            case 6:
              map.push_back(Fp::from_u32(x < y));
              compare_witness(x, y, true);
              break;
            case 7:
              map.push_back(Fp::from_u32(x == 0));
              if (x) non_zero_witness(x);
              else
                for (int i = 0; i < 8; i++) cx.push_aux(Fp::zero());
              cx.push_aux(Fp::from_u32(x == 0));
              break;
            case 8: map_word(product_witness(x, y)); break;
            default: {  // DivRem: b_non_zero, q, qb (Product), r (Diff), r_lt_b (LessThanWitness), qb_cmp_a (CompareWitness)
              uint64_t qq = x / y;
              non_zero_witness(y);
              aux_word(qq);
              uint64_t qb = product_witness(qq, y);
              uint64_t r = x - qb;
              aux_word(r);
              compare_witness(r, y, false);
              compare_witness(qb, x, true);
              map_word(qq);
              map_word(r);
            }
          }
          for (uint32_t i = 0; i < c.require_size; i++) cx.push_require();
          break;
        }
        int w = chip_width(c.kind);
        Fp st[40];
        for (int i = 0; i < w; i++) st[i] = map[op.xs[i]];
        std::vector<Fp> wit(p2_witness_cols(w));
        p2_permute(st, w, wit.data());
        for (int i = 0; i < 8; i++) map.push_back(st[i]);
        for (int i = 0; i < 8; i++) cx.push_aux(st[i]);
        for (Fp x : wit) cx.push_aux(x);
      } break;
      case OP_RANGE_U8:
        for (size_t i = 0; i < (op.xs.size() + 1) / 2; i++) cx.push_require();
        break;
      default: throw std::runtime_error("bad op");
    }
  }
  const CtrlNode& c = b.ctrl;
  if (c.kind == CTRL_RETURN) {
    if (cx.req_at != cx.requires_->size()) throw std::runtime_error("unused require hints");
    if (cx.depth_requires && cx.dreq_at != cx.depth_requires->size()) throw std::runtime_error("unused depth require hints");
    cx.row[cx.width - f.sel + c.ident] = Fp::one();
    return;
  }
  const Block* chosen = nullptr;
  for (const auto& br : c.branches) {
    if (c.kind == CTRL_CHOOSE) {
      for (Fp k : br.keys)
        if (k == map[c.vars[0]]) chosen = br.block.get();
    } else {
      bool all = true;
      for (size_t i = 0; i < c.vars.size(); i++) all &= br.keys[i] == map[c.vars[i]];
      if (all) chosen = br.block.get();
    }
    if (chosen) break;
  }
  if (!chosen) chosen = c.dflt.get();
  if (!chosen) throw std::runtime_error("No match");
  populate_block(top, q, f, *chosen, map, cx);
}

// FuncChip::generate_trace (trace.rs:72-135): row-major, height = next_pow2(real rows in shard).
static inline std::vector<Fp> func_trace(const Toplevel& top, const QueryRecord& q, uint32_t fi, uint32_t shard,
                                         uint32_t shard_size, size_t* height_out) {
  const FuncDef& f = top.funcs[fi];
  Range rg = shard_func_range(q, fi, shard, shard_size);
  size_t real = rg.len(), height = next_pow2(real), width = f.width();
  std::vector<Fp> rows(height * width, Fp::zero());
  for (size_t i = 0; i < height; i++) rows[i * width] = Fp::from_u32((uint32_t)(rg.start + i));
  const QueryMap& qm = q.func_queries[fi];
  std::string err;
  parallel_for(real, [&](size_t i) {
    try {
      const auto& args = qm.keys[rg.start + i];
      const QueryResult& res = qm.vals[rg.start + i];
      Fp* row = &rows[i * width];
      for (size_t j = 0; j < f.input_size; j++) row[1 + j] = args[j];
      for (size_t j = 0; j < f.output_size; j++) row[1 + f.input_size + j] = res.output[j];
      RowCtx cx{row, (size_t)1 + f.input_size + f.output_size, width, &res.requires_};
      cx.push_aux(Fp::from_u32(res.provide.nonce));
      cx.push_aux(Fp::from_u32(res.provide.count));
      if (f.partial) {  // trace.rs:110-121
        cx.depth_requires = &res.depth_requires;
        cx.depth = res.depth;
        for (int b = 0; b < 4; b++) cx.push_aux(Fp::from_u32((res.depth >> (8 * b)) & 255u));
        cx.push_depth_require();
        cx.push_depth_require();
      }
      std::vector<Fp> map(args);
      populate_block(top, q, f, f.body, map, cx);
    } catch (std::exception& e) {
      err = e.what();
    }
  });
  if (!err.empty()) throw std::runtime_error(err);
  *height_out = height;
  return rows;
}

// MemChip::generate_trace (memory.rs:30-68)
static inline std::vector<Fp> mem_trace(const QueryRecord& q, int mem_idx, size_t* height_out) {
  const QueryMap& mm = q.mem_queries[mem_idx];
  size_t len = MEM_TABLE_SIZES[mem_idx], width = 4 + len;
  size_t height = std::max<size_t>(next_pow2(mm.size()), 4);
  std::vector<Fp> rows(height * width, Fp::zero());
  for (size_t i = 0; i < mm.size(); i++) {
    Fp* r = &rows[i * width];
    r[0] = Fp::one();
    r[1] = Fp::from_u32((uint32_t)i + 1);
    r[2] = Fp::from_u32(mm.vals[i].provide.nonce);
    r[3] = Fp::from_u32(mm.vals[i].provide.count);
    for (size_t j = 0; j < len; j++) r[4 + j] = mm.keys[i][j];
  }
  *height_out = height;
  return rows;
}

// BytesChip::generate_trace (bytes/trace.rs:75-101): 2^16 x 13, all-zero when the record is empty
// or (lair_chip.rs:104-111) on shards other than 0.
static inline std::vector<Fp> bytes_trace(const QueryRecord& q, uint32_t shard) {
  const size_t H = 1 << 16, W = 13;
  std::vector<Fp> rows(H * W, Fp::zero());
  if (shard != 0 || q.bytes.empty()) return rows;
  for (size_t i = 0; i < H; i++) rows[i * W] = Fp::one();
  for (auto& kv : q.bytes)
    for (int k = 0; k < 6; k++) {
      rows[(size_t)kv.first * W + 1 + 2 * k] = Fp::from_u32(kv.second.rec[k].nonce);
      rows[(size_t)kv.first * W + 2 + 2 * k] = Fp::from_u32(kv.second.rec[k].count);
    }
  return rows;
}
// BytesChip::preprocessed_trace (bytes/trace.rs:49-67): (b1, b2, b1<b2, and, xor, or)
static inline std::vector<Fp> bytes_preprocessed() {
  const size_t H = 1 << 16, W = 6;
  std::vector<Fp> rows(H * W);
  for (size_t i = 0; i < H; i++) {
    uint32_t b1 = i & 255, b2 = i >> 8;
    Fp* r = &rows[i * W];
    r[0] = Fp::from_u32(b1); r[1] = Fp::from_u32(b2); r[2] = Fp::from_u32(b1 < b2);
    r[3] = Fp::from_u32(b1 & b2); r[4] = Fp::from_u32(b1 ^ b2); r[5] = Fp::from_u32(b1 | b2);
  }
  return rows;
}

}  // namespace orc
