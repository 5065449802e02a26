// Flat Lair program image: the form in which a Toplevel (src/lair/toplevel.rs:25-56) crosses the
// C ABI and in which both the host executor and the GPU trace kernel walk it.  All u32 words.
//
//   image   := MAGIC nfuncs nchips { chip: kind in out witness requires }* { func offset }* funcs...
//   func    := in out partial invertible aux sel max_map block
//   block   := len nops op* ctrl nret ident*            (len = words in the block, incl. itself)
//   op      := CONST f | ADD a b | SUB a b | MUL a b has_aux | INV a has_aux | NOT a has_aux
//            | ASSERT_EQ n a.. b.. | ASSERT_NE n a.. b.. | CONTAINS n a.. b
//            | CALL f n in.. nout partial | PREIMG f n out.. nin partial
//            | STORE n v.. | LOAD len ptr | EXTERN chip n in.. | RANGE_U8 n x..
//   ctrl    := RETURN ident n v..
//            | CHOOSE 1 var nbranch has_default { nkeys key* block }* [block]
//            | CHOOSE_MANY nv var* nbranch has_default { 1 key{nv} block }* [block]
// Operands a/b/v/x are indices into the per-row value stack ("map", src/lair/trace.rs:159).
// has_aux says whether the op owns a witness column (degree rule of src/lair/func_chip.rs:199-233).
#pragma once
#include <stdint.h>

namespace lsk {

constexpr uint32_t PROGRAM_MAGIC = 0x4c414952u;

enum : uint32_t {
  OP_CONST = 1, OP_ADD, OP_SUB, OP_MUL, OP_INV, OP_NOT, OP_ASSERT_EQ, OP_ASSERT_NE, OP_CONTAINS,
  OP_CALL, OP_PREIMG, OP_STORE, OP_LOAD, OP_EXTERN, OP_RANGE_U8,
  CTRL_RETURN = 100, CTRL_CHOOSE = 101, CTRL_CHOOSE_MANY = 102
};

constexpr int NUM_MEM_TABLES = 6;
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int mem_table_len(int idx) {
  return idx < 5 ? idx + 2 : 8;   // [2,3,4,5,6,8], src/lair/execute.rs:243-244
}
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int mem_index_from_len(int len) { return len == 8 ? 5 : (len >= 2 && len <= 6 ? len - 2 : -1); }

struct FuncHeader {   // the 7 words before a func's body block
  uint32_t in, out, partial, invertible, aux, sel, max_map;
};
constexpr uint32_t FUNC_HEADER_WORDS = 7;
constexpr uint32_t CHIP_WORDS = 5;

// number of words an op occupies, given a pointer to its opcode
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t op_words(const uint32_t* w) {
  switch (w[0]) {
    case OP_CONST: return 2;
    case OP_ADD: case OP_SUB: return 3;
    case OP_MUL: return 4;
    case OP_INV: case OP_NOT: return 3;
    case OP_ASSERT_EQ: case OP_ASSERT_NE: return 2 + 2 * w[1];
    case OP_CONTAINS: return 3 + w[1];
    case OP_CALL: case OP_PREIMG: return 5 + w[2];
    case OP_STORE: case OP_RANGE_U8: return 2 + w[1];
    case OP_LOAD: return 3;
    case OP_EXTERN: return 3 + w[2];
    default: return 0;
  }
}

}  // namespace lsk
