// Host side of the Lair front half, in C++ (the reference is compiled Rust; its toolchain is absent
// here): the `func!` surface as a builder API, expansion of `if` / `match` into
// `Choose` blocks, compilation to index-based bytecode, `Toplevel`, `FuncChip` column layout and the
// flat program image that crosses the C ABI (lsk_program_load).
//
// Mirrors (behaviour, not text):
//   src/lair/expr.rs, src/lair/macros.rs          surface IR (`FuncE`, `OpE`, `CtrlE`)
//   src/lair/toplevel.rs:423-573                  expansion of If / Match / MatchMany
//   src/lair/toplevel.rs:255-330,506-573,739-880  compilation (stack-index linking, return idents)
//   src/lair/bytecode.rs:12-146                   `Op`, `Block`, `Ctrl`, `Func`
//   src/lair/func_chip.rs:90-277                  `compute_layout_sizes`
//   src/lair/mod.rs:53-112                        `demo_toplevel`
// Static, once-per-program logic; the per-proof work is in the CUDA translation units.
#pragma once
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace lurk {
namespace host {

constexpr uint32_t FIELD_P = 2013265921u;
inline uint32_t F(int64_t x) {  // field_from_i32, src/lair/mod.rs:33-40
  int64_t r = x % (int64_t)FIELD_P;
  return (uint32_t)(r < 0 ? r + FIELD_P : r);
}

// ---- surface language (what `func!` produces) ------------------------------------------------------------
struct Var {  // a name, possibly denoting an array of `size` field elements
  std::string name;
  uint32_t size = 1;
  Var(const char* n) : name(n) {}
  Var(std::string n, uint32_t s = 1) : name(std::move(n)), size(s) {}
};
struct St;
using Stmts = std::vector<St>;
enum class SK : uint32_t {
  Const, Array, Add, Sub, Mul, Div, Inv, Not, Eq, AssertEq, AssertNe, Contains, Call, PreImg, Store, Load,
  ExternCall, RangeU8, Return, If, Match, MatchMany, Slice, Emit
};
struct Arm {
  std::vector<uint32_t> keys;  // Match: alternative scalar keys; MatchMany: one key tuple
  Stmts body;
};
struct St {
  SK kind;
  std::string tgt, a, b;            // target / operands (names)
  uint32_t value = 0;               // Const
  std::vector<uint32_t> values;     // Array
  std::vector<Var> outs;            // Call / PreImg / ExternCall / Load
  std::string callee;               // Call / PreImg / ExternCall
  std::vector<std::string> names;   // inputs / stored values / range-checked / returned variables
  std::vector<Arm> arms;            // Match / MatchMany; If: arms[0].body = then
  std::shared_ptr<Stmts> dflt;      // explicit else / default block (null: the statements that follow)
  bool exhaustive = false;
};

// builder helpers, named after the reference's macro forms
St const_(const std::string& tgt, int64_t value);
St array(const std::string& tgt, const std::vector<int64_t>& values);
St add(const std::string& tgt, const std::string& a, const std::string& b);
St sub(const std::string& tgt, const std::string& a, const std::string& b);
St mul(const std::string& tgt, const std::string& a, const std::string& b);
St div(const std::string& tgt, const std::string& a, const std::string& b);
St inv(const std::string& tgt, const std::string& a);
St not_(const std::string& tgt, const std::string& a);
St eq(const std::string& tgt, const std::string& a, const std::string& b);
St assert_eq(const std::string& a, const std::string& b);
St assert_ne(const std::string& a, const std::string& b);
St contains(const std::string& a, const std::string& b);
St call(const std::vector<Var>& outs, const std::string& f, const std::vector<std::string>& inps);
St preimg(const std::vector<Var>& outs, const std::string& f, const std::vector<std::string>& inps);
St store(const std::string& ptr, const std::vector<std::string>& vals);
St load(const std::vector<Var>& vals, const std::string& ptr);
St extern_call(const std::vector<Var>& outs, const std::string& chip, const std::vector<std::string>& inps);
St range_u8(const std::vector<std::string>& xs);
// `let (x, y: [7]) = (a, b);` -- OpE::Slice (src/lair/expr.rs:141-143): rebinds the concatenated values, no bytecode
St slice(const std::vector<Var>& outs, const std::vector<std::string>& inps);
// `emit(x, ...)` -- OpE::Emit: no columns, no constraints; checked and compiled to nothing (see lurk_b200/lair.py::emit)
St emit(const std::vector<std::string>& xs);
St ret(const std::vector<std::string>& vs);
St if_(const std::string& x, Stmts then);                 // the statements that follow form the else block
St if_else(const std::string& x, Stmts then, Stmts else_);
St match(const std::string& x, std::vector<std::pair<std::vector<int64_t>, Stmts>> arms, bool exhaustive = false);
St match_many(const std::string& x, std::vector<std::pair<std::vector<int64_t>, Stmts>> arms, bool exhaustive = false);

struct FuncE {  // src/lair/expr.rs:227-236
  std::string name;
  std::vector<Var> input_params;
  uint32_t output_size = 0;
  Stmts body;
  bool invertible = false, partial = false;
  uint32_t input_size() const {
    uint32_t n = 0;
    for (auto& v : input_params) n += v.size;
    return n;
  }
};
FuncE func(const std::string& name, std::vector<Var> params, uint32_t output_size, Stmts body, bool invertible = false,
           bool partial = false);

struct ChipInfo {  // what Lair needs to know about an extern chip (src/lair/chipset.rs:9-47)
  std::string name;
  uint32_t input_size = 0, output_size = 0, witness_size = 0, require_size = 0;
  uint32_t kind = 0;  // native chipset id (1 / 2 / 3 = Poseidon2 hasher of width 24 / 32 / 40)
};

// ---- bytecode (src/lair/bytecode.rs) ----------------------------------------------------------------------
enum class OK : uint32_t {  // the values are the opcodes of the program image (csrc/program.h)
  Const = 1, Add, Sub, Mul, Inv, Not, AssertEq, AssertNe, Contains, Call, PreImg, Store, Load, ExternCall, RangeU8
};
struct Op {
  OK kind;
  uint32_t a = 0, b = 0;          // scalar operands: value / var indices / func / chip / len
  std::vector<uint32_t> xs, ys;   // list operands
};
struct Block;
struct Ctrl {
  enum Kind { Return, Choose, ChooseMany } kind = Return;
  uint32_t ident = 0;                                                   // Return
  std::vector<uint32_t> vars;                                           // Return values / matched variable(s)
  std::vector<std::pair<std::vector<uint32_t>, std::unique_ptr<Block>>> branches;
  std::unique_ptr<Block> dflt;
};
struct Block {
  std::vector<Op> ops;
  Ctrl ctrl;
  std::vector<uint32_t> return_idents;
};
struct Func {
  std::string name;
  bool invertible = false, partial = false;
  uint32_t index = 0, input_size = 0, output_size = 0;
  Block body;
};
struct LayoutSizes {  // src/lair/func_chip.rs:11-26; max_map: bound of the per-row value stack
  uint32_t nonce = 1, input = 0, output = 0, aux = 0, sel = 0, max_map = 0;
  uint32_t total() const { return nonce + input + output + aux + sel; }
};

constexpr uint32_t DEPTH_W = 4, DEPTH_LESS_THAN_SIZE = 6, DEPTH_LESS_THAN_REQUIRES = 1;
extern const uint32_t MEM_TABLE_SIZES[6];

class Toplevel {  // src/lair/toplevel.rs:25-56
 public:
  Toplevel(const std::vector<FuncE>& funcs, std::vector<ChipInfo> chips = {});
  static Toplevel new_pure(const std::vector<FuncE>& funcs) { return Toplevel(funcs); }
  const Func& func_by_name(const std::string& name) const;
  const Func& func_by_index(uint32_t i) const { return funcs_[i]; }
  uint32_t num_funcs() const { return (uint32_t)funcs_.size(); }
  const std::vector<ChipInfo>& chips() const { return chips_; }
  LayoutSizes compute_layout_sizes(const Func& f) const;   // src/lair/func_chip.rs:90-277
  std::vector<uint32_t> serialize() const;                 // the image lsk_program_load takes

 private:
  std::vector<Func> funcs_;
  std::vector<ChipInfo> chips_;
  std::map<std::string, uint32_t> by_name_;
};

Toplevel demo_toplevel();  // src/lair/mod.rs:53-112

// Surface image (a Toplevel's FuncE list as u32 words; written by lurk_b200/lair.py::surface_image so that the
// Python test harness and this code compile the very same definitions)
std::pair<std::vector<FuncE>, std::vector<ChipInfo>> parse_surface(const uint32_t* w, size_t n);

// build_lair_chip_vector + Chip::new (src/lair/lair_chip.rs:196-231): every chip's Air::eval captured
// symbolically and compiled into the 'MACH' blob lsk_machine_new / lsk_vk_new take (host_air.cpp)
std::vector<uint32_t> chip_vector_blob(const Toplevel& top, uint32_t entry_func);

}  // namespace host
}  // namespace lurk
