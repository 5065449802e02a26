// C++ tests of the host layer, written the way the reference writes its own (src/lair/trace.rs:445-459,
// 516-552, 578-619): build a toplevel with the `func!`-style builder, check the column layouts the reference
// pins.  Host code only.   Build + run: see tests/test_host_cpp.py::test_cpp_layout_tests
#include <cstdio>
#include <cstdlib>

#include "host_lair.h"

using namespace lurk::host;

#define CHECK(cond)                                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
      exit(1);                                                       \
    }                                                                \
  } while (0)

static void expect_layout(const Toplevel& top, const char* name, uint32_t input, uint32_t aux, uint32_t sel, uint32_t output) {
  LayoutSizes ls = top.compute_layout_sizes(top.func_by_name(name));
  CHECK(ls.nonce == 1 && ls.input == input && ls.aux == aux && ls.sel == sel && ls.output == output);
}

int main() {
  // lair_layout_sizes_test (src/lair/trace.rs:445-459)
  Toplevel demo = demo_toplevel();
  expect_layout(demo, "factorial", 1, 8, 2, 1);
  CHECK(demo.compute_layout_sizes(demo.func_by_name("fib")).total() == 18);  // golden fib trace width, trace.rs:498-513

  // lair_match_trace_test (src/lair/trace.rs:517-552)
  FuncE test = func("test", {"n", "m"}, 1, {
      const_("one", 1),
      match("n", {{{0}, {ret({"one"})}},
                  {{1}, {ret({"m"})}},
                  {{2}, {mul("res", "m", "m"), ret({"res"})}},
                  {{3}, {mul("res", "m", "m"), mul("res2", "res", "res"), ret({"res2"})}}}),
      sub("pred", "n", "one"),
      call({"res"}, "test", {"pred", "m"}),
      ret({"res"}),
  });
  expect_layout(Toplevel::new_pure({test}), "test", 2, 10, 5, 1);

  // lair_inner_match_trace_test (src/lair/trace.rs:580-619)
  FuncE inner = func("test", {"n", "m"}, 1, {
      const_("zero", 0), const_("one", 1), const_("two", 2), const_("three", 3),
      match("n", {{{0}, {match("m", {{{0}, {ret({"zero"})}}, {{1}, {ret({"one"})}}}, true)}},
                  {{1}, {match("m", {{{0}, {ret({"two"})}}, {{1}, {ret({"three"})}}}, true)}}}, true),
  });
  expect_layout(Toplevel::new_pure({inner}), "test", 2, 2, 4, 1);

  // hasher chips: widths pinned at src/core/eval_direct.rs:2053-2055
  std::vector<ChipInfo> chips{{"hasher3", 24, 8, 457, 0, 1}, {"hasher4", 32, 8, 611, 0, 2}, {"hasher5", 40, 8, 763, 0, 3}};
  FuncE h3 = func("hash3", {Var("preimg", 24)}, 8, {extern_call({Var("img", 8)}, "hasher3", {"preimg"}), ret({"img"})});
  FuncE h4 = func("hash4", {Var("preimg", 32)}, 8, {extern_call({Var("img", 8)}, "hasher4", {"preimg"}), ret({"img"})});
  FuncE h5 = func("hash5", {Var("preimg", 40)}, 8, {extern_call({Var("img", 8)}, "hasher5", {"preimg"}), ret({"img"})});
  Toplevel hashes({h3, h4, h5}, chips);
  CHECK(hashes.compute_layout_sizes(hashes.func_by_name("hash3")).total() == 493);
  CHECK(hashes.compute_layout_sizes(hashes.func_by_name("hash4")).total() == 655);
  CHECK(hashes.compute_layout_sizes(hashes.func_by_name("hash5")).total() == 815);
  // the chip vector captures: Entrypoint + 3 funcs + 6 memory tables + byte table
  std::vector<uint32_t> blob = chip_vector_blob(hashes, 1);
  CHECK(blob[0] == 0x4d414348u && blob[1] == 11 && blob[2] == 40);

  // nonpartial_calls_partial (src/lair/execute.rs:1005-1023): "assertion failed: ctx.partial"
  bool threw = false;
  try {
    FuncE foo = func("foo", {"a"}, 1, {ret({"a"})}, false, true);
    FuncE bar = func("bar", {"a"}, 1, {call({"b"}, "foo", {"a"}), ret({"b"})});
    Toplevel t({foo, bar});
  } catch (const std::exception& e) {
    threw = std::string(e.what()).find("ctx.partial") != std::string::npos;
  }
  CHECK(threw);
  printf("host layout tests ok\n");
  return 0;
}
