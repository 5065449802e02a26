// ORACLE (test infrastructure, NOT product code): C entry points over lair.h for ctypes.
#include "lair.h"
#include <cstring>
using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (std::exception& e) { g_err = e.what(); return -1; }

extern "C" {
const char* orc_last_error() { return g_err.c_str(); }

void* orc_toplevel_new(const uint32_t* prog, size_t n) {
  try { return parse_toplevel(prog, n); } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
void orc_toplevel_free(void* t) { delete (Toplevel*)t; }
void* orc_record_new(void* t) { return new QueryRecord(*(Toplevel*)t); }
void orc_record_free(void* r) { delete (QueryRecord*)r; }

int orc_execute(void* t, void* r, uint32_t func, const uint32_t* args, size_t nargs, uint32_t* out, size_t nout) {
  ORC_TRY
  std::vector<Fp> a(nargs);
  for (size_t i = 0; i < nargs; i++) a[i] = Fp::from_u32(args[i]);
  auto o = execute(*(Toplevel*)t, func, a, *(QueryRecord*)r);
  if (o.size() != nout) throw std::runtime_error("output size mismatch");
  for (size_t i = 0; i < nout; i++) out[i] = o[i].to_u32();
  return 0;
  ORC_CATCH
}
// inject_inv_queries (execute.rs:299-319): n pairs (inp[in_size], out[out_size])
int orc_inject_inv_queries(void* t, void* r, uint32_t func, const uint32_t* inps, const uint32_t* outs, size_t n) {
  ORC_TRY
  const FuncDef& f = ((Toplevel*)t)->funcs[func];
  auto& m = ((QueryRecord*)r)->inv_func_queries[func];
  for (size_t i = 0; i < n; i++) {
    std::vector<Fp> in(f.input_size), out(f.output_size);
    for (size_t j = 0; j < f.input_size; j++) in[j] = Fp::from_u32(inps[i * f.input_size + j]);
    for (size_t j = 0; j < f.output_size; j++) out[j] = Fp::from_u32(outs[i * f.output_size + j]);
    m[QueryMap::raw(out)] = in;
  }
  return 0;
  ORC_CATCH
}
size_t orc_num_queries(void* r, uint32_t func) { return ((QueryRecord*)r)->func_queries[func].size(); }
size_t orc_num_mem(void* r, uint32_t mem_idx) { return ((QueryRecord*)r)->mem_queries[mem_idx].size(); }
uint32_t orc_num_shards(void* r, uint32_t shard_size) { return num_shards(*(QueryRecord*)r, shard_size); }
size_t orc_public_values(void* r, uint32_t* out, size_t cap) {
  auto& pv = ((QueryRecord*)r)->public_values;
  for (size_t i = 0; i < pv.size() && i < cap; i++) out[i] = pv[i].to_u32();
  return pv.size();
}
uint32_t orc_func_width(void* t, uint32_t func) { return ((Toplevel*)t)->funcs[func].width(); }
size_t orc_func_trace_height(void* r, uint32_t func, uint32_t shard, uint32_t shard_size) {
  return next_pow2(shard_func_range(*(QueryRecord*)r, func, shard, shard_size).len());
}
int orc_func_trace(void* t, void* r, uint32_t func, uint32_t shard, uint32_t shard_size, uint32_t* out) {
  ORC_TRY
  size_t h;
  auto rows = func_trace(*(Toplevel*)t, *(QueryRecord*)r, func, shard, shard_size, &h);
  for (size_t i = 0; i < rows.size(); i++) out[i] = rows[i].to_u32();
  return 0;
  ORC_CATCH
}
size_t orc_mem_trace_height(void* r, uint32_t mem_idx) {
  return std::max<size_t>(next_pow2(((QueryRecord*)r)->mem_queries[mem_idx].size()), 4);
}
int orc_mem_trace(void* r, uint32_t mem_idx, uint32_t* out) {
  ORC_TRY
  size_t h;
  auto rows = mem_trace(*(QueryRecord*)r, mem_idx, &h);
  for (size_t i = 0; i < rows.size(); i++) out[i] = rows[i].to_u32();
  return 0;
  ORC_CATCH
}
int orc_bytes_trace(void* r, uint32_t shard, uint32_t* out) {
  ORC_TRY
  auto rows = bytes_trace(*(QueryRecord*)r, shard);
  for (size_t i = 0; i < rows.size(); i++) out[i] = rows[i].to_u32();
  return 0;
  ORC_CATCH
}
int orc_bytes_preprocessed(uint32_t* out) {
  auto rows = bytes_preprocessed();
  for (size_t i = 0; i < rows.size(); i++) out[i] = rows[i].to_u32();
  return 0;
}
}  // extern "C"
