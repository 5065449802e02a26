// C-ABI entry points of the C++ host layer (host_lair.h / host_air.cpp): compile a Toplevel from its surface
// image, capture the chip vector's AIR.  Pure host code -- no device is touched.
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "host_lair.h"

using namespace lsk;
namespace lh = lurk::host;

static int give(const std::vector<uint32_t>& v, uint32_t** words, size_t* len) {
  *words = (uint32_t*)malloc(v.size() * 4 + 4);
  if (!*words) fail(LSK_ERR_ARG, "out of memory");
  memcpy(*words, v.data(), v.size() * 4);
  *len = v.size();
  return LSK_OK;
}

extern "C" {

int lsk_host_compile_toplevel(const uint32_t* surface, size_t n, uint32_t** words, size_t* len) {
  LSK_API_BEGIN
  if (!surface || !words || !len) fail(LSK_ERR_ARG, "null argument");
  auto parsed = lh::parse_surface(surface, n);
  lh::Toplevel top(parsed.first, parsed.second);
  return give(top.serialize(), words, len);
  LSK_API_END
}

int lsk_host_chip_vector(const uint32_t* surface, size_t n, uint32_t entry_func, uint32_t** words, size_t* len) {
  LSK_API_BEGIN
  if (!surface || !words || !len) fail(LSK_ERR_ARG, "null argument");
  auto parsed = lh::parse_surface(surface, n);
  lh::Toplevel top(parsed.first, parsed.second);
  if (entry_func >= top.num_funcs()) fail(LSK_ERR_ARG, "entry func out of range");
  return give(lh::chip_vector_blob(top, entry_func), words, len);
  LSK_API_END
}

int lsk_host_demo_toplevel(uint32_t** words, size_t* len) {
  LSK_API_BEGIN
  if (!words || !len) fail(LSK_ERR_ARG, "null argument");
  return give(lh::demo_toplevel().serialize(), words, len);
  LSK_API_END
}

void lsk_words_free(uint32_t* words) { free(words); }

}  // extern "C"
