// Internal helpers shared by the translation units of liblurkstark (not part of the C ABI).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/lurk_stark.h"

namespace lsk {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}
void set_last_error(const char* msg);

#define LSK_API_BEGIN try {
#define LSK_API_END                                   \
  }                                                   \
  catch (const lsk::Error& e) {                       \
    lsk::set_last_error(e.what());                    \
    return e.code;                                    \
  }                                                   \
  catch (const std::exception& e) {                   \
    lsk::set_last_error(e.what());                    \
    return LSK_ERR_ARG;                               \
  }

}  // namespace lsk
