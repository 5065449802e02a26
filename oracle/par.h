// ORACLE (test infrastructure): minimal std::thread parallel-for (stands in for rayon in the
// reference; OpenMP is not usable with this image's compiler wrapper).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

namespace orc {
static inline int num_threads() {
  const char* e = std::getenv("ORACLE_THREADS");
  int n = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
  return n > 0 ? n : 1;
}
// f(begin, end) over [0, n) split into contiguous chunks
template <class F>
static inline void parallel_chunks(size_t n, F f, size_t min_chunk = 256) {
  int nt = num_threads();
  if (nt <= 1 || n <= min_chunk) {
    f((size_t)0, n);
    return;
  }
  size_t chunks = std::min<size_t>(nt, (n + min_chunk - 1) / min_chunk);
  size_t per = (n + chunks - 1) / chunks;
  std::vector<std::thread> th;
  for (size_t c = 0; c < chunks; c++) {
    size_t b = c * per, e = std::min(n, b + per);
    if (b >= e) break;
    th.emplace_back([=] { f(b, e); });
  }
  for (auto& t : th) t.join();
}
template <class F>
static inline void parallel_for(size_t n, F f, size_t min_chunk = 256) {
  parallel_chunks(n, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) f(i); }, min_chunk);
}
}  // namespace orc
