// TEST INFRASTRUCTURE (oracle build only): tbb::parallel_for stand-in.
#pragma once
#include "blocked_range.h"
namespace tbb {
namespace povar_detail {
template <typename T>
inline size_t chunk_len(size_t n, int threads) {
  size_t target = static_cast<size_t>(threads) * 16;
  size_t len = (n + target - 1) / target;
  return len > 0 ? len : 1;
}
}  // namespace povar_detail

template <typename T, typename Body>
void parallel_for(const blocked_range<T>& range, const Body& body) {
  if (range.empty()) return;
  const int threads = povar_detail::active_threads();
  const size_t n = range.size();
  if (threads <= 1 || povar_detail::in_worker()) {
    body(range);
    return;
  }
  const size_t len = povar_detail::chunk_len<T>(n, threads);
  const size_t chunks = (n + len - 1) / len;
  const T base = range.begin();
  std::function<void(size_t, int)> job = [&](size_t c, int) {
    const size_t lo = c * len;
    const size_t hi = std::min(n, lo + len);
    body(blocked_range<T>(static_cast<T>(base + lo), static_cast<T>(base + hi)));
  };
  povar_detail::Pool::instance().run(threads, chunks, job);
}

// generic range (used with concurrent_unordered_map::range()): run inline.
template <typename Range, typename Body>
void parallel_for(const Range& range, const Body& body) {
  body(range);
}
}  // namespace tbb
