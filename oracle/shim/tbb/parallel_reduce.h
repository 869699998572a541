// TEST INFRASTRUCTURE (oracle build only): tbb::parallel_reduce stand-in.
// Functional form: per-slot accumulators started from `identity`, joined in
// slot order.  Body form: per-slot bodies made with the splitting constructor,
// joined in slot order.  One thread => a single call over the whole range.
#pragma once
#include <memory>
#include <vector>
#include "blocked_range.h"
#include "parallel_for.h"
namespace tbb {

template <typename T, typename Value, typename Func, typename Join>
Value parallel_reduce(const blocked_range<T>& range, const Value& identity,
                      const Func& func, const Join& join) {
  if (range.empty()) return identity;
  const int threads = povar_detail::active_threads();
  const size_t n = range.size();
  if (threads <= 1 || povar_detail::in_worker()) {
    return func(range, identity);
  }
  const size_t len = povar_detail::chunk_len<T>(n, threads);
  const size_t chunks = (n + len - 1) / len;
  const T base = range.begin();
  std::vector<Value> acc(static_cast<size_t>(threads), identity);
  std::vector<char> used(static_cast<size_t>(threads), 0);
  std::function<void(size_t, int)> job = [&](size_t c, int slot) {
    const size_t lo = c * len;
    const size_t hi = std::min(n, lo + len);
    acc[slot] = func(
        blocked_range<T>(static_cast<T>(base + lo), static_cast<T>(base + hi)),
        acc[slot]);
    used[slot] = 1;
  };
  povar_detail::Pool::instance().run(threads, chunks, job);
  Value res = identity;
  for (int s = 0; s < threads; ++s) {
    if (used[s]) res = join(res, acc[s]);
  }
  return res;
}

template <typename T, typename Body>
void parallel_reduce(const blocked_range<T>& range, Body& body) {
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
  std::vector<std::unique_ptr<Body>> bodies(static_cast<size_t>(threads));
  std::mutex mk;
  std::function<void(size_t, int)> job = [&](size_t c, int slot) {
    if (!bodies[slot]) {
      std::unique_lock<std::mutex> lk(mk);
      bodies[slot] = std::make_unique<Body>(body, split());
    }
    const size_t lo = c * len;
    const size_t hi = std::min(n, lo + len);
    (*bodies[slot])(
        blocked_range<T>(static_cast<T>(base + lo), static_cast<T>(base + hi)));
  };
  povar_detail::Pool::instance().run(threads, chunks, job);
  for (int s = 0; s < threads; ++s) {
    if (bodies[s]) body.join(*bodies[s]);
  }
}
}  // namespace tbb
