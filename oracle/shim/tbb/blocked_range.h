// TEST INFRASTRUCTURE (oracle build only): tbb::blocked_range stand-in.
#pragma once
#include <cstddef>
#include "povar_pool.h"
namespace tbb {
template <typename T>
class blocked_range {
 public:
  using const_iterator = T;
  blocked_range(T b, T e, size_t grain = 1) : b_(b), e_(e), grain_(grain) {}
  T begin() const { return b_; }
  T end() const { return e_; }
  size_t size() const { return static_cast<size_t>(e_ - b_); }
  bool empty() const { return !(b_ < e_); }
  size_t grainsize() const { return grain_; }
 private:
  T b_, e_;
  size_t grain_;
};
}  // namespace tbb
