// TEST INFRASTRUCTURE (oracle build only): tbb::global_control stand-in.
#pragma once
#include <cstddef>
#include "povar_pool.h"
namespace tbb {
class global_control {
 public:
  enum parameter { max_allowed_parallelism, thread_stack_size };
  global_control(parameter p, size_t value) : p_(p) {
    if (p_ == max_allowed_parallelism) {
      prev_ = povar_detail::limit_ref();
      povar_detail::limit_ref() = static_cast<int>(value);
    }
  }
  ~global_control() {
    if (p_ == max_allowed_parallelism) povar_detail::limit_ref() = prev_;
  }
  static size_t active_value(parameter p) {
    if (p == max_allowed_parallelism) {
      return static_cast<size_t>(povar_detail::active_threads());
    }
    return 0;
  }
 private:
  parameter p_;
  int prev_ = 0;
};
}  // namespace tbb
