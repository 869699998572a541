// TEST INFRASTRUCTURE (oracle build only): tbb::this_task_arena stand-in.
#pragma once
#include "povar_pool.h"
namespace tbb {
namespace this_task_arena {
inline int max_concurrency() { return povar_detail::hw_threads(); }
}  // namespace this_task_arena
}  // namespace tbb
