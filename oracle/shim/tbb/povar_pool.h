// TEST INFRASTRUCTURE (oracle build only) -- not product code.
//
// Minimal stand-in for the oneTBB entry points the reference uses, so that the
// UNMODIFIED reference sources under /root/reference compile in an image that
// has no TBB.  Scheduling only: no arithmetic of the path lives here.
//
//  * max_allowed_parallelism == 1  -> every body runs inline on the calling
//    thread over the whole range, in index order (bit-reproducible goldens).
//  * otherwise a persistent std::thread pool pulls fixed-size chunks from an
//    atomic cursor (dynamic scheduling, like TBB's auto partitioner).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace tbb {

struct split {};

namespace povar_detail {

inline int& limit_ref() {
  static int limit = 0;  // 0 = hardware concurrency
  return limit;
}

inline int hw_threads() {
  int n = static_cast<int>(std::thread::hardware_concurrency());
  return n > 0 ? n : 1;
}

inline int active_threads() {
  int l = limit_ref();
  int hw = hw_threads();
  return l > 0 ? std::min(l, hw) : hw;
}

inline bool& in_worker() {
  static thread_local bool flag = false;
  return flag;
}

class Pool {
 public:
  static Pool& instance() {
    static Pool p;
    return p;
  }

  // run job(chunk_index) for chunk_index in [0, num_chunks) on up to
  // `threads` threads (the caller participates).
  void run(int threads, size_t num_chunks,
           const std::function<void(size_t, int)>& job) {
    if (threads <= 1 || num_chunks <= 1 || in_worker()) {
      for (size_t c = 0; c < num_chunks; ++c) job(c, 0);
      return;
    }
    ensure_workers(threads - 1);
    {
      std::unique_lock<std::mutex> lk(mu_);
      job_ = &job;
      cursor_.store(0);
      num_chunks_ = num_chunks;
      wanted_ = threads - 1;
      pending_ = wanted_;
      ++epoch_;
    }
    cv_.notify_all();
    work(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  Pool() = default;
  ~Pool() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }

  void ensure_workers(int n) {
    while (static_cast<int>(workers_.size()) < n) {
      int id = static_cast<int>(workers_.size());
      workers_.emplace_back([this, id] { loop(id); });
    }
  }

  void work(int slot) {
    const std::function<void(size_t, int)>& job = *job_;
    for (;;) {
      size_t c = cursor_.fetch_add(1);
      if (c >= num_chunks_) break;
      job(c, slot);
    }
  }

  void loop(int id) {
    in_worker() = true;
    size_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
        if (id >= wanted_) continue;
      }
      work(id + 1);
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_one();
      }
    }
  }

  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(size_t, int)>* job_ = nullptr;
  std::atomic<size_t> cursor_{0};
  size_t num_chunks_ = 0;
  size_t epoch_ = 0;
  int wanted_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};

}  // namespace povar_detail
}  // namespace tbb
