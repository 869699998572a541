// TEST INFRASTRUCTURE (oracle build only): tbb::task_scheduler_observer
// stand-in.  The shim pool does not report worker entry/exit; observe() fires
// one synthetic entry so that peak concurrency reads 1 rather than 0.
#pragma once
namespace tbb {
class task_scheduler_observer {
 public:
  virtual ~task_scheduler_observer() = default;
  void observe(bool state = true) {
    if (state && !observing_) {
      observing_ = true;
      on_scheduler_entry(false);
    } else if (!state && observing_) {
      observing_ = false;
      on_scheduler_exit(false);
    }
  }
  virtual void on_scheduler_entry(bool /*is_worker*/) {}
  virtual void on_scheduler_exit(bool /*is_worker*/) {}
 private:
  bool observing_ = false;
};
}  // namespace tbb
