// TEST INFRASTRUCTURE (oracle build only): tbb::concurrent_unordered_map
// stand-in.  The reference guards every insertion with its own std::mutex per
// block key, but two DIFFERENT keys may be inserted concurrently, which a
// plain std::unordered_map does not tolerate -- so structural changes
// (emplace/find) take one internal lock; element references stay valid
// (node-based container), which is all the reference relies on.
#pragma once
#include <mutex>
#include <unordered_map>
#include <utility>
namespace tbb {
template <typename Key, typename T, typename Hash = std::hash<Key>,
          typename Eq = std::equal_to<Key>>
class concurrent_unordered_map {
  using Map = std::unordered_map<Key, T, Hash, Eq>;
 public:
  using iterator = typename Map::iterator;
  using const_iterator = typename Map::const_iterator;
  using value_type = typename Map::value_type;

  struct range_type {
    iterator b, e;
    iterator begin() const { return b; }
    iterator end() const { return e; }
  };

  concurrent_unordered_map() = default;
  concurrent_unordered_map(const concurrent_unordered_map& o) : map_(o.map_) {}
  concurrent_unordered_map& operator=(const concurrent_unordered_map& o) {
    map_ = o.map_;
    return *this;
  }

  iterator begin() { return map_.begin(); }
  iterator end() { return map_.end(); }
  const_iterator begin() const { return map_.begin(); }
  const_iterator end() const { return map_.end(); }
  bool empty() const { return map_.empty(); }
  size_t size() const { return map_.size(); }
  size_t count(const Key& k) const {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.count(k);
  }
  iterator find(const Key& k) {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.find(k);
  }
  const_iterator find(const Key& k) const {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.find(k);
  }
  T& at(const Key& k) {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.at(k);
  }
  const T& at(const Key& k) const {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.at(k);
  }
  template <typename... Args>
  std::pair<iterator, bool> emplace(Args&&... args) {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.emplace(std::forward<Args>(args)...);
  }
  template <typename... Args>
  iterator emplace_hint(const_iterator, Args&&... args) {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.emplace(std::forward<Args>(args)...).first;
  }
  void clear() { map_.clear(); }
  range_type range() { return range_type{map_.begin(), map_.end()}; }

 private:
  Map map_;
  mutable std::mutex mu_;
};
}  // namespace tbb
