// Engine: device state of one landmark shard + kernel sequencing for the PoVar hot path.
//
// Mirrors, per call, what LinearizorPowerVarproj / LinearizorSC do
// (/root/reference/src/rootba_povar/solver/linearizor_power_varproj.cpp, linearizor_sc.cpp):
//   linearize  -> landmark pass (Jl^T Jl, Jl^T r, column scales) + camera pass (Jp^T Jp) + scales
//   solve      -> Hll^-1, B^-1, b, then the power series (or PCG / Cholesky)
//   apply      -> back-substitution + camera update, model cost change
// The Jacobian blocks are never stored: see device_math.cuh.
#include "engine.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <thread>

namespace povar {

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: libpovar_b200.so stays loadable on a box without NCCL; the entry points
// are only needed when world_size > 1.
// ---------------------------------------------------------------------------------------------
struct Id128 {   // ncclUniqueId: 128 opaque bytes, passed by value
  char bytes[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* load_nccl(std::string* err) {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {getenv("POVAR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    if (err) *err = "cannot dlopen libnccl.so.2 (set POVAR_NCCL_LIB)";
    return nullptr;
  }
  api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclGetUniqueId"));
  api.CommInitRank =
      reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(api.lib, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
      dlsym(api.lib, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
    if (err) *err = "libnccl is missing expected symbols";
    dlclose(api.lib);
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

static std::mutex& comm_cache_mutex() {
  static std::mutex m;
  return m;
}
static std::map<std::string, void*>& comm_cache() {
  static std::map<std::string, void*> cache;
  return cache;
}
// Stream, events and the page-locked scratch of a handle are kept for the next handle on the same device instead of
// being destroyed: cudaHostAlloc / cudaFreeHost are the erratic part of povar_create / povar_destroy (3 ms and 0.5 ms
// as a rule, 130 ms for one cudaFreeHost in a trace, and the likely cause of end-to-end steps of 420 to 580 ms on
// some hosts).  A kit goes back only after its stream was synchronised.
struct HostKit {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double* pinned = nullptr;   // 256 bytes
};
static std::mutex& kit_mutex() {
  static std::mutex m;
  return m;
}
static std::vector<HostKit>& kit_cache() {
  static std::vector<HostKit> cache;
  return cache;
}
static bool take_kit(int device, HostKit* out) {
  std::lock_guard<std::mutex> lock(kit_mutex());
  auto& cache = kit_cache();
  for (size_t i = 0; i < cache.size(); ++i) {
    if (cache[i].device == device) {
      *out = cache[i];
      cache.erase(cache.begin() + static_cast<long>(i));
      return true;
    }
  }
  return false;
}
static void give_kit(const HostKit& kit) {
  std::lock_guard<std::mutex> lock(kit_mutex());
  kit_cache().push_back(kit);
}

struct PeerShared {
  void* mem = nullptr;            // this rank's receive buffer + flags (cudaMalloc, IPC-exported)
  std::vector<void*> opened;      // the peers' buffers as mapped here
  PeerExchange px{};
  bool ok = false;
  void release() {
    ok = false;
    for (void* p : opened) cudaIpcCloseMemHandle(p);
    opened.clear();
    if (mem) cudaFree(mem);
    mem = nullptr;
    cudaGetLastError();
  }
};
static std::map<std::string, PeerShared*>& peer_cache() {
  static std::map<std::string, PeerShared*> cache;
  return cache;
}
static std::string comm_key(const povar_comm_desc& c) {
  std::string key(reinterpret_cast<const char*>(c.nccl_id), 128);
  key += ":" + std::to_string(c.rank) + ":" + std::to_string(c.world_size) + ":" + std::to_string(c.device);
  return key;
}

static void release_rendezvous();   // host rendezvous segments (below)

// destroys every cached communicator; handles made with them must have been destroyed before
int nccl_finalize() {
  std::lock_guard<std::mutex> lock(comm_cache_mutex());
  NcclApi* api = comm_cache().empty() ? nullptr : load_nccl(nullptr);
  for (auto& kv : peer_cache()) {
    kv.second->release();
    delete kv.second;
  }
  peer_cache().clear();
  release_rendezvous();
  for (auto& kv : comm_cache()) {
    if (api && kv.second) api->CommDestroy(kv.second);
  }
  comm_cache().clear();
  return POVAR_OK;
}

int nccl_unique_id(uint8_t id[128], std::string* err) {
  NcclApi* api = load_nccl(err);
  if (!api) return POVAR_ERR_NCCL;
  const int rc = api->GetUniqueId(id);
  if (rc != 0) {
    if (err) *err = std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(rc) : "?");
    return POVAR_ERR_NCCL;
  }
  return POVAR_OK;
}

// ---------------------------------------------------------------------------------------------
// Host rendezvous: a POSIX shared-memory segment instead of an NCCL communicator for the ONE thing the
// communicator is needed for when the peer buffers carry every reduction -- swapping the CUDA IPC handles at
// set-up.  Selected by a communicator id that starts with kHostIdMagic (povar_comm_host_id).  It is what lets
// several ranks share one device (NCCL refuses two ranks on the same GPU): the sharded arithmetic, the IPC
// mapping and the tagged peer stores are then exactly those of a multi-GPU run, on a one-GPU box.
// ---------------------------------------------------------------------------------------------
static const char kHostIdMagic[] = "POVAR-SHM:";
constexpr int kRdvRounds = 64;       // rendezvous per id (one per distinct camera count)
constexpr int kRdvSlotWords = 160;   // >= kMaxPeers * 17 words
struct RdvRegion {
  std::atomic<unsigned int> arrived[kRdvRounds][2];
  unsigned int payload[kRdvRounds][2][kMaxPeers][kRdvSlotWords];
};
struct HostRendezvous {
  RdvRegion* region = nullptr;
  int next_round = 0;
  std::string name;
};
static std::map<std::string, HostRendezvous>& rdv_cache() {
  static std::map<std::string, HostRendezvous> cache;
  return cache;
}
static void release_rendezvous() {
  for (auto& kv : rdv_cache()) {
    if (kv.second.region) munmap(kv.second.region, sizeof(RdvRegion));
    shm_unlink(kv.first.c_str());
  }
  rdv_cache().clear();
}
static bool is_host_id(const uint8_t* id) { return std::memcmp(id, kHostIdMagic, sizeof(kHostIdMagic) - 1) == 0; }

int host_unique_id(uint8_t id[128]) {
  static std::atomic<unsigned int> counter{0};
  std::memset(id, 0, 128);
  const auto now = std::chrono::steady_clock::now().time_since_epoch().count();
  std::snprintf(reinterpret_cast<char*>(id), 128, "%s/povar_%d_%u_%llx", kHostIdMagic, static_cast<int>(getpid()),
                counter.fetch_add(1), static_cast<unsigned long long>(now));
  return POVAR_OK;
}

// comm_cache_mutex() held by the caller
static HostRendezvous* open_rendezvous(const uint8_t* id, std::string* err) {
  const std::string name(reinterpret_cast<const char*>(id) + sizeof(kHostIdMagic) - 1);
  auto& cache = rdv_cache();
  auto it = cache.find(name);
  if (it != cache.end()) return &it->second;
  const int fd = shm_open(name.c_str(), O_CREAT | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, sizeof(RdvRegion)) != 0) {   // a new segment reads as zeros; same size on every rank
    if (fd >= 0) close(fd);
    if (err) *err = "host rendezvous: shm_open(" + name + ") failed";
    return nullptr;
  }
  void* p = mmap(nullptr, sizeof(RdvRegion), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    if (err) *err = "host rendezvous: mmap failed";
    return nullptr;
  }
  HostRendezvous& r = cache[name];
  r.region = static_cast<RdvRegion*>(p);
  r.name = name;
  return &r;
}

// element-wise sum over the ranks of `words` unsigned ints (what the set-up does with ncclAllReduce otherwise)
static bool rendezvous_sum(HostRendezvous* r, int round, int phase, int rank, int world, unsigned int* data,
                           int words) {
  if (round >= kRdvRounds || words > kRdvSlotWords) return false;
  RdvRegion* g = r->region;
  std::memcpy(g->payload[round][phase][rank], data, sizeof(unsigned int) * words);
  g->arrived[round][phase].fetch_add(1, std::memory_order_acq_rel);
  const auto t0 = std::chrono::steady_clock::now();
  while (g->arrived[round][phase].load(std::memory_order_acquire) < static_cast<unsigned int>(world)) {
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return false;   // a rank never came
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
  for (int w = 0; w < words; ++w) data[w] = 0;
  for (int q = 0; q < world; ++q) {
    for (int w = 0; w < words; ++w) data[w] += g->payload[round][phase][q][w];
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// host-side index construction
// ---------------------------------------------------------------------------------------------
// Static chunking over a few host threads: povar_create walks the observation list a handful of times
// (validation, per-camera counts, sliced-ELL order) and at 5e6 observations a single thread spends more
// on that than the GPU on ten LM iterations.  Results never depend on the thread count.
static int g_host_threads_override = 0;   // povar_debug_sell_layout: tests compare thread counts
void set_host_threads_override(int n) { g_host_threads_override = n; }
static int host_threads(long long work) {
  static const int forced = getenv("POVAR_HOST_THREADS") ? atoi(getenv("POVAR_HOST_THREADS")) : 0;
  if (g_host_threads_override > 0) return g_host_threads_override;
  if (forced > 0) return forced;
  if (work < (1 << 18)) return 1;
  const unsigned hw = std::thread::hardware_concurrency();
  return static_cast<int>(std::min<unsigned>(hw == 0 ? 1 : hw, 16));
}
template <typename F>
static void parallel_chunks(int chunks, F&& body) {   // body(chunk index)
  if (chunks <= 1) {
    body(0);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(chunks - 1);
  for (int t = 1; t < chunks; ++t) pool.emplace_back([&body, t] { body(t); });
  body(0);
  for (auto& th : pool) th.join();
}

// Sliced ELL order of the landmarks with 1..32 observations (the landmark half of E0 gives a lane
// to each landmark and walks its observations serially, so the 32 landmarks of a slice should have the
// same degree).  Landmarks are first put in the order of their KEY camera (stable counting sort): the centre
// of the stretch of kSellKeySpan cameras that holds most of the landmark's observations (the first such
// stretch; for a landmark whose cameras span less than that, the middle between its first and its last
// camera).  Neighbouring slices then gather from one stretch of the camera table, as short as the tracks
// allow -- a block of the landmark half stages exactly that stretch in shared memory (kernels_series.cu); with
// the median camera as key the stretch was twice as long on banded scenes.  Inside windows of `window`
// landmarks of that order: stable sort by descending degree, kSellWidth landmarks per slice, slice length =
// largest degree in it.  The order of the observations INSIDE a landmark is untouched (camera ascending), so
// every H_l keeps its bits.
int sell_key(const int* cams, int deg, int span) {
  int best_i = 0, best_j = 0;
  for (int i = 0, j = 0; i < deg; ++i) {
    if (j < i) j = i;
    while (j + 1 < deg && cams[j + 1] - cams[i] < span) ++j;
    if (j - i > best_j - best_i) {
      best_i = i;
      best_j = j;
    }
  }
  return (cams[best_i] + cams[best_j]) / 2;
}

int sell_window(int landmarks, int max_window) {
  int w = max_window;
  while (w > 512 && static_cast<long long>(w) * 128 > landmarks) w /= 2;
  return w;
}

// Largest degree of a landmark of the sliced-ELL set.  A slice is walked row by row by ONE warp, so a slice of 32
// rows is a chain of 32 dependent steps (0.76 us each when the SM is not full); on a large shard every warp has
// dozens of rows anyway, on a small one (a venice-1778 shard on 8 GPUs: one slice of ~5 rows per warp) the few
// slices of 30 rows set the time of the whole launch (27 us where the median block needs 19).  There the
// landmarks with more than about twice the rows a warp has anyway go to the warp-per-landmark kernels instead
// (32 observations per step): 31 -> 22 us per launch with 12, measured on that shard.
int sell_max_degree(long long nnz, int sms) {
  const long long rows_per_warp = nnz / (32LL * 32 * sms);   // rows of one resident wave of warps
  if (rows_per_warp >= 16) return 32;
  long long t = (2 * rows_per_warp + 3) / 4 * 4;
  if (t < 8) t = 8;
  if (t > 32) t = 32;
  return static_cast<int>(t);
}

void build_sell(const std::vector<int>& lm_ptr, const int* obs_cam, int num_cams, int max_window,
                SellLayout* out, int max_deg) {
  const int L = static_cast<int>(lm_ptr.size()) - 1;
  out->slice_ptr.assign(1, 0);
  out->sell_lm.clear();
  out->slice_lo.clear();
  out->slice_hi.clear();
  out->long_lms.clear();
  out->rows = 0;
  if (L <= 0) return;
  const int T = host_threads(L);
  auto chunk_begin = [&](int t) { return static_cast<int>(static_cast<long long>(L) * t / T); };
  // key camera of every landmark (-1: not in the set), made once
  std::vector<int> keys(static_cast<size_t>(L));
  // landmarks with 1..32 observations by key camera: stable counting sort, one histogram per chunk
  // (chunk t's landmarks of a camera go after those of the chunks before it)
  std::vector<std::vector<int>> hist(T, std::vector<int>(static_cast<size_t>(num_cams), 0));
  std::vector<std::vector<int>> longs(T);
  parallel_chunks(T, [&](int t) {
    std::vector<int>& h = hist[t];
    for (int l = chunk_begin(t); l < chunk_begin(t + 1); ++l) {
      const int deg = lm_ptr[l + 1] - lm_ptr[l];
      keys[l] = -1;
      if (deg > max_deg) {
        longs[t].push_back(l);
      } else if (deg > 0) {
        keys[l] = sell_key(obs_cam + lm_ptr[l], deg, kSellKeySpan);
        h[keys[l]]++;
      }
    }
  });
  for (int t = 0; t < T; ++t) out->long_lms.insert(out->long_lms.end(), longs[t].begin(), longs[t].end());
  int n = 0;
  for (int c = 0; c < num_cams; ++c) {
    for (int t = 0; t < T; ++t) {
      const int cnt = hist[t][c];
      hist[t][c] = n;   // first position of chunk t's landmarks with key camera c
      n += cnt;
    }
  }
  std::vector<int> by_cam(static_cast<size_t>(n));
  parallel_chunks(T, [&](int t) {
    std::vector<int>& h = hist[t];
    for (int l = chunk_begin(t); l < chunk_begin(t + 1); ++l) {
      if (keys[l] >= 0) by_cam[h[keys[l]]++] = l;
    }
  });
  // windows of `window` landmarks of that order, each sorted (stably) by descending degree and cut into
  // slices: the windows are independent, and where a window's slices go is known up front.  A long window
  // means little padding (2.6 % at 4,096 on venice-1778, 16 % at 512) but its slices mix the keys of the whole
  // window; a shard with few landmarks per camera takes a shorter one, so that the cameras a block of the
  // landmark half meets stay one stretch of the table: `max_window`, halved while there are fewer than 128
  // windows, not below 512.
  const int window = sell_window(n, max_window);
  const int num_windows = (n + window - 1) / window;
  const int W = kSellWidth;
  const int slices_per_full = (window + W - 1) / W;
  const int last_cnt = n - (num_windows - 1) * window;
  const int num_slices = num_windows == 0 ? 0 : (num_windows - 1) * slices_per_full + (last_cnt + W - 1) / W;
  out->sell_lm.assign(static_cast<size_t>(num_slices) * W, -1);
  out->slice_lo.assign(static_cast<size_t>(num_slices), 0);
  out->slice_hi.assign(static_cast<size_t>(num_slices), 0);
  std::vector<int> slice_len(static_cast<size_t>(num_slices), 0);
  parallel_chunks(T, [&](int t) {
    int head[34];
    std::vector<int> order(window);
    const int wb = static_cast<int>(static_cast<long long>(num_windows) * t / T);
    const int we = static_cast<int>(static_cast<long long>(num_windows) * (t + 1) / T);
    for (int w = wb; w < we; ++w) {
      const int w0 = w * window, w1 = std::min(n, w0 + window);
      for (int k = 0; k < 34; ++k) head[k] = 0;
      for (int i = w0; i < w1; ++i) head[32 - (lm_ptr[by_cam[i] + 1] - lm_ptr[by_cam[i]]) + 1]++;
      for (int k = 0; k < 33; ++k) head[k + 1] += head[k];
      for (int i = w0; i < w1; ++i) order[head[32 - (lm_ptr[by_cam[i] + 1] - lm_ptr[by_cam[i]])]++] = by_cam[i];
      const int cnt = w1 - w0;
      int sl = w * slices_per_full;
      for (int i = 0; i < cnt; i += W, ++sl) {
        slice_len[sl] = lm_ptr[order[i] + 1] - lm_ptr[order[i]];   // the largest degree of the slice
        int lo = num_cams, hi = -1;
        for (int g = 0; g < W && i + g < cnt; ++g) {
          const int l = order[i + g];
          out->sell_lm[static_cast<size_t>(sl) * W + g] = l;
          lo = std::min(lo, obs_cam[lm_ptr[l]]);            // cameras ascend inside a landmark
          hi = std::max(hi, obs_cam[lm_ptr[l + 1] - 1]);
        }
        out->slice_lo[sl] = lo;
        out->slice_hi[sl] = hi;
      }
    }
  });
  out->slice_ptr.resize(static_cast<size_t>(num_slices) + 1);
  int rows = 0;
  for (int sl = 0; sl < num_slices; ++sl) {
    rows += slice_len[sl];
    out->slice_ptr[sl + 1] = rows;
  }
  out->rows = rows;
}

// ---- plan of the landmark half (LmPlan, povar_internal.h) ----
// Shared memory of one block: [mbarriers][warps x stages x stage_bytes of stream ring][window].
size_t landmark_half_smem(int warps, int stages, int stage_bytes, int win_cams, int rec_bytes) {
  return lm_bar_bytes(warps, stages) + static_cast<size_t>(warps) * stages * stage_bytes +
         static_cast<size_t>(win_cams) * rec_bytes;
}

LmPlanHost plan_landmark_half(const SellLayout& sell, int num_cams, int num_long, int rec_bytes, int stage_bytes,
                              int sms, int max_warps_per_sm, int reserve_bytes) {
  const int S = static_cast<int>(sell.slice_ptr.size()) - 1;
  constexpr size_t kSmemPerSm = 228 * 1024;   // sm_100: 228 KB per SM, 1 KB of it reserved per resident block
  struct Cand {
    int warps, stages, bps;
  };
  // preference: small blocks while the table still fits beside the rings (short launches on small shards,
  // less tail), then one block per SM with the deepest ring, then fewer warps around a larger window
  // (operations with many registers per lane ask for at most 16 warps per SM: 128 registers per thread)
  const Cand cands[] = {{8, 3, 4}, {16, 3, 2}, {32, 3, 1}, {32, 2, 1}, {24, 2, 1},
                        {8, 3, 2}, {16, 3, 1}, {16, 2, 1}};
  LmPlanHost best;
  int best_win = -1;
  for (const Cand& cd : cands) {
    if (cd.warps * cd.bps > max_warps_per_sm) continue;
    LmPlanHost h;
    LmPlan& p = h.p;
    p.warps = cd.warps;
    p.stages = cd.stages;
    p.blocks_per_sm = cd.bps;
    const long long resident = static_cast<long long>(sms) * cd.bps * cd.warps;   // warps of one wave
    p.ranges = static_cast<int>(std::min<long long>(resident, S));
    // ranges of (nearly) equal cost -- rows plus kSliceCost per slice (opening and closing a slice: its landmarks
    // in, its results out, about two rows' worth; blocks of many short slices were 20 % behind those of few
    // long ones with rows alone): range w starts at the first slice at or after cost w * total / ranges
    h.range_slice.assign(static_cast<size_t>(p.ranges) + 1, S);
    {
      const long long total = S > 0 ? sell.slice_ptr[S] + static_cast<long long>(kSliceCost) * S : 0;
      int sl = 0;
      for (int w = 0; w < p.ranges; ++w) {
        const long long target = total * w / p.ranges;
        while (sl < S && sell.slice_ptr[sl] + static_cast<long long>(kSliceCost) * sl < target) ++sl;
        h.range_slice[w] = sl;
      }
    }
    const int slice_blocks = (p.ranges + cd.warps - 1) / cd.warps;
    const int long_blocks = static_cast<int>(std::min<long long>(static_cast<long long>(sms) * cd.bps,
                                                                 (num_long + cd.warps - 1) / cd.warps));
    p.blocks = std::max(slice_blocks, long_blocks);
    // cameras every block meets
    std::vector<int> lo(static_cast<size_t>(std::max(p.blocks, 1)), num_cams), hi(lo.size(), -1);
    int need = 1;
    for (int b = 0; b < slice_blocks; ++b) {
      const int s0 = h.range_slice[static_cast<size_t>(b) * cd.warps];
      const int s1 = h.range_slice[std::min<size_t>(static_cast<size_t>(b + 1) * cd.warps, p.ranges)];
      for (int sl = s0; sl < s1; ++sl) {
        lo[b] = std::min(lo[b], sell.slice_lo[sl]);
        hi[b] = std::max(hi[b], sell.slice_hi[sl]);
      }
      if (hi[b] >= lo[b]) need = std::max(need, hi[b] - lo[b] + 1);
    }
    const size_t budget = kSmemPerSm / cd.bps - 1024 - static_cast<size_t>(reserve_bytes);
    const size_t fixed = landmark_half_smem(cd.warps, cd.stages, stage_bytes, 0, rec_bytes);
    const int fit = budget > fixed ? static_cast<int>((budget - fixed) / rec_bytes) : 0;
    p.win_cams = std::min(num_cams, std::min(need, fit));
    p.covered = fit >= std::min(need, num_cams) ? 1 : 0;
    h.blk_lo.assign(static_cast<size_t>(std::max(p.blocks, 1)), 0);
    for (int b = 0; b < p.blocks; ++b) {
      int start = 0;
      if (hi[b] >= lo[b]) {
        const int span = hi[b] - lo[b] + 1;
        start = span <= p.win_cams ? lo[b] : lo[b] + (span - p.win_cams) / 2;   // not covered: the middle
      }
      h.blk_lo[b] = std::max(0, std::min(start, num_cams - p.win_cams));
    }
    h.smem_bytes = landmark_half_smem(cd.warps, cd.stages, stage_bytes, p.win_cams, rec_bytes);
    if (p.covered) return h;
    if (p.win_cams > best_win) {
      best_win = p.win_cams;
      best = std::move(h);
    }
  }
  return best;
}

int choose_item_len(long long nnz) {
  // enough items to fill every SM with 64 warps a couple of times, at most 256 entries per item (measured on the
  // venice-1778 shape: 128 -> camera half 67.8 us, 256 -> 58.7 us, 512 -> 61.8 us)
  long long len = nnz / (static_cast<long long>(sm_count()) * 64 * 2);
  len = (len + 31) / 32 * 32;
  if (len < 32) len = 32;
  if (len > 256) len = 256;
  return static_cast<int>(len);
}

void build_items(const std::vector<int>& cam_ptr, int item_len, std::vector<int>* item_ptr,
                 std::vector<int>* item_cam, std::vector<int>* cam_item_ptr) {
  const int C = static_cast<int>(cam_ptr.size()) - 1;
  item_ptr->clear();
  item_cam->clear();
  cam_item_ptr->assign(C + 1, 0);
  item_ptr->push_back(0);
  for (int c = 0; c < C; ++c) {
    (*cam_item_ptr)[c] = static_cast<int>(item_cam->size());
    for (int b = cam_ptr[c]; b < cam_ptr[c + 1]; b += item_len) {
      item_cam->push_back(c);
      item_ptr->push_back(std::min(b + item_len, cam_ptr[c + 1]));
    }
  }
  (*cam_item_ptr)[C] = static_cast<int>(item_cam->size());
}


// ---------------------------------------------------------------------------------------------
int Engine::fail(int code, const std::string& what) {
  err_ = what;
  return code;
}

int Engine::check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return POVAR_OK;
  err_ = std::string(what) + ": " + cudaGetErrorString(e);
  return POVAR_ERR_CUDA;
}

#define PV_CUDA(call)                                   \
  do {                                                  \
    const int rc_ = check((call), #call);               \
    if (rc_ != POVAR_OK) return rc_;                    \
  } while (0)

double Engine::elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms * 1e-3;
}

// Stream-ordered allocation from the device's default memory pool, which is told to keep what it is
// given back: creating a handle after another one was destroyed costs no cudaMalloc.
template <typename T>
static int dev_alloc(std::vector<void*>& allocs, cudaStream_t stream, T** p, size_t n) {
  void* q = nullptr;
  if (n == 0) n = 1;
  const cudaError_t e = cudaMallocAsync(&q, n * sizeof(T), stream);
  if (e != cudaSuccess) return POVAR_ERR_CUDA;
  cudaMemsetAsync(q, 0, n * sizeof(T), stream);
  allocs.push_back(q);
  *p = static_cast<T*>(q);
  return POVAR_OK;
}

#define PV_ALLOC(ptr, n)                                                        \
  do {                                                                          \
    if (dev_alloc(allocs_, stream_, &(ptr), static_cast<size_t>(n)) != POVAR_OK) \
      return fail(POVAR_ERR_CUDA, std::string("cudaMalloc failed for ") + #ptr); \
  } while (0)

int Engine::create(const povar_problem_desc* desc, const povar_options* opt,
                   const povar_comm_desc* comm, Engine** out, std::string* err) {
  *out = nullptr;
  if (!desc || !opt || desc->num_cams <= 0 || desc->num_lms < 0 || desc->num_obs < 0 ||
      !desc->lm_ptr || !desc->cam_P || (desc->num_obs > 0 && (!desc->obs_cam || !desc->obs_uv))) {
    if (err) *err = "povar_create: invalid problem description";
    return POVAR_ERR_INVALID;
  }
  if (desc->num_obs >= (1LL << 31)) {
    if (err) *err = "povar_create: more than 2^31 observations in one shard";
    return POVAR_ERR_UNSUPPORTED;
  }
  const auto t_create = std::chrono::steady_clock::now();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    if (err) *err = "povar_create: no CUDA device (this library has no CPU fallback)";
    return POVAR_ERR_NO_DEVICE;
  }
  Engine* e = new Engine();
  e->opt_ = *opt;
  e->rank_ = comm ? comm->rank : 0;
  e->world_ = comm ? comm->world_size : 1;
  e->device_ = comm ? comm->device : 0;
  if (e->device_ < 0 || e->device_ >= ndev) {
    if (err) *err = "povar_create: device ordinal out of range";
    delete e;
    return POVAR_ERR_INVALID;
  }
  int rc = e->check(cudaSetDevice(e->device_), "cudaSetDevice");
  if (rc == POVAR_OK) {
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, e->device_) == cudaSuccess) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  HostKit kit;
  if (rc == POVAR_OK && take_kit(e->device_, &kit)) {
    e->stream_ = kit.stream;
    for (int i = 0; i < 8; ++i) e->ev_[i] = kit.ev[i];
    e->host_out_ = kit.pinned;
  } else {
    if (rc == POVAR_OK) rc = e->check(cudaStreamCreateWithFlags(&e->stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int i = 0; i < 8 && rc == POVAR_OK; ++i) rc = e->check(cudaEventCreate(&e->ev_[i]), "cudaEventCreate");
    if (rc == POVAR_OK) rc = e->check(cudaHostAlloc(reinterpret_cast<void**>(&e->host_out_), 256, cudaHostAllocDefault), "cudaHostAlloc");
  }
  if (rc == POVAR_OK && e->world_ > 1 && opt->solver_type_step_1 == POVAR_CHOLESKY) {
    // the direct solver's reduced camera system is not sharded (solver/linearizor_sc.cpp:121-128 is serial too)
    rc = e->fail(POVAR_ERR_UNSUPPORTED, "CHOLESKY is single-GPU: create the handle without a communicator");
  }
  if (rc == POVAR_OK && e->world_ > 1 && is_host_id(comm->nccl_id)) {
    // host rendezvous (povar_comm_host_id): no NCCL communicator at all, every reduction over the peer buffers
    e->comm_key_ = comm_key(*comm);
    std::lock_guard<std::mutex> lock(comm_cache_mutex());
    std::string herr;
    e->rdv_ = open_rendezvous(comm->nccl_id, &herr);
    if (!e->rdv_) rc = e->fail(POVAR_ERR_NCCL, herr);
  } else if (rc == POVAR_OK && e->world_ > 1) {
    std::string nerr;
    e->nccl_ = load_nccl(&nerr);
    if (!e->nccl_) {
      rc = e->fail(POVAR_ERR_NCCL, nerr);
    } else {
      // one communicator per (id, rank) and process: the step-1 and step-2 linearizors of a solve, and
      // every later handle made with the same descriptor, share it (ncclCommInitRank accepts an id once)
      const std::string key = comm_key(*comm);
      e->comm_key_ = key;
      std::lock_guard<std::mutex> lock(comm_cache_mutex());
      auto& cache = comm_cache();
      auto it = cache.find(key);
      if (it != cache.end()) {
        e->nccl_comm_ = it->second;
      } else {
        Id128 id;
        std::memcpy(id.bytes, comm->nccl_id, 128);
        const int nrc = e->nccl_->CommInitRank(&e->nccl_comm_, e->world_, id, e->rank_);
        if (nrc != 0) {
          e->nccl_comm_ = nullptr;
          rc = e->fail(POVAR_ERR_NCCL, "ncclCommInitRank failed");
        } else {
          cache[key] = e->nccl_comm_;
        }
      }
    }
  }
  if (getenv("POVAR_TRACE_CREATE") != nullptr) {
    std::fprintf(stderr, "povar_create: %-28s %8.3f ms\n", "device, stream, scratch",
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_create).count());
  }
  if (rc == POVAR_OK) rc = e->upload(desc);
  if (rc == POVAR_OK) rc = e->setup_peer_exchange();
  if (rc != POVAR_OK) {
    if (err) *err = e->err_;
    delete e;
    return rc;
  }
  *out = e;
  return POVAR_OK;
}

Engine::~Engine() {
  static const bool trace = getenv("POVAR_TRACE_CREATE") != nullptr;   // phase times of povar_destroy on stderr
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "povar_destroy: %-27s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  if (device_ >= 0) cudaSetDevice(device_);
  if (stream_) cudaStreamSynchronize(stream_);
  lap("synchronize");
  // the communicator belongs to the process-wide cache (povar_comm_finalize releases it)
  for (void*& g : series_graph_) {
    if (g) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g));
    g = nullptr;
  }
  lap("graphs");
  if (peer_owned_ && peer_) {
    peer_->release();
    delete peer_;
  }
  for (void* p : allocs_) cudaFreeAsync(p, stream_);
  lap("free (enqueue)");
  bool clean = stream_ != nullptr && host_out_ != nullptr && cudaStreamSynchronize(stream_) == cudaSuccess;
  for (auto& ev : ev_) clean = clean && ev != nullptr;
  lap("free (synchronize)");
  if (clean) {
    // stream, events and scratch wait for the next handle on this device (HostKit)
    HostKit kit;
    kit.device = device_;
    kit.stream = stream_;
    for (int i = 0; i < 8; ++i) kit.ev[i] = ev_[i];
    kit.pinned = host_out_;
    give_kit(kit);
  } else {
    if (host_out_) cudaFreeHost(host_out_);
    for (auto& ev : ev_) {
      if (ev) cudaEventDestroy(ev);
    }
    if (stream_) cudaStreamDestroy(stream_);
  }
  lap("stream, events, scratch");
}

int Engine::upload(const povar_problem_desc* desc) {
  const int C = desc->num_cams, L = desc->num_lms;
  const int nnz = static_cast<int>(desc->num_obs);
  C_ = C;
  L_ = L;
  nnz_ = nnz;
  // ---- host: the landmark table (one pass over lm_ptr) and the small tables; everything per observation --
  // the check of the camera indices and the per-camera counts included -- is made on the device
  // (kernels_index.cu)
  static const bool trace = getenv("POVAR_TRACE_CREATE") != nullptr;   // phase times of povar_create on stderr
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "povar_create: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  std::vector<int> lm_ptr(L + 1), cam_ptr(C + 1, 0);
  SellLayout sell;    // sliced-ELL order: the long landmarks from the pass below, the slices from the device
  int sell_n = 0;     // landmarks with 1..sell_max_deg observations
  const int sell_max_deg = sell_max_degree(nnz, sm_count());
  d_.ix.sell_max_deg = sell_max_deg;
  if (desc->lm_ptr[0] != 0 || desc->lm_ptr[L] != nnz) return fail(POVAR_ERR_INVALID, "lm_ptr does not span the observations");
  // (one thread: a million trivial iterations cost about a millisecond, less than starting threads does on a
  // virtualised host -- the threaded version of this pass took 6 to 11 ms on the bench boxes)
  for (int l = 0; l < L; ++l) {
    const int64_t b = desc->lm_ptr[l], e = desc->lm_ptr[l + 1];
    if (e < b || b < 0 || e > nnz) return fail(POVAR_ERR_INVALID, "lm_ptr is not monotone");
    lm_ptr[l] = static_cast<int>(b);
    if (e - b > sell_max_deg) sell.long_lms.push_back(l);
    else if (e > b) ++sell_n;
  }
  lm_ptr[L] = nnz;
  lap("landmark table");
  // the observation list starts travelling now: camera indices first (the device checks and counts them while
  // the image coordinates follow), and the copies run while the host derives the work items below
#define PV_UP(dst, src, n) PV_CUDA(cudaMemcpyAsync((dst), (src), (n), cudaMemcpyHostToDevice, stream_))
  DeviceIndex& ix = d_.ix;
  PV_ALLOC(ix.lm_ptr, L + 1);
  PV_ALLOC(ix.obs_cam, nnz);
  PV_ALLOC(ix.obs_lm, nnz);
  PV_ALLOC(ix.obs_uv, nnz);
  PV_UP(ix.lm_ptr, lm_ptr.data(), sizeof(int) * (L + 1));
  if (nnz > 0) PV_UP(ix.obs_cam, desc->obs_cam, sizeof(int) * static_cast<size_t>(nnz));
  {
    int* d_count = nullptr;
    const size_t keep = allocs_.size();
    PV_ALLOC(d_count, static_cast<size_t>(C) + 1);
    PV_CUDA(validate_obs(L, C, sm_count(), ix.lm_ptr, ix.obs_cam, d_count, lc()));
    std::vector<int> count(static_cast<size_t>(C) + 1);
    PV_CUDA(cudaMemcpyAsync(count.data(), d_count, sizeof(int) * count.size(), cudaMemcpyDeviceToHost, stream_));
    cudaEvent_t counted = nullptr;
    PV_CUDA(cudaEventCreateWithFlags(&counted, cudaEventDisableTiming));
    cudaEventRecord(counted, stream_);
    if (nnz > 0) {
      const cudaError_t ce = cudaMemcpyAsync(ix.obs_uv, desc->obs_uv, sizeof(double) * 2 * static_cast<size_t>(nnz),
                                             cudaMemcpyHostToDevice, stream_);
      if (ce != cudaSuccess) {
        cudaEventDestroy(counted);
        return fail(POVAR_ERR_CUDA, cudaGetErrorString(ce));
      }
    }
    const cudaError_t we = cudaEventSynchronize(counted);
    cudaEventDestroy(counted);
    if (we != cudaSuccess) return fail(POVAR_ERR_CUDA, cudaGetErrorString(we));
    while (allocs_.size() > keep) {
      cudaFreeAsync(allocs_.back(), stream_);
      allocs_.pop_back();
    }
    if (count[C] & 1) return fail(POVAR_ERR_INVALID, "camera index out of range");
    if (count[C] & 2) return fail(POVAR_ERR_INVALID, "observations of a landmark must have strictly ascending camera indices");
    for (int c = 0; c < C; ++c) cam_ptr[c + 1] = cam_ptr[c] + count[c];
    if (cam_ptr[C] != nnz) return fail(POVAR_ERR_INVALID, "camera counts do not add up to the observations");
  }
  lap("camera check + counts (device)");
  // the sliced-ELL order of the landmarks (rule: build_sell above, which the CPU tests check) on the device, as soon
  // as the copies above have landed; the host derives the camera work items meanwhile
  const int sell_slices = (sell_n + kSellWidth - 1) / kSellWidth;
  const int window = sell_window(sell_n, kSellWindow);
  std::vector<int> slice_len(static_cast<size_t>(sell_slices)), slice_lo(slice_len.size()), slice_hi(slice_len.size());
  PV_ALLOC(ix.sell_lm, static_cast<size_t>(sell_slices) * kSellWidth);
  if (sell_n > 0) {
    int *keys_a = nullptr, *keys_b = nullptr, *ids_a = nullptr, *ids_b = nullptr, *d_len = nullptr, *d_lo = nullptr,
        *d_hi = nullptr;
    char* sort_temp = nullptr;
    const size_t temp_bytes = sell_sort_temp_bytes(L, C, sell_n, window);
    const size_t keep = allocs_.size();
    PV_ALLOC(keys_a, L);
    PV_ALLOC(keys_b, L);
    PV_ALLOC(ids_a, L);
    PV_ALLOC(ids_b, L);
    PV_ALLOC(d_len, sell_slices);
    PV_ALLOC(d_lo, sell_slices);
    PV_ALLOC(d_hi, sell_slices);
    PV_ALLOC(sort_temp, temp_bytes);
    PV_CUDA(build_device_sell(L, C, sell_n, window, sell_max_deg, ix.lm_ptr, ix.obs_cam, keys_a, keys_b, ids_a, ids_b,
                              sort_temp, temp_bytes, ix.sell_lm, d_len, d_lo, d_hi, lc()));
    PV_CUDA(cudaMemcpyAsync(slice_len.data(), d_len, sizeof(int) * slice_len.size(), cudaMemcpyDeviceToHost, stream_));
    PV_CUDA(cudaMemcpyAsync(slice_lo.data(), d_lo, sizeof(int) * slice_lo.size(), cudaMemcpyDeviceToHost, stream_));
    PV_CUDA(cudaMemcpyAsync(slice_hi.data(), d_hi, sizeof(int) * slice_hi.size(), cudaMemcpyDeviceToHost, stream_));
    while (allocs_.size() > keep) {
      cudaFreeAsync(allocs_.back(), stream_);
      allocs_.pop_back();
    }
  }
  std::vector<int> item_ptr, item_cam, cam_item_ptr;
  build_items(cam_ptr, choose_item_len(nnz), &item_ptr, &item_cam, &cam_item_ptr);
  lap("items (+ device: sliced-ELL order)");
  PV_CUDA(cudaStreamSynchronize(stream_));
  sell.slice_ptr.assign(static_cast<size_t>(sell_slices) + 1, 0);
  for (int sl = 0; sl < sell_slices; ++sl) sell.slice_ptr[sl + 1] = sell.slice_ptr[sl] + slice_len[sl];
  sell.rows = sell.slice_ptr[sell_slices];
  sell.slice_lo.swap(slice_lo);
  sell.slice_hi.swap(slice_hi);
  lap("wait for the order");
  if (static_cast<long long>(sell.rows) * kSellWidth >= (1LL << 31)) {
    return fail(POVAR_ERR_UNSUPPORTED, "sliced-ELL layout exceeds 2^31 slots in one shard");
  }

  ix.C = C;
  ix.L = L;
  ix.nnz = nnz;
  ix.num_items = static_cast<int>(item_cam.size());
  ix.num_slices = static_cast<int>(sell.slice_ptr.size()) - 1;
  ix.sell_slots = static_cast<long long>(kSellWidth) * sell.rows;
  ix.num_long = static_cast<int>(sell.long_lms.size());
  const size_t slots = static_cast<size_t>(ix.sell_slots);
  PV_ALLOC(ix.cam_ptr, C + 1);
  PV_ALLOC(ix.csc_lm, nnz);
  PV_ALLOC(ix.csc_uv, nnz);
  PV_ALLOC(ix.item_cam, item_cam.size());
  PV_ALLOC(ix.item_ptr, item_ptr.size());
  PV_ALLOC(ix.cam_item_ptr, C + 1);
  PV_ALLOC(ix.slice_ptr, sell.slice_ptr.size());
  PV_ALLOC(ix.sell_cam, slots);
  PV_ALLOC(ix.sell_uv, slots);
  PV_ALLOC(ix.sell_cam_e0, slots);
  PV_ALLOC(ix.sell_uv_e0, slots);
  PV_ALLOC(ix.sell_row_e0, slots + 32);
  PV_ALLOC(ix.obs_slot, nnz);
  PV_ALLOC(ix.long_lm, sell.long_lms.size());
  if (nnz > 0) PV_UP(ix.item_cam, item_cam.data(), sizeof(int) * item_cam.size());
  PV_UP(ix.slice_ptr, sell.slice_ptr.data(), sizeof(int) * sell.slice_ptr.size());
  // landmark half: ranges of slices per warp, windows of cameras per block (kernels_series.cu), per model
  LmPlanHost lm_plans[5];   // alive until the synchronisation at the end of this function
  for (int m = 0; m < 5; ++m) {
    // [0..2] landmark half of a term (pOSE, joint, pOSE + HUBER); [3], [4] once-per-trial walks: more registers
    // per lane (16 warps per SM), no long landmarks, 256 bytes of static shared memory for the block sums
    const int rec_bytes = 8 * (m == 1 ? kCamRecJoint : m == 3 ? kCamTab1 : m == 4 ? kCamTab2 : kCamRecPose);
    const int stage_bytes = (m == 1 || m == 2) ? kStageWide : (m == 3 ? kStageLin : kStagePose);
    lm_plans[m] = m < 3 ? plan_landmark_half(sell, C, ix.num_long, rec_bytes, stage_bytes, sm_count(), 32, 0)
                        : plan_landmark_half(sell, C, 0, rec_bytes, stage_bytes, sm_count(), 16, 256);
    const LmPlanHost& h = lm_plans[m];
    d_.plan[m] = h.p;
    PV_ALLOC(d_.plan[m].range_slice, h.range_slice.size());
    PV_ALLOC(d_.plan[m].blk_lo, h.blk_lo.size());
    PV_UP(d_.plan[m].range_slice, h.range_slice.data(), sizeof(int) * h.range_slice.size());
    PV_UP(d_.plan[m].blk_lo, h.blk_lo.data(), sizeof(int) * h.blk_lo.size());
  }
  if (ix.num_long > 0) PV_UP(ix.long_lm, sell.long_lms.data(), sizeof(int) * sell.long_lms.size());
  PV_UP(ix.cam_ptr, cam_ptr.data(), sizeof(int) * (C + 1));
  PV_UP(ix.item_ptr, item_ptr.data(), sizeof(int) * item_ptr.size());
  PV_UP(ix.cam_item_ptr, cam_item_ptr.data(), sizeof(int) * (C + 1));
  {
    // scratch of the device-side build, returned to the pool right after
    int *iota = nullptr, *keys_out = nullptr, *perm = nullptr, *lm_slot = nullptr;
    char* sort_temp = nullptr;
    const size_t temp_bytes = index_sort_temp_bytes(nnz, C);
    const size_t keep = allocs_.size();
    PV_ALLOC(iota, nnz);
    PV_ALLOC(keys_out, nnz);
    PV_ALLOC(perm, nnz);
    PV_ALLOC(lm_slot, L);
    PV_ALLOC(sort_temp, temp_bytes);
    PV_CUDA(build_device_index(ix, iota, keys_out, perm, lm_slot, sort_temp, temp_bytes, lc()));
    while (allocs_.size() > keep) {
      cudaFreeAsync(allocs_.back(), stream_);
      allocs_.pop_back();
    }
  }
  // ---- state and work arrays
  const size_t C12 = static_cast<size_t>(C) * 12, C144 = static_cast<size_t>(C) * 144;
  PV_ALLOC(d_.P, C12);
  PV_ALLOC(d_.P_bak, C12);
  PV_ALLOC(P_prev_, C12);
  PV_ALLOC(d_.X, static_cast<size_t>(L) * 4);
  PV_ALLOC(d_.X_bak, static_cast<size_t>(L) * 4);
  PV_ALLOC(d_.pose_scale, C12);
  PV_ALLOC(d_.lm_scale, static_cast<size_t>(L) * 4);
  PV_ALLOC(d_.lm_hraw, static_cast<size_t>(L) * 10);
  PV_ALLOC(d_.lm_graw, static_cast<size_t>(L) * 4);
  PV_ALLOC(d_.hll_inv, static_cast<size_t>(L) * 6);
  PV_ALLOC(d_.lm_rec, static_cast<size_t>(L) * kLmRec);
  PV_ALLOC(d_.lm_fold, static_cast<size_t>(L) * 10);
  PV_ALLOC(d_.cam_rec, static_cast<size_t>(C) * kCamRecStride);
  PV_ALLOC(d_.cam_tab, static_cast<size_t>(C) * kCamTab2);
  {
    const size_t plane = static_cast<size_t>(ix.num_slices) * kSellWidth;
    PV_ALLOC(d_.sell_hraw, plane * 10);
    PV_ALLOC(d_.sell_graw, plane * 4);
    PV_ALLOC(d_.sell_scale, plane * 4);
    PV_ALLOC(d_.sell_hinv, plane * 6);
    PV_ALLOC(d_.sell_step, plane * 4);
  }
  PV_ALLOC(d_.sell_x, static_cast<size_t>(ix.num_slices) * 4 * kSellWidth);
  PV_ALLOC(d_.sell_fold, static_cast<size_t>(ix.num_slices) * 10 * kSellWidth);
  PV_ALLOC(d_.kron, static_cast<size_t>(C) * kKron);
  PV_ALLOC(d_.kron2, static_cast<size_t>(C) * kKron);
  PV_ALLOC(d_.item_kron, static_cast<size_t>(ix.num_items) * kKron);
  PV_ALLOC(d_.item_part, static_cast<size_t>(ix.num_items) * 12);
  PV_ALLOC(d_.cam_raw, C12);
  PV_ALLOC(d_.Bmat, C144);
  PV_ALLOC(d_.Binv, C144);
  PV_ALLOC(d_.Mprec, C144);
  PV_ALLOC(d_.b, C12);
  PV_ALLOC(d_.vec_tmp, C12);
  PV_ALLOC(d_.vec_acc, C12);
  PV_ALLOC(d_.vec_y, C12);
  PV_ALLOC(d_.vec_x, C12);
  PV_ALLOC(d_.cg_r, C12);
  PV_ALLOC(d_.cg_p, C12);
  PV_ALLOC(d_.cg_z, C12);
  PV_ALLOC(d_.cg_q, C12);
  PV_ALLOC(d_.cg_x, C12);
  PV_ALLOC(d_.norm_part, static_cast<size_t>(C) * 4);
  PV_ALLOC(d_.cost_part, static_cast<size_t>(cost_blocks(d_)) * 8);
  PV_ALLOC(d_.cost_out, 1);
  PV_ALLOC(d_.scalar_part, static_cast<size_t>(scalar_blocks(d_)) + 8);
  PV_ALLOC(d_.scalar_out, 8);
  PV_ALLOC(d_.trial_out, 16);
  PV_ALLOC(d_.flags, 4);
  PV_ALLOC(d_.ctl, 1);
  PV_ALLOC(d_.cg, 1);
  PV_UP(d_.P, desc->cam_P, sizeof(double) * C12);
#undef PV_UP
  lap("allocate + enqueue uploads");
  // the host tables above live on this stack frame
  PV_CUDA(cudaStreamSynchronize(stream_));
  lap("device index build + sync");
  return POVAR_OK;
}

// Peer-memory exchange for the per-term camera sums: one cudaMalloc per rank (receive buffer + exchange
// counter), exported with CUDA IPC, handles swapped through the NCCL communicator (or the host rendezvous),
// mapped with peer access over NVLink.  Every rank takes the same decisions (the inputs of each decision are
// all-reduced), because a rank that fell back to NCCL alone would deadlock the others.  The mapping is made
// once per (communicator, camera count) and process -- like the communicator itself -- and shared by later
// handles: ranks create their handles in the same order, so cache hits are symmetric.
//
// The number of an exchange -- its tag and the parity of the slots it uses -- is a DEVICE-side counter that
// lives next to the receive buffer and is advanced by the kernel that performed the exchange (k_term16,
// k_peer_allreduce): terms that are enqueued but skipped after the series converged do not count, every rank
// performs the same sequence of exchanges, and the count survives the handle (the slots keep their tags).
static constexpr size_t kPeerCounterBytes = 256;
int Engine::sum_setup_words(unsigned int* host, int words, int phase, int round) {
  if (rdv_) {
    if (!rendezvous_sum(rdv_, round, phase, rank_, world_, host, words)) {
      return fail(POVAR_ERR_NCCL, "host rendezvous: a rank did not arrive (peer exchange setup)");
    }
    return POVAR_OK;
  }
  unsigned int* dev = nullptr;
  PV_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dev), sizeof(unsigned int) * words, stream_));
  PV_CUDA(cudaMemcpyAsync(dev, host, sizeof(unsigned int) * words, cudaMemcpyHostToDevice, stream_));
  if (nccl_->AllReduce(dev, dev, words, /*ncclUint32*/ 3, /*ncclSum*/ 0, nccl_comm_, stream_) != 0) {
    return fail(POVAR_ERR_NCCL, "ncclAllReduce failed (peer exchange setup)");
  }
  PV_CUDA(cudaMemcpyAsync(host, dev, sizeof(unsigned int) * words, cudaMemcpyDeviceToHost, stream_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  PV_CUDA(cudaFreeAsync(dev, stream_));
  return POVAR_OK;
}

int Engine::setup_peer_exchange() {
  const char* env = getenv("POVAR_PEER_EXCHANGE");
  const bool forbid = env != nullptr && std::strcmp(env, "0") == 0;
  if (forbid && rdv_) return fail(POVAR_ERR_NCCL, "the host rendezvous needs the peer exchange (POVAR_PEER_EXCHANGE=0 set)");
  if (forbid) return POVAR_OK;
  if (world_ > kMaxPeers || C_ <= 0) {
    if (rdv_) return fail(POVAR_ERR_UNSUPPORTED, "host rendezvous: more ranks than the peer exchange supports");
    return POVAR_OK;
  }
  if (world_ == 1) {
    // POVAR_PEER_EXCHANGE=self: a single GPU runs the exchange protocol against its own buffer (isolates the
    // protocol's cost from the NVLink hop)
    if (env == nullptr || std::strcmp(env, "self") != 0) return POVAR_OK;
    PeerShared* ps = new PeerShared();
    const size_t bytes = 16 * 2 * static_cast<size_t>(C_) * 12;
    if (cudaMalloc(&ps->mem, bytes + kPeerCounterBytes) != cudaSuccess ||
        cudaMemset(ps->mem, 0, bytes + kPeerCounterBytes) != cudaSuccess) {
      delete ps;
      return fail(POVAR_ERR_CUDA, "cudaMalloc failed for the self exchange buffer");
    }
    ps->px.recv[0] = static_cast<double*>(ps->mem);
    ps->px.rank = 0;
    ps->px.world = 1;
    ps->px.stride = static_cast<unsigned long long>(C_) * 12;
    ps->px.count = reinterpret_cast<unsigned int*>(static_cast<char*>(ps->mem) + bytes);
    ps->ok = true;
    peer_ = ps;
    peer_owned_ = true;
    peer_ok_ = true;
    return POVAR_OK;
  }
  if (env != nullptr && std::strcmp(env, "self") == 0) env = nullptr;
  const bool must = env != nullptr || rdv_ != nullptr;   // asked for explicitly, or no NCCL to fall back to
  const std::string key = comm_key_ + ":" + std::to_string(C_);
  std::lock_guard<std::mutex> lock(comm_cache_mutex());
  auto& cache = peer_cache();
  auto it = cache.find(key);
  if (it != cache.end()) {
    peer_ = it->second;
    peer_ok_ = peer_->ok;
    if (!peer_ok_ && must) {
      return fail(POVAR_ERR_NCCL, "peer exchange requested but CUDA IPC / peer access is unavailable");
    }
    return POVAR_OK;
  }
  PeerShared* ps = new PeerShared();
  cache[key] = ps;
  peer_ = ps;
  const int round = rdv_ ? rdv_->next_round++ : 0;
  // 16-byte tagged slots, [2 parities][world][60*C]; zero = "exchange 0", never waited for
  const size_t stride = static_cast<size_t>(C_) * kKron;
  const size_t recv_pad = 16 * 2 * static_cast<size_t>(world_) * stride;
  unsigned int ok = 1;
  cudaIpcMemHandle_t mine{};
  if (cudaMalloc(&ps->mem, recv_pad + kPeerCounterBytes) != cudaSuccess) {
    ps->mem = nullptr;
    ok = 0;
  }
  if (ok && cudaMemsetAsync(ps->mem, 0, recv_pad + kPeerCounterBytes, stream_) != cudaSuccess) ok = 0;
  if (ok && cudaStreamSynchronize(stream_) != cudaSuccess) ok = 0;   // zeroed before any peer can store into it
  if (ok && cudaIpcGetMemHandle(&mine, ps->mem) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // swap: [world][16 words of handle] + [world] ok flags, summed as uint32 (everything else is zero)
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  const int words = world_ * 17;
  std::vector<unsigned int> host(words, 0u);
  if (ok) std::memcpy(host.data() + 16 * rank_, &mine, 64);
  host[16 * world_ + rank_] = ok;
  {
    const int rc = sum_setup_words(host.data(), words, 0, round);
    if (rc != POVAR_OK) return rc;
  }
  unsigned int all_ok = 1;
  for (int r = 0; r < world_; ++r) all_ok &= host[16 * world_ + r];
  std::vector<void*> base(world_, nullptr);
  if (all_ok) {
    for (int r = 0; r < world_; ++r) {
      if (r == rank_) {
        base[r] = ps->mem;
        continue;
      }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, host.data() + 16 * r, 64);
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        all_ok = 0;
        break;
      }
      ps->opened.push_back(ptr);
      base[r] = ptr;
    }
  }
  // second round: did everybody map everybody?
  unsigned int mapped = all_ok;
  {
    const int rc = sum_setup_words(&mapped, 1, 1, round);
    if (rc != POVAR_OK) return rc;
  }
  if (mapped != static_cast<unsigned int>(world_)) {
    ps->release();   // the entry stays, marked unavailable: later handles do not try again
    if (must) {   // asked for explicitly: do not hide it
      return fail(POVAR_ERR_NCCL, "peer exchange requested but CUDA IPC / peer access is unavailable");
    }
    return POVAR_OK;   // ncclAllReduce per term instead
  }
  for (int r = 0; r < world_; ++r) {
    ps->px.recv[r] = static_cast<double*>(base[r]);
  }
  ps->px.rank = rank_;
  ps->px.world = world_;
  ps->px.stride = stride;
  ps->px.count = reinterpret_cast<unsigned int*>(static_cast<char*>(ps->mem) + recv_pad);
  ps->ok = true;
  peer_ok_ = true;
  return POVAR_OK;
}

const PeerExchange* Engine::exchange() const { return peer_ok_ ? &peer_->px : nullptr; }

int Engine::allreduce(double* buf, size_t n, bool skip_when_done) {
  if (world_ <= 1) return POVAR_OK;
  if (peer_ok_ && n <= peer_->px.stride && peer_small_) {
    launch_peer_allreduce(d_, buf, n, peer_->px, lc(), skip_when_done);
    return POVAR_OK;
  }
  if (!nccl_comm_) return fail(POVAR_ERR_NCCL, "reduction needs an NCCL communicator (host rendezvous handle)");
  // ncclDouble = 8, ncclSum = 0
  const int rc = nccl_->AllReduce(buf, buf, n, 8, 0, nccl_comm_, stream_);
  if (rc != 0) return fail(POVAR_ERR_NCCL, "ncclAllReduce failed");
  return POVAR_OK;
}

void Engine::set_model(bool joint, double alpha) {
  mp_.c1 = std::sqrt(1.0 - alpha);
  mp_.c2 = std::sqrt(alpha);
  mp_.robust_norm = opt_.robust_norm;
  mp_.huber = opt_.huber_parameter;
  mp_.jacobi_eps = opt_.jacobi_scaling_epsilon > 0 ? opt_.jacobi_scaling_epsilon : kEpsSqrtHost;
  (void)joint;
}

// ---------------------------------------------------------------------------------------------
int Engine::init_varproj(double alpha) {
  PV_CUDA(cudaSetDevice(device_));
  set_model(false, alpha);
  launch_init_varproj(d_, mp_, lc());
  PV_CUDA(cudaGetLastError());
  return POVAR_OK;
}

// cost kernels + (sharded) the in-place sum over the ranks: the result is trial_out[0..7] on the device
int Engine::enqueue_cost(bool joint, double alpha, bool reduce) {
  set_model(joint, joint ? opt_.alpha : alpha);
  launch_cost(d_, mp_, joint, lc());
  return reduce ? allreduce(d_.trial_out, 8) : POVAR_OK;
}

void Engine::decode_cost(const double* v, povar_residual_info* out) {
  out->error_all = v[0];
  out->residual_sum_all = v[1];
  out->error_valid = v[2];
  out->residual_sum_valid = v[3];
  out->num_obs_all = static_cast<long long>(v[4] + 0.5);
  out->num_obs_valid = static_cast<long long>(v[5] + 0.5);
  out->is_numerically_valid = v[6] > 0 ? 0 : 1;
}

int Engine::cost(bool joint, double alpha, povar_residual_info* out) {
  PV_CUDA(cudaSetDevice(device_));
  PV_CUDA(cudaEventRecord(ev_[0], stream_));
  {
    const int rc = enqueue_cost(joint, alpha);
    if (rc != POVAR_OK) return rc;
  }
  PV_CUDA(cudaGetLastError());
  double* v = host_out_ + 8;
  PV_CUDA(cudaMemcpyAsync(v, d_.trial_out, 8 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  PV_CUDA(cudaEventRecord(ev_[1], stream_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  times_.residual += elapsed(ev_[0], ev_[1]);
  decode_cost(v, out);
  return POVAR_OK;
}

int Engine::linearize(bool joint, double alpha, bool defer_check) {
  PV_CUDA(cudaSetDevice(device_));
  set_model(joint, joint ? opt_.alpha : alpha);
  joint_lin_ = joint ? 1 : 0;
  have_solve_ = false;
  dim_ = joint ? 11 : 12;
  const bool power = joint ? true
                           : (opt_.solver_type_step_1 == POVAR_POWER_VARPROJ ||
                              opt_.solver_type_step_1 == POVAR_POWER_SCHUR_COMPLEMENT);
  // per-observation coefficients the term kernels stream (kernels_series.cu): first use allocates
  if (joint && !d_.obs_d) {
    PV_ALLOC(d_.obs_d, static_cast<size_t>(nnz_) * 3);
    PV_ALLOC(d_.csc_d, static_cast<size_t>(nnz_) * 3);
    PV_ALLOC(d_.sell_d, static_cast<size_t>(d_.ix.sell_slots) * 3);
  }
  if (!joint && opt_.robust_norm == POVAR_NORM_HUBER && !d_.obs_w) {
    PV_ALLOC(d_.obs_w, static_cast<size_t>(nnz_));
    PV_ALLOC(d_.csc_w, static_cast<size_t>(nnz_));
    PV_ALLOC(d_.sell_w, static_cast<size_t>(d_.ix.sell_slots));
  }
  PV_CUDA(cudaEventRecord(ev_[6], stream_));
  PV_CUDA(cudaMemsetAsync(d_.flags, 0, 4 * sizeof(int), stream_));
  // landmark side: Jl^T Jl, Jl^T r, column scales (LinearizorSC step 1 does not scale Jl:
  // solver/linearizor_sc.cpp:174-203)
  launch_lin_landmark(d_, mp_, joint, power, lc());
  // camera side: Jp^T Jp as Kronecker sums, then the pose scales
  launch_kron(d_, mp_, joint, KRON_HPP, lc());
  launch_reduce_items(d_, d_.item_kron, kKron, d_.kron, false, lc());
  {
    const int rc = allreduce(d_.kron, static_cast<size_t>(C_) * kKron);
    if (rc != POVAR_OK) return rc;
  }
  launch_cam_scale(d_, mp_, lc());
  launch_cam_rec_static(d_, joint, lc());
  // numerical-failure flag of this shard as a double, summed over the ranks in place
  launch_flag_to_double(d_, lc());
  {
    const int rc = allreduce(d_.trial_out + 9, 1);
    if (rc != POVAR_OK) return rc;
  }
  PV_CUDA(cudaEventRecord(ev_[7], stream_));
  PV_CUDA(cudaGetLastError());
  if (defer_check) {
    // the caller goes on to a trial(): the flag comes back with that trial's results (one synchronisation)
    lin_check_pending_ = true;
    return POVAR_OK;
  }
  double* v = host_out_ + 8;
  PV_CUDA(cudaMemcpyAsync(v + 9, d_.trial_out + 9, sizeof(double), cudaMemcpyDeviceToHost, stream_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  times_.linearize += elapsed(ev_[6], ev_[7]);
  if (v[9] > 0) return fail(POVAR_NUM_LINEARIZATION, "did not expect numerical failure during linearization");
  return POVAR_OK;
}

// raw_c = sum over the observations of camera c of the camera half of E0 (or of b), all ranks.
// skip_when_done: the launches return at once when ctl->done is set (inside a power series or a CG solve).
// fused_reduce: the caller's term kernel adds the item partials (and exchanges them) itself.
int Engine::e0_product(bool joint, bool skip_when_done, bool fused_reduce) {
  // the passes gather y from the camera records (cam_rec), written by whoever made y
  launch_e0_landmark_v2(d_, mp_, joint, skip_when_done, lc());
  launch_passB_e0_v2(d_, mp_, joint, skip_when_done, lc());
  if (fused_reduce) return POVAR_OK;
  launch_reduce_items(d_, d_.item_part, 12, d_.cam_raw, skip_when_done, lc());
  return allreduce(d_.cam_raw, static_cast<size_t>(C_) * 12, skip_when_done);
}

int Engine::solve_power(bool joint, double lambda) {
  const bool poba = !joint && opt_.solver_type_step_1 == POVAR_POWER_SCHUR_COMPLEMENT;
  const double lambda_lm = joint ? lambda : (poba ? lambda : 0.0);
  PV_CUDA(cudaEventRecord(ev_[0], stream_));
  // prepare_Hb_*: Hll^-1 and Hll^-1 Jl^T r per landmark; B^-1 per camera; b
  launch_prep_landmark(d_, joint, lambda_lm, false, lc());
  launch_cam_binv(d_, joint, lambda, lc());
  launch_passB(d_, mp_, joint, lc());
  launch_reduce_items(d_, d_.item_part, 12, d_.cam_raw, false, lc());
  {
    const int rc = allreduce(d_.cam_raw, static_cast<size_t>(C_) * 12);
    if (rc != POVAR_OK) return rc;
  }
  PV_CUDA(cudaEventRecord(ev_[1], stream_));
  // solve_*: accum = B^-1(-b); tmp = B^-1 E0 tmp; accum += tmp; early exit on zeta < eta
  {
    const int rc = enqueue_series(joint);
    if (rc != POVAR_OK) return rc;
  }
  PV_CUDA(cudaEventRecord(ev_[2], stream_));
  PV_CUDA(cudaGetLastError());
  return POVAR_OK;
}

// One power series is b -> x0, series start, then per term landmark half, camera half, term kernel -- the same
// arguments in every solve of a model (the damping enters before, the stopping rule is applied on the device, the
// number of a peer exchange and of a term are device-side counters).  It is built once per model as a CUDA graph
// with a conditional WHILE node whose body is one term: k_series_start and the last block of k_term16 set the
// condition (cudaGraphSetConditional), so a solve is ONE launch that runs exactly the terms the reference's
// stopping rule asks for.  (Until this round the graph held all m terms and the converged ones returned at once:
// 430 skipped launches of 2-4 us per venice-1778 solve.)  The first solve of a model runs eagerly -- it sets
// function attributes -- and so does a series whose exchange is ncclAllReduce (no peer memory).
int Engine::enqueue_series(bool joint) {
  const int m = opt_.power_sc_iterations;
  const int which = joint ? 1 : 0;
  const bool graphable = term_mode() != kTermRaw;
  if (!graphable || series_calls_[which]++ == 0) {
    launch_finish_b(d_, joint, lc());
    launch_series_start(d_, opt_.r_tolerance, m, lc());
    for (int i = 1; i <= m; ++i) {
      const int rc = e0_product(joint, true, term_mode() != kTermRaw);
      if (rc != POVAR_OK) return rc;
      launch_series_term(d_, joint, i, opt_.eta, opt_.r_tolerance, term_mode(), exchange(), lc());
    }
    series_terms_counted_ = true;   // the launches above are in launches_ (skipped ones included)
    return POVAR_OK;
  }
  if (!series_graph_[which]) {
    const long long before = launches_;
    cudaGraph_t graph = nullptr;
    cudaGraphConditionalHandle handle = 0;
    PV_CUDA(cudaGraphCreate(&graph, 0));
    auto bail = [&](cudaError_t e, const char* what) {
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(stream_, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
        cudaGraph_t dummy = nullptr;
        cudaStreamEndCapture(stream_, &dummy);
      }
      cudaGraphDestroy(graph);
      cudaGetLastError();
      launches_ = before;
      return e != cudaSuccess ? check(e, what) : fail(POVAR_ERR_CUDA, what);
    };
    cudaError_t e = cudaGraphConditionalHandleCreate(&handle, graph, 0, 0);   // set by k_series_start in every run
    if (e != cudaSuccess) return bail(e, "cudaGraphConditionalHandleCreate");
    SeriesLoop loop;
    loop.handle = static_cast<unsigned long long>(handle);
    loop.active = 1;
    loop.max_terms = m;
    // prefix: b -> x0, norms, loop condition
    e = cudaStreamBeginCaptureToGraph(stream_, graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return bail(e, "cudaStreamBeginCaptureToGraph");
    launch_finish_b(d_, joint, lc());
    launch_series_start(d_, opt_.r_tolerance, m, lc(), loop);
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    const cudaGraphNode_t* deps = nullptr;
    size_t num_deps = 0;
    e = cudaStreamGetCaptureInfo(stream_, &status, nullptr, nullptr, &deps, &num_deps);
    if (e != cudaSuccess || status != cudaStreamCaptureStatusActive) return bail(e, "cudaStreamGetCaptureInfo");
    cudaGraphNodeParams params{};
    params.type = cudaGraphNodeTypeConditional;
    params.conditional.handle = handle;
    params.conditional.type = cudaGraphCondTypeWhile;
    params.conditional.size = 1;
    cudaGraphNode_t loop_node = nullptr;
    e = cudaGraphAddNode(&loop_node, graph, deps, num_deps, &params);
    if (e != cudaSuccess) return bail(e, "cudaGraphAddNode(conditional)");
    e = cudaStreamUpdateCaptureDependencies(stream_, &loop_node, 1, cudaStreamSetCaptureDependencies);
    if (e != cudaSuccess) return bail(e, "cudaStreamUpdateCaptureDependencies");
    cudaGraph_t same = nullptr;
    e = cudaStreamEndCapture(stream_, &same);
    if (e != cudaSuccess) return bail(e, "cudaStreamEndCapture");
    // body: one term
    cudaGraph_t body = params.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(stream_, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return bail(e, "cudaStreamBeginCaptureToGraph(body)");
    const int rc = e0_product(joint, true, true);
    launch_series_term(d_, joint, 0, opt_.eta, opt_.r_tolerance, term_mode(), exchange(), lc(), loop);
    e = cudaStreamEndCapture(stream_, &same);
    if (rc != POVAR_OK || e != cudaSuccess) return bail(e, "stream capture of a power-series term failed");
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      launches_ = before;
      return check(e, "cudaGraphInstantiate");
    }
    series_graph_[which] = exec;
    launches_ = before;   // nothing ran during the capture
  }
  PV_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(series_graph_[which]), stream_));
  launches_ += 2;                  // b -> x0 and the series start; the terms are counted when the solve reports them
  series_terms_counted_ = false;
  return POVAR_OK;
}

int Engine::finish_solve(bool joint, double* inc, int32_t* iterations) {
  SeriesCtl h{};
  PV_CUDA(cudaMemcpyAsync(&h, d_.ctl, sizeof(h), cudaMemcpyDeviceToHost, stream_));
  if (inc) {
    PV_CUDA(cudaMemcpyAsync(inc, d_.vec_acc, sizeof(double) * static_cast<size_t>(C_) * (joint ? 11 : 12),
                            cudaMemcpyDeviceToHost, stream_));
  }
  PV_CUDA(cudaStreamSynchronize(stream_));
  times_.prepare += elapsed(ev_[0], ev_[1]);
  times_.reduced_solve += elapsed(ev_[1], ev_[2]);
  if (iterations) *iterations = h.iterations;
  count_series_terms(h.term);
  if (h.peer_timeout) return fail(POVAR_ERR_NCCL, "peer exchange of the camera sums timed out (a rank is gone?)");
  if (h.nonfinite) return POVAR_NUM_NONFINITE_INC;
  return POVAR_OK;
}

// kernels the loop of the series graph ran: three per term (the host learns the count with the solve's results)
void Engine::count_series_terms(int terms) {
  if (series_terms_counted_) return;
  series_terms_counted_ = true;
  launches_ += 3LL * terms;
}

int Engine::enqueue_solve(bool joint, double lambda) {
  if (static_cast<int>(joint) != joint_lin_) return fail(POVAR_ERR_INVALID, "solve called without a matching linearize");
  have_solve_ = true;
  lambda_ = lambda;
  if (joint) return opt_.solver_type_step_2 == POVAR_RIPOBA ? solve_power(true, lambda) : solve_pcg(true, lambda);
  switch (opt_.solver_type_step_1) {
    case POVAR_POWER_VARPROJ:
    case POVAR_POWER_SCHUR_COMPLEMENT:
      return solve_power(false, lambda);
    case POVAR_PCG:
      return solve_pcg(false, lambda);
    case POVAR_CHOLESKY:
      return solve_cholesky(lambda);
    default:
      return fail(POVAR_ERR_INVALID, "unknown solver_type_step_1");
  }
}

int Engine::solve(bool joint, double lambda, double* inc, int32_t* iterations) {
  PV_CUDA(cudaSetDevice(device_));
  const int rc = enqueue_solve(joint, lambda);
  if (rc != POVAR_OK) return rc;
  return finish_solve(joint, inc, iterations);
}

// b (and B, B^-1) exactly as the power solvers build them; shared by PCG / RIPCG / CHOLESKY
int Engine::prepare_reduced_system(bool joint, double lambda, double lambda_lm) {
  launch_prep_landmark(d_, joint, lambda_lm, true, lc());
  launch_cam_binv(d_, joint, lambda, lc());
  launch_passB(d_, mp_, joint, lc());
  launch_reduce_items(d_, d_.item_part, 12, d_.cam_raw, false, lc());
  const int rc = allreduce(d_.cam_raw, static_cast<size_t>(C_) * 12);
  if (rc != POVAR_OK) return rc;
  launch_finish_b(d_, joint, lc());
  return POVAR_OK;
}

// out = S p = B p - E0 p   (the reduced camera system applied implicitly)
int Engine::schur_product(bool joint, const double* p, double* out, bool skip_when_done) {
  const int n = C_ * (joint ? 11 : 12);
  launch_make_y(d_, joint, p, d_.vec_y, lc());
  const int rc = e0_product(joint, skip_when_done, false);
  if (rc != POVAR_OK) return rc;
  launch_e0_finish(d_, joint, d_.vec_x, lc());
  launch_block_matvec(d_, joint ? 11 : 12, d_.Bmat, p, out, lc());
  launch_axpby(d_, n, 1.0, out, -1.0, d_.vec_x, out, lc());
  return POVAR_OK;
}

// ConjugateGradientsSolver::solve / solve_joint (cg/conjugate_gradient.hpp:114-489) with the block-Jacobi
// preconditioner (cg/preconditioner.hpp:70-144), x0 = 0, r_tolerance = -1, then x = -x
// (solver/linearizor_base.cpp:102-147).  The scalars (rho, beta, alpha, Q) and every termination test of the
// reference's loop -- its NaN behaviour included -- live on the device (k_cg_scalar): the host enqueues
// iterations a few at a time and reads the control block once per batch; after termination the remaining
// launches of a batch return at once.
int Engine::solve_pcg(bool joint, double lambda) {
  const int D = joint ? 11 : 12;
  const int n = C_ * D;
  const size_t bytes = sizeof(double) * static_cast<size_t>(n);
  PV_CUDA(cudaEventRecord(ev_[0], stream_));
  // LinearizorSC step 1 has no landmark damping (landmark_block.hpp:360-374); step 2 has (:414-431)
  int rc = prepare_reduced_system(joint, lambda, joint ? lambda : 0.0);
  if (rc != POVAR_OK) return rc;
  launch_kron(d_, mp_, joint, KRON_SDIAG, lc());
  launch_reduce_items(d_, d_.item_kron, kKron, d_.kron2, false, lc());
  rc = allreduce(d_.kron2, static_cast<size_t>(C_) * kKron);
  if (rc != POVAR_OK) return rc;
  launch_cam_precond(d_, joint, d_.kron2, lc());
  PV_CUDA(cudaEventRecord(ev_[1], stream_));

  const double* b = d_.b;
  double *x = d_.cg_x, *r = d_.cg_r, *p = d_.cg_p, *z = d_.cg_z, *q = d_.cg_q, *tmp = d_.vec_tmp;
  const int min_it = opt_.min_linear_solver_iterations, max_it = opt_.max_linear_solver_iterations;
  PV_CUDA(cudaMemsetAsync(x, 0, bytes, stream_));
  PV_CUDA(cudaMemsetAsync(p, 0, bytes, stream_));
  PV_CUDA(cudaMemcpyAsync(r, b, bytes, cudaMemcpyDeviceToDevice, stream_));   // r = b - S*0
  launch_dot_partials(d_, n, b, b, lc());
  launch_cg_scalar(d_, CG_BEGIN, 0, opt_.eta, min_it, max_it, lc());          // |b| = 0: done at once
  SeriesCtl* hctl = reinterpret_cast<SeriesCtl*>(host_out_);
  constexpr int kBatch = 4;
  for (int it = 1; it <= max_it;) {
    for (int k = 0; k < kBatch && it <= max_it; ++k, ++it) {
      launch_block_matvec(d_, D, d_.Mprec, r, z, lc());                       // z = M^-1 r
      launch_dot_partials(d_, n, r, z, lc());
      launch_cg_scalar(d_, CG_RHO, it, opt_.eta, min_it, max_it, lc());
      launch_cg_update(d_, CG_UPDATE_P, n, it, z, p, lc());                   // p = z + beta p
      rc = schur_product(joint, p, q, true);                                  // q = S p
      if (rc != POVAR_OK) return rc;
      launch_dot_partials(d_, n, p, q, lc());
      launch_cg_scalar(d_, CG_PQ, it, opt_.eta, min_it, max_it, lc());
      launch_cg_update(d_, CG_UPDATE_X, n, it, p, x, lc());                   // x += alpha p
      if (it % 10 == 0) {                                                     // residual_reset_period
        rc = schur_product(joint, x, tmp, true);
        if (rc != POVAR_OK) return rc;
        launch_axpby(d_, n, 1.0, b, -1.0, tmp, r, lc());                      // r = b - S x
      } else {
        launch_cg_update(d_, CG_UPDATE_R, n, it, q, r, lc());                 // r -= alpha q
      }
      launch_axpby(d_, n, 1.0, b, 1.0, r, tmp, lc());
      launch_dot_partials(d_, n, x, tmp, lc());
      launch_cg_scalar(d_, CG_ZETA, it, opt_.eta, min_it, max_it, lc());      // Q, zeta < eta, max iterations
    }
    PV_CUDA(cudaMemcpyAsync(hctl, d_.ctl, sizeof(SeriesCtl), cudaMemcpyDeviceToHost, stream_));
    PV_CUDA(cudaStreamSynchronize(stream_));
    if (hctl->done) break;
  }
  launch_axpby(d_, n, -1.0, x, 0.0, nullptr, d_.vec_acc, lc());   // "negate the pose increment"
  launch_finite_check(d_, n, d_.vec_acc, lc());
  PV_CUDA(cudaEventRecord(ev_[2], stream_));
  PV_CUDA(cudaGetLastError());
  return POVAR_OK;
}

// CHOLESKY, step 1 (solver/linearizor_sc.cpp:121-128, sc/linearization_sc.hpp:236-245): the reference
// factorises the sparse reduced camera system with Eigen::SimplicialLLT; here S is assembled densely without
// atomics and factorised by the hand-written blocked LL^T of kernels_chol.cu (FP64 tensor-core tile updates).
// The solution of S x = -b is unique, so ordering / blocking do not change the result beyond rounding.
int Engine::solve_cholesky(double lambda) {
  const int n = 12 * C_;
  const int n_pad = chol_padded(n);
  if (static_cast<long long>(n_pad) * n_pad * 8 > (96LL << 30)) {
    return fail(POVAR_ERR_UNSUPPORTED, "CHOLESKY: dense reduced system would exceed 96 GB");
  }
  PV_CUDA(cudaEventRecord(ev_[0], stream_));
  int rc = prepare_reduced_system(false, lambda, 0.0);
  if (rc != POVAR_OK) return rc;
  if (!d_.dense_S) {
    PV_ALLOC(d_.dense_S, static_cast<size_t>(n_pad) * n_pad);
    PV_ALLOC(chol_linv_, static_cast<size_t>(n_pad) * 64);
    PV_ALLOC(chol_rhs_, static_cast<size_t>(n_pad));
  }
  PV_CUDA(cudaMemsetAsync(d_.dense_S, 0, sizeof(double) * static_cast<size_t>(n_pad) * n_pad, stream_));
  launch_schur_lower(d_, mp_, d_.dense_S, n_pad, lc());
  PV_CUDA(cudaEventRecord(ev_[1], stream_));
  PV_CUDA(cudaMemsetAsync(chol_rhs_, 0, sizeof(double) * static_cast<size_t>(n_pad), stream_));
  launch_axpby(d_, n, -1.0, d_.b, 0.0, nullptr, chol_rhs_, lc());
  int* info = d_.flags + 2;
  launch_cholesky_factor(d_.dense_S, n_pad, chol_linv_, info, lc());
  launch_cholesky_solve(d_.dense_S, n_pad, chol_linv_, chol_rhs_, info, lc());
  PV_CUDA(cudaMemcpyAsync(d_.vec_acc, chol_rhs_, sizeof(double) * static_cast<size_t>(n), cudaMemcpyDeviceToDevice, stream_));
  int hinfo = 0;
  PV_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  SeriesCtl h{};
  h.done = 1;
  h.iterations = 0;   // the reference reports 0 linear solver iterations for the direct solve
  PV_CUDA(cudaMemcpyAsync(d_.ctl, &h, sizeof(h), cudaMemcpyHostToDevice, stream_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  if (hinfo != 0) {
    // not positive definite: SimplicialLLT would report NumericalIssue; surface it as an invalid step
    h.nonfinite = 1;
    PV_CUDA(cudaMemcpyAsync(d_.ctl, &h, sizeof(h), cudaMemcpyHostToDevice, stream_));
    PV_CUDA(cudaStreamSynchronize(stream_));
  } else {
    launch_finite_check(d_, n, d_.vec_acc, lc());
  }
  PV_CUDA(cudaEventRecord(ev_[2], stream_));
  PV_CUDA(cudaGetLastError());
  return POVAR_OK;
}

// the factorisation and substitution kernels on a caller-supplied symmetric matrix (tests)
int debug_cholesky(int n, const double* A, const double* b, double* x, int* info_out) {
  if (n <= 0 || !A || !b || !x) return POVAR_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return POVAR_ERR_NO_DEVICE;
  const int n_pad = chol_padded(n);
  double *S = nullptr, *linv = nullptr, *r = nullptr;
  int* info = nullptr;
  int rc = POVAR_OK;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess) rc = POVAR_ERR_CUDA;
    return e == cudaSuccess;
  };
  if (ok(cudaSetDevice(0)) && ok(cudaMalloc(&S, sizeof(double) * static_cast<size_t>(n_pad) * n_pad)) &&
      ok(cudaMalloc(&linv, sizeof(double) * static_cast<size_t>(n_pad) * 64)) &&
      ok(cudaMalloc(&r, sizeof(double) * static_cast<size_t>(n_pad))) && ok(cudaMalloc(&info, sizeof(int))) &&
      ok(cudaMemset(S, 0, sizeof(double) * static_cast<size_t>(n_pad) * n_pad)) &&
      ok(cudaMemset(r, 0, sizeof(double) * static_cast<size_t>(n_pad))) &&
      ok(cudaMemcpy2D(S, sizeof(double) * n_pad, A, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice)) &&
      ok(cudaMemcpy(r, b, sizeof(double) * n, cudaMemcpyHostToDevice))) {
    std::vector<double> ones(static_cast<size_t>(n_pad - n), 1.0);
    if (n_pad > n) {
      ok(cudaMemcpy2D(S + static_cast<size_t>(n) * n_pad + n, sizeof(double) * (n_pad + 1), ones.data(), sizeof(double),
                      sizeof(double), n_pad - n, cudaMemcpyHostToDevice));
    }
    LaunchCfg lc{};
    launch_cholesky_factor(S, n_pad, linv, info, lc);
    launch_cholesky_solve(S, n_pad, linv, r, info, lc);
    int hinfo = 0;
    ok(cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost));
    ok(cudaMemcpy(x, r, sizeof(double) * n, cudaMemcpyDeviceToHost));
    ok(cudaGetLastError());
    if (info_out) *info_out = hinfo;
  }
  cudaFree(S);
  cudaFree(linv);
  cudaFree(r);
  cudaFree(info);
  return rc;
}

// back-substitution and camera update + (sharded) the sum of l_diff over the ranks: trial_out[8] on the device
int Engine::enqueue_apply(bool joint, double alpha, bool reduce) {
  if (static_cast<int>(joint) != joint_lin_ || !have_solve_) {
    return fail(POVAR_ERR_INVALID, "apply called without a matching linearize + solve");
  }
  set_model(joint, joint ? opt_.alpha : alpha);
  const size_t C12 = static_cast<size_t>(C_) * 12;
  if (joint) {
    // apply_joint (linearizor_power_varproj.cpp:276-308): landmarks first, then P += s o (Pi inc)
    launch_make_y(d_, true, d_.vec_acc, d_.vec_y, lc());
    launch_backsub_joint(d_, mp_, d_.vec_y, lc());
    launch_cam_update_joint(d_, d_.vec_y, lc());
  } else if (opt_.solver_type_step_1 == POVAR_POWER_SCHUR_COMPLEMENT) {
    // linearizor_power_varproj.cpp:260-270
    launch_make_y(d_, false, d_.vec_acc, d_.vec_y, lc());
    launch_backsub_poba(d_, mp_, d_.vec_y, lc());
    launch_cam_update_pose(d_, d_.vec_acc, lc());
  } else {
    // VarPro flavours (linearizor_power_varproj.cpp:250-259, linearizor_sc.cpp:69-89):
    // cameras first, then the closed-form landmark re-solve
    PV_CUDA(cudaMemcpyAsync(P_prev_, d_.P, sizeof(double) * C12, cudaMemcpyDeviceToDevice, stream_));
    launch_cam_update_pose(d_, d_.vec_acc, lc());
    DeviceState tmp = d_;
    tmp.P_bak = P_prev_;
    launch_backsub_varpro(tmp, mp_, d_.vec_acc, lc());
  }
  return reduce ? allreduce(d_.trial_out + 8, 1) : POVAR_OK;
}

int Engine::apply(bool joint, double alpha, double* l_diff) {
  PV_CUDA(cudaSetDevice(device_));
  PV_CUDA(cudaEventRecord(ev_[0], stream_));
  {
    const int rc = enqueue_apply(joint, alpha);
    if (rc != POVAR_OK) return rc;
  }
  PV_CUDA(cudaGetLastError());
  double* v = host_out_ + 8;
  PV_CUDA(cudaMemcpyAsync(v + 8, d_.trial_out + 8, sizeof(double), cudaMemcpyDeviceToHost, stream_));
  PV_CUDA(cudaEventRecord(ev_[1], stream_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  times_.back_substitution += elapsed(ev_[0], ev_[1]);
  if (l_diff) *l_diff = v[8];
  return POVAR_OK;
}

int Engine::trial(bool joint, double alpha, double lambda, int32_t* iterations, double* l_diff,
                  povar_residual_info* ri) {
  PV_CUDA(cudaSetDevice(device_));
  int rc = enqueue_solve(joint, lambda);   // records ev_[0] (start), ev_[1] (prepared), ev_[2] (solved)
  if (rc != POVAR_OK) return rc;
  rc = backup(joint ? POVAR_STATE_JOINT : POVAR_STATE_POSE);
  if (rc != POVAR_OK) return rc;
  // sharded: the model decrease and the cost scalars of the trial travel together (trial_out[0..9), one exchange)
  rc = enqueue_apply(joint, alpha, /*reduce=*/false);
  if (rc != POVAR_OK) return rc;
  if (joint) {   // solver/bal_bundle_adjustment.cpp:700-705
    launch_normalize_cams(d_, lc());
    launch_normalize_joint(d_, lc());
  }
  PV_CUDA(cudaEventRecord(ev_[3], stream_));
  rc = enqueue_cost(joint, alpha, /*reduce=*/false);
  if (rc != POVAR_OK) return rc;
  rc = allreduce(d_.trial_out, 9);
  if (rc != POVAR_OK) return rc;
  PV_CUDA(cudaGetLastError());
  static_assert(sizeof(SeriesCtl) <= 64, "SeriesCtl shares the pinned read-back buffer");
  SeriesCtl* hctl = reinterpret_cast<SeriesCtl*>(host_out_);
  double* v = host_out_ + 8;
  PV_CUDA(cudaMemcpyAsync(hctl, d_.ctl, sizeof(SeriesCtl), cudaMemcpyDeviceToHost, stream_));
  PV_CUDA(cudaMemcpyAsync(v, d_.trial_out, 16 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  PV_CUDA(cudaEventRecord(ev_[4], stream_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  times_.prepare += elapsed(ev_[0], ev_[1]);
  times_.reduced_solve += elapsed(ev_[1], ev_[2]);
  times_.back_substitution += elapsed(ev_[2], ev_[3]);
  times_.residual += elapsed(ev_[3], ev_[4]);
  if (lin_check_pending_) {
    lin_check_pending_ = false;
    times_.linearize += elapsed(ev_[6], ev_[7]);
    if (v[9] > 0) return fail(POVAR_NUM_LINEARIZATION, "did not expect numerical failure during linearization");
  }
  if (iterations) *iterations = hctl->iterations;
  count_series_terms(hctl->term);
  if (l_diff) *l_diff = v[8];
  if (ri) decode_cost(v, ri);
  if (hctl->peer_timeout) return fail(POVAR_ERR_NCCL, "peer exchange of the camera sums timed out (a rank is gone?)");
  if (hctl->nonfinite) return POVAR_NUM_NONFINITE_INC;
  return POVAR_OK;
}

int Engine::backup(int which) {
  (void)which;
  PV_CUDA(cudaSetDevice(device_));
  PV_CUDA(cudaMemcpyAsync(d_.P_bak, d_.P, sizeof(double) * static_cast<size_t>(C_) * 12, cudaMemcpyDeviceToDevice, stream_));
  PV_CUDA(cudaMemcpyAsync(d_.X_bak, d_.X, sizeof(double) * static_cast<size_t>(L_) * 4, cudaMemcpyDeviceToDevice, stream_));
  return POVAR_OK;
}

int Engine::restore(int which) {
  (void)which;
  PV_CUDA(cudaSetDevice(device_));
  PV_CUDA(cudaMemcpyAsync(d_.P, d_.P_bak, sizeof(double) * static_cast<size_t>(C_) * 12, cudaMemcpyDeviceToDevice, stream_));
  PV_CUDA(cudaMemcpyAsync(d_.X, d_.X_bak, sizeof(double) * static_cast<size_t>(L_) * 4, cudaMemcpyDeviceToDevice, stream_));
  return POVAR_OK;
}

int Engine::to_homogeneous() {
  PV_CUDA(cudaSetDevice(device_));
  launch_to_homogeneous(d_, lc());
  launch_normalize_cams(d_, lc());
  PV_CUDA(cudaGetLastError());
  return POVAR_OK;
}

int Engine::normalize_joint() {
  PV_CUDA(cudaSetDevice(device_));
  launch_normalize_cams(d_, lc());
  launch_normalize_joint(d_, lc());
  PV_CUDA(cudaGetLastError());
  return POVAR_OK;
}

int Engine::get_state(int which, double* cam_P, double* lms) {
  PV_CUDA(cudaSetDevice(device_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  if (cam_P) PV_CUDA(cudaMemcpy(cam_P, d_.P, sizeof(double) * static_cast<size_t>(C_) * 12, cudaMemcpyDeviceToHost));
  if (lms && L_ > 0) {
    // straight into the caller's buffer (no staging vector: its page faults cost more than the copy);
    // step-1 landmarks are [x y z 1] on the device and [x y z] for the caller: a pitched copy drops w
    if (which == POVAR_STATE_JOINT) {
      PV_CUDA(cudaMemcpy(lms, d_.X, sizeof(double) * 4 * static_cast<size_t>(L_), cudaMemcpyDeviceToHost));
    } else {
      PV_CUDA(cudaMemcpy2D(lms, 3 * sizeof(double), d_.X, 4 * sizeof(double), 3 * sizeof(double),
                           static_cast<size_t>(L_), cudaMemcpyDeviceToHost));
    }
  }
  return POVAR_OK;
}

int Engine::set_state(int which, const double* cam_P, const double* lms) {
  PV_CUDA(cudaSetDevice(device_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  if (cam_P) PV_CUDA(cudaMemcpy(d_.P, cam_P, sizeof(double) * static_cast<size_t>(C_) * 12, cudaMemcpyHostToDevice));
  if (lms) {
    std::vector<double> h(static_cast<size_t>(L_) * 4);
    for (int l = 0; l < L_; ++l) {
      if (which == POVAR_STATE_JOINT) {
        for (int k = 0; k < 4; ++k) h[4 * static_cast<size_t>(l) + k] = lms[4 * static_cast<size_t>(l) + k];
      } else {
        for (int k = 0; k < 3; ++k) h[4 * static_cast<size_t>(l) + k] = lms[3 * static_cast<size_t>(l) + k];
        h[4 * static_cast<size_t>(l) + 3] = 1.0;
      }
    }
    PV_CUDA(cudaMemcpy(d_.X, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
  }
  return POVAR_OK;
}

int64_t Engine::debug_read(const char* name, double* out, int64_t capacity) {
  if (cudaSetDevice(device_) != cudaSuccess) return POVAR_ERR_CUDA;
  cudaStreamSynchronize(stream_);
  const std::string n(name ? name : "");
  const double* dsrc = nullptr;
  const int* isrc = nullptr;
  int64_t count = 0;
  const int D = dim_;
  if (n == "P") dsrc = d_.P, count = 12LL * C_;
  else if (n == "X") dsrc = d_.X, count = 4LL * L_;
  else if (n == "pose_scale") dsrc = d_.pose_scale, count = 12LL * C_;
  else if (n == "lm_scale") dsrc = d_.lm_scale, count = 4LL * L_;
  else if (n == "lm_hraw" || n == "lm_graw" || n == "hll_inv") {
    // by landmark for the caller: the landmarks of the sliced-ELL set keep these as lane-major planes
    // [slice][component][32], the others (more than 32 observations) by landmark
    const int K = n == "lm_hraw" ? 10 : (n == "lm_graw" ? 4 : 6);
    count = static_cast<int64_t>(K) * L_;
    if (!out) return count;
    if (capacity < count) {
      fail(POVAR_ERR_INVALID, "debug_read: buffer too small");
      return POVAR_ERR_INVALID;
    }
    const double* by_lm = K == 10 ? d_.lm_hraw : (K == 4 ? d_.lm_graw : d_.hll_inv);
    const double* planes = K == 10 ? d_.sell_hraw : (K == 4 ? d_.sell_graw : d_.sell_hinv);
    const size_t slots = static_cast<size_t>(kSellWidth) * d_.ix.num_slices;
    std::vector<double> hp(slots * K);
    std::vector<int> hlm(slots);
    if (cudaMemcpy(out, by_lm, sizeof(double) * count, cudaMemcpyDeviceToHost) != cudaSuccess) return POVAR_ERR_CUDA;
    if (slots > 0) {
      if (cudaMemcpy(hp.data(), planes, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(hlm.data(), d_.ix.sell_lm, sizeof(int) * slots, cudaMemcpyDeviceToHost) != cudaSuccess) {
        return POVAR_ERR_CUDA;
      }
    }
    for (size_t i = 0; i < slots; ++i) {
      if (hlm[i] < 0) continue;
      const size_t sl = i / kSellWidth, lane = i % kSellWidth;
      for (int k = 0; k < K; ++k) out[static_cast<size_t>(hlm[i]) * K + k] = hp[(sl * K + k) * kSellWidth + lane];
    }
    return count;
  }
  else if (n == "lm_rec") dsrc = d_.lm_rec, count = static_cast<int64_t>(kLmRec) * L_;
  else if (n == "kron") dsrc = d_.kron, count = static_cast<int64_t>(kKron) * C_;
  else if (n == "b_mat") dsrc = d_.Bmat, count = 144LL * C_;
  else if (n == "b_inv") dsrc = d_.Binv, count = 144LL * C_;
  else if (n == "b") dsrc = d_.b, count = static_cast<int64_t>(D) * C_;
  else if (n == "inc") dsrc = d_.vec_acc, count = static_cast<int64_t>(D) * C_;
  else if (n == "cam_raw") dsrc = d_.cam_raw, count = 12LL * C_;
  else if (n == "vec_y") dsrc = d_.vec_y, count = 12LL * C_;
  else if (n == "lm_ptr") isrc = d_.ix.lm_ptr, count = L_ + 1;
  else if (n == "obs_cam") isrc = d_.ix.obs_cam, count = nnz_;
  else if (n == "obs_lm") isrc = d_.ix.obs_lm, count = nnz_;
  else if (n == "cam_ptr") isrc = d_.ix.cam_ptr, count = C_ + 1;
  else if (n == "csc_lm") isrc = d_.ix.csc_lm, count = nnz_;
  else if (n == "item_ptr") isrc = d_.ix.item_ptr, count = d_.ix.num_items + 1;
  else if (n == "item_cam") isrc = d_.ix.item_cam, count = d_.ix.num_items;
  else if (n == "slice_ptr") isrc = d_.ix.slice_ptr, count = d_.ix.num_slices + 1;
  else if (n == "sell_lm") isrc = d_.ix.sell_lm, count = static_cast<int64_t>(kSellWidth) * d_.ix.num_slices;
  else if (n == "sell_cam") isrc = d_.ix.sell_cam, count = d_.ix.sell_slots;
  else if (n == "sell_cam_e0") isrc = d_.ix.sell_cam_e0, count = d_.ix.sell_slots;
  else if (n == "sell_max_deg") {
    if (out && capacity >= 1) out[0] = d_.ix.sell_max_deg;
    return 1;
  }
  else if (n == "obs_slot") isrc = d_.ix.obs_slot, count = nnz_;
  else {
    fail(POVAR_ERR_INVALID, "debug_read: unknown array '" + n + "'");
    return POVAR_ERR_INVALID;
  }
  if (!out) return count;
  if (capacity < count) {
    fail(POVAR_ERR_INVALID, "debug_read: buffer too small");
    return POVAR_ERR_INVALID;
  }
  if (dsrc) {
    if (cudaMemcpy(out, dsrc, sizeof(double) * count, cudaMemcpyDeviceToHost) != cudaSuccess) return POVAR_ERR_CUDA;
  } else {
    std::vector<int> h(count);
    if (cudaMemcpy(h.data(), isrc, sizeof(int) * count, cudaMemcpyDeviceToHost) != cudaSuccess) return POVAR_ERR_CUDA;
    for (int64_t i = 0; i < count; ++i) out[i] = h[i];
  }
  return count;
}

// E0 x with the current linearisation and the Hll^-1 of the last solve
int Engine::right_mul_e0(bool joint, const double* x, double* out) {
  PV_CUDA(cudaSetDevice(device_));
  if (static_cast<int>(joint) != joint_lin_ || !have_solve_) {
    return fail(POVAR_ERR_INVALID, "right_mul_e0: needs a matching linearize + solve");
  }
  const size_t n = static_cast<size_t>(C_) * (joint ? 11 : 12);
  PV_CUDA(cudaMemcpyAsync(d_.vec_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
  launch_make_y(d_, joint, d_.vec_x, d_.vec_y, lc());
  {
    const int rc = e0_product(joint, false, false);
    if (rc != POVAR_OK) return rc;
  }
  launch_e0_finish(d_, joint, d_.vec_x, lc());
  PV_CUDA(cudaGetLastError());
  PV_CUDA(cudaMemcpyAsync(out, d_.vec_x, sizeof(double) * n, cudaMemcpyDeviceToHost, stream_));
  PV_CUDA(cudaStreamSynchronize(stream_));
  return POVAR_OK;
}

// `terms` full power-series terms (landmark pass, camera pass, item reduction, allreduce, B^-1
// and norms) on the current linearisation, no early exit: the SpMV measurement of bench.py.
int Engine::bench_power_terms(bool joint, int terms, double* seconds_per_term) {
  PV_CUDA(cudaSetDevice(device_));
  if (static_cast<int>(joint) != joint_lin_ || !have_solve_) {
    return fail(POVAR_ERR_INVALID, "bench_power_terms: needs a matching linearize + solve");
  }
  if (terms <= 0) return fail(POVAR_ERR_INVALID, "bench_power_terms: terms must be positive");
  launch_finish_b(d_, joint, lc());
  launch_series_start(d_, -1.0, terms, lc());
  PV_CUDA(cudaEventRecord(ev_[0], stream_));
  for (int i = 1; i <= terms; ++i) {
    const int rc = e0_product(joint, true, term_mode() != kTermRaw);
    if (rc != POVAR_OK) return rc;
    launch_series_term(d_, joint, i, /*eta=*/-1.0, /*r_tolerance=*/-1.0, term_mode(), exchange(), lc());
  }
  PV_CUDA(cudaEventRecord(ev_[1], stream_));
  PV_CUDA(cudaGetLastError());
  PV_CUDA(cudaStreamSynchronize(stream_));
  if (seconds_per_term) *seconds_per_term = elapsed(ev_[0], ev_[1]) / terms;
  return POVAR_OK;
}

// average device time of each kernel of a power-series term, launched `reps` times back to back:
// [0] landmark half, [1] camera half, [2] item reduction, [3] B^-1 / accumulate / norms / test
int Engine::bench_power_kernels(bool joint, int reps, double* seconds) {
  PV_CUDA(cudaSetDevice(device_));
  if (static_cast<int>(joint) != joint_lin_ || !have_solve_) {
    return fail(POVAR_ERR_INVALID, "bench_power_kernels: needs a matching linearize + solve");
  }
  if (reps <= 0 || !seconds) return fail(POVAR_ERR_INVALID, "bench_power_kernels: bad arguments");
  launch_finish_b(d_, joint, lc());
  launch_series_start(d_, -1.0, reps, lc());
  for (int k = 0; k < 4; ++k) {
    PV_CUDA(cudaEventRecord(ev_[0], stream_));
    for (int i = 0; i < reps; ++i) {
      switch (k) {
        case 0:
          launch_e0_landmark_v2(d_, mp_, joint, true, lc());
          break;
        case 1:
          launch_passB_e0_v2(d_, mp_, joint, true, lc());
          break;
        case 2: launch_reduce_items(d_, d_.item_part, 12, d_.cam_raw, true, lc()); break;
        default: launch_series_term(d_, joint, i + 1, -1.0, -1.0, term_mode(), exchange(), lc()); break;
      }
    }
    PV_CUDA(cudaEventRecord(ev_[1], stream_));
    PV_CUDA(cudaGetLastError());
    PV_CUDA(cudaStreamSynchronize(stream_));
    seconds[k] = elapsed(ev_[0], ev_[1]) / reps;
  }
  return POVAR_OK;
}

}  // namespace povar
