// The landmark-major walk over the sliced-ELL observation order with everything the inner loop reads staged in
// shared memory by the TMA unit (cp.async.bulk + mbarrier).  One kernel template, k_sell_walk<Op, ...>; the
// operations are the landmark half of a power-series term (kernels_series.cu) and the once-per-trial landmark
// passes -- linearisation and back-substitution (kernels_landmark.cu).
//
// A warp walks a contiguous range of slices (LmPlan::range_slice: equal numbers of rows per warp); a slice is
// 32 landmarks of (nearly) equal degree, one per lane; the lane visits the observations of its landmark in
// camera order and keeps its sums in registers: no tile table, no reduction, no exchange.
//   * per-camera records.  The landmarks are ordered by the centre of their cameras, so the slices of one
//     block meet a WINDOW of cameras (LmPlan::blk_lo, made on the host from the exact camera ranges of the
//     block's slices); the block stages the records of that window once, and every lane reads the record of
//     its observation's camera with LDS.128.  A record is an odd number of 16-byte units, so consecutive
//     records start in different banks.  A camera outside the window (only when the window a block needs does
//     not fit; never for a small C) is read from the same table in global memory, so the result does not
//     depend on the window;
//   * the observation stream (camera index and model coefficients, Op::kStage bytes per row of 32 slots):
//     every warp owns a ring of D rows; one lane refills the stage of row r with row r + D as soon as the warp
//     has used it.  No registers and no scoreboards are tied up while the data travels (round 2's first version
//     kept the next rows in registers: the compiler put the scoreboard waits of those loads at the head of the
//     loop and the latency of the stream was exposed in every iteration, profiles/r2_summary.md).
//
// An operation is a struct passed by value:
//   static constexpr int kRec;      doubles per camera record (kRec / 2 odd)
//   static constexpr int kStage;    bytes per row of the stream (multiple of 128)
//   static constexpr int kWarpsPerSm;   32, or 16 for an operation that needs more than 64 registers per lane
//                                       (its plan is made with that limit)
//   struct Lane;                    per-lane state of the open slice
//   bool skip() const;              every thread: leave at once (a converged series)
//   void init(Lane&) const;         every thread, before the walk (sums that outlive a slice)
//   void issue(ix, row, stage, bar) const;    one lane: bulk copies of row `row` into `stage`, completing on bar
//   void open(Lane&, ix, slice, lane, last_slice) const;
//   void obs(Lane&, rec, stage, lane, row) const;   rec: the record of the slot's camera (shared or global)
//   void close(Lane&, ix, slice, lane) const;
//   void finish(Lane&, ix, win, table) const; after the walk, every thread of the block (long landmarks,
//                                             block-wide sums)
#pragma once

#include <cuda_runtime.h>

#include "povar_internal.h"

namespace povar {

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same for data that are read once per kernel (the observation streams): evict_first in L2 (device_math.cuh)
__device__ __forceinline__ void bulk_stream_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}

// the staged stretch of the per-camera table
struct CamWindow {
  const double* smem;   // records of cameras lo .. lo + n - 1
  int lo, n;
};

// the usual stream of the walks that evaluate the observation model themselves: camera index + (u, v)
__device__ __forceinline__ void issue_cam_uv(const DeviceIndex& ix, int row, unsigned char* stage,
                                             unsigned long long* bar) {
  const size_t slot = kSellWidth * static_cast<size_t>(row);
  mbar_expect_tx(bar, kStagePose);
  bulk_stream_g2s(stage, ix.sell_cam + slot, 128u, bar);
  bulk_stream_g2s(stage + 128, ix.sell_uv + slot, 512u, bar);
}

#ifdef POVAR_WALK_TRACE
// tuning builds only (python -m povar_b200.build with -DPOVAR_WALK_TRACE): per block, the latest time any of its
// warps passed [0] kernel entry, [1] window staged, [2] slices walked, [3] finish() done (globaltimer ns)
__device__ unsigned long long g_walk_trace[4 * 1024];
__device__ __forceinline__ void walk_stamp(int which) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if ((threadIdx.x & 31) == 0 && blockIdx.x < 1024) atomicMax(&g_walk_trace[4 * blockIdx.x + which], t);
}
#define POVAR_WALK_STAMP(k) walk_stamp(k)
#else
#define POVAR_WALK_STAMP(k)
#endif

template <class Op, int W, int D, int BPS>
__global__ void __launch_bounds__(32 * W, BPS)
k_sell_walk(DeviceIndex ix, LmPlan plan, int win_cams, const double* __restrict__ table, Op op) {
  extern __shared__ __align__(128) unsigned char walk_smem[];
  if (op.skip()) return;
  POVAR_WALK_STAMP(0);
  constexpr int kStage = Op::kStage;
  constexpr int kRec = Op::kRec;
  constexpr int kBars = lm_bar_bytes(W, D);
  static_assert(kRec % 2 == 0 && (kRec / 2) % 2 == 1, "records are an odd number of 16-byte units");
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int range = static_cast<int>(blockIdx.x) * W + wib;
  // shared memory: [mbarriers: window, then D per warp][rings][window]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(walk_smem);
  unsigned long long* my_bars = bars + 1 + wib * D;
  unsigned char* my_ring = walk_smem + kBars + static_cast<size_t>(wib) * D * kStage;
  CamWindow win;
  win.smem = reinterpret_cast<const double*>(walk_smem + kBars + static_cast<size_t>(W) * D * kStage);
  win.n = min(win_cams, ix.C);
  win.lo = min(__ldg(plan.blk_lo + blockIdx.x), ix.C - win.n);
  if (threadIdx.x == 0) mbar_init(&bars[0], 1);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < D; ++s) mbar_init(&my_bars[s], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = static_cast<unsigned>(win.n) * (kRec * 8);
    mbar_expect_tx(&bars[0], bytes);
    const char* src = reinterpret_cast<const char*>(table + kRec * static_cast<size_t>(win.lo));
    char* dst = reinterpret_cast<char*>(const_cast<double*>(win.smem));
    for (unsigned off = 0; off < bytes; off += 32768u) {
      bulk_copy_g2s(dst + off, src + off, min(32768u, bytes - off), &bars[0]);
    }
  }
  int s0 = 0, s1 = 0;
  if (range < plan.ranges) {
    s0 = __ldg(plan.range_slice + range);
    s1 = __ldg(plan.range_slice + range + 1);
  }
  typename Op::Lane st;
  op.init(st);
  if (s0 < s1) {
    const int row_first = __ldg(ix.slice_ptr + s0);
    const int row_end = __ldg(ix.slice_ptr + s1);        // one past the last row of this warp
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < D; ++s) {
        if (row_first + s < row_end) op.issue(ix, row_first + s, my_ring + s * kStage, &my_bars[s]);
      }
    }
    // first rows after the slices of this warp, 32 at a time: lane j keeps the one of slice hdr_base + j
    int hdr_base = s0;
    int hdr = __ldg(ix.slice_ptr + min(s0 + lane, s1 - 1) + 1);
    int sl = s0, row1 = 0;
    auto open_slice = [&]() {
      if (sl - hdr_base >= 32) {
        hdr_base = sl;
        hdr = __ldg(ix.slice_ptr + min(sl + lane, s1 - 1) + 1);
      }
      row1 = __shfl_sync(0xffffffffu, hdr, sl - hdr_base);
      op.open(st, ix, sl, lane, s1 - 1);
    };
    open_slice();
    mbar_wait(&bars[0], 0);                   // the window is in shared memory
    POVAR_WALK_STAMP(1);
    unsigned phase = 0;
    for (int row = row_first; row < row_end; row += D) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int r = row + i;
        if (r < row_end) {           // warp-uniform
          const unsigned char* stage = my_ring + i * kStage;
          mbar_wait(&my_bars[i], phase);
          const int c = reinterpret_cast<const int*>(stage)[lane];
          if (r == row1) {           // warp-uniform: the previous slice is complete
            op.close(st, ix, sl, lane);
            ++sl;
            open_slice();
          }
          if (c >= 0) {
            const unsigned rel = static_cast<unsigned>(c - win.lo);
            if (rel < static_cast<unsigned>(win.n)) {
              op.obs(st, reinterpret_cast<const double2*>(win.smem + kRec * static_cast<size_t>(rel)), stage, lane, r);
            } else {
              op.obs(st, reinterpret_cast<const double2*>(table + kRec * static_cast<size_t>(c)), stage, lane, r);
            }
          }
          // every lane holds what it needs of its slot in registers: the stage can take row r + D
          __syncwarp();
          if (lane == 0 && r + D < row_end) op.issue(ix, r + D, my_ring + i * kStage, &my_bars[i]);
        }
      }
      phase ^= 1u;
    }
    op.close(st, ix, sl, lane);
  } else {
    mbar_wait(&bars[0], 0);   // nobody leaves while the copy is in flight
  }
  POVAR_WALK_STAMP(2);
  op.finish(st, ix, win, table);
  POVAR_WALK_STAMP(3);
}

// launch with the shape the plan chose (plan_landmark_half, engine.cu)
template <class Op, int W, int D, int BPS>
void launch_sell_walk_cfg(const DeviceIndex& ix, const LmPlan& plan, int win_cams, const double* table, const Op& op,
                          cudaStream_t stream) {
  const size_t smem = lm_bar_bytes(W, D) + static_cast<size_t>(W) * D * Op::kStage +
                      static_cast<size_t>(win_cams) * Op::kRec * 8;
  auto kernel = k_sell_walk<Op, W, D, BPS>;
  static const cudaError_t attr = [&]() {
    // the opt-in maximum covers static and dynamic shared memory together
    cudaFuncAttributes fa{};
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e == cudaSuccess) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024 - static_cast<int>(fa.sharedSizeBytes));
    }
    if (e == cudaSuccess) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    return e;
  }();
  (void)attr;
  kernel<<<plan.blocks, 32 * W, smem, stream>>>(ix, plan, win_cams, table, op);
}

// false: nothing to do (no landmarks in this shard)
template <class Op>
bool launch_sell_walk(const DeviceIndex& ix, const LmPlan& plan, int window_cap, const double* table, const Op& op,
                      cudaStream_t stream) {
  if (plan.blocks == 0) return false;
  // tests cap the window so that the global-memory path for cameras outside it is exercised
  const int win = window_cap > 0 && window_cap < plan.win_cams ? window_cap : plan.win_cams;
#define POVAR_WALK(W, D, BPS)                                                     \
  if (plan.warps == W && plan.stages == D && plan.blocks_per_sm == BPS) {         \
    launch_sell_walk_cfg<Op, W, D, BPS>(ix, plan, win, table, op, stream);        \
    return true;                                                                  \
  }
  if constexpr (Op::kWarpsPerSm == 32) {
    POVAR_WALK(8, 3, 4)
    POVAR_WALK(16, 3, 2)
    POVAR_WALK(32, 3, 1)
    POVAR_WALK(32, 2, 1)
    POVAR_WALK(24, 2, 1)
  } else {   // 16: up to 128 registers per lane
    POVAR_WALK(8, 3, 2)
    POVAR_WALK(16, 3, 1)
  }
  POVAR_WALK(16, 2, 1)
#undef POVAR_WALK
  return false;   // plan_landmark_half makes no other shape
}

}  // namespace povar
