// The two halves of one power-series term,  raw_c = sum_l Jp^T Jl Hll^-1 Jl^T Jp y.  Replaces
// right_mul_e0_pOSE / right_mul_e0_joint
// (/root/reference/src/rootba_povar/sc/linearization_power_varproj.hpp:364-453).
//
// Both observation models have  Jp_raw = K (x) X^T  and  Jl_raw = K M  with a small K that depends on
// the observation only (step 1: K_i(u, v, c1, c2) 4x3, M = P[:, 0:3]; step 2: K_i = d_i 2x3, M = P), so
//   landmark half:  G_l = sum_i M^T (K^T W K) (Y_c X_l),   H_l = fold_l G_l
//   camera half:    raw_c = sum_i ((K^T W K) (M H_l)) (x) X_l
// where Y_c is y_c as a 3x4 matrix and fold_l = S (Pi) Hll^-1 (Pi^T) S is made once per solve
// (k_prep_landmark).  Step 2 streams sqrt(w) d_i, stored at the linearisation point.
//
// Landmark half: one LANE per landmark (sliced ELL, 32 landmarks of nearly equal degree per slice, the k-th
// observations of the 32 landmarks in one row), the camera records in shared memory.  History, because it
// explains the shape: with one observation per lane and the records in global memory every LDG.128 touched
// 32 cache lines (round 1, v1); four lanes per observation reading 64-byte chunks fixed the gathers but spent
// 4x the instructions on redundant arithmetic and a 3-value exchange per observation -- that kernel ran at
// half the issue rate of its 32 warps per SM and scaled with nothing but the warp count (this round's
// profiles).  From shared memory a whole 176-byte record per lane is cheap (eleven LDS.128, conflicts limited
// by the odd record stride), nothing is computed twice and nothing is exchanged.
// Camera half: two lanes per entry, the landmark records gathered from L2, one 256-bit load per lane.
#include <cuda_runtime.h>

#include <cstdlib>

#include "device_math.cuh"
#include "povar_internal.h"
#include "sell_walk.cuh"

namespace povar {

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

inline void count(const LaunchCfg& lc, int n = 1) {
  if (lc.launch_counter) *lc.launch_counter += n;
}

__device__ __forceinline__ double2 ldg2(const double* __restrict__ p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

// m = K^T W K q for the step-1 model (pose_jp_mul followed by pose_jpT_coef)
__device__ __forceinline__ void pose_normal_coef(double q0, double q1, double q2, double u, double v,
                                                 double w, double c1, double c2, double (&m)[3]) {
  const double a0 = w * (c1 * (q0 - u * q2)), a1 = w * (c1 * (q1 - v * q2));
  const double a2 = w * (c2 * q0), a3 = w * (c2 * q1);
  m[0] = c1 * a0 + c2 * a2;
  m[1] = c1 * a1 + c2 * a3;
  m[2] = -c1 * (u * a0 + v * a1);
}

// m = d'^T d' q for the step-2 model, d' = sqrt(w) [[d0 0 d02], [0 d0 d12]]
__device__ __forceinline__ void joint_normal_coef(double q0, double q1, double q2, double d0, double d02,
                                                  double d12, double (&m)[3]) {
  const double a0 = d0 * q0 + d02 * q2, a1 = d0 * q1 + d12 * q2;
  m[0] = d0 * a0;
  m[1] = d0 * a1;
  m[2] = d02 * a0 + d12 * a1;
}

struct ObsCoef {
  double a, b, c;   // step 1: u, v, w;  step 2: sqrt(w) (1/z, -x/z^2, -y/z^2)
};

// ------------------------------------------------------------------------------------------
// landmark half: the sliced-ELL walk of sell_walk.cuh with the camera records [y_c | M_c] (CamRec) staged per
// block and the stream [camera index | (u, v) (| w)] resp. [camera index | d] in the per-warp rings.  The
// per-slice landmark data (X, fold) are lane-major planes written by the linearisation walk and k_prep_sell
// (kernels_landmark.cu): one coalesced line per component; they are pulled into L2 one slice ahead.
// ------------------------------------------------------------------------------------------
// G += M^T (K^T W K) (Y x) for one observation whose camera record is `rec` (shared or global memory)
template <bool JOINT>
__device__ __forceinline__ void landmark_obs(const double2* __restrict__ rec, const double (&x)[4], const ObsCoef& k,
                                             double c1, double c2, double (&G)[4]) {
  double q[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double2 a = rec[2 * r], b = rec[2 * r + 1];
    q[r] = a.x * x[0] + a.y * x[1] + b.x * x[2] + b.y * x[3];
  }
  double m[3];
  if (JOINT) {
    joint_normal_coef(q[0], q[1], q[2], k.a, k.b, k.c, m);
    // M = P, row-major 3 x 4, units 6..11
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double2 a = rec[6 + 2 * r], b = rec[7 + 2 * r];
      G[0] += a.x * m[r];
      G[1] += a.y * m[r];
      G[2] += b.x * m[r];
      G[3] += b.y * m[r];
    }
  } else {
    pose_normal_coef(q[0], q[1], q[2], k.a, k.b, k.c, c1, c2, m);
    // M = P[:, 0:3], row-major 3 x 3 (+ pad), units 6..10:  (M00 M01) (M02 M10) (M11 M12) (M20 M21) (M22 -)
    const double2 u0 = rec[6], u1 = rec[7], u2 = rec[8], u3 = rec[9], u4 = rec[10];
    G[0] += u0.x * m[0] + u1.y * m[1] + u3.x * m[2];
    G[1] += u0.y * m[0] + u2.x * m[1] + u3.y * m[2];
    G[2] += u1.x * m[0] + u2.y * m[1] + u4.x * m[2];
  }
}

// the record of camera c (c >= 0): from the window if it is there, else from global memory
template <bool JOINT>
__device__ __forceinline__ void landmark_obs_at(const CamWindow& w, const double* __restrict__ cam_rec, int c,
                                                const double (&x)[4], const ObsCoef& k, double c1, double c2,
                                                double (&G)[4]) {
  constexpr int kStride = CamRec::stride(JOINT);
  const unsigned rel = static_cast<unsigned>(c - w.lo);
  if (rel < static_cast<unsigned>(w.n)) {
    landmark_obs<JOINT>(reinterpret_cast<const double2*>(w.smem + kStride * static_cast<size_t>(rel)), x, k, c1, c2, G);
  } else {
    landmark_obs<JOINT>(reinterpret_cast<const double2*>(cam_rec + kStride * static_cast<size_t>(c)), x, k, c1, c2, G);
  }
}

// H = fold G  (fold packed symmetric: 3x3 in step 1, 4x4 in step 2)
template <bool JOINT>
__device__ __forceinline__ void fold_apply(const double (&F)[10], const double (&G)[4], double (&H)[4]) {
  if (JOINT) {
    H[0] = F[0] * G[0] + F[1] * G[1] + F[2] * G[2] + F[3] * G[3];
    H[1] = F[1] * G[0] + F[4] * G[1] + F[5] * G[2] + F[6] * G[3];
    H[2] = F[2] * G[0] + F[5] * G[1] + F[7] * G[2] + F[8] * G[3];
    H[3] = F[3] * G[0] + F[6] * G[1] + F[8] * G[2] + F[9] * G[3];
  } else {
    H[0] = F[0] * G[0] + F[1] * G[1] + F[2] * G[2];
    H[1] = F[1] * G[0] + F[3] * G[1] + F[4] * G[2];
    H[2] = F[2] * G[0] + F[4] * G[1] + F[5] * G[2];
    H[3] = 0.0;
  }
}

// landmarks with more than 32 observations (outside the sliced-ELL set): one warp per landmark, taken by
// the warps of the same blocks after their slices; lane g takes observations g, g + 32, ... of the CSR list,
// fixed-tree sum over the lanes
template <bool JOINT, bool HASW>
__device__ __forceinline__ void long_landmark_warp(const DeviceIndex& ix, const CamWindow& win, int which,
                                                   const double* __restrict__ X,
                                                   const double* __restrict__ cam_rec,
                                                   const double* __restrict__ obs_d,
                                                   const double* __restrict__ obs_w, double c1, double c2,
                                                   const double* __restrict__ lm_fold,
                                                   double* __restrict__ lm_rec) {
  const int lane = threadIdx.x & 31;
  const int lm = __ldg(ix.long_lm + which);
  const int ob = __ldg(ix.lm_ptr + lm), oe = __ldg(ix.lm_ptr + lm + 1);
  double x[4];
  load4(X + 4 * static_cast<size_t>(lm), x);
  double G[4] = {0.0, 0.0, 0.0, 0.0};
  for (int o = ob + lane; o < oe; o += 32) {
    const int c = __ldg(ix.obs_cam + o);
    ObsCoef k;
    if (JOINT) {
      const double* dp = obs_d + 3 * static_cast<size_t>(o);
      k.a = __ldg(dp);
      k.b = __ldg(dp + 1);
      k.c = __ldg(dp + 2);
    } else {
      const double2 uv = ix.obs_uv[o];
      k.a = uv.x;
      k.b = uv.y;
      k.c = HASW ? __ldg(obs_w + o) : 1.0;
    }
    landmark_obs_at<JOINT>(win, cam_rec, c, x, k, c1, c2, G);
  }
  warp_allreduce<4>(G);
  double F[10], H[4];
#pragma unroll
  for (int i = 0; i < 10; ++i) F[i] = (JOINT || i < 6) ? __ldg(lm_fold + 10 * static_cast<size_t>(lm) + i) : 0.0;
  fold_apply<JOINT>(F, G, H);
  if (lane < 4) {
    lm_rec[kLmRec * static_cast<size_t>(lm) + (lane < 2 ? kLmRecH0 + lane : kLmRecH2 + lane - 2)] =
        lane == 0 ? H[0] : (lane == 1 ? H[1] : (lane == 2 ? H[2] : H[3]));
  }
}

template <bool JOINT, bool HASW>
struct E0LandmarkOp {
  static constexpr int kRec = CamRec::stride(JOINT);
  static constexpr int kStage = (JOINT || HASW) ? kStageWide : kStagePose;
  static constexpr int kWarpsPerSm = 32;
  static_assert(kRec == (JOINT ? kCamRecJoint : kCamRecPose), "plan_landmark_half sizes the window with these");
  const double* X;
  const double* sell_d;
  const double* sell_w;
  double c1, c2;
  const double* lm_fold;
  const double* sell_x;
  const double* sell_fold;
  double* lm_rec;
  const double* obs_d;
  const double* obs_w;
  const SeriesCtl* ctl;
  SeriesCtl* queue;   // work queue of the long landmarks (nullptr: round robin)

  struct Lane {
    int lm;
    double x[4], G[4];
  };

  __device__ __forceinline__ bool skip() const { return ctl != nullptr && ctl->done; }
  __device__ __forceinline__ void init(Lane&) const {}

  __device__ __forceinline__ void issue(const DeviceIndex& ix, int row, unsigned char* stage,
                                        unsigned long long* bar) const {
    const size_t slot = kSellWidth * static_cast<size_t>(row);
    mbar_expect_tx(bar, kStage);
    bulk_stream_g2s(stage, ix.sell_cam_e0 + slot, 128u, bar);
    if (JOINT) {
      bulk_stream_g2s(stage + 128, sell_d + 3 * slot, 768u, bar);
    } else {
      bulk_stream_g2s(stage + 128, ix.sell_uv_e0 + slot, 512u, bar);
      if (HASW) bulk_stream_g2s(stage + 640, sell_w + slot, 256u, bar);
    }
  }

  __device__ __forceinline__ void open(Lane& st, const DeviceIndex& ix, int sl, int lane, int last) const {
    st.lm = __ldcs(ix.sell_lm + kSellWidth * static_cast<size_t>(sl) + lane);
    const double* xp = sell_x + 4 * kSellWidth * static_cast<size_t>(sl) + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      st.x[k] = __ldcs(xp + k * kSellWidth);
      st.G[k] = 0.0;
    }
    // towards L2, one 128-byte line per lane: the fold of this slice (read when it closes) and the
    // landmarks of the next one
    constexpr int kFoldLines = JOINT ? 20 : 12;
    if (lane < kFoldLines) {
      prefetch_l2(sell_fold + 10 * kSellWidth * static_cast<size_t>(sl) + 16 * lane);
    } else if (lane - kFoldLines < 8 && sl < last) {
      prefetch_l2(sell_x + 4 * kSellWidth * static_cast<size_t>(sl + 1) + 16 * (lane - kFoldLines));
    } else if (lane == 31 && sl < last) {
      prefetch_l2(ix.sell_lm + kSellWidth * static_cast<size_t>(sl + 1));
    }
  }

  __device__ __forceinline__ void obs(Lane& st, const double2* __restrict__ rec, const unsigned char* stage, int lane,
                                      int /*row*/) const {
    ObsCoef k;
    if (JOINT) {
      const double* dp = reinterpret_cast<const double*>(stage + 128) + lane;
      k.a = dp[0];
      k.b = dp[kSellWidth];
      k.c = dp[2 * kSellWidth];
    } else {
      const double2 uv = reinterpret_cast<const double2*>(stage + 128)[lane];
      k.a = uv.x;
      k.b = uv.y;
      k.c = HASW ? reinterpret_cast<const double*>(stage + 640)[lane] : 1.0;
    }
    landmark_obs<JOINT>(rec, st.x, k, c1, c2, st.G);
  }

  __device__ __forceinline__ void close(Lane& st, const DeviceIndex& /*ix*/, int sl, int lane) const {
    double F[10], H[4];
    const double* fp = sell_fold + 10 * kSellWidth * static_cast<size_t>(sl) + lane;
#pragma unroll
    for (int i = 0; i < 10; ++i) F[i] = (JOINT || i < 6) ? __ldcs(fp + i * kSellWidth) : 0.0;
    fold_apply<JOINT>(F, st.G, H);
    if (st.lm >= 0) {
      // evict_last: the camera half gathers these records next (a streaming store sent half of them to DRAM).
      // Measured and dropped: whole sectors [X | H] as two 256-bit stores, which spares the read-back of
      // half-written sectors that leave L2 (33 MB per launch) but costs the store path more: 78.9 -> 81.6 us
      double* out = lm_rec + kLmRec * static_cast<size_t>(st.lm);
      const unsigned long long keep = l2_keep();
      store2(out + kLmRecH0, H[0], H[1], keep);
      store2(out + kLmRecH2, H[2], H[3], keep);
    }
  }

  // Long landmarks (one warp each).  Inside a series they are handed out through a counter to the warps as they
  // finish their slices, so the blocks that are done early take them (which warp makes a landmark's sum does not
  // change it); outside a series (no control block) they are dealt round robin.
  __device__ __forceinline__ void finish(Lane& /*st*/, const DeviceIndex& ix, const CamWindow& win,
                                         const double* __restrict__ cam_rec) const {
    if (ix.num_long == 0) return;
    if (queue == nullptr) {
      const int warps = static_cast<int>(blockDim.x >> 5);
      const int total_warps = static_cast<int>(gridDim.x) * warps;
      for (int k = static_cast<int>(blockIdx.x) * warps + static_cast<int>(threadIdx.x >> 5); k < ix.num_long;
           k += total_warps) {
        long_landmark_warp<JOINT, HASW>(ix, win, k, X, cam_rec, obs_d, obs_w, c1, c2, lm_fold, lm_rec);
      }
      return;
    }
    const int lane = threadIdx.x & 31;
    for (;;) {
      unsigned int k = 0;
      if (lane == 0) k = atomicAdd(&queue->long_next, 1u);
      k = __shfl_sync(kFullMask, k, 0);
      if (k >= static_cast<unsigned int>(ix.num_long)) break;
      long_landmark_warp<JOINT, HASW>(ix, win, static_cast<int>(k), X, cam_rec, obs_d, obs_w, c1, c2, lm_fold, lm_rec);
    }
    // the last block to leave puts the queue back for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&queue->long_left, 1u) == gridDim.x - 1) {
        queue->long_next = 0u;
        queue->long_left = 0u;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// camera half.  One warp per work item (a run of CSC entries of one camera, kernels_camera.cu),
// two lanes per entry, sixteen entries per step, two steps per trip.  Lane j owns X[2j..2j+1]
// and H[2j..2j+1] of the landmark record (one 256-bit load, kLmRec) and the matching two columns of M; the
// 3-vector M H is completed with one exchange, and each lane accumulates its six entries of m (x) X.  The
// landmark indices of the next trip are loaded before the records of this one are used.
// The kernel waits for the record gather (long_scoreboard; L2, half of it DRAM).  Measured and dropped: a ring
// of register stages refilled right after use (ptxas puts the loads of all stages on one scoreboard, so every
// step waits for the youngest load: 63 -> 118 us); the gathers as LDGSTS copies into a per-warp shared-memory ring,
// four or six steps deep (copying in and reading out costs more L1 wavefronts than the depth wins: 75 / 85 us);
// four blocks per SM at 64 registers (40 bytes of spills: 79 us).
// ------------------------------------------------------------------------------------------
template <bool JOINT, bool HASW, int kSteps, int kOcc = (kSteps <= 2 ? 3 : 2)>
__global__ void __launch_bounds__(kBlock, kOcc)
k_passB_e0_v2(DeviceIndex ix, const double* __restrict__ P, const double* __restrict__ lm_rec,
              const double* __restrict__ csc_d, const double* __restrict__ csc_w, double c1, double c2,
              double* __restrict__ item_part, const SeriesCtl* __restrict__ ctl) {
  if (ctl != nullptr && ctl->done) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= ix.num_items) return;
  const int pr = lane >> 1, j = lane & 1;
  const int c = __ldg(ix.item_cam + warp);
  const int eb = __ldg(ix.item_ptr + warp), ee = __ldg(ix.item_ptr + warp + 1);
  const unsigned long long keep = l2_keep(), stream = l2_stream();   // records / entry lists (device_math.cuh)
  int lmn[kSteps];
#pragma unroll
  for (int s = 0; s < kSteps; ++s) lmn[s] = ldg1(ix.csc_lm + min(eb + 16 * s + pr, ee - 1), stream);
  double Ma[3], Mb[3];   // M[r][2j], M[r][2j+1]
  {
    const double* p = P + 12 * static_cast<size_t>(c);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double2 t = ldg2(p + 4 * r + 2 * j);
      Ma[r] = t.x;
      Mb[r] = (JOINT || j == 0) ? t.y : 0.0;   // step 1: M = P[:, 0:3]
    }
  }
  double acc[3][2];
#pragma unroll
  for (int k = 0; k < 3; ++k) acc[k][0] = acc[k][1] = 0.0;
  for (int e0 = eb; e0 < ee; e0 += 16 * kSteps) {
    double2 xe[kSteps], he[kSteps];
    ObsCoef kc[kSteps];
    bool act[kSteps];
#pragma unroll
    for (int s = 0; s < kSteps; ++s) {
      const int e = e0 + 16 * s + pr;
      act[s] = e < ee;
      const int ec = act[s] ? e : ee - 1;
      double t[4];   // this lane's half of the record: X[2j], X[2j+1], H[2j], H[2j+1]
      load4_256(lm_rec + kLmRec * static_cast<size_t>(lmn[s]) + 4 * j, t, keep);
      xe[s] = make_double2(t[0], t[1]);
      he[s] = make_double2(t[2], t[3]);
      if (JOINT) {
        const double* dp = csc_d + 3 * static_cast<size_t>(ec);
        kc[s].a = ldg1(dp, stream);
        kc[s].b = ldg1(dp + 1, stream);
        kc[s].c = ldg1(dp + 2, stream);
      } else {
        const double2 uv = load2(ix.csc_uv + ec, stream);
        kc[s].a = uv.x;
        kc[s].b = uv.y;
        kc[s].c = HASW ? ldg1(csc_w + ec, stream) : 1.0;
      }
    }
    if (e0 + 16 * kSteps < ee) {
#pragma unroll
      for (int s = 0; s < kSteps; ++s) lmn[s] = ldg1(ix.csc_lm + min(e0 + 16 * kSteps + 16 * s + pr, ee - 1), stream);
    }
#pragma unroll
    for (int s = 0; s < kSteps; ++s) {
      double v[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double part = Ma[r] * he[s].x + Mb[r] * he[s].y;
        v[r] = part + __shfl_xor_sync(kFullMask, part, 1);
      }
      double m[3];
      if (JOINT) {
        joint_normal_coef(v[0], v[1], v[2], kc[s].a, kc[s].b, kc[s].c, m);
      } else {
        pose_normal_coef(v[0], v[1], v[2], kc[s].a, kc[s].b, kc[s].c, c1, c2, m);
      }
      if (act[s]) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          acc[k][0] += m[k] * xe[s].x;
          acc[k][1] += m[k] * xe[s].y;
        }
      }
    }
  }
#pragma unroll
  for (int off = 2; off < 32; off <<= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      acc[k][0] += __shfl_xor_sync(kFullMask, acc[k][0], off);
      acc[k][1] += __shfl_xor_sync(kFullMask, acc[k][1], off);
    }
  }
  if (lane < 2) {
    double* out = item_part + 12 * static_cast<size_t>(warp) + 2 * j;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      *reinterpret_cast<double2*>(out + 4 * k) = make_double2(acc[k][0], acc[k][1]);
    }
  }
}

// matrix part of the camera records (after a linearisation: P is fixed until the next one)
template <bool JOINT>
__global__ void __launch_bounds__(kBlock)
k_cam_rec_static(int C, const double* __restrict__ P, double* __restrict__ cam_rec) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * 4) return;
  const int c = idx >> 2, n = idx & 3;
  const double* p = P + 12 * static_cast<size_t>(c);
  double* r = cam_rec + CamRec::stride(JOINT) * static_cast<size_t>(c);
  if (JOINT || n < 3) {
    r[CamRec::m_index(JOINT, 0, n)] = p[n];
    r[CamRec::m_index(JOINT, 1, n)] = p[4 + n];
    r[CamRec::m_index(JOINT, 2, n)] = p[8 + n];
  } else {
    r[CamRec::stride(JOINT) - 1] = 0.0;   // the pad of the step-1 record
  }
  if (JOINT && n == 0) r[24] = r[25] = 0.0;
}

template <bool JOINT, bool HASW>
void launch_landmark_half(const DeviceState& d, const ModelParams& mp, const SeriesCtl* ctl,
                          const LaunchCfg& lc) {
  const E0LandmarkOp<JOINT, HASW> op{d.X,         d.sell_d, d.sell_w, mp.c1,   mp.c2, d.lm_fold, d.sell_x,
                                     d.sell_fold, d.lm_rec, d.obs_d,  d.obs_w, ctl,   ctl != nullptr ? d.ctl : nullptr};
  if (launch_sell_walk(d.ix, d.plan[JOINT ? 1 : (HASW ? 2 : 0)], d.debug_window_cams, d.cam_rec, op, lc.stream)) {
    count(lc);
  }
}

}  // namespace

// tuning builds (-DPOVAR_WALK_TRACE): stamps of the last walk kernels of this translation unit, reset on read
int debug_walk_trace(unsigned long long* out, int n) {
#ifdef POVAR_WALK_TRACE
  if (cudaMemcpyFromSymbol(out, g_walk_trace, sizeof(unsigned long long) * n) != cudaSuccess) return POVAR_ERR_CUDA;
  static unsigned long long zeros[4 * 1024];
  if (cudaMemcpyToSymbol(g_walk_trace, zeros, sizeof(zeros)) != cudaSuccess) return POVAR_ERR_CUDA;
  return POVAR_OK;
#else
  (void)out;
  (void)n;
  return POVAR_ERR_UNSUPPORTED;
#endif
}

void launch_cam_rec_static(const DeviceState& d, bool joint, const LaunchCfg& lc) {
  const int blocks = (d.ix.C * 4 + kBlock - 1) / kBlock;
  if (joint) {
    k_cam_rec_static<true><<<blocks, kBlock, 0, lc.stream>>>(d.ix.C, d.P, d.cam_rec);
  } else {
    k_cam_rec_static<false><<<blocks, kBlock, 0, lc.stream>>>(d.ix.C, d.P, d.cam_rec);
  }
  count(lc);
}

void launch_e0_landmark_v2(const DeviceState& d, const ModelParams& mp, bool joint, bool in_series,
                           const LaunchCfg& lc) {
  const SeriesCtl* ctl = in_series ? d.ctl : nullptr;
  const bool hasw = !joint && mp.robust_norm == NORM_HUBER;
  if (joint) launch_landmark_half<true, false>(d, mp, ctl, lc);
  else if (hasw) launch_landmark_half<false, true>(d, mp, ctl, lc);
  else launch_landmark_half<false, false>(d, mp, ctl, lc);
}

void launch_passB_e0_v2(const DeviceState& d, const ModelParams& mp, bool joint, bool in_series,
                        const LaunchCfg& lc) {
  if (d.ix.num_items == 0) return;
  const int blocks = (d.ix.num_items + kWarps - 1) / kWarps;
  const SeriesCtl* ctl = in_series ? d.ctl : nullptr;
  const bool hasw = !joint && mp.robust_norm == NORM_HUBER;
  // two steps of sixteen entries in flight per trip (more was measured slower: registers, profiles/r1_summary.md)
#define POVAR_PASSB(J, W, S, CD, CW)                                                                   \
  k_passB_e0_v2<J, W, S><<<blocks, kBlock, 0, lc.stream>>>(d.ix, d.P, d.lm_rec, CD, CW, mp.c1, mp.c2, \
                                                           d.item_part, ctl)
  if (joint) POVAR_PASSB(true, false, 2, d.csc_d, nullptr);
  else if (hasw) POVAR_PASSB(false, true, 2, nullptr, d.csc_w);
  else POVAR_PASSB(false, false, 2, nullptr, nullptr);
#undef POVAR_PASSB
  count(lc);
}

}  // namespace povar
