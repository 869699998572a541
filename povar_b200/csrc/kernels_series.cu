// The two halves of one power-series term,  raw_c = sum_l Jp^T Jl Hll^-1 Jl^T Jp y,  in the
// lane-group layout.  Replaces right_mul_e0_pOSE / right_mul_e0_joint
// (/root/reference/src/rootba_povar/sc/linearization_power_varproj.hpp:364-453).
//
// Why not one observation per lane (kernels_landmark.cu k_e0_landmark, kernels_camera.cu k_passB):
// every observation needs ~170 bytes of its camera (landmark half) or 64 bytes of its landmark
// (camera half) from a table that lives in L1/L2.  With one observation per lane each LDG.128
// touches 32 different cache lines; L1TEX, not HBM, bounded those kernels (tools/ubench_gather.cu,
// profiles/).  Here a small group of lanes shares an observation and reads consecutive 16-byte
// chunks of the record (whole sectors), the arithmetic is split along the same lines (each lane
// owns one row / column of the 3x4 blocks), and the only cross-lane traffic is a 3-value exchange
// per observation.
//
// Both observation models have  Jp_raw = K (x) X^T  and  Jl_raw = K M  with a small K that depends on
// the observation only (step 1: K_i(u, v, c1, c2) 4x3, M = P[:, 0:3]; step 2: K_i = d_i 2x3, M = P), so
//   landmark half:  G_l = sum_i M^T (K^T W K) (Y_c X_l),   H_l = fold_l G_l
//   camera half:    raw_c = sum_i ((K^T W K) (M H_l)) (x) X_l
// where Y_c is y_c as a 3x4 matrix and fold_l = S (Pi) Hll^-1 (Pi^T) S is made once per solve
// (k_prep_landmark).  Step 2 streams sqrt(w) d_i, stored at the linearisation point.
#include <cuda_runtime.h>

#include <cstdlib>

#include "device_math.cuh"
#include "povar_internal.h"

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

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
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

// One observation of the landmark half, seen from one lane of its 4-lane group.  L0..L2 are the
// lane's three chunks of the camera record (CamRec): lanes 0..2 make q_sub = y_sub . X, everybody
// receives q, and the products with M are accumulated where the entries of M live:
//   lanes 0..2:  acc[0] += M[0][sub] m0 + M[1][sub] m1
//   lane 3:      acc[n] += M[2][n] m2  (n = 0..3),  acc[3] += M[0][3] m0 + M[1][3] m1
// (combined once per landmark by group_totals).
struct ObsCoef {
  double a, b, c;   // step 1: u, v, w;  step 2: sqrt(w) (1/z, -x/z^2, -y/z^2)
};

template <bool JOINT>
__device__ __forceinline__ void landmark_obs(const double2& L0, const double2& L1, const double2& L2,
                                             const double (&x)[4], const ObsCoef& k, double c1,
                                             double c2, int sub, int gb, bool act, double (&acc)[4]) {
  const double q = L0.x * x[0] + L0.y * x[1] + L1.x * x[2] + L1.y * x[3];
  const double q0 = __shfl_sync(kFullMask, q, gb);
  const double q1 = __shfl_sync(kFullMask, q, gb + 1);
  const double q2 = __shfl_sync(kFullMask, q, gb + 2);
  double m[3];
  if (JOINT) {
    joint_normal_coef(q0, q1, q2, k.a, k.b, k.c, m);
  } else {
    pose_normal_coef(q0, q1, q2, k.a, k.b, k.c, c1, c2, m);
  }
  if (act) {
    const double e0 = L2.x * m[0] + L2.y * m[1];
    acc[0] += sub < 3 ? e0 : L0.x * m[2];
    acc[1] += L0.y * m[2];
    acc[2] += L1.x * m[2];
    acc[3] += L1.y * m[2] + e0;
  }
}

// G_l on every lane of the group from the per-lane accumulators above
template <bool JOINT>
__device__ __forceinline__ void group_totals(const double (&acc)[4], int gb, double (&G)[4]) {
  G[0] = __shfl_sync(kFullMask, acc[0], gb) + __shfl_sync(kFullMask, acc[0], gb + 3);
  G[1] = __shfl_sync(kFullMask, acc[0], gb + 1) + __shfl_sync(kFullMask, acc[1], gb + 3);
  G[2] = __shfl_sync(kFullMask, acc[0], gb + 2) + __shfl_sync(kFullMask, acc[2], gb + 3);
  G[3] = JOINT ? __shfl_sync(kFullMask, acc[3], gb + 3) : 0.0;
}

// H[sub] = row `sub` of fold_l times G  (fold packed symmetric: 3x3 in step 1, 4x4 in step 2)
template <bool JOINT>
__device__ __forceinline__ double fold_row(const double* __restrict__ f, int sub, const double (&G)[4]) {
  if (JOINT) {
    double F[10];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double2 t = ldg2(f + 2 * k);
      F[2 * k] = t.x;
      F[2 * k + 1] = t.y;
    }
    const double r0 = sub == 0 ? F[0] : (sub == 1 ? F[1] : (sub == 2 ? F[2] : F[3]));
    const double r1 = sub == 0 ? F[1] : (sub == 1 ? F[4] : (sub == 2 ? F[5] : F[6]));
    const double r2 = sub == 0 ? F[2] : (sub == 1 ? F[5] : (sub == 2 ? F[7] : F[8]));
    const double r3 = sub == 0 ? F[3] : (sub == 1 ? F[6] : (sub == 2 ? F[8] : F[9]));
    return r0 * G[0] + r1 * G[1] + r2 * G[2] + r3 * G[3];
  }
  double F[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double2 t = ldg2(f + 2 * k);
    F[2 * k] = t.x;
    F[2 * k + 1] = t.y;
  }
  const double r0 = sub == 0 ? F[0] : (sub == 1 ? F[1] : F[2]);
  const double r1 = sub == 0 ? F[1] : (sub == 1 ? F[3] : F[4]);
  const double r2 = sub == 0 ? F[2] : (sub == 1 ? F[4] : F[5]);
  return sub < 3 ? r0 * G[0] + r1 * G[1] + r2 * G[2] : 0.0;
}

// ------------------------------------------------------------------------------------------
// landmark half, sliced-ELL order (DeviceIndex::slice_ptr ...): a warp walks a contiguous range of
// slices; a slice is eight landmarks of (nearly) equal degree, four lanes per landmark.  The group
// visits the observations of its landmark in camera order and keeps its part of G_l in registers:
// no tile table, no shared memory, no segmented reduction.
//
// Software pipeline: the rows of a warp's range are contiguous and every slice has an even number
// of rows, so rows alternate between two register sets A / B regardless of slice boundaries; the
// record of row r + 1 is requested before row r is computed, camera indices run two rows ahead, and
// the observation stream is pulled into L2 kStreamAhead rows ahead.
// ------------------------------------------------------------------------------------------
struct RowData {
  double2 L0, L1, L2;
  ObsCoef k;
  bool act;
};

template <bool JOINT, bool HASW>
__device__ __forceinline__ void load_row(const DeviceIndex& ix, const double* __restrict__ cam_rec,
                                         const double* __restrict__ sell_d,
                                         const double* __restrict__ sell_w, int row, int c, int grp,
                                         int sub, RowData& d) {
  const size_t slot = 8 * static_cast<size_t>(row) + grp;
  d.act = c >= 0;
  const double* r = cam_rec + CamRec::kStride * static_cast<size_t>(d.act ? c : 0) + 2 * sub;
  d.L0 = ldg2(r);
  d.L1 = ldg2(r + 8);
  d.L2 = ldg2(r + 16);
  if (JOINT) {
    const double* dp = sell_d + 3 * slot;
    d.k.a = __ldcs(dp);
    d.k.b = __ldcs(dp + 1);
    d.k.c = __ldcs(dp + 2);
  } else {
    const double2 uv = __ldcs(ix.sell_uv + slot);
    d.k.a = uv.x;
    d.k.b = uv.y;
    d.k.c = HASW ? __ldcs(sell_w + slot) : 1.0;
  }
}

// landmarks with more than 32 observations (outside the sliced-ELL set): one warp per landmark (the
// first blocks of the grid of k_e0_landmark_sell, so that they overlap with the slices),
// group g takes observations g, g + 8, ... of the CSR list, fixed-tree sum over the groups
template <bool JOINT, bool HASW>
__device__ __forceinline__ void long_landmark_warp(const DeviceIndex& ix, int warp,
                                                   const double* __restrict__ X,
                                                   const double* __restrict__ cam_rec,
                                                   const double* __restrict__ obs_d,
                                                   const double* __restrict__ obs_w, double c1, double c2,
                                                   const double* __restrict__ lm_fold,
                                                   double* __restrict__ lm_rec) {
  const int lane = threadIdx.x & 31;
  const int grp = lane >> 2, sub = lane & 3, gb = lane & ~3;
  if (warp >= ix.num_long) return;
  const int lm = __ldg(ix.long_lm + warp);
  const int ob = __ldg(ix.lm_ptr + lm), oe = __ldg(ix.lm_ptr + lm + 1);
  double x[4];
  load4(X + 4 * static_cast<size_t>(lm), x);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int base = ob; base < oe; base += 8) {
    const bool act = base + grp < oe;
    const int o = act ? base + grp : ob;
    const int c = __ldg(ix.obs_cam + o);
    const double* r = cam_rec + CamRec::kStride * static_cast<size_t>(c) + 2 * sub;
    const double2 A0 = ldg2(r), A1 = ldg2(r + 8), A2 = ldg2(r + 16);
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
    landmark_obs<JOINT>(A0, A1, A2, x, k, c1, c2, sub, gb, act, acc);
  }
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n] += __shfl_xor_sync(kFullMask, acc[n], off);
  }
  double G[4];
  group_totals<JOINT>(acc, gb, G);
  const double H = fold_row<JOINT>(lm_fold + 10 * static_cast<size_t>(lm), sub, G);
  if (grp == 0) lm_rec[kLmRec * static_cast<size_t>(lm) + 4 + sub] = H;
}

constexpr int kStreamAhead = 24;   // rows

// per-warp state of the slice being accumulated
template <bool JOINT>
struct SliceState {
  int lm, row1;        // landmark of this group (-1: none), first row after the slice
  int lm_n, row1_n;    // the same for the next slice (loaded one slice ahead)
  double x[4];
  double fr[4];        // row `sub` of fold_l
  double acc[4];
};

template <bool JOINT, bool HASW, int NR>
__global__ void __launch_bounds__(kBlock, NR == 1 ? 4 : 3)
k_e0_landmark_sell(DeviceIndex ix, const double* __restrict__ X, const double* __restrict__ cam_rec,
                   const double* __restrict__ sell_d, const double* __restrict__ sell_w, double c1,
                   double c2, const double* __restrict__ lm_fold, double* __restrict__ lm_rec,
                   const double* __restrict__ obs_d, const double* __restrict__ obs_w,
                   const SeriesCtl* __restrict__ ctl, int slices_per_warp, int stream_ahead, int long_blocks) {
  if (ctl != nullptr && ctl->done) return;
  if (static_cast<int>(blockIdx.x) < long_blocks) {
    long_landmark_warp<JOINT, HASW>(ix, blockIdx.x * kWarps + (threadIdx.x >> 5), X, cam_rec, obs_d, obs_w,
                                    c1, c2, lm_fold, lm_rec);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int grp = lane >> 2, sub = lane & 3, gb = lane & ~3;
  // Blocks that share an SM should work on neighbouring slices (same stretch of the camera table in
  // L1).  With a grid of 148 k blocks, all resident, blocks b, b + 148, ... land on the same SM.
  const int bid = static_cast<int>(blockIdx.x) - long_blocks, nblk = static_cast<int>(gridDim.x) - long_blocks;
  const int per_sm = nblk / 148;
  const int chunk = (per_sm * 148 == nblk) ? (bid % 148) * per_sm + bid / 148 : bid;
  const int warp = chunk * kWarps + (threadIdx.x >> 5);
  const int s0 = warp * slices_per_warp;
  if (s0 >= ix.num_slices) return;
  const int s1 = min(s0 + slices_per_warp, ix.num_slices);
  const int row_first = __ldg(ix.slice_ptr + s0);
  const int row_end = __ldg(ix.slice_ptr + s1);        // one past the last row of this warp
  const int row_last = row_end - 1;
  // indices of the packed symmetric fold matrix that make up row `sub`
  const int f0 = sub;
  const int f1 = JOINT ? (sub == 0 ? 1 : sub + 3) : (sub == 0 ? 1 : sub + 2);
  const int f2 = JOINT ? (sub == 0 ? 2 : (sub == 1 ? 5 : sub + 5)) : (sub == 0 ? 2 : sub + 3);
  const int f3 = sub == 0 ? 3 : (sub == 1 ? 6 : sub + 6);

  SliceState<JOINT> st;
  int sl = s0;
  auto open_slice = [&]() {
    // st.lm / st.row1 are set; fetch this slice's landmark data and the header of the next slice
    const int sn = min(sl + 1, s1 - 1);
    st.lm_n = __ldg(ix.sell_lm + 8 * sn + grp);
    st.row1_n = __ldg(ix.slice_ptr + sn + 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) st.x[k] = st.fr[k] = st.acc[k] = 0.0;
    if (st.lm >= 0) {
      // streamed once per term: keep them from displacing the camera records in L1
      const double2* xp = reinterpret_cast<const double2*>(X + 4 * static_cast<size_t>(st.lm));
      const double2 xa = __ldcs(xp), xb = __ldcs(xp + 1);
      st.x[0] = xa.x;
      st.x[1] = xa.y;
      st.x[2] = xb.x;
      st.x[3] = xb.y;
      const double* f = lm_fold + 10 * static_cast<size_t>(st.lm);
      if (JOINT || sub < 3) {
        st.fr[0] = __ldcs(f + f0);
        st.fr[1] = __ldcs(f + f1);
        st.fr[2] = __ldcs(f + f2);
        if (JOINT) st.fr[3] = __ldcs(f + f3);
      }
    }
    if (st.lm_n >= 0 && sub == 0) {
      prefetch_l2(X + 4 * static_cast<size_t>(st.lm_n));
      prefetch_l2(lm_fold + 10 * static_cast<size_t>(st.lm_n));
    }
  };
  auto close_slice = [&]() {
    double G[4];
    group_totals<JOINT>(st.acc, gb, G);
    if (st.lm >= 0) {
      const double H = st.fr[0] * G[0] + st.fr[1] * G[1] + st.fr[2] * G[2] + st.fr[3] * G[3];
      __stcs(lm_rec + kLmRec * static_cast<size_t>(st.lm) + 4 + sub, H);
    }
    ++sl;
    st.lm = st.lm_n;
    st.row1 = st.row1_n;
  };

  st.lm = __ldg(ix.sell_lm + 8 * s0 + grp);
  st.row1 = __ldg(ix.slice_ptr + s0 + 1);
  // ring of NR rows in flight; camera indices run kCamAhead rows ahead of the records
  constexpr int kCamAhead = 4;
  constexpr int kUnroll = NR * kCamAhead;   // both rings keep static indices
  RowData ring[NR];
  int camq[kCamAhead];                      // camq[j]: camera index of row (next record row) + j
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int r = min(row_first + i, row_last);
    const int c = __ldcs(ix.sell_cam + 8 * static_cast<size_t>(r) + grp);
    load_row<JOINT, HASW>(ix, cam_rec, sell_d, sell_w, r, c, grp, sub, ring[i]);
  }
#pragma unroll
  for (int j = 0; j < kCamAhead; ++j) {
    camq[j] = __ldcs(ix.sell_cam + 8 * static_cast<size_t>(min(row_first + NR + j, row_last)) + grp);
  }
  open_slice();
  for (int row = row_first; row < row_end; row += kUnroll) {
    // pull the observation stream towards L2 ahead of the loads (one 128-byte line of uv per row)
    if (stream_ahead > 0 && lane < kUnroll) {
      const int rp = row + stream_ahead + lane;
      if (rp <= row_last) {
        if (JOINT) {
          prefetch_l2(sell_d + 24 * static_cast<size_t>(rp));
          prefetch_l2(sell_d + 24 * static_cast<size_t>(rp) + 16);
        } else {
          prefetch_l2(ix.sell_uv + 8 * static_cast<size_t>(rp));
        }
        if ((lane & 3) == 0) prefetch_l2(ix.sell_cam + 8 * static_cast<size_t>(rp));
      }
    }
#pragma unroll
    for (int i = 0; i < kUnroll; ++i) {
      const int r = row + i;
      if (r < row_end) {           // warp-uniform
        if (r == st.row1) {        // warp-uniform: the previous slice is complete
          close_slice();
          open_slice();
        }
        RowData& cur = ring[i % NR];
        landmark_obs<JOINT>(cur.L0, cur.L1, cur.L2, st.x, cur.k, c1, c2, sub, gb, cur.act, st.acc);
        load_row<JOINT, HASW>(ix, cam_rec, sell_d, sell_w, min(r + NR, row_last), camq[i % kCamAhead],
                              grp, sub, cur);
        camq[i % kCamAhead] =
            __ldcs(ix.sell_cam + 8 * static_cast<size_t>(min(r + NR + kCamAhead, row_last)) + grp);
      }
    }
  }
  close_slice();
}

// ------------------------------------------------------------------------------------------
// camera half.  One warp per work item (a run of CSC entries of one camera, kernels_camera.cu),
// two lanes per entry, sixteen entries per step, two steps per trip.  Lane j owns X[2j..2j+1]
// and H[2j..2j+1] of the landmark record and the matching two columns of M; the 3-vector M H is
// completed with one exchange, and each lane accumulates its six entries of m (x) X.  The
// landmark indices of the next trip are loaded before the records of this one are used.
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
  int lmn[kSteps];
#pragma unroll
  for (int s = 0; s < kSteps; ++s) lmn[s] = __ldg(ix.csc_lm + min(eb + 16 * s + pr, ee - 1));
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
      const double* rp = lm_rec + kLmRec * static_cast<size_t>(lmn[s]) + 2 * j;
      xe[s] = ldg2(rp);
      he[s] = ldg2(rp + 4);
      if (JOINT) {
        const double* dp = csc_d + 3 * static_cast<size_t>(ec);
        kc[s].a = __ldg(dp);
        kc[s].b = __ldg(dp + 1);
        kc[s].c = __ldg(dp + 2);
      } else {
        const double2 uv = ix.csc_uv[ec];
        kc[s].a = uv.x;
        kc[s].b = uv.y;
        kc[s].c = HASW ? __ldg(csc_w + ec) : 1.0;
      }
    }
    if (e0 + 16 * kSteps < ee) {
#pragma unroll
      for (int s = 0; s < kSteps; ++s) lmn[s] = __ldg(ix.csc_lm + min(e0 + 16 * kSteps + 16 * s + pr, ee - 1));
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
  double* r = cam_rec + CamRec::kStride * static_cast<size_t>(c);
  const bool used = JOINT || n < 3;
  r[CamRec::m_index(0, n)] = used ? p[n] : 0.0;
  r[CamRec::m_index(1, n)] = used ? p[4 + n] : 0.0;
  r[CamRec::m_index(2, n)] = used ? p[8 + n] : 0.0;
}

template <bool JOINT, bool HASW>
void launch_landmark_half(const DeviceState& d, const ModelParams& mp, const SeriesCtl* ctl,
                          const LaunchCfg& lc) {
  if (d.ix.num_slices == 0 && d.ix.num_long == 0) return;
  // one wave of persistent-style blocks for the slices: 148 SMs x 4 resident blocks x 8 warps,
  // contiguous slice ranges per warp; in front of them one warp per long landmark.  Small problems and
  // small shards (a venice-1778 shard on 8 GPUs has 15 k slices) keep the wave full with short ranges:
  // a warp walks its rows one dependent gather after the other, so the launch lasts as long as the
  // longest range (64 us for 16 slices, whatever the problem size).
  const long long full = 148LL * 4 * kWarps;
  long long per_warp = (d.ix.num_slices + full - 1) / full;
  static const int min_per_warp = getenv("POVAR_SELL_MIN_SLICES") ? atoi(getenv("POVAR_SELL_MIN_SLICES")) : 2;
  if (per_warp < min_per_warp) per_warp = min_per_warp;
  const long long warps = (d.ix.num_slices + per_warp - 1) / per_warp;
  int blocks = static_cast<int>((warps + kWarps - 1) / kWarps);
  if (blocks > 148) blocks = (blocks + 147) / 148 * 148;   // whole multiples of 148: see the chunk map
  const int long_blocks = (d.ix.num_long + kWarps - 1) / kWarps;
  static const int nr = getenv("POVAR_SELL_NR") ? atoi(getenv("POVAR_SELL_NR")) : 1;
  static const int ahead = getenv("POVAR_SELL_AHEAD") ? atoi(getenv("POVAR_SELL_AHEAD")) : kStreamAhead;
#define POVAR_SELL_LAUNCH(NRV)                                                                     \
  {                                                                                                \
    static const cudaError_t carve_##NRV = cudaFuncSetAttribute(                                   \
        k_e0_landmark_sell<JOINT, HASW, NRV>, cudaFuncAttributePreferredSharedMemoryCarveout,      \
        cudaSharedmemCarveoutMaxL1); /* no shared memory: all of it to L1 (the camera table) */    \
    (void)carve_##NRV;                                                                             \
    k_e0_landmark_sell<JOINT, HASW, NRV><<<blocks + long_blocks, kBlock, 0, lc.stream>>>(          \
        d.ix, d.X, d.cam_rec, d.sell_d, d.sell_w, mp.c1, mp.c2, d.lm_fold, d.lm_rec, d.obs_d,      \
        d.obs_w, ctl, static_cast<int>(per_warp), ahead, long_blocks);                             \
  }
  if (nr == 2) POVAR_SELL_LAUNCH(2)
  else POVAR_SELL_LAUNCH(1)
#undef POVAR_SELL_LAUNCH
  count(lc);
}

}  // namespace

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
  // steps (of sixteen entries) a warp keeps in flight per trip: the pass is bound by the number of
  // landmark records in flight (Little's law), registers permitting
  static const int steps = getenv("POVAR_PASSB_STEPS") ? atoi(getenv("POVAR_PASSB_STEPS")) : 2;
#define POVAR_PASSB(J, W, S, CD, CW)                                                                   \
  k_passB_e0_v2<J, W, S><<<blocks, kBlock, 0, lc.stream>>>(d.ix, d.P, d.lm_rec, CD, CW, mp.c1, mp.c2, \
                                                           d.item_part, ctl)
  if (steps == 12) {   // two steps, four blocks per SM (64 registers)
    if (joint) {
      k_passB_e0_v2<true, false, 2, 4><<<blocks, kBlock, 0, lc.stream>>>(d.ix, d.P, d.lm_rec, d.csc_d, nullptr, mp.c1,
                                                                         mp.c2, d.item_part, ctl);
    } else {
      k_passB_e0_v2<false, false, 2, 4><<<blocks, kBlock, 0, lc.stream>>>(d.ix, d.P, d.lm_rec, nullptr, nullptr,
                                                                          mp.c1, mp.c2, d.item_part, ctl);
    }
  } else if (joint) {
    if (steps >= 4) POVAR_PASSB(true, false, 4, d.csc_d, nullptr);
    else if (steps == 3) POVAR_PASSB(true, false, 3, d.csc_d, nullptr);
    else POVAR_PASSB(true, false, 2, d.csc_d, nullptr);
  } else if (hasw) {
    POVAR_PASSB(false, true, 2, nullptr, d.csc_w);
  } else {
    if (steps >= 4) POVAR_PASSB(false, false, 4, nullptr, nullptr);
    else if (steps == 3) POVAR_PASSB(false, false, 3, nullptr, nullptr);
    else POVAR_PASSB(false, false, 2, nullptr, nullptr);
  }
#undef POVAR_PASSB
  count(lc);
}

}  // namespace povar
