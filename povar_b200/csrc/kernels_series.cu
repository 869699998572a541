// The two halves of one power-series term,  raw_c = sum_l Jp^T Jl Hll^-1 Jl^T Jp y,  in the
// lane-group layout.  Replaces right_mul_e0_pOSE / right_mul_e0_joint
// (/root/reference/src/rootba_povar/sc/linearization_power_varproj.hpp:364-453).
//
// Why not one observation per lane (kernels_landmark.cu k_e0_landmark, kernels_camera.cu k_passB):
// every observation needs ~170 bytes of its camera (landmark half) or 64 bytes of its landmark
// (camera half) from a table that lives in L1/L2.  With one observation per lane each LDG.128
// touches 32 different cache lines; L1TEX serves about one line per clock, and that -- not HBM --
// bounded the old kernels (tools/ubench_gather.cu, profiles/r1_summary.md).  Here a small group of
// lanes shares an observation and reads consecutive 16-byte chunks of the record, the arithmetic
// is split along the same lines (each lane owns one row / column of the 3x4 blocks), and the only
// cross-lane traffic is a 3-value exchange per observation.
//
// Both observation models have  Jp_raw = K (x) X^T  and  Jl_raw = K M  with a small K that depends on
// the observation only (step 1: K_i(u, v, c1, c2) 4x3, M = P[:, 0:3]; step 2: K_i = d_i 2x3, M = P), so
//   landmark half:  G_l = sum_i M^T (K^T W K) (Y_c X_l),   H_l = fold_l G_l
//   camera half:    raw_c = sum_i ((K^T W K) (M H_l)) (x) X_l
// where Y_c is y_c as a 3x4 matrix and fold_l = S (Pi) Hll^-1 (Pi^T) S is made once per solve
// (k_prep_landmark).  Step 2 streams sqrt(w) d_i, stored at the linearisation point (obs_d / csc_d).
#include <cuda_runtime.h>

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

// ------------------------------------------------------------------------------------------
// landmark half.  One warp per tile (kernels_landmark.cu), four lanes per observation, eight
// observations per step.  Lane `sub` owns y_sub (row sub of Y_c) and column sub of M: it makes
// q_sub = y_sub . X, receives the other two q, and produces component sub of M^T m.  The
// per-observation components go through a small shared-memory buffer; one lane per landmark
// adds them in observation order and applies fold_l.
// ------------------------------------------------------------------------------------------
template <bool JOINT, bool HASW>
__global__ void __launch_bounds__(kBlock, 2)
k_e0_landmark_v2(DeviceIndex ix, const double* __restrict__ X, const double* __restrict__ cam_rec,
                 const double* __restrict__ obs_d, const double* __restrict__ obs_w, double c1,
                 double c2, const double* __restrict__ lm_fold, double* __restrict__ lm_rec,
                 const SeriesCtl* __restrict__ ctl, const int4* __restrict__ tile_info, int num_tiles,
                 int tiles_per_block) {
  if (ctl != nullptr && ctl->done) return;
  constexpr int NV = JOINT ? 4 : 3;
  using R = CamRec<JOINT>;
  __shared__ double gbuf_all[kWarps][4][33];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane >> 2, sub = lane & 3, gb = lane & ~3;
  double (*gbuf)[33] = gbuf_all[wib];
  const int tile_begin = blockIdx.x * tiles_per_block;
  const int tile_end = min(tile_begin + tiles_per_block, num_tiles);
  for (int tile = tile_begin + wib; tile < tile_end; tile += kWarps) {
    const int4 ti = __ldg(tile_info + tile);
    const int tb = ti.x, n = ti.y, lm_first = ti.z, nl = ti.w;
    double gtot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) gtot[k] = 0.0;
    for (int cb = 0; cb < n; cb += 32) {
      const int m_obs = min(32, n - cb);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int oi = 8 * s + grp;
        const bool act = oi < m_obs;
        const int o = tb + cb + (act ? oi : 0);
        const int c = __ldg(ix.obs_cam + o);
        const int lm = __ldg(ix.obs_lm + o);
        double x[4];
        load4(X + 4 * static_cast<size_t>(lm), x);
        const double* r = cam_rec + R::kStride * static_cast<size_t>(c);
        double2 ya = make_double2(0.0, 0.0), yb = ya, ma = ya, mb = ya;
        if (sub < 3) {
          ya = ldg2(r + 2 * R::chunk(0, sub));
          yb = ldg2(r + 2 * R::chunk(1, sub));
        }
        if (JOINT || sub < 3) {
          ma = ldg2(r + 2 * R::chunk(2, sub));
          mb = ldg2(r + 2 * R::chunk(3, sub));
        }
        const double q = ya.x * x[0] + ya.y * x[1] + yb.x * x[2] + yb.y * x[3];
        const double q0 = __shfl_sync(kFullMask, q, gb);
        const double q1 = __shfl_sync(kFullMask, q, gb + 1);
        const double q2 = __shfl_sync(kFullMask, q, gb + 2);
        double m[3];
        if (JOINT) {
          const double* dp = obs_d + 3 * static_cast<size_t>(o);
          joint_normal_coef(q0, q1, q2, __ldg(dp), __ldg(dp + 1), __ldg(dp + 2), m);
        } else {
          const double2 uv = ix.obs_uv[o];
          const double w = HASW ? __ldg(obs_w + o) : 1.0;
          pose_normal_coef(q0, q1, q2, uv.x, uv.y, w, c1, c2, m);
        }
        const double g = ma.x * m[0] + ma.y * m[1] + mb.x * m[2];
        if (sub < NV) gbuf[sub][oi] = act ? g : 0.0;
      }
      __syncwarp();
      if (n <= 32) {
        // one lane per landmark of the tile (the span may contain landmarks without observations)
        for (int lb = 0; lb < nl; lb += 32) {
          const int li = lb + lane;
          if (li < nl) {
            const int lm = lm_first + li;
            const int b = __ldg(ix.lm_ptr + lm) - tb, e = __ldg(ix.lm_ptr + lm + 1) - tb;
            double G[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) G[k] = 0.0;
            for (int o = b; o < e; ++o) {
#pragma unroll
              for (int k = 0; k < NV; ++k) G[k] += gbuf[k][o];
            }
            const double* f = lm_fold + 10 * static_cast<size_t>(lm);
            double H[4];
            if (JOINT) {
              double F[10];
#pragma unroll
              for (int k = 0; k < 5; ++k) {
                const double2 t = ldg2(f + 2 * k);
                F[2 * k] = t.x;
                F[2 * k + 1] = t.y;
              }
              H[0] = F[0] * G[0] + F[1] * G[1] + F[2] * G[2] + F[3] * G[3];
              H[1] = F[1] * G[0] + F[4] * G[1] + F[5] * G[2] + F[6] * G[3];
              H[2] = F[2] * G[0] + F[5] * G[1] + F[7] * G[2] + F[8] * G[3];
              H[3] = F[3] * G[0] + F[6] * G[1] + F[8] * G[2] + F[9] * G[3];
            } else {
              double F[6];
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const double2 t = ldg2(f + 2 * k);
                F[2 * k] = t.x;
                F[2 * k + 1] = t.y;
              }
              const double G3[3] = {G[0], G[1], G[2]};
              double H3[3];
              sym3_mul(F, G3, H3);
              H[0] = H3[0];
              H[1] = H3[1];
              H[2] = H3[2];
              H[3] = 0.0;
            }
            double2* out = reinterpret_cast<double2*>(lm_rec + kLmRec * static_cast<size_t>(lm) + 4);
            out[0] = make_double2(H[0], H[1]);
            out[1] = make_double2(H[2], H[3]);
          }
        }
      } else {
        // one long landmark: fixed-tree sum of this chunk, chunks added in order
        double v[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = gbuf[k][lane];
        warp_allreduce<NV>(v);
#pragma unroll
        for (int k = 0; k < NV; ++k) gtot[k] += v[k];
      }
      __syncwarp();
    }
    if (n > 32 && lane == 0) {
      const double* f = lm_fold + 10 * static_cast<size_t>(lm_first);
      double H[4];
      if (JOINT) {
        H[0] = f[0] * gtot[0] + f[1] * gtot[1] + f[2] * gtot[2] + f[3] * gtot[NV - 1];
        H[1] = f[1] * gtot[0] + f[4] * gtot[1] + f[5] * gtot[2] + f[6] * gtot[NV - 1];
        H[2] = f[2] * gtot[0] + f[5] * gtot[1] + f[7] * gtot[2] + f[8] * gtot[NV - 1];
        H[3] = f[3] * gtot[0] + f[6] * gtot[1] + f[8] * gtot[2] + f[9] * gtot[NV - 1];
      } else {
        const double F[6] = {f[0], f[1], f[2], f[3], f[4], f[5]};
        const double G3[3] = {gtot[0], gtot[1], gtot[2]};
        double H3[3];
        sym3_mul(F, G3, H3);
        H[0] = H3[0];
        H[1] = H3[1];
        H[2] = H3[2];
        H[3] = 0.0;
      }
      double* out = lm_rec + kLmRec * static_cast<size_t>(lm_first) + 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] = H[k];
    }
  }
}


// ------------------------------------------------------------------------------------------
// landmark half, sliced-ELL order (DeviceIndex::slice_ptr ...): one warp per slice of eight
// landmarks of (nearly) equal degree, four lanes per landmark.  The group walks the observations
// of its landmark in camera order and keeps component `sub` of G_l in a register: no tile table,
// no shared memory, no segmented reduction.  Two rows are in flight per trip.
// ------------------------------------------------------------------------------------------
template <bool JOINT, bool HASW>
__global__ void __launch_bounds__(kBlock, 3)
k_e0_landmark_sell(DeviceIndex ix, const double* __restrict__ X, const double* __restrict__ cam_rec,
                   const double* __restrict__ sell_d, const double* __restrict__ sell_w, double c1,
                   double c2, const double* __restrict__ lm_fold, double* __restrict__ lm_rec,
                   const SeriesCtl* __restrict__ ctl, int slices_per_block) {
  if (ctl != nullptr && ctl->done) return;
  constexpr int NV = JOINT ? 4 : 3;
  using R = CamRec<JOINT>;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int grp = lane >> 2, sub = lane & 3, gb = lane & ~3;
  const int s_begin = blockIdx.x * slices_per_block;
  const int s_end = min(s_begin + slices_per_block, ix.num_slices);
  for (int sl = s_begin + wib; sl < s_end; sl += kWarps) {
    const int row0 = __ldg(ix.slice_ptr + sl), row1 = __ldg(ix.slice_ptr + sl + 1);
    const int lm = __ldg(ix.sell_lm + 8 * sl + grp);
    double x[4] = {0.0, 0.0, 0.0, 0.0};
    if (lm >= 0) load4(X + 4 * static_cast<size_t>(lm), x);
    double G = 0.0;
    for (int row = row0; row < row1; row += 2) {
      int c[2];
      bool act[2];
      size_t slot[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool in = row + u < row1;
        slot[u] = 8 * static_cast<size_t>(in ? row + u : row) + grp;
        c[u] = __ldg(ix.sell_cam + slot[u]);
        act[u] = in && c[u] >= 0;
      }
      double2 ya[2], yb[2], ma[2], mb[2], uv[2];
      double d[2][3], w[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const double* r = cam_rec + R::kStride * static_cast<size_t>(act[u] ? c[u] : 0);
        ya[u] = yb[u] = ma[u] = mb[u] = make_double2(0.0, 0.0);
        if (sub < 3) {
          ya[u] = ldg2(r + 2 * R::chunk(0, sub));
          yb[u] = ldg2(r + 2 * R::chunk(1, sub));
        }
        if (JOINT || sub < 3) {
          ma[u] = ldg2(r + 2 * R::chunk(2, sub));
          mb[u] = ldg2(r + 2 * R::chunk(3, sub));
        }
        if (JOINT) {
          const double* dp = sell_d + 3 * slot[u];
          d[u][0] = __ldg(dp);
          d[u][1] = __ldg(dp + 1);
          d[u][2] = __ldg(dp + 2);
        } else {
          uv[u] = ix.sell_uv[slot[u]];
          w[u] = HASW ? __ldg(sell_w + slot[u]) : 1.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const double q = ya[u].x * x[0] + ya[u].y * x[1] + yb[u].x * x[2] + yb[u].y * x[3];
        const double q0 = __shfl_sync(kFullMask, q, gb);
        const double q1 = __shfl_sync(kFullMask, q, gb + 1);
        const double q2 = __shfl_sync(kFullMask, q, gb + 2);
        double m[3];
        if (JOINT) {
          joint_normal_coef(q0, q1, q2, d[u][0], d[u][1], d[u][2], m);
        } else {
          pose_normal_coef(q0, q1, q2, uv[u].x, uv[u].y, w[u], c1, c2, m);
        }
        const double g = ma[u].x * m[0] + ma[u].y * m[1] + mb[u].x * m[2];
        if (act[u]) G += g;
      }
    }
    // H_l = fold_l G_l: every lane of the group gets G, lane `sub` makes and stores H[sub]
    const double G0 = __shfl_sync(kFullMask, G, gb), G1 = __shfl_sync(kFullMask, G, gb + 1);
    const double G2 = __shfl_sync(kFullMask, G, gb + 2), G3 = __shfl_sync(kFullMask, G, gb + 3);
    if (lm >= 0) {
      const double* f = lm_fold + 10 * static_cast<size_t>(lm);
      double H;
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
        H = r0 * G0 + r1 * G1 + r2 * G2 + r3 * G3;
      } else {
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
        H = sub < 3 ? r0 * G0 + r1 * G1 + r2 * G2 : 0.0;
      }
      lm_rec[kLmRec * static_cast<size_t>(lm) + 4 + sub] = H;
    }
  }
}

// ------------------------------------------------------------------------------------------
// camera half.  One warp per work item (a run of CSC entries of one camera, kernels_camera.cu),
// two lanes per entry, sixteen entries per step.  Lane j owns X[2j..2j+1] and H[2j..2j+1] of the
// landmark record and the matching two columns of M; the 3-vector M H is completed with one
// exchange, and each lane accumulates its six entries of m (x) X.
// ------------------------------------------------------------------------------------------
template <bool JOINT, bool HASW>
__global__ void __launch_bounds__(kBlock)
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
  for (int e0 = eb; e0 < ee; e0 += 64) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int e = e0 + 16 * s + pr;
      const bool act = e < ee;
      const int ec = act ? e : eb;
      const int lm = __ldg(ix.csc_lm + ec);
      const double* rp = lm_rec + kLmRec * static_cast<size_t>(lm);
      const double2 xe = ldg2(rp + 2 * j), he = ldg2(rp + 4 + 2 * j);
      double v[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double part = Ma[r] * he.x + Mb[r] * he.y;
        v[r] = part + __shfl_xor_sync(kFullMask, part, 1);
      }
      double m[3];
      if (JOINT) {
        const double* dp = csc_d + 3 * static_cast<size_t>(ec);
        joint_normal_coef(v[0], v[1], v[2], __ldg(dp), __ldg(dp + 1), __ldg(dp + 2), m);
      } else {
        const double2 uv = ix.csc_uv[ec];
        const double w = HASW ? __ldg(csc_w + ec) : 1.0;
        pose_normal_coef(v[0], v[1], v[2], uv.x, uv.y, w, c1, c2, m);
      }
      if (act) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          acc[k][0] += m[k] * xe.x;
          acc[k][1] += m[k] * xe.y;
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
  using R = CamRec<JOINT>;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * 4) return;
  const int c = idx >> 2, n = idx & 3;
  if (!JOINT && n == 3) return;
  const double* p = P + 12 * static_cast<size_t>(c);
  double* r = cam_rec + R::kStride * static_cast<size_t>(c);
  r[R::m_index(0, n)] = p[n];
  r[R::m_index(1, n)] = p[4 + n];
  r[R::m_index(2, n)] = p[8 + n];
  r[R::m_index(2, n) + 1] = 0.0;
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

template <bool JOINT, bool HASW>
static void launch_tiles_v2(const DeviceState& d, const ModelParams& mp, const SeriesCtl* ctl,
                            const int4* tile_info, int num_tiles, const LaunchCfg& lc) {
  if (num_tiles == 0) return;
  // contiguous tile ranges per block, about four waves of 148 SMs x 2 resident blocks
  const long long target = 148LL * 2 * 4;
  long long per_block = (num_tiles + target - 1) / target;
  per_block = (per_block + kWarps - 1) / kWarps * kWarps;
  const int blocks = static_cast<int>((num_tiles + per_block - 1) / per_block);
  k_e0_landmark_v2<JOINT, HASW><<<blocks, kBlock, 0, lc.stream>>>(
      d.ix, d.X, d.cam_rec, d.obs_d, d.obs_w, mp.c1, mp.c2, d.lm_fold, d.lm_rec, ctl, tile_info, num_tiles,
      static_cast<int>(per_block));
  count(lc);
}

template <bool JOINT, bool HASW>
static void launch_sell(const DeviceState& d, const ModelParams& mp, const SeriesCtl* ctl,
                        const LaunchCfg& lc) {
  if (d.ix.num_slices > 0) {
    // contiguous slice ranges per block (a sorting window holds long and short slices: the warps of
    // a block interleave over it), about three waves of 148 SMs x 3 resident blocks
    const long long target = 148LL * 3 * 3;
    long long per_block = (d.ix.num_slices + target - 1) / target;
    per_block = (per_block + kWarps - 1) / kWarps * kWarps;
    const int blocks = static_cast<int>((d.ix.num_slices + per_block - 1) / per_block);
    k_e0_landmark_sell<JOINT, HASW><<<blocks, kBlock, 0, lc.stream>>>(
        d.ix, d.X, d.cam_rec, d.sell_d, d.sell_w, mp.c1, mp.c2, d.lm_fold, d.lm_rec, ctl,
        static_cast<int>(per_block));
    count(lc);
  }
  // landmarks with more than 32 observations: one warp each, tile kernel
  launch_tiles_v2<JOINT, HASW>(d, mp, ctl, d.ix.long_tile_info, d.ix.num_long_tiles, lc);
}

// layout 0: sliced ELL (+ tile kernel for long landmarks); layout 1: tile kernel for everything
void launch_e0_landmark_v2(const DeviceState& d, const ModelParams& mp, bool joint, bool in_series,
                           int layout, const LaunchCfg& lc) {
  const SeriesCtl* ctl = in_series ? d.ctl : nullptr;
  const bool hasw = !joint && mp.robust_norm == NORM_HUBER;
  if (layout == 1) {
    if (joint) launch_tiles_v2<true, false>(d, mp, ctl, d.ix.tile_info, d.ix.num_tiles, lc);
    else if (hasw) launch_tiles_v2<false, true>(d, mp, ctl, d.ix.tile_info, d.ix.num_tiles, lc);
    else launch_tiles_v2<false, false>(d, mp, ctl, d.ix.tile_info, d.ix.num_tiles, lc);
  } else {
    if (joint) launch_sell<true, false>(d, mp, ctl, lc);
    else if (hasw) launch_sell<false, true>(d, mp, ctl, lc);
    else launch_sell<false, false>(d, mp, ctl, lc);
  }
}

void launch_passB_e0_v2(const DeviceState& d, const ModelParams& mp, bool joint, bool in_series,
                        const LaunchCfg& lc) {
  if (d.ix.num_items == 0) return;
  const int blocks = (d.ix.num_items + kWarps - 1) / kWarps;
  const SeriesCtl* ctl = in_series ? d.ctl : nullptr;
  const bool hasw = !joint && mp.robust_norm == NORM_HUBER;
  if (joint) {
    k_passB_e0_v2<true, false><<<blocks, kBlock, 0, lc.stream>>>(d.ix, d.P, d.lm_rec, d.csc_d, nullptr,
                                                                 mp.c1, mp.c2, d.item_part, ctl);
  } else if (hasw) {
    k_passB_e0_v2<false, true><<<blocks, kBlock, 0, lc.stream>>>(d.ix, d.P, d.lm_rec, nullptr, d.csc_w,
                                                                 mp.c1, mp.c2, d.item_part, ctl);
  } else {
    k_passB_e0_v2<false, false><<<blocks, kBlock, 0, lc.stream>>>(d.ix, d.P, d.lm_rec, nullptr, nullptr,
                                                                  mp.c1, mp.c2, d.item_part, ctl);
  }
  count(lc);
}

}  // namespace povar
