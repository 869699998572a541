// Kernels of the explicit-Schur-complement solvers: PCG / RIPCG vector algebra and the dense reduced
// camera system for CHOLESKY.
//
// The reference forms S = Hpp - sum_l Hpl Hll^-1 Hlp as a hash map of 12x12 (11x11) blocks
// (sc/landmark_block.hpp:360-472, cg/block_sparse_matrix.hpp:152-345) and multiplies with it.  On the
// GPU PCG/RIPCG apply S implicitly,  S p = B p - E0 p,  with the same matrix-free E0 product as the
// power series (kernels_landmark.cu / kernels_camera.cu); only the block-Jacobi preconditioner needs
// the diagonal blocks of S (k_kron<KRON_SDIAG>).  CHOLESKY (step 1 only, sc/linearization_sc.hpp:236-245)
// does need S: it is assembled densely here and factorised by cuSOLVER.
#include <cuda_runtime.h>

#include "device_math.cuh"
#include "povar_internal.h"

namespace povar {

namespace {

constexpr int kBlock = 256;

inline void count(const LaunchCfg& lc, int n = 1) {
  if (lc.launch_counter) *lc.launch_counter += n;
}

__global__ void __launch_bounds__(kBlock)
k_axpby(int n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
        double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (y != nullptr) ? a * x[i] + b * y[i] : a * x[i];
}

// two-stage deterministic dot product: per-block partials, then one block adds them in order
__global__ void __launch_bounds__(kBlock)
k_dot_part(int n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ part) {
  __shared__ double smem[kBlock / 32];
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc[0] += x[i] * y[i];
  block_reduce<1>(acc, smem);
  if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
}

__global__ void __launch_bounds__(kBlock)
k_dot_final(int nblocks, const double* __restrict__ part, double* out) {
  __shared__ double smem[kBlock / 32];
  double acc[1] = {0.0};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc[0] += part[b];
  block_reduce<1>(acc, smem);
  if (threadIdx.x == 0) out[0] = acc[0];
}

__global__ void __launch_bounds__(kBlock)
k_finite_check(int n, const double* __restrict__ x, SeriesCtl* ctl) {
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    bad = bad || !isfinite(x[i]);
  }
  if (bad) atomicExch(&ctl->nonfinite, 1);
}

// S -= (s_i s_j^T) o [ (Wk_i^T Hll^-1 Wk_j) (x) (Xt Xt^T) ] for every pair of observations (i, j) of a
// landmark, Wk = w Z^T K (3x3) with Z = Jl_raw o lm_scale and K the 4x3 coefficient matrix of Jp_raw.
// One warp per landmark, lanes over the 144 entries of a block; FP64 atomics (the only place in the
// library where the summation order is not fixed -- the reference's own order depends on its hash map).
__device__ __forceinline__ void pose_wk(const Cam3x4& cam, double u, double v, const double (&x)[4],
                                        const double (&sl)[4], double c1, double c2, const Robust& rb,
                                        double (&Wk)[3][3]) {
  PoseObs ob;
  ob.eval(cam, u, v, x, c1, c2, rb);
  const double w = ob.sw * ob.sw;
  const double K[4][3] = {{c1, 0.0, -c1 * u}, {0.0, c1, -c1 * v}, {c2, 0.0, 0.0}, {0.0, c2, 0.0}};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      Wk[a][k] = w * sl[a] * (ob.T[0][a] * K[0][k] + ob.T[1][a] * K[1][k] + ob.T[2][a] * K[2][k] +
                              ob.T[3][a] * K[3][k]);
    }
  }
}

__global__ void __launch_bounds__(kBlock)
k_dense_schur(DeviceIndex ix, const double* __restrict__ P, const double* __restrict__ X, double c1,
              double c2, Robust rb, const double* __restrict__ lm_scale, const double* __restrict__ hll_inv,
              const double* __restrict__ pose_scale, double* __restrict__ S, long long ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int l = warp; l < ix.L; l += nwarps) {
    const int ob = ix.lm_ptr[l], oe = ix.lm_ptr[l + 1];
    double x[4], sl[4], inv[6];
    load4(X + 4 * static_cast<size_t>(l), x);
    load4(lm_scale + 4 * static_cast<size_t>(l), sl);
    {
      const double* hi = hll_inv + 6 * static_cast<size_t>(l);
#pragma unroll
      for (int k = 0; k < 6; ++k) inv[k] = hi[k];
    }
    for (int i = ob; i < oe; ++i) {
      const int ci = ix.obs_cam[i];
      Cam3x4 cam_i;
      load_cam(P, ci, cam_i);
      const double2 uvi = ix.obs_uv[i];
      double Wi[3][3], HWi[3][3];
      pose_wk(cam_i, uvi.x, uvi.y, x, sl, c1, c2, rb, Wi);
      // HWi = Hll^-1 Wi (column by column)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double col[3] = {Wi[0][k], Wi[1][k], Wi[2][k]};
        double out[3];
        sym3_mul(inv, col, out);
        HWi[0][k] = out[0];
        HWi[1][k] = out[1];
        HWi[2][k] = out[2];
      }
      for (int j = ob; j < oe; ++j) {
        const int cj = ix.obs_cam[j];
        Cam3x4 cam_j;
        load_cam(P, cj, cam_j);
        const double2 uvj = ix.obs_uv[j];
        double Wj[3][3];
        pose_wk(cam_j, uvj.x, uvj.y, x, sl, c1, c2, rb, Wj);
        // Q = Wi^T Hll^-1 Wj = HWi^T Wj  (Hll^-1 symmetric)
        double Q[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int b2 = 0; b2 < 3; ++b2) {
            Q[a][b2] = HWi[0][a] * Wj[0][b2] + HWi[1][a] * Wj[1][b2] + HWi[2][a] * Wj[2][b2];
          }
        }
        const double* si = pose_scale + 12 * static_cast<size_t>(ci);
        const double* sj = pose_scale + 12 * static_cast<size_t>(cj);
        for (int e = lane; e < 144; e += 32) {
          const int r = e / 12, cc = e % 12;
          const double val = si[r] * sj[cc] * Q[r >> 2][cc >> 2] * x[r & 3] * x[cc & 3];
          atomicAdd(S + (static_cast<long long>(ci) * 12 + r) * ld + (static_cast<long long>(cj) * 12 + cc), -val);
        }
      }
    }
  }
}

// S_cc += Bmat_c (which already holds (s s^T) o Jp^T Jp + lambda I)
__global__ void __launch_bounds__(kBlock)
k_dense_add_diag(int C, const double* __restrict__ Bmat, double* __restrict__ S, long long ld) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * 144) return;
  const int c = idx / 144, e = idx % 144, r = e / 12, cc = e % 12;
  S[(static_cast<long long>(c) * 12 + r) * ld + (static_cast<long long>(c) * 12 + cc)] += Bmat[idx];
}

}  // namespace

void launch_axpby(const DeviceState& d, int n, double a, const double* x, double b, const double* y,
                  double* out, const LaunchCfg& lc) {
  (void)d;
  k_axpby<<<(n + kBlock - 1) / kBlock, kBlock, 0, lc.stream>>>(n, a, x, b, y, out);
  count(lc);
}

void launch_dot(const DeviceState& d, int n, const double* x, const double* y, int slot, const LaunchCfg& lc) {
  int blocks = (n + kBlock - 1) / kBlock;
  const int cap = scalar_blocks(d);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_dot_part<<<blocks, kBlock, 0, lc.stream>>>(n, x, y, d.scalar_part);
  k_dot_final<<<1, kBlock, 0, lc.stream>>>(blocks, d.scalar_part, d.scalar_out + slot);
  count(lc, 2);
}

void launch_finite_check(const DeviceState& d, int n, const double* x, const LaunchCfg& lc) {
  int blocks = (n + kBlock - 1) / kBlock;
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  k_finite_check<<<blocks, kBlock, 0, lc.stream>>>(n, x, d.ctl);
  count(lc);
}

void launch_dense_schur(const DeviceState& d, const ModelParams& mp, double* S, const LaunchCfg& lc) {
  const Robust rb = {mp.robust_norm, mp.huber};
  const long long ld = 12LL * d.ix.C;
  long long blocks = (static_cast<long long>(d.ix.L) + (kBlock / 32) - 1) / (kBlock / 32);
  if (blocks > static_cast<long long>(sm_count()) * 8 * 4) blocks = static_cast<long long>(sm_count()) * 8 * 4;
  if (blocks < 1) blocks = 1;
  k_dense_schur<<<static_cast<int>(blocks), kBlock, 0, lc.stream>>>(d.ix, d.P, d.X, mp.c1, mp.c2, rb, d.lm_scale,
                                                                    d.hll_inv, d.pose_scale, S, ld);
  const int n = d.ix.C * 144;
  k_dense_add_diag<<<(n + kBlock - 1) / kBlock, kBlock, 0, lc.stream>>>(d.ix.C, d.Bmat, S, ld);
  count(lc, 2);
}

}  // namespace povar
