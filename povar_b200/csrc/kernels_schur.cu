// Kernels of the explicit-Schur-complement solvers: PCG / RIPCG vector algebra (CHOLESKY: kernels_chol.cu).
//
// The reference forms S = Hpp - sum_l Hpl Hll^-1 Hlp as a hash map of 12x12 (11x11) blocks
// (sc/landmark_block.hpp:360-472, cg/block_sparse_matrix.hpp:152-345) and multiplies with it.  On the
// GPU PCG/RIPCG apply S implicitly,  S p = B p - E0 p,  with the same matrix-free E0 product as the
// power series (kernels_landmark.cu / kernels_camera.cu); only the block-Jacobi preconditioner needs
// the diagonal blocks of S (k_kron<KRON_SDIAG>).
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

}  // namespace povar
