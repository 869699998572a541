// Per-observation models and small dense helpers shared by the landmark-major and
// camera-major kernels.  Everything is FP64, as in the reference.
//
// The kernels are matrix-free: the reference materialises [Jp | Jl | r] per observation
// (sc/landmark_block.hpp:135-225, 512 B/obs) and streams it in every product; here the
// blocks are functions of (camera matrix, landmark, observation, weights, scalings) and are
// re-evaluated in registers, so a product reads 20 B/obs plus cache-resident gathers.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

namespace povar {

constexpr double kEpsSqrt = 1e-5;  // Sophus::Constants<double>::epsilonSqrt()
constexpr unsigned kFullMask = 0xffffffffu;

enum : int { NORM_NONE = 0, NORM_HUBER = 1, NORM_CAUCHY = 2 };

struct Robust {
  int norm;
  double huber;
};

// compute_error_weight, bal/bal_bundle_adjustment_helper.cpp:50-74
__device__ __forceinline__ void error_weight(const Robust& rb, double res_sq, double& error,
                                             double& weight) {
  if (rb.norm == NORM_HUBER) {
    const double th = rb.huber;
    const double hw = res_sq < th * th ? 1.0 : th / sqrt(res_sq);
    error = 0.5 * (2.0 - hw) * hw * res_sq;
    weight = hw;
  } else if (rb.norm == NORM_CAUCHY) {
    error = log(1.0 + res_sq);
    weight = 1.0;  // CAUCHY changes the reported cost only
  } else {
    error = 0.5 * res_sq;
    weight = 1.0;
  }
}

__device__ __forceinline__ double robust_sqrt_weight(const Robust& rb, double res_sq) {
  if (rb.norm == NORM_HUBER) {
    const double th = rb.huber;
    const double hw = res_sq < th * th ? 1.0 : th / sqrt(res_sq);
    return sqrt(hw);
  }
  return 1.0;
}

__device__ __forceinline__ bool finite3(double a, double b, double c) {
  return isfinite(a) && isfinite(b) && isfinite(c);
}

// ---- loads ------------------------------------------------------------------------------
struct Cam3x4 {
  double r0[4], r1[4], r2[4];
};

__device__ __forceinline__ void load4(const double* __restrict__ p, double (&v)[4]) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  v[0] = a.x;
  v[1] = a.y;
  v[2] = b.x;
  v[3] = b.y;
}

// 32 bytes with one 256-bit load (LDG.E.256, sm_100): a gather of 32-byte pieces costs one L1 wavefront per piece
// instead of the two of a pair of 128-bit loads.  p must be 32-byte aligned; read-only data path.
__device__ __forceinline__ void load4_256(const double* __restrict__ p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p));
}

// L2 eviction priorities.  A power-series term streams ~290 MB of observation data (read once per kernel) past
// the 64 MB of landmark records that its two halves hand to each other and gather from five times: the records
// are written and read with evict_last, the streams with evict_first, so that the gathers of the camera half
// find the records in L2 (126 MB) instead of fetching half of them from DRAM.
__device__ __forceinline__ unsigned long long l2_keep() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_stream() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void load4_256(const double* __restrict__ p, double (&v)[4], unsigned long long policy) {
  asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
               : "l"(p), "l"(policy));
}
__device__ __forceinline__ double2 load2(const double2* __restrict__ p, unsigned long long policy) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ double ldg1(const double* __restrict__ p, unsigned long long policy) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ int ldg1(const int* __restrict__ p, unsigned long long policy) {
  int v;
  asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void store2(double* p, double a, double b, unsigned long long policy) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(policy) : "memory");
}
// one full 32-byte sector per lane instead of two half-written ones
__device__ __forceinline__ void store4_256(double* p, double a, double b, double c, double d,
                                           unsigned long long policy) {
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d),
               "l"(policy)
               : "memory");
}

__device__ __forceinline__ void load_cam(const double* __restrict__ P, int c, Cam3x4& m) {
  const double* p = P + 12 * static_cast<size_t>(c);
  load4_256(p, m.r0);
  load4_256(p + 4, m.r1);
  load4_256(p + 8, m.r2);
}

__device__ __forceinline__ double dot4(const double (&a)[4], const double (&b)[4]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
}

// ---- step 1: pOSE observation model (helper.cpp:243-313) ----------------------------------
// T = [c1 (P0 - u P2); c1 (P1 - v P2); c2 P0; c2 P1],  r = T [X;1] - [0 0 c2 u c2 v],
// Jl_raw = T[:, 0:3],  Jp_raw rows = c1 [Xt 0 -u Xt], c1 [0 Xt -v Xt], c2 [Xt 0 0], c2 [0 Xt 0].
struct PoseObs {
  double T[4][3];   // Jl_raw
  double r[4];      // raw residual
  double sw;        // sqrt of the robust weight
  __device__ __forceinline__ void eval(const Cam3x4& P, double u, double v, const double (&X)[4],
                                       double c1, double c2, const Robust& rb) {
    double t3[4];   // 4th column of T
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      T[0][k] = c1 * (P.r0[k] - P.r2[k] * u);
      T[1][k] = c1 * (P.r1[k] - P.r2[k] * v);
      T[2][k] = c2 * P.r0[k];
      T[3][k] = c2 * P.r1[k];
    }
    t3[0] = c1 * (P.r0[3] - P.r2[3] * u);
    t3[1] = c1 * (P.r1[3] - P.r2[3] * v);
    t3[2] = c2 * P.r0[3];
    t3[3] = c2 * P.r1[3];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r[q] = T[q][0] * X[0] + T[q][1] * X[1] + T[q][2] * X[2] + t3[q] * X[3];
    }
    r[2] -= c2 * u;
    r[3] -= c2 * v;
    const double res_sq = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
    sw = robust_sqrt_weight(rb, res_sq);
  }
  __device__ __forceinline__ double res_sq() const {
    return r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
  }
};

// a = Jp_raw * y for a 12-vector y (three 4-blocks), given Xt = [X;1]
__device__ __forceinline__ void pose_jp_mul(const double (&X)[4], double u, double v, double c1,
                                            double c2, const double (&y0)[4], const double (&y1)[4],
                                            const double (&y2)[4], double (&a)[4]) {
  const double q0 = dot4(X, y0), q1 = dot4(X, y1), q2 = dot4(X, y2);
  a[0] = c1 * (q0 - u * q2);
  a[1] = c1 * (q1 - v * q2);
  a[2] = c2 * q0;
  a[3] = c2 * q1;
}

// Jp_raw^T t = m (x) Xt with the 3-vector m below
__device__ __forceinline__ void pose_jpT_coef(const double (&t)[4], double u, double v, double c1,
                                              double c2, double (&m)[3]) {
  m[0] = c1 * t[0] + c2 * t[2];
  m[1] = c1 * t[1] + c2 * t[3];
  m[2] = -c1 * (u * t[0] + v * t[1]);
}

// ---- step 2: projective observation model (helper.cpp:315-380, bal_camera.hpp:116-167) ------
// pc = P Xh, r = (pc0/pc2 - u, pc1/pc2 - v), d = [[1/z 0 -x/z^2],[0 1/z -y/z^2]],
// Jp_raw = d (I3 (x) Xh^T), Jl_raw = d P.
struct JointObs {
  double iz, d02, d12;   // d
  double r[2];
  double sw;
  bool valid;            // |z| >= sqrt(eps)
  __device__ __forceinline__ void eval(const Cam3x4& P, double u, double v, const double (&X)[4],
                                       const Robust& rb) {
    const double x = dot4(P.r0, X), y = dot4(P.r1, X), z = dot4(P.r2, X);
    iz = 1.0 / z;
    d02 = -x / (z * z);
    d12 = -y / (z * z);
    r[0] = x / z - u;
    r[1] = y / z - v;
    valid = fabs(z) >= kEpsSqrt;
    sw = robust_sqrt_weight(rb, r[0] * r[0] + r[1] * r[1]);
  }
  __device__ __forceinline__ double res_sq() const { return r[0] * r[0] + r[1] * r[1]; }
  // Jl_raw rows (2 x 4)
  __device__ __forceinline__ void jl_rows(const Cam3x4& P, double (&j0)[4], double (&j1)[4]) const {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      j0[k] = iz * P.r0[k] + d02 * P.r2[k];
      j1[k] = iz * P.r1[k] + d12 * P.r2[k];
    }
  }
  // a = Jp_raw * y
  __device__ __forceinline__ void jp_mul(const double (&X)[4], const double (&y0)[4],
                                         const double (&y1)[4], const double (&y2)[4],
                                         double (&a)[2]) const {
    const double q0 = dot4(X, y0), q1 = dot4(X, y1), q2 = dot4(X, y2);
    a[0] = iz * q0 + d02 * q2;
    a[1] = iz * q1 + d12 * q2;
  }
  // Jp_raw^T t = m (x) Xh
  __device__ __forceinline__ void jpT_coef(const double (&t)[2], double (&m)[3]) const {
    m[0] = iz * t[0];
    m[1] = iz * t[1];
    m[2] = d02 * t[0] + d12 * t[1];
  }
};

// ---- 3x3 symmetric inverse by cofactors (Eigen's 3x3 inverse, LU/InverseImpl.h:156-169) ------
// packing of a symmetric 3x3: [00 01 02 11 12 22]
__device__ __forceinline__ void inv3_sym(const double (&a)[6], double (&inv)[6]) {
  const double a00 = a[0], a01 = a[1], a02 = a[2], a11 = a[3], a12 = a[4], a22 = a[5];
  const double c00 = a11 * a22 - a12 * a12;
  const double c01 = a02 * a12 - a01 * a22;   // cofactor(0,1) = -(a01 a22 - a12 a02)
  const double c02 = a01 * a12 - a02 * a11;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  const double id = 1.0 / det;
  inv[0] = c00 * id;
  inv[1] = c01 * id;
  inv[2] = c02 * id;
  inv[3] = (a00 * a22 - a02 * a02) * id;
  inv[4] = (a01 * a02 - a00 * a12) * id;
  inv[5] = (a00 * a11 - a01 * a01) * id;
}

__device__ __forceinline__ void sym3_mul(const double (&s)[6], const double (&x)[3], double (&y)[3]) {
  y[0] = s[0] * x[0] + s[1] * x[1] + s[2] * x[2];
  y[1] = s[1] * x[0] + s[3] * x[1] + s[4] * x[2];
  y[2] = s[2] * x[0] + s[4] * x[1] + s[5] * x[2];
}

// packing of a symmetric 4x4: [00 01 02 03 11 12 13 22 23 33]
__device__ __forceinline__ int sym4_index(int i, int j) {
  // i <= j
  return i * 4 - (i * (i - 1)) / 2 + (j - i);
}

// ---- kernel_COD as a Householder reflector (helper.cpp:201-216) ---------------------
// For a 1 x N row m:  p = first argmax |m_j|;  r = m with entries 0 and p exchanged;
// H = I - tau w w^T the reflector of r (Eigen's makeHouseholder);  Pi = swap_rows(0,p)( H[:, 1:] ).
template <int N>
struct Reflector {
  double w[N];   // w[0] = 1
  double tau;
  int p;
  __device__ __forceinline__ void make(const double (&m)[N]) {
    int pp = 0;
    double best = fabs(m[0]);
#pragma unroll
    for (int j = 1; j < N; ++j) {
      const double a = fabs(m[j]);
      if (a > best) {
        best = a;
        pp = j;
      }
    }
    p = pp;
    double r0 = m[0];
    double tail2 = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j == pp) {
        r0 = m[j];
      }
    }
    // r[j] for j >= 1 is m[j], except r[p] = m[0]
    w[0] = 1.0;
#pragma unroll
    for (int j = 1; j < N; ++j) {
      const double rj = (j == pp) ? m[0] : m[j];
      w[j] = rj;
      tail2 += rj * rj;
    }
    if (tail2 <= DBL_MIN) {
      tau = 0.0;
#pragma unroll
      for (int j = 1; j < N; ++j) w[j] = 0.0;
    } else {
      double beta = sqrt(r0 * r0 + tail2);
      if (r0 >= 0.0) beta = -beta;
      const double inv = 1.0 / (r0 - beta);
#pragma unroll
      for (int j = 1; j < N; ++j) w[j] = w[j] * inv;
      tau = (beta - r0) / beta;
    }
  }
  // out[N] = Pi x, x has N-1 entries
  __device__ __forceinline__ void apply(const double* x, double (&out)[N]) const {
    double dot = 0.0;
#pragma unroll
    for (int j = 1; j < N; ++j) dot += w[j] * x[j - 1];
    const double td = tau * dot;
    out[0] = -td;
#pragma unroll
    for (int j = 1; j < N; ++j) out[j] = x[j - 1] - td * w[j];
    const double o0 = out[0];
#pragma unroll
    for (int j = 1; j < N; ++j) {
      if (j == p) {
        out[0] = out[j];
        out[j] = o0;
      }
    }
  }
  // out[N-1] = Pi^T v
  __device__ __forceinline__ void apply_t(const double (&v)[N], double* out) const {
    double vv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) vv[j] = v[j];
    const double v0 = vv[0];
#pragma unroll
    for (int j = 1; j < N; ++j) {
      if (j == p) {
        vv[0] = vv[j];
        vv[j] = v0;
      }
    }
    double dot = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) dot += w[j] * vv[j];
    const double td = tau * dot;
#pragma unroll
    for (int j = 1; j < N; ++j) out[j - 1] = vv[j] - td * w[j];
  }
};

// ---- per-camera record gathered by the landmark-major half of E0 (kernels_series.cu) -----------
// One lane handles one observation and reads the whole record of its camera from shared memory in 16-byte
// units: y (the vector the product is applied to, three 4-blocks, 12 doubles), then M = P[:, 0:3] row-major
// (step 1, 9 doubles + 1 pad) or M = P (step 2, 12 doubles + 2 pad).  The stride is an ODD number of units
// (11 resp. 13): unit j of camera c falls into bank group (stride * c + j) mod 8, so the eight lanes of a
// shared-memory wavefront -- eight different cameras, same j -- spread over the bank groups like their camera
// numbers mod 8 instead of all landing on one.
struct CamRec {
  __host__ __device__ static constexpr int stride(bool joint) { return joint ? 26 : 22; }   // doubles
  __host__ __device__ static constexpr int y_index(int k, int j) { return 4 * k + j; }
  __host__ __device__ static constexpr int m_index(bool joint, int r, int n) {
    return joint ? 12 + 4 * r + n : 12 + 3 * r + n;
  }
};

// ---- warp helpers ----------------------------------------------------------------------------
__device__ __forceinline__ double shfl_double(double v, int src) {
  return __shfl_sync(kFullMask, v, src);
}

template <int NV>
__device__ __forceinline__ void warp_allreduce(double (&v)[NV]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] += __shfl_xor_sync(kFullMask, v[k], off);
  }
}

// all-reduce over contiguous lane segments [seg_first, seg_last]; every lane of a segment
// passes the same bounds.  Fixed summation tree => bit-reproducible.
template <int NV>
__device__ __forceinline__ void segment_allreduce(double (&v)[NV], int lane, int seg_first,
                                                  int seg_last) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double t = __shfl_up_sync(kFullMask, v[k], d);
      if (lane - d >= seg_first) v[k] += t;
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = __shfl_sync(kFullMask, v[k], seg_last);
}

// deterministic block sum of NV doubles per thread; result valid in thread 0.
// smem must hold NV * (blockDim.x / 32) doubles.
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  warp_allreduce<NV>(v);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarps; ++w) {
#pragma unroll
      for (int k = 0; k < NV; ++k) v[k] += smem[w * NV + k];
    }
  }
  __syncthreads();
}

}  // namespace povar
