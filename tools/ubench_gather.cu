// Micro-benchmark behind DESIGN.md section 4: what does a gather of per-camera / per-landmark
// records cost on B200, as a function of how the lanes of a warp are mapped onto the records?
// (L1TEX processes one 128-byte line per wavefront; a per-lane gather touches 32 lines per
// instruction, a cooperative load of whole lines touches 4.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench_gather tools/ubench_gather.cu
//   tools/ubench_gather            (prints one line per variant)
//
// Index distribution = povar_b200/synthetic.py at the venice-1778 shape: 1,778 cameras, ~1 M
// landmarks, ~5 M observations, each landmark seen from cameras spread over a window of half the
// trajectory.  Not part of the product; nothing here is linked into libpovar_b200.so.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      std::exit(1);                                                                  \
    }                                                                                \
  } while (0)

constexpr int kBlock = 256;

__device__ __forceinline__ double2 ldg2(const double* p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

// ---- camera-record gathers (landmark-major pass) ---------------------------------------------
// G0: one observation per lane, 12 x LDG.128 from two 96-byte record tables (today's kernel)
__global__ void __launch_bounds__(kBlock) g0_lane(int nnz, const int* __restrict__ obs_cam,
                                                  const double* __restrict__ P,
                                                  const double* __restrict__ Y, double* out) {
  double acc = 0.0;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < nnz; o += gridDim.x * blockDim.x) {
    const int c = __ldg(obs_cam + o);
    const double* p = P + 12 * (size_t)c;
    const double* y = Y + 12 * (size_t)c;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double2 a = ldg2(p + 2 * k), b = ldg2(y + 2 * k);
      acc += a.x * b.y + a.y * b.x;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// G1: eight lanes per observation, 256-byte aligned records, 2 x LDG.128 per lane (whole lines)
template <int SHFL>
__global__ void __launch_bounds__(kBlock) g1_coop8(int nnz, const int* __restrict__ obs_cam,
                                                   const double* __restrict__ R, double* out) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double acc = 0.0;
  for (int base = warp * 32; base < nnz; base += nwarps * 32) {
    double2 a[8], b[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int o = min(base + 4 * s + grp, nnz - 1);
      const int c = __ldg(obs_cam + o);
      const double* r = R + 32 * (size_t)c + 2 * sub;
      a[s] = ldg2(r);
      b[s] = ldg2(r + 16);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      double q = a[s].x * b[s].y + a[s].y * b[s].x;
      if (SHFL) {
        // the slice layout exchanges three doubles per step
        const double q0 = __shfl_sync(0xffffffffu, q, (lane & 24) + 0);
        const double q1 = __shfl_sync(0xffffffffu, q, (lane & 24) + 1);
        const double q2 = __shfl_sync(0xffffffffu, q, (lane & 24) + 2);
        q = q0 * a[s].x + q1 * a[s].y + q2 * b[s].x;
      }
      acc += q;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// G2: sixteen lanes per record stage 176 bytes (11 chunks) into shared memory with cp.async,
// then one observation per lane reads its record back with 11 x LDS.128 (odd chunk stride)
__global__ void __launch_bounds__(kBlock) g2_stage(int nnz, const int* __restrict__ obs_cam,
                                                   const double* __restrict__ R, double* out) {
  __shared__ __align__(16) double stage[kBlock / 32][32 * 22];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double* st = stage[wib];
  const int half = lane >> 4, chunk = lane & 15;
  double acc = 0.0;
  for (int base = warp * 32; base < nnz; base += nwarps * 32) {
    const int c_mine = __ldg(obs_cam + min(base + lane, nnz - 1));
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int rec = 2 * i + half;
      const int c = __shfl_sync(0xffffffffu, c_mine, rec);
      if (chunk < 11) {
        const double* src = R + 32 * (size_t)c + 2 * chunk;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(st + rec * 22 + 2 * chunk);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
      }
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    const double2* mine = reinterpret_cast<const double2*>(st + lane * 22);
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const double2 v = mine[k];
      acc += v.x * v.y;
    }
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// G4: four lanes per observation, 4 x LDG.128 per lane, record = four 64-byte planes
__global__ void __launch_bounds__(kBlock) g4_coop4(int nnz, const int* __restrict__ obs_cam,
                                                   const double* __restrict__ R, double* out) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, sub = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double acc = 0.0;
  for (int base = warp * 32; base < nnz; base += nwarps * 32) {
    double2 a[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int o = min(base + 8 * s + grp, nnz - 1);
      const int c = __ldg(obs_cam + o);
      const double* r = R + 32 * (size_t)c + 2 * sub;
#pragma unroll
      for (int k = 0; k < 4; ++k) a[s][k] = ldg2(r + 8 * k);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += a[s][k].x * a[s][k].y;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- landmark-record gathers (camera-major pass) ---------------------------------------------
// H0: one entry per lane, 4 x LDG.128 from 64-byte records (today's kernel)
__global__ void __launch_bounds__(kBlock) h0_lane(int nnz, const int* __restrict__ csc_lm,
                                                  const double* __restrict__ rec, double* out) {
  double acc = 0.0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) {
    const int l = __ldg(csc_lm + e);
    const double* r = rec + 8 * (size_t)l;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double2 a = ldg2(r + 2 * k);
      acc += a.x * a.y;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// H1: four lanes per entry, one LDG.128 per lane
template <int SHFL>
__global__ void __launch_bounds__(kBlock) h1_coop4(int nnz, const int* __restrict__ csc_lm,
                                                   const double* __restrict__ rec, double* out) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, sub = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double acc = 0.0;
  for (int base = warp * 64; base < nnz; base += nwarps * 64) {
    double2 a[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int e = min(base + 8 * s + grp, nnz - 1);
      const int l = __ldg(csc_lm + e);
      a[s] = ldg2(rec + 8 * (size_t)l + 2 * sub);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      double q = a[s].x * a[s].y;
      if (SHFL) {
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        const double q2 = __shfl_xor_sync(0xffffffffu, a[s].x * q, 1);
        q += q2;
      }
      acc += q;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// H3: 48-byte records (step 1: X (3), H (3)), three lanes of four active
__global__ void __launch_bounds__(kBlock) h3_coop4_48(int nnz, const int* __restrict__ csc_lm,
                                                      const double* __restrict__ rec, double* out) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, sub = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double acc = 0.0;
  for (int base = warp * 64; base < nnz; base += nwarps * 64) {
    double2 a[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int e = min(base + 8 * s + grp, nnz - 1);
      const int l = __ldg(csc_lm + e);
      a[s] = sub < 3 ? ldg2(rec + 6 * (size_t)l + 2 * sub) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) acc += a[s].x * a[s].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// streaming reference: the bytes of the observation stream alone (24 B / observation)
__global__ void __launch_bounds__(kBlock) s0_stream(int nnz, const int* __restrict__ obs_cam,
                                                    const int* __restrict__ obs_lm,
                                                    const double2* __restrict__ uv, double* out) {
  double acc = 0.0;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < nnz; o += gridDim.x * blockDim.x) {
    const double2 v = uv[o];
    acc += v.x * __ldg(obs_cam + o) + v.y * __ldg(obs_lm + o);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static double time_us(F&& launch, int reps = 20) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1e3 / reps;
}

int main(int argc, char** argv) {
  const int C = argc > 1 ? std::atoi(argv[1]) : 1778;
  const int L = argc > 2 ? std::atoi(argv[2]) : 993923;
  const double mean_deg = 5.03;
  std::mt19937_64 rng(1005);
  std::poisson_distribution<int> pois(mean_deg - 2.0);
  std::uniform_int_distribution<int> ucam(0, C - 1);
  std::uniform_real_distribution<double> u01(0.0, 1.0);
  std::vector<int> obs_cam, obs_lm;
  obs_cam.reserve((size_t)(L * mean_deg * 1.05));
  obs_lm.reserve(obs_cam.capacity());
  const int half = std::max(6, (int)(0.25 * C + 0.5));
  for (int l = 0; l < L; ++l) {
    const int deg = std::min(C, 2 + pois(rng)), near = ucam(rng), width = 2 * half + 1;
    std::vector<int> pick;
    for (int i = 0; i < deg; ++i) {
      const int s0 = (int)((long long)i * width / deg), s1 = (int)((long long)(i + 1) * width / deg);
      int c = near + s0 + (int)(u01(rng) * std::max(s1 - s0, 1)) - half;
      if (c < 0) c = -c - 1;
      if (c > C - 1) c = 2 * (C - 1) - c + 1;
      c = std::min(std::max(c, 0), C - 1);
      pick.push_back(c);
    }
    std::sort(pick.begin(), pick.end());
    pick.erase(std::unique(pick.begin(), pick.end()), pick.end());
    for (int c : pick) {
      obs_cam.push_back(c);
      obs_lm.push_back(l);
    }
  }
  const int nnz = (int)obs_cam.size();
  std::vector<int> cam_ptr(C + 1, 0), csc_lm(nnz);
  for (int o = 0; o < nnz; ++o) cam_ptr[obs_cam[o] + 1]++;
  for (int c = 0; c < C; ++c) cam_ptr[c + 1] += cam_ptr[c];
  {
    std::vector<int> fill(cam_ptr.begin(), cam_ptr.end() - 1);
    for (int o = 0; o < nnz; ++o) csc_lm[fill[obs_cam[o]]++] = obs_lm[o];
  }
  std::printf("C=%d L=%d nnz=%d\n", C, L, nnz);

  int *d_cam, *d_lm, *d_csc;
  double *d_P, *d_Y, *d_R, *d_rec, *d_rec48, *d_out;
  double2* d_uv;
  const int grid = 148 * 8;
  CK(cudaMalloc(&d_cam, sizeof(int) * nnz));
  CK(cudaMalloc(&d_lm, sizeof(int) * nnz));
  CK(cudaMalloc(&d_csc, sizeof(int) * nnz));
  CK(cudaMalloc(&d_uv, sizeof(double2) * (size_t)nnz));
  CK(cudaMalloc(&d_P, sizeof(double) * 12 * C));
  CK(cudaMalloc(&d_Y, sizeof(double) * 12 * C));
  CK(cudaMalloc(&d_R, sizeof(double) * 32 * C));
  CK(cudaMalloc(&d_rec, sizeof(double) * 8 * (size_t)L));
  CK(cudaMalloc(&d_rec48, sizeof(double) * 6 * (size_t)L));
  CK(cudaMalloc(&d_out, sizeof(double) * grid * kBlock));
  CK(cudaMemcpy(d_cam, obs_cam.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_lm, obs_lm.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_csc, csc_lm.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_uv, 0, sizeof(double2) * (size_t)nnz));
  CK(cudaMemset(d_P, 0, sizeof(double) * 12 * C));
  CK(cudaMemset(d_Y, 0, sizeof(double) * 12 * C));
  CK(cudaMemset(d_R, 0, sizeof(double) * 32 * C));
  CK(cudaMemset(d_rec, 0, sizeof(double) * 8 * (size_t)L));
  CK(cudaMemset(d_rec48, 0, sizeof(double) * 6 * (size_t)L));

  struct Row {
    const char* name;
    double us;
  };
  std::vector<Row> rows;
  rows.push_back({"S0 stream 24 B/obs only", time_us([&] { s0_stream<<<grid, kBlock>>>(nnz, d_cam, d_lm, d_uv, d_out); })});
  rows.push_back({"G0 camera rec, 1 lane/obs, 12 x LDG.128 (today)", time_us([&] { g0_lane<<<grid, kBlock>>>(nnz, d_cam, d_P, d_Y, d_out); })});
  rows.push_back({"G1 camera rec, 8 lanes/obs, whole lines", time_us([&] { g1_coop8<0><<<grid, kBlock>>>(nnz, d_cam, d_R, d_out); })});
  rows.push_back({"G1s same + 3 double shuffles per step", time_us([&] { g1_coop8<1><<<grid, kBlock>>>(nnz, d_cam, d_R, d_out); })});
  rows.push_back({"G2 camera rec, cp.async staging + 11 x LDS.128", time_us([&] { g2_stage<<<grid, kBlock>>>(nnz, d_cam, d_R, d_out); })});
  rows.push_back({"G4 camera rec, 4 lanes/obs, 64-byte planes", time_us([&] { g4_coop4<<<grid, kBlock>>>(nnz, d_cam, d_R, d_out); })});
  rows.push_back({"H0 landmark rec 64 B, 1 lane/entry (today)", time_us([&] { h0_lane<<<grid, kBlock>>>(nnz, d_csc, d_rec, d_out); })});
  rows.push_back({"H1 landmark rec 64 B, 4 lanes/entry", time_us([&] { h1_coop4<0><<<grid, kBlock>>>(nnz, d_csc, d_rec, d_out); })});
  rows.push_back({"H1s same + 3 double shuffles per step", time_us([&] { h1_coop4<1><<<grid, kBlock>>>(nnz, d_csc, d_rec, d_out); })});
  rows.push_back({"H3 landmark rec 48 B, 3 of 4 lanes", time_us([&] { h3_coop4_48<<<grid, kBlock>>>(nnz, d_csc, d_rec48, d_out); })});
  for (const Row& r : rows) {
    std::printf("%-52s %9.1f us  %6.2f ns/kobs\n", r.name, r.us, r.us * 1e6 / nnz);
  }
  return 0;
}
