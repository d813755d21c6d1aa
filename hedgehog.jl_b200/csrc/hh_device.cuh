// hh_device.cuh — device-side building blocks shared by every kernel of libhedgehog_mc:
// counter-based Philox4x32-10, the Box-Muller normal pair, and block reductions.
//
// RNG convention (restated on the CPU in oracle/hh_oracle.c: hho_normal_pair):
//   key     = per-trajectory seed (hh_sim.seeds[i]) or hh_sim.base_seed
//   counter = (idx_lo, idx_hi, block, stream), idx = global trajectory index (0 with per-path seeds)
//   one Philox block -> 128 bits -> (u1, u2) -> Box-Muller pair (z1, z2)
// Paths therefore never share state and any shard can be re-run bit-identically on any GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hh {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct u32x4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;  // IMAD.WIDE.U32
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;  // LOP3
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return u32x4{c0, c1, c2, c3};
}

// (0,1] and [0,1) uniforms with 53 bits, exactly as the oracle builds them.
__device__ __forceinline__ double u01_open_low(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)((x >> 11) + 1) * 0x1.0p-53;
}
__device__ __forceinline__ double u01_open_high(uint32_t lo, uint32_t hi) {
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)(x >> 11) * 0x1.0p-53;
}

// One Box-Muller pair in binary64.
__device__ __forceinline__ void normal_pair(uint64_t key, uint64_t idx, uint32_t block, uint32_t stream, double &z1,
                                            double &z2) {
  const u32x4 w = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), block, stream, (uint32_t)key,
                                (uint32_t)(key >> 32));
  const double u1 = u01_open_low(w.x, w.y);
  const double u2 = u01_open_high(w.z, w.w);
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z1 = r * c;
  z2 = r * s;
}

// Deterministic block-wide sum of one double per thread (fixed tree: shuffle, then warp 0).
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *scratch /* >= THREADS/32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < THREADS / 32 ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

}  // namespace hh
