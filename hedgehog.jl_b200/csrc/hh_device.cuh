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
    uint32_t lo0, hi0, lo1, hi1;  // one IMAD.WIDE.U32 each
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(kPhiloxM0));
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(kPhiloxM1));
    const uint32_t n0 = hi1 ^ c1 ^ k0;  // LOP3
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c1 = lo1;
    c3 = lo0;
    c0 = n0;
    c2 = n2;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return u32x4{c0, c1, c2, c3};
}

// Philox with a precomputed key schedule: in base_seed mode the ten round keys are uniform across the grid and
// sit in the kernel arguments (constant bank), so the rounds are 2 IMAD.WIDE + 2 LOP3 with no key arithmetic
// and no registers for the schedule.
struct PhiloxRoundKeys {
  uint32_t k0[10], k1[10];
};
__host__ __device__ __forceinline__ PhiloxRoundKeys philox_round_keys(uint64_t key) {
  PhiloxRoundKeys rk;
  uint32_t a = (uint32_t)key, b = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    rk.k0[r] = a;
    rk.k1[r] = b;
    a += kPhiloxW0;
    b += kPhiloxW1;
  }
  return rk;
}
__device__ __forceinline__ u32x4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  const PhiloxRoundKeys &rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t lo0, hi0, lo1, hi1;
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(kPhiloxM0));
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(kPhiloxM1));
    const uint32_t n0 = hi1 ^ c1 ^ rk.k0[r];
    const uint32_t n2 = hi0 ^ c3 ^ rk.k1[r];
    c1 = lo1;
    c3 = lo0;
    c0 = n0;
    c2 = n2;
  }
  return u32x4{c0, c1, c2, c3};
}

// 52-bit uniforms from one Philox block, exactly as the oracle builds them (oracle/hh_oracle.c: hho_normal_pair):
//   u1 = 1 - n1 2^-52 in [2^-52, 1 - 2^-52],  u2 = n2 2^-52 in [0, 1)
__device__ __forceinline__ double u01_for_log(uint32_t w0, uint32_t w1) {
  return 2.0 - __hiloint2double((int)(0x3FF00000u | (w1 & 0xFFFFFu)), (int)(w0 | 1u));
}
__device__ __forceinline__ double u01_for_angle(uint32_t w2, uint32_t w3) {
  return __hiloint2double((int)(0x3FF00000u | (w3 & 0xFFFFFu)), (int)w2) - 1.0;
}

// One Box-Muller pair in binary64 through the CUDA math library: the reference implementation of the
// mapping; the pricing kernels use the table-driven version in hh_fastnormal.cuh.
__device__ __forceinline__ void normal_pair_libm(const u32x4 w, double &z1, double &z2) {
  const double u1 = u01_for_log(w.x, w.y);
  const double u2 = u01_for_angle(w.z, w.w);
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
