// pipes.cu — where do the cycles of the Heston step go? Times the step with pieces removed.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../hedgehog.jl_b200/csrc -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "hh_paths.cuh"
using namespace hh;

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, PhiloxRoundKeys rk, HestonFolded f, double a11, double a12, int M) {
  __shared__ FastNormalTables tb;
  load_fast_tables(&tb);
  __syncthreads();
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  double x = 4.6, v = 0.04;
  uint32_t acc = 0;
  for (int n = 0; n < M; ++n) {
    u32x4 w;
    if (MODE == 0 || MODE == 2) {            // real Philox
      w = philox4x32_10_rk(idx, 0u, (uint32_t)n, 0u, rk);
    } else {                                 // cheap bits (4 IMAD + few LOP)
      w.x = idx * 0x9E3779B9u + n * 0x85EBCA6Bu; w.y = w.x * 0xC2B2AE35u; w.z = w.y ^ (w.x >> 15); w.w = w.z * 0x27D4EB2Fu;
    }
    if (MODE == 0 || MODE == 1) {            // real FP64 work
      double z1, z2;
      fast_normal_pair(&tb, w.x, w.y, w.z, w.w, z1, z2);
      const double dW1 = fma(a12, z2, a11 * z1);
      const double dW2 = fma(f.b22, z2, f.b21 * z1);
      heston_em_step_fast(f, true, x, v, dW1, dW2);
    } else {
      acc ^= w.x ^ w.y ^ w.z ^ w.w;
    }
  }
  if (x + v + acc == 123.456) out[0] = x;
}

template <int MODE>
float run(int M, int blocks) {
  double *d; cudaMalloc(&d, 8);
  upload_fast_tables(0, 0);
  PhiloxRoundKeys rk = philox_round_keys(42);
  HestonFolded f{0.03 / 252, -0.5 / 252, -2.0 / 252, 0.08 / 252, -0.0132, 0.0135};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(d, rk, f, 0.063, 0.0, M);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, rk, f, 0.063, 0.0, M);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const int M = 252 * 16;
  for (int occ = 2; occ <= 6; occ += 2) {
    int blocks = 148 * occ;
    float t0 = run<0>(M, blocks), t1 = run<1>(M, blocks), t2 = run<2>(M, blocks);
    double warp_steps = (double)blocks * 8 * M;
    double cyc = 1.965e6 * 592;  // SMSP-cycles per ms
    printf("blocks/SM=%d  full %.2f ms (%.0f cyc/warp-step)  fp64-only %.2f (%.0f)  philox-only %.2f (%.0f)\n", occ, t0,
           t0 * cyc / warp_steps, t1, t1 * cyc / warp_steps, t2, t2 * cyc / warp_steps);
  }
  return 0;
}
