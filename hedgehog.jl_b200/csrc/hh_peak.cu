// hh_peak.cu — FP64 roofline denominator: DFMA-chain microbenchmark (8 independent chains per thread,
// everything in registers). 2 FLOP per DFMA. Used by bench.py because MEASURED_PEAKS.json has no FP64 entry.
#include "hh_ctx.h"

namespace hh {

constexpr int kPeakIters = 4096;
constexpr int kPeakChains = 8;

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, double a, double b) {
  double x[kPeakChains];
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) x[c] = (double)(threadIdx.x + c) * 1e-3;
#pragma unroll 1
  for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep)
#pragma unroll
      for (int c = 0; c < kPeakChains; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) s += x[c];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

int fp64_peak(hh_ctx *ctx, double *tflops, double *ms_out) {
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  HH_CUDA(ctx, ctx->d_misc.ensure(64));
  cudaStream_t st = ctx->stream;
  const int blocks = ctx->sm_count * 8;
  for (int w = 0; w < 2; ++w) dfma_peak_kernel<<<blocks, 256, 0, st>>>(ctx->d_misc.as<double>(), 0.999999, 1e-9);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    HH_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(ctx->d_misc.as<double>(), 0.999999, 1e-9);
    HH_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    HH_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    HH_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms < best) best = ms;
  }
  const double flops = 2.0 * (double)blocks * 256.0 * kPeakIters * 4.0 * kPeakChains;
  if (tflops) *tflops = flops / (best * 1e-3) * 1e-12;
  if (ms_out) *ms_out = best;
  return HH_OK;
}

}  // namespace hh
