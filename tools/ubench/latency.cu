// Dependent-issue latency of FP64 instructions on sm_100a: ONE warp per SM, one dependent chain, clock64 around it.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu ; run: ./latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, long long *cycles, int iters, double a, double b) {
  double x = a + threadIdx.x * 1e-9;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) x = fma(x, b, a);        // DFMA
    if (OP == 1) x = x * b;               // DMUL
    if (OP == 2) x = x + b;               // DADD
    if (OP == 3) x = fmax(x * b, a);      // DMUL + DSETP/SEL (max)
    if (OP == 4) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = (double)f; }  // conversions round trip
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  double *out;
  long long *cyc, h[4];
  cudaMalloc(&out, 4 * 32 * sizeof(double));
  cudaMalloc(&cyc, 4 * sizeof(long long));
  const int iters = 1 << 16;
  const char *names[] = {"DFMA", "DMUL", "DADD", "DMUL+max", "F2F round trip + FFMA"};
  for (int op = 0; op < 5; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) chain<0><<<1, 32>>>(out, cyc, iters, 1.0, 0.999999);
      if (op == 1) chain<1><<<1, 32>>>(out, cyc, iters, 1.0, 0.999999);
      if (op == 2) chain<2><<<1, 32>>>(out, cyc, iters, 1.0, 0.999999);
      if (op == 3) chain<3><<<1, 32>>>(out, cyc, iters, 1.0, 0.999999);
      if (op == 4) chain<4><<<1, 32>>>(out, cyc, iters, 1.0, 0.999999);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("%-24s %.2f cycles per dependent iteration\n", names[op], (double)h[0] / iters);
  }
  return 0;
}
