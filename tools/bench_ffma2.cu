// Micro-benchmark: FP32 FMA throughput of FFMA vs the packed FFMA2 (fma.rn.f32x2, sm_100) on one B200.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* sink, int iters) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        else a[i] = __ffma2_rn(a[i], m, c);
      }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  if (s == 123.456f) sink[0] = s;
}
int main() {
  float* sink; cudaMalloc(&sink, 16);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount * 8, iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, 256>>>(sink, iters); else k<1><<<blocks, 256>>>(sink, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double flops = 2.0 * 2 * 64 * (double)iters * blocks * 256;      // 64 float2 FMAs per iteration per thread
    printf("%s: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2 (packed)" : "FFMA  (scalar)", best, flops / best / 1e9);
  }
  return 0;
}
