// Micro-benchmark: write-bandwidth ceilings of the store patterns available to fg_trajgen on one B200.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/bench_store.cu -o build/bench_store
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// A: each warp owns a 2 x CH-byte stage and stores chunk after chunk with TMA (the fg_trajgen pattern, no compute)
template <int CH>
__global__ void __launch_bounds__(256) k_tma(char* a, char* b, long long nchunks, int spin) {
  extern __shared__ __align__(128) char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  char* st = sm + warp * 2 * 4096;
  for (int i = lane; i < 2 * 4096 / 4; i += 32) ((float*)st)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long long stride = (long long)gridDim.x * 8;
  for (long long c = (long long)blockIdx.x * 8 + warp; c < nchunks; c += stride) {
    if (lane == 0) {
      bulk_wait_read0();
      bulk_store_s2g(a + c * CH, st, CH);
      bulk_store_s2g(b + c * CH, st + 4096, CH);
      bulk_commit();
    }
    if (spin) { float x = lane; for (int i = 0; i < spin; ++i) x = fmaf(x, 1.0001f, 0.5f); if (x == 12345.f) st[0] = 1; }
    __syncwarp();
  }
  if (lane == 0) bulk_wait_all();
}

// C: plain coalesced float4 stores, grid-stride
__global__ void __launch_bounds__(256) k_stg(float4* a, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) a[i] = make_float4(1, 2, 3, 4);
}

// D: each warp writes its chunk with coalesced STG.128 from shared memory (no TMA)
template <int CH>
__global__ void __launch_bounds__(256) k_stg_chunks(char* a, char* b, long long nchunks) {
  extern __shared__ __align__(128) char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  char* st = sm + warp * 2 * 4096;
  for (int i = lane; i < 2 * 4096 / 4; i += 32) ((float*)st)[i] = (float)i;
  __syncwarp();
  const long long stride = (long long)gridDim.x * 8;
  for (long long c = (long long)blockIdx.x * 8 + warp; c < nchunks; c += stride) {
    float4* ga = (float4*)(a + c * CH); float4* gb = (float4*)(b + c * CH);
    for (int i = lane; i < CH / 16; i += 32) { ga[i] = ((float4*)st)[i]; gb[i] = ((float4*)(st + 4096))[i]; }
  }
}

template <typename F> float timeit(F f, int reps = 10) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) f(i);
  CK(cudaDeviceSynchronize());
  std::vector<float> ts;
  for (int i = 0; i < reps; ++i) { cudaEventRecord(e0); f(i); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); ts.push_back(ms); }
  std::sort(ts.begin(), ts.end());
  return ts[reps / 2];
}

int main() {
  const long long n = 1 << 18;                 // chunks (envs)
  const size_t bytes = (size_t)n * 4096;
  char* buf[4];
  for (auto& p : buf) CK(cudaMalloc(&p, bytes));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  auto report = [&](const char* name, float ms, double b) { printf("%-44s %.4f ms  %.0f GB/s\n", name, ms, b / ms / 1e6); };
  report("memset 2 x 1.07 GB", timeit([&](int i) { cudaMemsetAsync(buf[(i & 1) * 2], 0, bytes); cudaMemsetAsync(buf[(i & 1) * 2 + 1], 0, bytes); }), 2.0 * bytes);
  for (int bps : {2, 4, 8}) {
    char nm[80]; snprintf(nm, 80, "STG.128 grid-stride, %d blocks/SM", bps);
    report(nm, timeit([&](int i) { k_stg<<<sms * bps, 256>>>((float4*)buf[i & 1], bytes / 16); }), (double)bytes);
  }
  CK(cudaFuncSetAttribute(k_tma<4000>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_tma<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_tma<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_stg_chunks<4000>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int bps : {1, 2, 3}) {
    for (int spin : {0, 200, 400}) {
      char nm[80]; snprintf(nm, 80, "TMA 2x4000 B/warp-iter, %d blk/SM, spin %d", bps, spin);
      report(nm, timeit([&](int i) { k_tma<4000><<<sms * bps, 256, 65536>>>(buf[(i & 1) * 2], buf[(i & 1) * 2 + 1], n, spin); }), 2.0 * n * 4000);
    }
  }
  report("TMA 2x4096 B/warp-iter, 3 blk/SM", timeit([&](int i) { k_tma<4096><<<sms * 3, 256, 65536>>>(buf[(i & 1) * 2], buf[(i & 1) * 2 + 1], n, 0); }), 2.0 * n * 4096);
  report("TMA 2x2048 B/warp-iter, 3 blk/SM", timeit([&](int i) { k_tma<2048><<<sms * 3, 256, 65536>>>(buf[(i & 1) * 2], buf[(i & 1) * 2 + 1], 2 * n, 0); }), 2.0 * 2 * n * 2048);
  report("STG.128 from smem, 2x4000 B chunks, 3 blk/SM", timeit([&](int i) { k_stg_chunks<4000><<<sms * 3, 256, 65536>>>(buf[(i & 1) * 2], buf[(i & 1) * 2 + 1], n); }), 2.0 * n * 4000);
  return 0;
}
