// Micro-benchmark 2: which store flavour / work assignment reaches the memset write rate?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ void st4(float4* p, float4 v) {
  if constexpr (MODE == 0) *p = v;
  else if constexpr (MODE == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else if constexpr (MODE == 2) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else if constexpr (MODE == 3) asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256) k_stride(float4* a, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) st4<MODE>(a + i, make_float4(1, 2, 3, 4));
}
template <int MODE>
__global__ void __launch_bounds__(256) k_blocked(float4* a, long long n4) {
  const long long per = (n4 + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = min(n4, lo + per);
  for (long long i = lo + threadIdx.x; i < hi; i += 256) st4<MODE>(a + i, make_float4(1, 2, 3, 4));
}
// one block per 256*16*U bytes (non persistent, like a plain elementwise kernel)
template <int U>
__global__ void __launch_bounds__(256) k_flat(float4* a, long long n4) {
  const long long base = ((long long)blockIdx.x * U) * 256 + threadIdx.x;
#pragma unroll
  for (int u = 0; u < U; ++u) { const long long i = base + (long long)u * 256; if (i < n4) a[i] = make_float4(1, 2, 3, 4); }
}

__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_s2g_hint(void* gdst, const void* ssrc, unsigned bytes, uint64_t pol) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(s), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// MODE 0: round-robin envs over warps; 1: each warp a contiguous env range; 2: round-robin + evict_first hint;
// 3: each BLOCK a contiguous range, warps round-robin inside it
template <int MODE>
__global__ void __launch_bounds__(256) k_tma(char* a, char* b, long long n) {
  extern __shared__ __align__(128) char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  char* st = sm + warp * 2 * 4096;
  for (int i = lane; i < 2 * 4096 / 4; i += 32) ((float*)st)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint64_t pol = 0;
  if (MODE == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  long long lo, hi, step;
  const long long W = (long long)gridDim.x * 8, wid = (long long)blockIdx.x * 8 + warp;
  if (MODE == 1) { const long long per = (n + W - 1) / W; lo = wid * per; hi = min(n, lo + per); step = 1; }
  else if (MODE == 3) { const long long per = (n + gridDim.x - 1) / gridDim.x; lo = per * blockIdx.x + warp; hi = min(n, per * (blockIdx.x + 1)); step = 8; }
  else { lo = wid; hi = n; step = W; }
  if (lane == 0) {
    for (long long c = lo; c < hi; c += step) {
      bulk_wait_read0();
      if (MODE == 2) { bulk_store_s2g_hint(a + c * 4000, st, 4000, pol); bulk_store_s2g_hint(b + c * 4000, st + 4096, 4000, pol); }
      else { bulk_store_s2g(a + c * 4000, st, 4000); bulk_store_s2g(b + c * 4000, st + 4096, 4000); }
      bulk_commit();
    }
    bulk_wait_all();
  }
}

__global__ void __launch_bounds__(256) k_stgc(char* a, char* b, long long n) {
  extern __shared__ __align__(128) char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  char* st = sm + warp * 2 * 4096;
  for (int i = lane; i < 2 * 4096 / 4; i += 32) ((float*)st)[i] = (float)i;
  __syncwarp();
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  for (long long c = per * blockIdx.x + warp; c < min(n, per * (blockIdx.x + 1)); c += 8) {
    float4* ga = (float4*)(a + c * 4000); float4* gb = (float4*)(b + c * 4000);
    for (int i = lane; i < 250; i += 32) { ga[i] = ((float4*)st)[i]; gb[i] = ((float4*)(st + 4096))[i]; }
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
  const long long n = 1 << 18;
  const size_t bytes = (size_t)n * 4096;
  char* buf[4];
  for (auto& p : buf) CK(cudaMalloc(&p, bytes));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  auto report = [&](const char* name, float ms, double b) { printf("%-52s %.4f ms  %.0f GB/s\n", name, ms, b / ms / 1e6); };
  const long long n4 = 2 * bytes / 16;       // buf[0],buf[1] are not contiguous: use one 2.1 GB span instead
  float4* big; CK(cudaMalloc(&big, 2 * bytes)); float4* big2; CK(cudaMalloc(&big2, 2 * bytes));
  report("memset 2.1 GB", timeit([&](int i) { cudaMemsetAsync(i & 1 ? big : big2, 0, 2 * bytes); }), 2.0 * bytes);
#define RUN(name, call) report(name, timeit([&](int i) { float4* dst = (i & 1) ? big : big2; call; }), 2.0 * bytes)
  RUN("stride default 4 blk/SM", (k_stride<0><<<sms * 4, 256>>>(dst, n4)));
  RUN("stride .cs", (k_stride<1><<<sms * 4, 256>>>(dst, n4)));
  RUN("stride .wt", (k_stride<2><<<sms * 4, 256>>>(dst, n4)));
  RUN("stride .cg", (k_stride<3><<<sms * 4, 256>>>(dst, n4)));
  RUN("blocked default 4 blk/SM", (k_blocked<0><<<sms * 4, 256>>>(dst, n4)));
  RUN("blocked .cs", (k_blocked<1><<<sms * 4, 256>>>(dst, n4)));
  RUN("blocked default 16 blk/SM", (k_blocked<0><<<sms * 16, 256>>>(dst, n4)));
  RUN("flat U=1 (one float4 per thread)", (k_flat<1><<<(unsigned)((n4 + 255) / 256), 256>>>(dst, n4)));
  RUN("flat U=4", (k_flat<4><<<(unsigned)((n4 + 1023) / 1024), 256>>>(dst, n4)));
  RUN("flat U=16", (k_flat<16><<<(unsigned)((n4 + 4095) / 4096), 256>>>(dst, n4)));
  CK(cudaFuncSetAttribute(k_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
#define RUNT(name, M, bps) report(name, timeit([&](int i) { k_tma<M><<<sms * bps, 256, 65536>>>(buf[(i & 1) * 2], buf[(i & 1) * 2 + 1], n); }), 2.0 * n * 4000)
  RUNT("TMA round-robin 2 blk/SM", 0, 2);
  RUNT("TMA contiguous range per warp 2 blk/SM", 1, 2);
  RUNT("TMA round-robin evict_first 2 blk/SM", 2, 2);
  RUNT("TMA contiguous range per block 2 blk/SM", 3, 2);
  RUNT("TMA contiguous range per block 1 blk/SM", 3, 1);
#define RUNTD(name, epb) report(name, timeit([&](int i) { k_tma<3><<<(unsigned)(n / epb), 256, 65536>>>(buf[(i & 1) * 2], buf[(i & 1) * 2 + 1], n); }), 2.0 * n * 4000)
  RUNTD("TMA dynamic blocks of 16 envs", 16);
  RUNTD("TMA dynamic blocks of 32 envs", 32);
  RUNTD("TMA dynamic blocks of 64 envs", 64);
  RUNTD("TMA dynamic blocks of 256 envs", 256);
  CK(cudaFuncSetAttribute(k_stgc, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int epb : {32, 64}) { char nm[80]; snprintf(nm, 80, "STG.128 from smem, dynamic blocks of %d envs", epb);
    report(nm, timeit([&](int i) { k_stgc<<<(unsigned)(n / epb), 256, 65536>>>(buf[(i & 1) * 2], buf[(i & 1) * 2 + 1], n); }), 2.0 * n * 4000); }
  return 0;
}
