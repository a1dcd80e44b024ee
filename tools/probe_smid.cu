// How does the hardware block scheduler spread a sub-wave grid over the SMs?  Launches grids shaped like the rollout
// kernel's (blocks x threads, ~128 registers, 14 KB dynamic smem), every block records its SM id and spins ~50 us.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(int* smid, long long spin) {
  extern __shared__ float sm[];
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  if (threadIdx.x == 0) smid[blockIdx.x] = (int)id;
  const long long t0 = clock64();
  float acc = threadIdx.x;
  while (clock64() - t0 < spin) acc = acc * 1.0001f + 0.5f;
  if (acc == 12345.f) sm[threadIdx.x] = acc;
}
template <int THREADS, int MINB>
void run(int blocks) {
  int* d; cudaMalloc(&d, blocks * sizeof(int));
  cudaFuncSetAttribute(k<THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * 1024);
  k<THREADS, MINB><<<blocks, THREADS, 14 * 1024>>>(d, 100000);
  cudaDeviceSynchronize();
  std::vector<int> h(blocks); cudaMemcpy(h.data(), d, blocks * sizeof(int), cudaMemcpyDeviceToHost);
  std::vector<int> per(256, 0); for (int s : h) per[s]++;
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  std::vector<int> hist(64, 0); for (int s = 0; s < nsm; ++s) hist[per[s]]++;
  printf("%d blocks x %d threads (min %d blocks/SM) on %d SMs: blocks-per-SM histogram:", blocks, THREADS, MINB, nsm);
  for (int c = 0; c < 64; ++c) if (hist[c]) printf("  %d SMs x %d", hist[c], c);
  printf("   [%s]\n", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<128, 4>(512); run<64, 8>(1024); run<32, 16>(2048); run<128, 4>(8192);
  return 0;
}
