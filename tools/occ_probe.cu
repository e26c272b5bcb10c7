// co-residency probe: do two CTAs that each allocate 256 TMEM columns run on one SM at the same time?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
__global__ void __launch_bounds__(256, 2) k(unsigned long long* rec, uint32_t ncols, long long spin) {
  extern __shared__ float s[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(a), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  __syncthreads();
  unsigned long long t0, t1; uint32_t smid;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  long long c0 = clock64();
  while (clock64() - c0 < spin) { s[threadIdx.x] += 1.0f; }
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) { rec[blockIdx.x * 3] = t0; rec[blockIdx.x * 3 + 1] = t1; rec[blockIdx.x * 3 + 2] = smid; }
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t t = slot;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(t), "r"(ncols) : "memory");
  }
}
int main() {
  const int grid = 296, smem = 100000;
  unsigned long long* d; cudaMalloc(&d, grid * 3 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (uint32_t ncols : {256u, 512u}) {
    k<<<grid, 256, smem>>>(d, ncols, 2000000);   // ~1 ms spin
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned long long> h(grid * 3);
    cudaMemcpy(h.data(), d, grid * 3 * 8, cudaMemcpyDeviceToHost);
    int overlaps = 0;
    for (int i = 0; i < grid; ++i)
      for (int j = i + 1; j < grid; ++j)
        if (h[i * 3 + 2] == h[j * 3 + 2]) {
          unsigned long long lo = std::max(h[i * 3], h[j * 3]), hi = std::min(h[i * 3 + 1], h[j * 3 + 1]);
          if (hi > lo && hi - lo > 500000) overlaps++;   // > 0.5 ms of common residency
        }
    unsigned long long tmin = ~0ull, tmax = 0;
    for (int i = 0; i < grid; ++i) { tmin = std::min(tmin, h[i * 3]); tmax = std::max(tmax, h[i * 3 + 1]); }
    printf("ncols %u: %s, pairs of CTAs co-resident on one SM: %d (of %d possible), total %.2f ms\n", ncols,
           cudaGetErrorString(e), overlaps, grid / 2, (tmax - tmin) / 1e6);
  }
  return 0;
}
