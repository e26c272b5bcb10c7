// barrier_probe.cu -- latency of device-wide barrier variants on a co-resident grid (1 CTA/SM, 512 threads)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
template <int V>
__global__ void __launch_bounds__(512, 1) k(unsigned* ctr, float* buf, long long* out, int iters) {
  cg::cluster_group cluster = cg::this_cluster();
  unsigned epoch = 0; const unsigned n = gridDim.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    buf[blockIdx.x * 512 + threadIdx.x] = (float)it;           // some global stores to drain, like a real phase
    if (V == 0) {
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory"); while (ld_acq(ctr) < epoch * n) {} }
      __syncthreads();
    } else if (V == 1) {
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; __threadfence(); atomicAdd(ctr, 1u); while (*(volatile unsigned*)ctr < epoch * n) {} __threadfence(); }
      __syncthreads();
    } else if (V == 2) {   // relaxed polling + one acquire fence
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory"); while (ld_rlx(ctr) < epoch * n) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
      __syncthreads();
    } else if (V == 3) {   // one arrival per cluster
      cluster.sync();
      if (cluster.block_rank() == 0 && threadIdx.x == 0) { epoch++; asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory"); while (ld_acq(ctr) < epoch * (n / cluster.num_blocks())) {} }
      cluster.sync();
    } else if (V == 4) {   // cooperative groups grid.sync()
      cg::this_grid().sync();
    }
  }
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}
template <int V> void run(const char* name, int nblk, int cs) {
  unsigned* ctr; float* buf; long long* out;
  cudaMalloc(&ctr, 256); cudaMalloc(&buf, 148 * 512 * 4); cudaMalloc(&out, 148 * 8);
  cudaMemset(ctr, 0, 256);
  int iters = 2000;
  void* args[] = {&ctr, &buf, &out, &iters};
  cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  cfg.gridDim = dim3(nblk); cfg.blockDim = dim3(512); cfg.attrs = at; cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)k<V>, args);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, nblk * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < nblk; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-34s ctas %3d: %7.1f cycles / barrier (%.2f us @1.965GHz) %s %s\n", name, nblk, (double)mx / iters, mx / iters / 1965.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e), e2 == cudaSuccess ? "" : cudaGetErrorString(e2));
  cudaFree(ctr); cudaFree(buf); cudaFree(out);
}
int main() {
  for (int n : {128, 148}) {
    run<0>("red.release + ld.acquire poll", n, 2);
    run<1>("threadfence+atomicAdd+volatile", n, 2);
    run<2>("red.release + relaxed poll + fence", n, 2);
    run<3>("cluster.sync + one arrival/cluster", n, 2);
    run<4>("cooperative_groups grid.sync", n, 2);
  }
  return 0;
}
