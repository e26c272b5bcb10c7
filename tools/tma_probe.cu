// tma_probe.cu -- how fast can cp.async.bulk (1-D TMA) stream HBM into shared memory, per SM and per chip,
// as a function of stage size, ring depth and CTAs per SM?  (Design input for the streaming memory kernel.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// mode 0: TMA ring, one thread waits + re-issues.  mode 1: plain LDG.128 by 256 threads, `ns` loads in flight each.
__global__ void __launch_bounds__(256) probe(const float4* src, size_t bytes_per_cta, int stage_bytes, int ns, int mode, float* sink, int reps = 1) {
  extern __shared__ float4 smem[];
  __shared__ uint64_t bars[32];
  const char* base = reinterpret_cast<const char*>(src) + (size_t)blockIdx.x * bytes_per_cta;
  const int nwin = (int)(bytes_per_cta / stage_bytes);
  const int nstage = nwin * reps;
  if (mode == 0) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < ns; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s_u32(&bars[i])));
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      for (int q = 0; q < nstage + ns; ++q) {
        if (q >= ns) mbar_wait(&bars[(q - ns) % ns], ((q - ns) / ns) & 1);
        if (q < nstage) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(&bars[q % ns])), "r"(stage_bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                       ::"r"(s_u32(reinterpret_cast<char*>(smem) + (size_t)(q % ns) * stage_bytes)), "l"(base + (size_t)(q % nwin) * stage_bytes),
                         "r"(stage_bytes), "r"(s_u32(&bars[q % ns])) : "memory");
        }
      }
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(base) + threadIdx.x;
    const size_t n4 = bytes_per_cta / 16 / 256;
    float acc = 0.f;
    for (size_t i = 0; i < n4; i += ns) {
      float4 v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) if (u < ns && i + u < n4) v[u] = __ldcg(p + (i + u) * 256);
#pragma unroll
      for (int u = 0; u < 16; ++u) if (u < ns && i + u < n4) acc += v[u].x + v[u].w;
    }
    if (acc == 12345.f) *sink = acc;
  }
}

int main() {
  const size_t total = 4ull << 30;
  float4* buf; float* sink;
  cudaMalloc(&buf, total); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, total);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("mode ctas/sm stageKB ns inflightKB/cta  GB/s(chip)  GB/s(per SM)\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
      for (int sb : {2048, 8192, 16384, 32768})
        for (int ns : {2, 4, 8, 16}) {
          if (mode == 1 && (sb != 8192)) continue;
          const int smem = mode == 0 ? sb * ns : 0;
          if (smem > 100 * 1024) continue;
          const int grid = 148 * per_sm;
          size_t per_cta = (total / grid) / (64 * 1024) * (64 * 1024);
          if (per_cta > (16u << 20)) per_cta = 16u << 20;
          for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            probe<<<grid, 256, smem>>>(buf, per_cta, sb, ns, mode, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
          }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          cudaError_t err = cudaGetLastError();
          const double gbs = (double)per_cta * grid / (ms * 1e-3) / 1e9;
          printf("%s %d %5.0f %2d %6.0f  %8.0f  %6.1f %s\n", mode ? "ldg" : "tma", per_sm, sb / 1024.0, ns,
                 mode ? ns * 4.0 : sb * ns / 1024.0, gbs, gbs / 148, err ? cudaGetErrorString(err) : "");
        }
  // L2-resident source: every CTA streams the same 32 MiB window over and over (TMA ring, 16 KiB stages)
  printf("L2-resident source (32 MiB window):\nctas/sm stageKB ns  GB/s(chip)  GB/s(per CTA)  implied latency us\n");
  for (int per_sm = 1; per_sm <= 2; ++per_sm)
    for (int sb : {8192, 16384})
      for (int ns : {1, 2, 4}) {
        const int grid = 148 * per_sm;
        const size_t per_cta = (32u << 20) / grid / sb * sb;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          probe<<<grid, 256, sb * ns>>>(buf, per_cta, sb, ns, 0, sink, 64);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double gbs = 64.0 * per_cta * grid / (ms * 1e-3) / 1e9;
        printf("%d %5.0f %2d  %8.0f  %6.1f  %5.2f\n", per_sm, sb / 1024.0, ns, gbs, gbs / grid, sb * ns / (gbs / grid * 1e3));
      }
  return 0;
}
