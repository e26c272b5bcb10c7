// umma_probe.cu -- standalone check of the tcgen05 building blocks used by the
// in-loop controller GEMMs (not part of the product library):
//   D[128 x N] (TMEM, fp32) = A[128 x K] * B[N x K]^T  with bf16x3 split operands,
//   mode 0: A from shared memory (SS), mode 1: A resident in TMEM (TS).
// B is K-major SWIZZLE_128B in shared memory, written by ordinary stores.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ntm_tracker_b200/csrc/ntm_b200_umma.cuh"

using namespace ntm_b200::umma;

constexpr int NT = 512;

// A [128][K] fp32, B [N][K] fp32 -> out [128][N]
__global__ void __launch_bounds__(NT, 1) probe_kernel(const float* A, const float* B, float* out, int N,
                                                      int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: B_hi, B_lo tiles [K/64 atoms][N rows][128 B]; A_hi, A_lo tiles [K/64][128][128 B] (mode 0)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int katoms = K / 64;
  uint8_t* sBhi = smem;
  uint8_t* sBlo = sBhi + katoms * N * 128;
  uint8_t* sAhi = sBlo + katoms * N * 128;
  uint8_t* sAlo = sAhi + katoms * 128 * 128;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) mbar_init(&mbar, 1);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem;                 // N columns
  const uint32_t tA = tmem + 256;           // K/2 hi columns then K/2 lo columns

  // ---- B -> swizzled bf16 hi/lo tiles ----
  for (int i = tid; i < N * K; i += NT) {
    const int n = i / K, k = i % K;
    const float v = B[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const uint32_t off = sw128_offset(n, k, N);
    *reinterpret_cast<__nv_bfloat16*>(sBhi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(sBlo + off) = lo;
  }
  if (mode == 0) {
    for (int i = tid; i < 128 * K; i += NT) {
      const int m = i / K, k = i % K;
      const float v = A[i];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      const uint32_t off = sw128_offset(m, k, 128);
      *reinterpret_cast<__nv_bfloat16*>(sAhi + off) = hi;
      *reinterpret_cast<__nv_bfloat16*>(sAlo + off) = lo;
    }
  } else if (warp < 4) {
    // row m = 32*warp + lane; 8 columns (16 k) per tcgen05.st
    const int m = 32 * warp + lane;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float v0 = A[m * K + k0 + 2 * q], v1 = A[m * K + k0 + 2 * q + 1];
        split_pack_bf16(v0, v1, hi[q], lo[q]);
      }
      const uint32_t lane_addr = (uint32_t)(32 * warp) << 16;
      tmem_st_x8(tA + lane_addr + k0 / 2, hi);
      tmem_st_x8(tA + lane_addr + K / 2 + k0 / 2, lo);
    }
    tmem_wait_st();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(128, N);
      uint32_t accum = 0;
      for (int k0 = 0; k0 < K; k0 += 16) {
        const int atom = k0 / 64, kin = k0 % 64;
        const uint64_t dBhi = make_sw128_desc(sBhi + atom * N * 128 + kin * 2);
        const uint64_t dBlo = make_sw128_desc(sBlo + atom * N * 128 + kin * 2);
        if (mode == 0) {
          const uint64_t dAhi = make_sw128_desc(sAhi + atom * 128 * 128 + kin * 2);
          const uint64_t dAlo = make_sw128_desc(sAlo + atom * 128 * 128 + kin * 2);
          mma_ss(tD, dAhi, dBhi, idesc, accum); accum = 1;
          mma_ss(tD, dAhi, dBlo, idesc, accum);
          mma_ss(tD, dAlo, dBhi, idesc, accum);
        } else {
          mma_ts(tD, tA + k0 / 2, dBhi, idesc, accum); accum = 1;
          mma_ts(tD, tA + k0 / 2, dBlo, idesc, accum);
          mma_ts(tD, tA + K / 2 + k0 / 2, dBhi, idesc, accum);
        }
      }
      mma_commit(&mbar);
    }
    __syncwarp();
  }
  mbar_wait(&mbar, 0);
  tcgen05_fence_after();
  if (warp < 4) {
    const int m = 32 * warp + lane;
    const uint32_t lane_addr = (uint32_t)(32 * warp) << 16;
    for (int n0 = 0; n0 < N; n0 += 8) {
      uint32_t v[8];
      tmem_ld_x8(tD + lane_addr + n0, v);
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q) out[m * N + n0 + q] = __uint_as_float(v[q]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  const int Ns[3] = {64, 80, 16};
  const int Ks[2] = {64, 192};
  int fails = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int ni = 0; ni < 3; ++ni)
      for (int ki = 0; ki < 2; ++ki) {
        const int N = Ns[ni], K = Ks[ki];
        std::vector<float> A(128 * K), B(N * K), out(128 * N, 0.f);
        srand(1 + N + K);
        for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
        for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
        float *dA, *dB, *dO;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dO, 0, out.size() * 4);
        const int smem = 1024 + (K / 64) * (2 * N * 128 + 2 * 128 * 128);
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        probe_kernel<<<1, NT, smem>>>(dA, dB, dO, N, K, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N %d K %d: CUDA error %s\n", mode, N, K, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * (double)B[n * K + k];
            maxerr = fmax(maxerr, fabs(r - out[m * N + n]));
            maxref = fmax(maxref, fabs(r));
          }
        const bool ok = maxerr < 2e-5 * fmax(maxref, 1e-3);
        printf("mode %s N %3d K %3d: max abs err %.3e (max |ref| %.3e) %s\n", mode ? "TS" : "SS", N, K,
               maxerr, maxref, ok ? "OK" : "FAIL");
        fails += !ok;
        cudaFree(dA); cudaFree(dB); cudaFree(dO);
      }
  return fails ? 1 : 0;
}
