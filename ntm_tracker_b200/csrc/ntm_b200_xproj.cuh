// ntm_b200_xproj.cuh -- hoisted input projection of the controller LSTM.
//
// xw[r, :] = x[r, :] @ W_x + b      r = 0 .. B*T-1,  W_x = rows [0, D) of the layer-0
// BasicLSTMCell weights (the `inputs` part of tf.concat([inputs, read_prev], 1) @ W,
// ntm_cell.py:101-105).  It has no dependence on the recurrent state, so it is
// taken out of the T-step loop and done as one GEMM over all B*T rows.
//
// Round-1 implementation: fp32 SIMT tiled GEMM (64x64 tile, 4x4 per thread).
#pragma once
#include <cuda_runtime.h>

namespace ntm_b200 {

constexpr int XP_BM = 64, XP_BN = 64, XP_BK = 16, XP_THREADS = 256;

__global__ void __launch_bounds__(XP_THREADS) xproj_kernel(const float* __restrict__ A,
                                                           const float* __restrict__ Bm,
                                                           const float* __restrict__ bias,
                                                           float* __restrict__ Cm, long long rows,
                                                           int K, int ncols) {
  __shared__ __align__(16) float As[XP_BK][XP_BM + 4];
  __shared__ __align__(16) float Bs[XP_BK][XP_BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * XP_BM;
  const int n0 = blockIdx.y * XP_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += XP_BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + j * XP_THREADS;
      const int r = idx >> 4, kk = idx & 15;
      float v = 0.0f;
      if (m0 + r < rows && k0 + kk < K) v = __ldg(A + (m0 + r) * (long long)K + k0 + kk);
      As[kk][r] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + j * XP_THREADS;
      const int kk = idx >> 6, c = idx & 63;
      float v = 0.0f;
      if (k0 + kk < K && n0 + c < ncols) v = __ldg(Bm + (long long)(k0 + kk) * ncols + n0 + c);
      Bs[kk][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < XP_BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = m0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < ncols) Cm[r * (long long)ncols + c] = acc[i][j] + __ldg(bias + c);
    }
  }
}

inline int launch_xproj(const float* x, const float* w0, const float* b0, float* xw, long long rows,
                        int D, int ncols, cudaStream_t stream) {
  dim3 grid((unsigned)((rows + XP_BM - 1) / XP_BM), (ncols + XP_BN - 1) / XP_BN);
  xproj_kernel<<<grid, XP_THREADS, 0, stream>>>(x, w0, b0, xw, rows, D, ncols);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace ntm_b200
