// ntm_b200_xproj_tc.cuh -- hoisted input projection on the 5th-gen tensor cores.
//
//   xw[r, :] = x[r, :] @ W_x + b     r = 0 .. B*T-1     (ntm_cell.py:101-105, the `inputs` rows of
//                                                         the layer-0 BasicLSTMCell weights)
//
// Persistent kernel, one CTA per SM.  A CTA owns one 128-column tile of W_x for its whole life:
// the tile is the A operand of tcgen05.mma, split a = hi + lo in bf16 (~2^-18), "hi" resident in
// TENSOR MEMORY (A-from-TMEM MMAs), "lo" resident in shared memory (K-major SWIZZLE_128B).  The CTAs
// that share a row block run concurrently, so x is read from HBM once and from L2 by the rest.
// Per row block of 128 frames the fp32 rows are converted on the fly into a double-buffered bf16
// hi/lo B operand (N = rows) while the previous K block's MMAs run; D[col][row] accumulates in
// TMEM over K = hi*hi + hi*lo + lo*hi and leaves through tcgen05.ld (+ bias) as coalesced stores.
// (x rows are D*4 = 2056 B apart at D = 514 -- not 16-byte aligned -- so TMA cannot stream them;
// the loads are 8-byte vector loads.)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ntm_b200_umma.cuh"

namespace ntm_b200 {

constexpr int XT_THREADS = 512;
constexpr int XT_ROWS = 128;          // frames per row block (MMA N)
constexpr int XT_MAX_KATOMS = 10;     // K <= 640

struct XprojTcArgs {
  const float* x; const float* w; const float* bias; float* out;
  long long rows; int D; int ncols; int ldw;
  int ntiles, ngroups, katoms;
};

__global__ void __launch_bounds__(XT_THREADS, 1) xproj_tc_kernel(const XprojTcArgs a) {
  using namespace umma;
  extern __shared__ __align__(16) uint8_t xt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xt_smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % a.ntiles, group = blockIdx.x / a.ntiles;
  const int katoms = a.katoms, Kpad = katoms * 64;
  const int ksteps = (a.D + 15) / 16;                     // 16-wide MMA k-steps that hold real data
  uint8_t* sWlo = smem;                                   // [katoms][128 cols][128 B]
  uint8_t* sB = sWlo + (size_t)katoms * 128 * 128;        // 2 buffers x {hi, lo} x [128 rows][128 B]
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar[2];
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tAcc = tmem;                             // 128 accumulator columns (rows of the block)
  const uint32_t tWhi = tmem + XT_ROWS;                   // Kpad / 2 columns

  // ---- one-time: this CTA's weight tile, hi -> TMEM, lo -> shared memory ----
  {
    const int j = tile * 128 + 32 * (warp & 3) + lane;    // weight column = TMEM lane = smem row
    const bool jok = j < a.ncols;
    const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    const int row = 32 * (warp & 3) + lane;
    for (int q = warp >> 2; q < Kpad / 16; q += XT_THREADS / 128) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = q * 16 + 2 * e;
        const float v0 = (jok && k < a.D) ? __ldg(a.w + (size_t)k * a.ldw + j) : 0.0f;
        const float v1 = (jok && k + 1 < a.D) ? __ldg(a.w + (size_t)(k + 1) * a.ldw + j) : 0.0f;
        split_pack_bf16(v0, v1, hi[e], lo[e]);
      }
      tmem_st_x8(tWhi + lane_addr + q * 8, hi);
      // the same 16 k as two 16-byte chunks of the swizzled lo tile
      *reinterpret_cast<uint4*>(sWlo + sw128_offset(row, q * 16, 128)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(sWlo + sw128_offset(row, q * 16 + 8, 128)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
    tmem_wait_st();
  }
  const int jcol = tile * 128 + 32 * (warp & 3) + lane;
  const float bias = (jcol < a.ncols) ? __ldg(a.bias + jcol) : 0.0f;
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  const uint32_t idesc = make_idesc_bf16_f32(128, XT_ROWS);
  uint32_t uses[2] = {0, 0};
  const long long nrb = (a.rows + XT_ROWS - 1) / XT_ROWS;
  const int nkb = (ksteps * 16 + 63) / 64;                // 64-wide K blocks with real data
  for (long long rb = group; rb < nrb; rb += a.ngroups) {
    const long long r0 = rb * XT_ROWS;
    for (int kb = 0; kb < nkb; ++kb) {
      const int buf = kb & 1;
      uint8_t* sBhi = sB + (size_t)buf * 2 * XT_ROWS * 128;
      uint8_t* sBlo = sBhi + XT_ROWS * 128;
      if (uses[buf] > 0) {                                // the MMAs that last read this buffer are done
        mbar_wait(&mbar[buf], (uses[buf] - 1) & 1u);
        tcgen05_fence_after();
      }
      // stage x[r0 .. r0+128) x [kb*64 .. +64): 1024 chunks of 8 k, two per thread
#pragma unroll
      for (int it = 0; it < (XT_ROWS * 8) / XT_THREADS; ++it) {
        const int ci = tid + it * XT_THREADS;
        const int r = ci >> 3, c = ci & 7;
        const int k = kb * 64 + c * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.0f;
        if (r0 + r < a.rows) {
          const float* src = a.x + (r0 + r) * (long long)a.D + k;
          if (((a.D & 1) == 0) && k + 8 <= a.D) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 t = __ldg(reinterpret_cast<const float2*>(src) + e);
              v[2 * e] = t.x; v[2 * e + 1] = t.y;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (k + e < a.D) v[e] = __ldg(src + e);
          }
        }
        uint4 h, l;
        split_pack_bf16(v[0], v[1], h.x, l.x);
        split_pack_bf16(v[2], v[3], h.y, l.y);
        split_pack_bf16(v[4], v[5], h.z, l.z);
        split_pack_bf16(v[6], v[7], h.w, l.w);
        const uint32_t off = sw128_offset(r, c * 8, XT_ROWS);
        *reinterpret_cast<uint4*>(sBhi + off) = h;
        *reinterpret_cast<uint4*>(sBlo + off) = l;
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncthreads();
      tcgen05_fence_after();
      if (warp == 0) {
        if (elect_one()) {
          const int ks_in_block = min(4, ksteps - kb * 4);
          for (int s = 0; s < ks_in_block; ++s) {
            const int kk = kb * 64 + s * 16;
            const uint64_t dBhi = make_sw128_desc(sBhi + s * 32);
            const uint64_t dBlo = make_sw128_desc(sBlo + s * 32);
            const uint64_t dWlo = make_sw128_desc(sWlo + (size_t)kb * 128 * 128 + s * 32);
            const uint32_t acc0 = (kb == 0 && s == 0) ? 0u : 1u;
            mma_ts(tAcc, tWhi + kk / 2, dBhi, idesc, acc0);
            mma_ts(tAcc, tWhi + kk / 2, dBlo, idesc, 1u);
            mma_ss(tAcc, dWlo, dBhi, idesc, 1u);
          }
          mma_commit(&mbar[buf]);
        }
        __syncwarp();
      }
      uses[buf] += 1;
    }
    // ---- epilogue: wait for every outstanding MMA of this row block, then TMEM -> xw ----
    for (int b2 = 0; b2 < 2; ++b2)
      if (uses[b2] > 0) mbar_wait(&mbar[b2], (uses[b2] - 1) & 1u);
    tcgen05_fence_after();
    {
      const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
      const int cq0 = (warp >> 2) * (XT_ROWS / (XT_THREADS / 128));   // 32 rows per warp group
      for (int c = 0; c < XT_ROWS / (XT_THREADS / 128); c += 8) {
        uint32_t v[8];
        tmem_ld_x8(tAcc + lane_addr + cq0 + c, v);
        tmem_wait_ld();
        if (jcol < a.ncols) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const long long r = r0 + cq0 + c + e;
            if (r < a.rows) a.out[r * (long long)a.ncols + jcol] = __uint_as_float(v[e]) + bias;
          }
        }
      }
    }
    tcgen05_fence_before();
    __syncthreads();       // all TMEM reads done before the next row block's first MMA overwrites D
    tcgen05_fence_after();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Returns 0 on success, 1 on launch failure, -1 if the shape is outside what the kernel supports
// (the caller then uses the SIMT kernel).
inline int launch_xproj_tc(const float* x, const float* w0, const float* b0, float* xw, long long rows,
                           int D, int ncols, int nsm, cudaStream_t stream) {
  XprojTcArgs a{};
  a.x = x; a.w = w0; a.bias = b0; a.out = xw; a.rows = rows; a.D = D; a.ncols = ncols; a.ldw = ncols;
  a.katoms = (D + 63) / 64;
  if (a.katoms > XT_MAX_KATOMS) return -1;
  a.ntiles = (ncols + 127) / 128;
  if (a.ntiles > nsm) return -1;
  const long long nrb = (rows + XT_ROWS - 1) / XT_ROWS;
  a.ngroups = (int)(nrb < (long long)(nsm / a.ntiles) ? nrb : (long long)(nsm / a.ntiles));
  if (a.ngroups < 1) a.ngroups = 1;
  const int smem = 1024 + a.katoms * 128 * 128 + 2 * 2 * XT_ROWS * 128;
  static int configured = 0;
  if (configured < smem) {
    if (cudaFuncSetAttribute(xproj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return -1;
    }
    configured = smem;
  }
  xproj_tc_kernel<<<a.ntiles * a.ngroups, XT_THREADS, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace ntm_b200
