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

#include "ntm_b200_params.h"
#include "ntm_b200_umma.cuh"

namespace ntm_b200 {

constexpr int XT_THREADS = 512;
constexpr int XT_ROWS = 128;          // frames per row block (MMA N)
constexpr int XT_MAX_KATOMS = 10;     // K <= 640

// out[ks][r, j] = x[r, k0 .. k0+kw) @ w[k0 .. k0+kw, j] (+ bias[j] when bias != null; K-slice ks starts
// at k0 = ks * kw).  x rows are ldx floats apart, out rows ldo floats apart, slice ks writes its own
// slab at out + ks * slab (the consumer sums the slabs in slice order: deterministic split-K).
struct XprojTcArgs {
  const float* x; const float* w; const float* bias; float* out;
  long long rows; int D; int ncols; int ldw;
  int ntiles, ngroups, katoms;
  int ldx, ldo, kslices, kw;   // D = total K; kw = slice width (multiple of 64 when kslices > 1)
  long long slab;
  int vec2;                    // x rows and slice starts are 8-byte aligned
};

__global__ void __launch_bounds__(XT_THREADS, 1) xproj_tc_kernel(const XprojTcArgs a0) {
  // this CTA's K-slice as a self-contained problem
  XprojTcArgs a = a0;
  const int kslice = (blockIdx.x / a0.ntiles) % a0.kslices;
  {
    const int k0 = kslice * a0.kw;
    a.D = min(a0.kw, a0.D - k0);
    a.x = a0.x + k0;
    a.w = a0.w + (size_t)k0 * a0.ldw;
    a.out = a0.out + (size_t)kslice * a0.slab;
    a.katoms = (a.D + 63) / 64;
  }
  using namespace umma;
  extern __shared__ __align__(16) uint8_t xt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xt_smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % a.ntiles, group = blockIdx.x / (a.ntiles * a.kslices);
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
  const float bias = (a.bias != nullptr && jcol < a.ncols) ? __ldg(a.bias + jcol) : 0.0f;
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
          const float* src = a.x + (r0 + r) * (long long)a.ldx + k;
          if (a.vec2 && k + 8 <= a.D) {
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
            if (r < a.rows) a.out[r * (long long)a.ldo + jcol] = __uint_as_float(v[e]) + bias;
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
// (the caller then uses the SIMT kernel).  General form: split-K over `kslices` slabs.
inline int launch_gemm_tc(const float* x, int ldx, const float* w, int ldw, const float* bias, float* out,
                          int ldo, long long slab, long long rows, int K, int ncols, int kslices, int nsm,
                          cudaStream_t stream) {
  XprojTcArgs a{};
  a.x = x; a.w = w; a.bias = bias; a.out = out; a.rows = rows; a.D = K; a.ncols = ncols; a.ldw = ldw;
  a.ldx = ldx; a.ldo = ldo; a.kslices = kslices; a.slab = slab;
  a.kw = kslices > 1 ? ((K + kslices - 1) / kslices + 63) / 64 * 64 : K;
  if (kslices > 1 && (long long)(kslices - 1) * a.kw >= K) return -1;   // an empty last slice
  a.katoms = (a.kw + 63) / 64;
  if (a.katoms > XT_MAX_KATOMS) return -1;
  a.vec2 = ((ldx & 1) == 0 && (kslices == 1 || (a.kw & 1) == 0) && (reinterpret_cast<uintptr_t>(x) & 7) == 0) ? 1 : 0;
  a.ntiles = (ncols + 127) / 128;
  const int units = a.ntiles * kslices;
  if (units > nsm) return -1;
  const long long nrb = (rows + XT_ROWS - 1) / XT_ROWS;
  a.ngroups = (int)(nrb < (long long)(nsm / units) ? nrb : (long long)(nsm / units));
  if (a.ngroups < 1) a.ngroups = 1;
  const int smem = 1024 + a.katoms * 128 * 128 + 2 * 2 * XT_ROWS * 128;
  static int configured[MAX_DEVICES] = {0};     // cudaFuncSetAttribute is per device
  {
    std::lock_guard<std::mutex> lk(config_mutex());
    const int dev = current_device_slot();
    if (configured[dev] < smem) {
      if (cudaFuncSetAttribute(xproj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
        cudaGetLastError();
        return -1;
      }
      configured[dev] = smem;
    }
  }
  xproj_tc_kernel<<<units * a.ngroups, XT_THREADS, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Smallest number of K-slices whose (64-rounded) width fits the resident weight tile.
inline int gemm_tc_kslices(int K) {
  int ks = 1;
  while (((K + ks - 1) / ks + 63) / 64 > XT_MAX_KATOMS) ++ks;
  return ks;
}

inline int launch_xproj_tc(const float* x, const float* w0, const float* b0, float* xw, long long rows,
                           int D, int ncols, int nsm, cudaStream_t stream) {
  return launch_gemm_tc(x, D, w0, ncols, b0, xw, ncols, 0, rows, D, ncols, 1, nsm, stream);
}

}  // namespace ntm_b200
