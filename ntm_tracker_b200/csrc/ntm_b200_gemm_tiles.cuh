// ntm_b200_gemm_tiles.cuh -- tile-record tcgen05 GEMM of the training path:
//
//   out[ks][r, j] = sum_{k in slice ks} ROW[r, k] * COL[j, k]        r < nrows, j < ncols
//
// the dense contractions of the backward pass (what tf.gradients derives for the two _linear projections and
// the BasicLSTMCell matmuls, direct_offset_output.py:611-613 through ntm_cell.py:101-105,124-130,220): per
// timestep the data gradients  d_h = d_raw @ [W_addr|W_out]^T  and  d_[read|h] = d_z @ W_lstm^T, and once per
// training step the weight gradients  X^T @ dZ  over all (t, b).
//
// fp32 accuracy on the bf16 tensor pipe by operand splitting (a = hi + lo, D += hi*hi + hi*lo + lo*hi, fp32
// accumulate in TMEM), exactly like the forward GEMMs (ntm_b200_gemm_ws.cuh).  BOTH operands arrive as tile
// records: for each (block of 128 rows, 64-wide K atom) one contiguous 32 KiB record = the K-major
// SWIZZLE_128B shared-memory image [128 rows][128 B] of the hi half followed by the lo half, written by the
// pack kernels below (any fp32 source with an (i, k) -> address map, so X^T needs no transposed copy).
//
// One CTA per (row block, column block, K slice); warp roles:
//   warp 0      producer: two 32 KiB cp.async.bulk per K atom (row record + column record) into a 3-slot ring
//   warp 1      tcgen05.mma issuer (one lane): 12 MMAs of 128x128x16 per atom, commits release the slots
//   warps 2..5  epilogue: TMEM -> registers -> global (lane = column j: coalesced rows)
// The K slices write their own slabs; the consumer sums them in slice order (deterministic).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ntm_b200_gemm_ws.cuh"
#include "ntm_b200_params.h"
#include "ntm_b200_umma.cuh"

namespace ntm_b200 {
namespace gemmt {

constexpr int THREADS = 192;
constexpr int ATOM_BYTES = gemmws::ATOM_BYTES;     // [128 rows][128 B]
constexpr int REC_BYTES = gemmws::REC_BYTES;       // hi | lo
constexpr int NSLOT = 3;                           // ring slots of (row record + column record) = 64 KiB each

struct Args {
  const uint8_t* rowT;   // [nrb][KAtot][hi | lo]   output rows r   (MMA B operand, N = 128)
  const uint8_t* colT;   // [ncb][KAtot][hi | lo]   output columns j (MMA A operand, M = 128)
  float* out;            // slab ks at out + ks * slab; rows ldo floats apart
  long long slab;
  int ldo, nrows, ncols, nrb, ncb, KAtot, KA, kslices;
};

static __global__ void __launch_bounds__(THREADS, 1) gemm_tiles_kernel(const Args a) {
  using namespace umma;
  extern __shared__ __align__(16) uint8_t gt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gt_smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb = blockIdx.x % a.ncb;
  const int rb = (blockIdx.x / a.ncb) % a.nrb;
  const int ks = blockIdx.x / (a.ncb * a.nrb);
  const int ka0 = ks * a.KA;
  const int nka = min(a.KA, a.KAtot - ka0);
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t full[NSLOT], empty[NSLOT], acc_full;
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 32) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_trigger();      // (programmatic dependent launch along the reverse-time loop: barrier init and the TMEM
  pdl_wait();         // allocation above overlap the predecessor's tail; operands and outputs are touched after this)

  if (warp == 0) {
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int k = 0; k < nka; ++k) {
        mbar_wait(&empty[slot], phase ^ 1u);                       // passes on a fresh barrier
        gemmws::mbar_expect_tx(&full[slot], 2u * REC_BYTES);
        uint8_t* dst = smem + (size_t)slot * 2 * REC_BYTES;
        gemmws::bulk_g2s(dst, a.rowT + ((size_t)rb * a.KAtot + ka0 + k) * REC_BYTES, REC_BYTES, &full[slot]);
        gemmws::bulk_g2s(dst + REC_BYTES, a.colT + ((size_t)cb * a.KAtot + ka0 + k) * REC_BYTES, REC_BYTES, &full[slot]);
        if (++slot == (uint32_t)NSLOT) { slot = 0; phase ^= 1u; }
      }
    }
    __syncwarp();   // the warp reaches the final CTA barrier as one
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(128, 128);
      const uint32_t ring_lo = sw128_desc_lo(smem);     // descriptor low words: slot / half / 16-k step are plain adds
      uint32_t slot = 0, phase = 0;
      for (int k = 0; k < nka; ++k) {
        mbar_wait(&full[slot], phase);
        tcgen05_fence_after();
        const uint32_t rhi = ring_lo + slot * (uint32_t)(2 * REC_BYTES >> 4);   // row operand (B of the MMA)
        const uint32_t rlo = rhi + (uint32_t)(ATOM_BYTES >> 4);
        const uint32_t chi = rhi + (uint32_t)(REC_BYTES >> 4);                  // column operand (A of the MMA)
        const uint32_t clo = chi + (uint32_t)(ATOM_BYTES >> 4);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          mma_ss_lo(tmem, chi + 2u * s, rhi + 2u * s, idesc, (k == 0 && s == 0) ? 0u : 1u);
          mma_ss_lo(tmem, chi + 2u * s, rlo + 2u * s, idesc, 1u);
          mma_ss_lo(tmem, clo + 2u * s, rhi + 2u * s, idesc, 1u);
        }
        mma_commit(&empty[slot]);
        if (++slot == (uint32_t)NSLOT) { slot = 0; phase ^= 1u; }
      }
      mma_commit(&acc_full);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;                            // TMEM lane quarter this warp may access
    const uint32_t lane_addr = (uint32_t)(32 * qd) << 16;
    const int j = cb * 128 + 32 * qd + lane;            // output column = TMEM lane
    float* outp = a.out + (size_t)ks * a.slab + j;
    const long long r0 = (long long)rb * 128;
    mbar_wait(&acc_full, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem + lane_addr + c0, v);
      tmem_wait_ld();
      if (j < a.ncols) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const long long r = r0 + c0 + e;
          if (r < a.nrows) outp[r * a.ldo] = __uint_as_float(v[e]);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- packing: any fp32 source -> tile records ----
// Element (i, k) of the operand (i: row of the operand = output row or column, k: contraction index) is
//   src[i * si + kmap(k) * sk],   kmap(k) = k                        (pT == 0)
//                                 kmap(k) = (k % pB) * pT + k / pB   (k = t * pB + b  ->  row b * pT + t of a
//                                                                    batch-major [pB, pT, .] source)
// and lands at row i_off + i, contraction position k_off + k (k_off % 8 == 0) of the tile set.  Only the
// addressed 8-element chunks are written (zero-filled past nK inside the last chunk): the tile buffer is
// zeroed once when it is laid out, so rows / k beyond the operand stay valid (zero) MMA input.
// i_fast: adjacent threads take adjacent i (for sources contiguous in i, si == 1), else adjacent k chunks.
struct PackArgs {
  const float* src; long long si, sk; int pB, pT;
  int nI, nK; uint8_t* tiles; int KAtot; int i_off, k_off; int i_fast;
};
static __global__ void pack_tiles_kernel(const PackArgs a) {
  const int nch = (a.nK + 7) >> 3;
  const long long total = (long long)a.nI * nch;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int i, ch;
    if (a.i_fast) { i = (int)(idx % a.nI); ch = (int)(idx / a.nI); }
    else { ch = (int)(idx % nch); i = (int)(idx / nch); }
    const int k = ch * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kk = k + e;
      float x = 0.0f;
      if (kk < a.nK) {
        const long long km = a.pT > 0 ? (long long)(kk % a.pB) * a.pT + kk / a.pB : kk;
        x = __ldg(a.src + (long long)i * a.si + km * a.sk);
      }
      v[e] = x;
    }
    gemmws::store_split8(a.tiles, a.KAtot, (long long)a.i_off + i, a.k_off + k, v);
  }
}

struct Plan {
  int nrows, ncols, K, KAtot, KA, kslices, nrb, ncb;
  size_t row_bytes, col_bytes;
};
// kslices_want: how many K slices to aim for (>= 1); slices are evened out in whole atoms.
inline Plan make_plan(int nrows, int ncols, int K, int kslices_want) {
  Plan p{};
  p.nrows = nrows; p.ncols = ncols; p.K = K;
  p.KAtot = (K + 63) / 64;
  int ks = kslices_want < 1 ? 1 : kslices_want;
  if (ks > p.KAtot) ks = p.KAtot;
  p.KA = (p.KAtot + ks - 1) / ks;
  p.kslices = (p.KAtot + p.KA - 1) / p.KA;
  p.nrb = (nrows + 127) / 128;
  p.ncb = (ncols + 127) / 128;
  p.row_bytes = (size_t)p.nrb * p.KAtot * REC_BYTES;
  p.col_bytes = (size_t)p.ncb * p.KAtot * REC_BYTES;
  return p;
}
inline int smem_bytes() { return 1024 + NSLOT * 2 * REC_BYTES; }

static inline cudaError_t launch(const Plan& p, const uint8_t* rowT, const uint8_t* colT, float* out, int ldo, long long slab,
                          cudaStream_t stream, bool pdl = false) {
  Args a{};
  a.rowT = rowT; a.colT = colT; a.out = out; a.slab = slab; a.ldo = ldo; a.nrows = p.nrows; a.ncols = p.ncols;
  a.nrb = p.nrb; a.ncb = p.ncb; a.KAtot = p.KAtot; a.KA = p.KA; a.kslices = p.kslices;
  const int smem = smem_bytes();
  static int configured[MAX_DEVICES] = {0};     // cudaFuncSetAttribute is per device
  {
    std::lock_guard<std::mutex> lk(config_mutex());
    const int dev = current_device_slot();
    if (configured[dev] < smem) {
      cudaError_t e = cudaFuncSetAttribute(gemm_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
      configured[dev] = smem;
    }
  }
  return launch_chain(gemm_tiles_kernel, (unsigned)(p.nrb * p.ncb * p.kslices), (unsigned)THREADS, (size_t)smem, stream, pdl, a);
}

static inline cudaError_t pack(const float* src, long long si, long long sk, int pB, int pT, int nI, int nK, uint8_t* tiles,
                        int KAtot, int i_off, int k_off, bool i_fast, int nsm, cudaStream_t stream) {
  PackArgs a{};
  a.src = src; a.si = si; a.sk = sk; a.pB = pB; a.pT = pT; a.nI = nI; a.nK = nK; a.tiles = tiles; a.KAtot = KAtot;
  a.i_off = i_off; a.k_off = k_off; a.i_fast = i_fast ? 1 : 0;
  const long long total = (long long)nI * ((nK + 7) / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 8ll * nsm) blocks = 8ll * nsm;
  if (blocks < 1) blocks = 1;
  pack_tiles_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace gemmt
}  // namespace ntm_b200
