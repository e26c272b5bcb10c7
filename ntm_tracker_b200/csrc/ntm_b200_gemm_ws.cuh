// ntm_b200_gemm_ws.cuh -- warp-specialised tcgen05 GEMM of the streaming mode:
//
//   out[ks][r, j] = sum_{k in slice ks} act[r, k] * W[k, j]  (+ bias[j])       r < rows, j < ncols
//
// the controller projection [read | h] @ W_rh and the head-parameter projection h @ [W_addr | W_out]
// (ntm_cell.py:101-105, 113-130, 220) for ALL sequences of the shard at once.
//
// fp32 accuracy on the bf16 tensor pipe: both operands are split a = hi + lo (bf16 each, ~2^-18
// relative) and D += hi*hi + hi*lo + lo*hi in fp32.  The split is done ONCE by the producers, not in
// this kernel: activations arrive as ready-made operand tiles -- for each (128-row block, 64-wide K atom)
// a contiguous 32 KiB record = the K-major SWIZZLE_128B shared-memory image of the hi half followed by
// the lo half -- written by the memory kernel (read vectors) and the LSTM kernel (hidden state); weights
// are packed the same way once per call.
//
// One persistent CTA per SM owns one (128-column weight tile, K-slice of KA atoms): weight hi half
// resident in TMEM (A-from-TMEM MMAs), lo half resident in shared memory.  Roles:
//   warp 0      producer: one 32 KiB cp.async.bulk per (row block, K atom) into a ring of 3 slots (7 when the
//               K slice is <= 256 wide: then the weight lo half lives in TMEM as well and shared memory is all ring)
//   warp 1      tcgen05.mma issuer (one elected lane); commits release ring slots / publish accumulators
//   warps 2..9  epilogue: TMEM -> registers -> global, double-buffered against the next row block's MMAs (two warps
//               per TMEM lane quarter, 64 accumulator columns each)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "ntm_b200_params.h"
#include "ntm_b200_umma.cuh"

namespace ntm_b200 {
namespace gemmws {

#ifndef NTM_KA_MAX
#define NTM_KA_MAX 8
#endif
constexpr int KA_MAX = NTM_KA_MAX;   // K atoms (64 wide) per slice: 256 TMEM columns of weight hi halves at 8
constexpr int NSLOT_MAX = 7;         // activation ring slots (32 KiB each): 3 next to a resident weight lo half in
                                     // shared memory, 7 when both weight halves fit TMEM (K slice <= 256)
constexpr int THREADS = 320;          // producer warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int ATOM_BYTES = 16384;    // [128 rows][128 B]
constexpr int REC_BYTES = 2 * ATOM_BYTES;

struct Args {
  const uint8_t* act;    // [row blocks][KAtot][hi 16 KiB | lo 16 KiB]
  const uint32_t* whi;   // [tiles][slices][KA * 4 16-k steps][128 cols][8] packed bf16 pairs (k, k+1)
  const uint8_t* wlo;    // [tiles][slices][KA][16 KiB swizzled image]
  const float* bias;     // [ncols] or null
  float* out;            // slab ks at out + ks * slab, rows ldo apart
  long long rows, slab;
  int ldo, ncols, ntiles, kslices, ngroups, KAtot, KA;
  int nslot;             // ring slots
  int wlo_tmem;          // 1: weight lo half in TMEM too (KA <= 4; `wlo` then holds packed words like `whi`)
  int exp;               // experiment switches (EnvSwitches::exp)
  long long* prof;       // development (NTM_B200_EXP bit 8): [grid][8] ns stamps / wait sums, or null
};
__device__ __forceinline__ long long ws_gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(umma::smem_u32(dst)), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}

static __global__ void __launch_bounds__(THREADS, 1) gemm_ws_kernel(const Args a) {
  using namespace umma;
  extern __shared__ __align__(16) uint8_t ws_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws_smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % a.ntiles;
  const int ks = (blockIdx.x / a.ntiles) % a.kslices;
  const int group = blockIdx.x / (a.ntiles * a.kslices);
  const int ka0 = ks * a.KA;
  const int nka = min(a.KA, a.KAtot - ka0);           // atoms of this slice
  const int NSLOT = a.nslot;
  uint8_t* sWlo = smem;                                // [KA][16 KiB] (absent when the lo half lives in TMEM)
  uint8_t* sRing = smem + (a.wlo_tmem ? 0 : (size_t)a.KA * ATOM_BYTES);   // [NSLOT][hi | lo]
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t full[NSLOT_MAX], empty[NSLOT_MAX], acc_full[2], acc_empty[2], wbar, whibar;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(&wbar, 1);
    mbar_init(&whibar, 8);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_trigger();       // the next kernel of the chain may become resident (it waits for this grid before reading)
  const uint32_t tmem = tmem_slot;
  if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 8 + 0] = ws_gtimer();
  const uint32_t tWhi = tmem + 256;                    // accumulators: columns 0..127 and 128..255
  const uint32_t tWlo = tWhi + 128;                    // (wlo_tmem: KA <= 4, so each half takes <= 128 columns)
  const long long nrb = (a.rows + 127) / 128;
  const size_t wrec = (size_t)tile * a.kslices + ks;   // this CTA's weight record

  if (warp == 0) {
    // ------------------------------------------------ producer ------------------------------------
    if (!a.wlo_tmem) {      // weight lo half: one 16 KiB bulk copy per K atom, each from its own lane (~240 ns per issue)
      if (lane == 0) mbar_expect_tx(&wbar, (uint32_t)nka * ATOM_BYTES);
      for (int k = lane; k < nka; k += 32)
        bulk_g2s(sWlo + (size_t)k * ATOM_BYTES, a.wlo + (wrec * a.KA + k) * ATOM_BYTES, ATOM_BYTES, &wbar);
    }
    // Two issuing lanes: lane 0 brings the hi half of a record, lane 1 the lo half (a bulk copy costs its
    // issuing thread ~240 ns and one thread sustains ~64 GB/s; two copies in flight per slot halve the time a
    // slot spends filling).  Lane 0 alone arms the barrier with the whole record's byte count: the phase cannot
    // complete before that arrive, whatever the order in which the two copies land.
    // (four lanes with a quarter record each: the copy's flight time is latency + bytes / 64 GB/s per issuing
    // thread, and with three slots that time is what the MMA warp ends up waiting for)
    const int nissue = (a.exp & 1) ? 2 : 4;
    if (lane < nissue) {
      uint32_t slot = 0, phase = 0;      // ring position kept incrementally (no division on this thread's path)
      const uint32_t bytes = (uint32_t)REC_BYTES / (uint32_t)nissue;
      const size_t half = (size_t)lane * bytes;
      long long wsum = 0;
      pdl_wait();                        // activations are the previous kernel's output
      for (long long rb = group; rb < nrb; rb += a.ngroups) {
        for (int k = 0; k < nka; ++k) {
          const long long tw = a.prof ? ws_gtimer() : 0;
          mbar_wait(&empty[slot], phase ^ 1u);                     // passes on a fresh barrier
          if (a.prof) wsum += ws_gtimer() - tw;
          if (lane == 0) mbar_expect_tx(&full[slot], REC_BYTES);
          bulk_g2s(sRing + (size_t)slot * REC_BYTES + half, a.act + ((size_t)rb * a.KAtot + ka0 + k) * REC_BYTES + half, bytes,
                   &full[slot]);
          if (++slot == (uint32_t)NSLOT) { slot = 0; phase ^= 1u; }
        }
      }
      if (a.prof && lane == 0) a.prof[(size_t)blockIdx.x * 8 + 4] = wsum;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer ----------------------------------
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(128, 128);
      // descriptor low words: ring slot / weight atom / 16-k step are plain adds of (bytes >> 4)
      const uint32_t ring_lo = sw128_desc_lo(sRing), wlo_lo = sw128_desc_lo(sWlo);
      if (!a.wlo_tmem) mbar_wait(&wbar, 0);   // weight lo half in shared memory (bulk copies)
      mbar_wait(&whibar, 0);     // weight hi (and lo) half in TMEM (stored by the four epilogue warps)
      tcgen05_fence_after();
      long long wfull = 0, wacc = 0;
      if (a.prof) a.prof[(size_t)blockIdx.x * 8 + 1] = ws_gtimer();
      uint32_t slot = 0, phase = 0, it = 0;
      for (long long rb = group; rb < nrb; rb += a.ngroups, ++it) {
        const uint32_t ab = it & 1u;
        const long long ta = a.prof ? ws_gtimer() : 0;
        mbar_wait(&acc_empty[ab], ((it >> 1) & 1u) ^ 1u);          // epilogue drained this accumulator
        if (a.prof) wacc += ws_gtimer() - ta;
        tcgen05_fence_after();
        const uint32_t tAcc = tmem + ab * 128;
        for (int k = 0; k < nka; ++k) {
          const long long tf = a.prof ? ws_gtimer() : 0;
          mbar_wait(&full[slot], phase);
          if (a.prof) wfull += ws_gtimer() - tf;
          tcgen05_fence_after();
          const uint32_t bhi = ring_lo + slot * (uint32_t)(REC_BYTES >> 4), blo = bhi + (uint32_t)(ATOM_BYTES >> 4);
          const uint32_t wlo = wlo_lo + (uint32_t)k * (uint32_t)(ATOM_BYTES >> 4);
          const uint32_t tA = tWhi + (uint32_t)k * 32u, tAl = tWlo + (uint32_t)k * 32u;
#pragma unroll
          for (int s = 0; s < 4; ++s) {     // 16-k steps: 32 bytes along the swizzled row, 8 TMEM columns
            mma_ts_lo(tAcc, tA + 8u * s, bhi + 2u * s, idesc, (k == 0 && s == 0) ? 0u : 1u);
            mma_ts_lo(tAcc, tA + 8u * s, blo + 2u * s, idesc, 1u);
            if (a.wlo_tmem) mma_ts_lo(tAcc, tAl + 8u * s, bhi + 2u * s, idesc, 1u);
            else mma_ss_lo(tAcc, wlo + 2u * s, bhi + 2u * s, idesc, 1u);
          }
          mma_commit(&empty[slot]);                                // slot reusable once these MMAs have read it
          if (++slot == (uint32_t)NSLOT) { slot = 0; phase ^= 1u; }
        }
        mma_commit(&acc_full[ab]);
      }
      if (a.prof) {
        a.prof[(size_t)blockIdx.x * 8 + 2] = ws_gtimer();
        a.prof[(size_t)blockIdx.x * 8 + 5] = wfull;
        a.prof[(size_t)blockIdx.x * 8 + 6] = wacc;
      }
    }
    __syncwarp();   // the warp reaches the final CTA barrier as one (a partial warp must not be counted as arrived)
  } else {
    // ------------------------------------------------ epilogue warps ------------------------------
    const int qd = warp & 3;                            // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;                  // two warps share a quarter: accumulator columns [64*chalf, +64)
    const uint32_t lane_addr = (uint32_t)(32 * qd) << 16;
    const int jl = 32 * qd + lane;                      // weight column within the tile = TMEM lane
    const int jcol = tile * 128 + jl;
    // weight hi half -> TMEM (once)
    {
      // 16 k = 8 packed words per step; SB steps' loads are issued together (the staging is a chain of L2 round
      // trips otherwise, and at small batches it is most of the kernel; 8 or 16 steps per batch measure the same 4.5 us)
#ifndef NTM_STAGE_SB
#define NTM_STAGE_SB 4
#endif
      constexpr int SB = NTM_STAGE_SB;
      auto stage_weights = [&](const uint32_t* src, uint32_t tdst) {   // src: this lane's 8 words of 16-k step 0
        const int nq = nka * 4;
        for (int q0 = SB * chalf; q0 < nq; q0 += 2 * SB) {    // the quarter's two warps alternate blocks of SB steps
          uint4 w4[2 * SB];
#pragma unroll
          for (int i = 0; i < SB; ++i) {
            const int q = min(q0 + i, nq - 1);
            w4[2 * i] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)q * 1024));
            w4[2 * i + 1] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)q * 1024 + 4));
          }
#pragma unroll
          for (int i = 0; i < SB; ++i) {
            if (q0 + i < nq) {
              const uint32_t v[8] = {w4[2 * i].x, w4[2 * i].y, w4[2 * i].z, w4[2 * i].w,
                                     w4[2 * i + 1].x, w4[2 * i + 1].y, w4[2 * i + 1].z, w4[2 * i + 1].w};
              tmem_st_x8(tdst + (q0 + i) * 8, v);
            }
          }
        }
      };
      // packed words [weight record][16-k step][128 columns][8 words]: a warp's load covers 1 KiB contiguous
      stage_weights(a.whi + (wrec * (size_t)(a.KA * 4) * 128 + jl) * 8, tWhi + lane_addr);
      if (a.wlo_tmem)
        stage_weights(reinterpret_cast<const uint32_t*>(a.wlo) + (wrec * (size_t)(a.KA * 4) * 128 + jl) * 8, tWlo + lane_addr);
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&whibar);
    }
    pdl_wait();   // the output buffer may still be read by the kernel before this one
    const float bias = (a.bias != nullptr && jcol < a.ncols) ? __ldg(a.bias + jcol) : 0.0f;
    float* outp = a.out + (size_t)ks * a.slab + jcol;
    uint32_t it = 0;
    for (long long rb = group; rb < nrb; rb += a.ngroups, ++it) {
      const uint32_t ab = it & 1u;
      mbar_wait(&acc_full[ab], (it >> 1) & 1u);
      tcgen05_fence_after();
      const long long r0 = rb * 128;
      const uint32_t tAcc = tmem + ab * 128 + lane_addr;
      if (a.wlo_tmem) {
        // short K (few MMAs per row block): the epilogue is on the critical path, and a tcgen05.ld queues
        // behind the MMAs of the next row block -- two 32-column loads in flight per wait (38.5 vs 42 us at C3)
#pragma unroll 1
        for (int c0 = 64 * chalf; c0 < 64 * chalf + 64; c0 += 64) {
          uint32_t v0[32], v1[32];
          tmem_ld_x32(tAcc + c0, v0);
          tmem_ld_x32(tAcc + c0 + 32, v1);
          tmem_wait_ld();
          // the accumulator is in registers: hand it back BEFORE the 64 row stores (the MMA warp waited for those
          // stores: 5.8 of its 20 us per launch)
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
          if (jcol < a.ncols) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const long long r = r0 + c0 + e;
              if (r < a.rows) outp[r * a.ldo] = __uint_as_float(v0[e]) + bias;
            }
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const long long r = r0 + c0 + 32 + e;
              if (r < a.rows) outp[r * a.ldo] = __uint_as_float(v1[e]) + bias;
            }
          }
        }
      } else {
        // long K: the epilogue hides behind the next row block's MMAs; narrow loads disturb their
        // A-from-TMEM operand reads least (57 vs 64 us at C3 with the wide ones)
#pragma unroll 1
        for (int c0 = 64 * chalf; c0 < 64 * chalf + 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tAcc + c0, v);
          tmem_wait_ld();
          if (jcol < a.ncols) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const long long r = r0 + c0 + e;
              if (r < a.rows) outp[r * a.ldo] = __uint_as_float(v[e]) + bias;
            }
          }
        }
      }
      if (!a.wlo_tmem) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 8 + 3] = ws_gtimer();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Producer-side store: 8 consecutive k (k % 8 == 0) of row r into the operand tiles.
__device__ __forceinline__ void store_split8(uint8_t* tiles, int KAtot, long long r, int k, const float (&v)[8]) {
  uint4 h, l;
  umma::split_pack_bf16(v[0], v[1], h.x, l.x);
  umma::split_pack_bf16(v[2], v[3], h.y, l.y);
  umma::split_pack_bf16(v[4], v[5], h.z, l.z);
  umma::split_pack_bf16(v[6], v[7], h.w, l.w);
  const int row = (int)(r & 127);
  uint8_t* rec = tiles + ((size_t)(r >> 7) * KAtot + (k >> 6)) * REC_BYTES +
                 (row >> 3) * 1024 + (row & 7) * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(rec) = h;
  *reinterpret_cast<uint4*>(rec + ATOM_BYTES) = l;
}

// ---- packing kernels ----
// fp32 [rows, K] (row stride ld) -> operand tiles [row block][KAtot][hi | lo]; rows / k beyond the matrix
// are written as zeros so the whole record is valid MMA input.  One thread per 8 consecutive k.
static __global__ void pack_act_tiles_kernel(const float* __restrict__ x, long long rows, int K, int ld, uint8_t* tiles,
                                      int KAtot, int k_off) {
  const long long nrb = (rows + 127) / 128;
  const int KAsrc = (K + 63) / 64;
  const long long total = nrb * 128 * (long long)KAsrc * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(i & 7);
    long long t = i >> 3;
    const int row = (int)(t & 127); t >>= 7;
    const int ka = (int)(t % KAsrc);
    const long long rb = t / KAsrc;
    const long long r = rb * 128 + row;
    const int k = ka * 64 + chunk * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (r < rows && k + e < K) ? x[r * (long long)ld + k + e] : 0.0f;
    uint4 h, l;
    umma::split_pack_bf16(v[0], v[1], h.x, l.x);
    umma::split_pack_bf16(v[2], v[3], h.y, l.y);
    umma::split_pack_bf16(v[4], v[5], h.z, l.z);
    umma::split_pack_bf16(v[6], v[7], h.w, l.w);
    const int kk = k_off + k;                       // position in the destination K axis (multiple of 8)
    if (kk >= KAtot * 64) continue;
    uint8_t* rec = tiles + ((size_t)rb * KAtot + (kk >> 6)) * REC_BYTES +
                   (row >> 3) * 1024 + (row & 7) * 128 + ((((kk & 63) >> 3) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(rec) = h;
    *reinterpret_cast<uint4*>(rec + ATOM_BYTES) = l;
  }
}

// W [K, ncols] (row stride ldw) -> per (tile, slice): hi words [KA*4 steps][128 cols][8], lo swizzled images [KA][16 KiB].
static __global__ void pack_weight_tiles_kernel(const float* __restrict__ w, int K, int ncols, int ldw, uint32_t* whi,
                                         uint8_t* wlo, int ntiles, int kslices, int KA, int wlo_words) {
  const long long total = (long long)ntiles * kslices * 128 * KA * 8;      // 8-k chunks
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int jl = (int)(i & 127);                   // column fastest: coalesced reads of W rows
    long long t = i >> 7;
    const int ch = (int)(t % (KA * 8)); t /= (KA * 8);
    const int ks = (int)(t % kslices);
    const int tile = (int)(t / kslices);
    const int j = tile * 128 + jl;
    const int kl = ch * 8;                           // k within the slice
    const int k = ks * KA * 64 + kl;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (j < ncols && k + e < K) ? w[(size_t)(k + e) * ldw + j] : 0.0f;
    uint4 h, l;
    umma::split_pack_bf16(v[0], v[1], h.x, l.x);
    umma::split_pack_bf16(v[2], v[3], h.y, l.y);
    umma::split_pack_bf16(v[4], v[5], h.z, l.z);
    umma::split_pack_bf16(v[6], v[7], h.w, l.w);
    const size_t wrec = (size_t)tile * kslices + ks;
    const size_t widx = ((wrec * (size_t)(KA * 4) + (ch >> 1)) * 128 + jl) * 8 + (ch & 1) * 4;
    *reinterpret_cast<uint4*>(whi + widx) = h;
    if (wlo_words) {   // lo half destined for TMEM: same packed-word layout as the hi half
      *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(wlo) + widx) = l;
    } else {
      uint8_t* img = wlo + (wrec * KA + (kl >> 6)) * ATOM_BYTES +
                     (jl >> 3) * 1024 + (jl & 7) * 128 + ((((kl & 63) >> 3) ^ (jl & 7)) << 4);
      *reinterpret_cast<uint4*>(img) = l;
    }
  }
}

struct Plan {
  int K, ncols, KAtot, KA, kslices, ntiles, ngroups, nslot, wlo_tmem;
  size_t whi_bytes, wlo_bytes, act_bytes;
};
inline Plan make_plan(int K, int ncols, long long rows, int nsm) {
  Plan p{};
  p.K = K; p.ncols = ncols;
  p.KAtot = (K + 63) / 64;
  p.KA = p.KAtot < KA_MAX ? p.KAtot : KA_MAX;
  p.kslices = (p.KAtot + p.KA - 1) / p.KA;
  p.KA = (p.KAtot + p.kslices - 1) / p.kslices;      // even out the slices
  p.ntiles = (ncols + 127) / 128;
  const long long nrb = (rows + 127) / 128;
  const int units = p.ntiles * p.kslices;
  long long g = nsm / (units > 0 ? units : 1);
  if (g > nrb) g = nrb;
  if (g < 1) g = 1;
  p.ngroups = (int)g;
  p.wlo_tmem = (p.KA <= 4) ? 1 : 0;                  // hi + lo halves: 2 * KA * 32 <= 256 TMEM columns
  // ring slots: whatever the 227 KiB leave next to the resident weight lo half (3 at KA = 8, 4 at KA = 6)
  p.nslot = p.wlo_tmem ? NSLOT_MAX : std::min(NSLOT_MAX, (B200_SMEM_OPTIN - 1024 - p.KA * ATOM_BYTES) / REC_BYTES);
  p.whi_bytes = (size_t)units * 128 * p.KA * 32 * 4;
  p.wlo_bytes = (size_t)units * p.KA * ATOM_BYTES;
  p.act_bytes = (size_t)nrb * p.KAtot * REC_BYTES;
  return p;
}
inline bool plan_ok(const Plan& p, int nsm) { return p.ntiles * p.kslices <= nsm && p.KA <= KA_MAX; }

inline int smem_bytes(const Plan& p) { return 1024 + (p.wlo_tmem ? 0 : p.KA * ATOM_BYTES) + p.nslot * REC_BYTES; }

static inline cudaError_t launch(const Plan& p, const uint8_t* act, const uint32_t* whi, const uint8_t* wlo, const float* bias,
                          float* out, int ldo, long long slab, long long rows, cudaStream_t stream, int exp = 0,
                          long long* prof = nullptr, bool pdl = false) {
  Args a{};
  a.prof = prof;
  a.act = act; a.whi = whi; a.wlo = wlo; a.bias = bias; a.out = out; a.rows = rows; a.slab = slab; a.ldo = ldo;
  a.ncols = p.ncols; a.ntiles = p.ntiles; a.kslices = p.kslices; a.ngroups = p.ngroups; a.KAtot = p.KAtot; a.KA = p.KA;
  a.nslot = p.nslot; a.wlo_tmem = p.wlo_tmem; a.exp = exp;
  const int smem = smem_bytes(p);
  static int configured[MAX_DEVICES] = {0};     // cudaFuncSetAttribute is per device
  {
    std::lock_guard<std::mutex> lk(config_mutex());
    const int dev = current_device_slot();
    if (configured[dev] < smem) {
      cudaError_t e = cudaFuncSetAttribute(gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
      configured[dev] = smem;
    }
  }
  return launch_chain(gemm_ws_kernel, (unsigned)(p.ntiles * p.kslices * p.ngroups), THREADS, (size_t)smem, stream, pdl, a);
}

}  // namespace gemmws
}  // namespace ntm_b200
