// ntm_b200_seq_kernel.cuh -- device code of the persistent NTM sequence kernel.
//
// Included by ntm_b200_k512.cu and ntm_b200_k256.cu, which define
//   NTM_NT        threads per CTA            (512 / 256)
//   NTM_MIN_CTAS  co-resident CTAs per SM    (1 / 2)
//   NTM_KNS       namespace of this build    (k512 / k256)
// What it replaces (paths relative to the reference root):
//   NTMCell.__call__ .......... ntm_cell.py:53-253
//   batched_smooth_cosine_similarity / batched_circular_convolution ... ops.py:135-242
//   LoopNTMTracker.__call__ ... ntm_tracker_new.py:13-64
// Design notes: DESIGN.md s3-4.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "ntm_b200_params.h"
#include "ntm_b200_umma.cuh"

namespace cg = cooperative_groups;

namespace ntm_b200 {
namespace NTM_KNS {

constexpr int NT = NTM_NT;       // threads per CTA
constexpr int NWARP = NT / 32;

// ------------------------------------------------------------------ helpers --
// Activations on the SFU: exp via one ex2 (2 ulp) and an IEEE reciprocal instead of libm's expf /
// tanhf / division sequences.  Absolute error <= ~1.2e-7 (the 1 - 2/(1+e^2x) form cancels for small
// |x|, which costs relative, not absolute, accuracy) -- three orders below the 1e-4 parity budget.
__device__ __forceinline__ float exp_f(float x) { return exp2f(x * 1.4426950408889634f); }
__device__ __forceinline__ float sigmoid_f(float x) { return __frcp_rn(1.0f + exp_f(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - 2.0f * __frcp_rn(1.0f + exp_f(2.0f * x)); }
// softplus(x) = max(x, 0) + log(1 + e^-|x|): never overflows; one ex2 + one lg2 (abs error ~1e-7)
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.0f) + __logf(1.0f + exp_f(-fabsf(x))); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// v + sum over K-slices ks (ascending: fixed summation order) of p[ks * stride], with the loads of
// each group of four issued before any of them is consumed (L2 latency overlapped).
__device__ __forceinline__ float sum_slabs(const float* p, size_t stride, int KS, float v) {
  constexpr int INFL = 24;                  // L2 loads in flight per round (one round for KS <= 24)
  for (int ks = 0; ks < KS; ks += INFL) {
    float a[INFL];
#pragma unroll
    for (int i = 0; i < INFL; ++i) a[i] = (ks + i < KS) ? __ldcg(p + (size_t)(ks + i) * stride) : 0.0f;
#pragma unroll
    for (int i = 0; i < INFL; ++i)
      if (ks + i < KS) v += a[i];
  }
  return v;
}
__device__ __forceinline__ float4 sum_slabs4(const float4* p, size_t stride4, int KS, float4 v) {
  int ks = 0;
  for (; ks + 4 <= KS; ks += 4) {
    const float4 a = __ldcg(p + (size_t)ks * stride4), b = __ldcg(p + (size_t)(ks + 1) * stride4);
    const float4 c = __ldcg(p + (size_t)(ks + 2) * stride4), d = __ldcg(p + (size_t)(ks + 3) * stride4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
    v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
  }
  for (; ks < KS; ++ks) {
    const float4 a = __ldcg(p + (size_t)ks * stride4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  return v;
}

// Phase-cycle accounting for the bench harness (thread 0 of each CTA; null = off).
__device__ __forceinline__ void mark_slot(long long* row, long long& tmark, int slot) {
  if (row != nullptr && threadIdx.x == 0) {
    const long long now = clock64();
    row[slot] += now - tmark;
    tmark = now;
  }
}

// Device-wide barrier over all CTAs of the (co-resident) grid.  Monotonic
// counter, zeroed by the host before launch.  A bounded spin turns a lost CTA
// into an error flag instead of a hung GPU.
__device__ __forceinline__ void grid_sync(unsigned* ctr, int* err, unsigned& epoch, unsigned nblk) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    // release-add / acquire-poll: the release is cumulative over the CTA's writes ordered
    // before it by the bar.sync above, so no separate (much slower) MEMBAR.SC is needed.
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    const unsigned target = epoch * nblk;
    long long t0 = clock64();
    unsigned spins = 0;
    while (ld_acquire_u32(ctr) < target) {
      if (((++spins) & 0x3ffu) == 0) {
        if (*reinterpret_cast<volatile int*>(err) != 0) break;
        if (clock64() - t0 > 4000000000ll) {   // ~2 s at 1.9 GHz
          atomicExch(err, 1);
          break;
        }
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------ phases A / C: skinny GEMM --
// part[ks][b][j] = sum_{k in slice ks} act[b][k] * Wt[k][j]   for all resident b.
// CTA unit = (K-slice, group of JW 64-column tiles); the activation slice is
// staged once in shared memory ([Gpad][KW], read back as warp-broadcast float4
// along k); each warp owns 64 columns (two per lane, coalesced float2 weight
// reads straight from L2, each weight read once per CTA-unit row tile) x TB
// sequences (register accumulators).  Summation order is k-ascending within a
// slice and slice-ascending in the consumer, i.e. fixed: results are
// bit-reproducible run to run.
template <int TB>
__device__ __forceinline__ void gemm_warp_tile(const float* __restrict__ wp, int ldw, const float* sp,
                                               int KW, int kn, bool jok, float* pp, int NCs) {
  float acc0[TB], acc1[TB];
#pragma unroll
  for (int i = 0; i < TB; ++i) { acc0[i] = 0.0f; acc1[i] = 0.0f; }
  float2 w[4], nw[4];
  auto loadw = [&](int kk, float2* d) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d[q] = (jok && kk + q < kn) ? __ldg(reinterpret_cast<const float2*>(wp + (size_t)(kk + q) * ldw))
                                  : make_float2(0.0f, 0.0f);
  };
  loadw(0, w);
  for (int kk = 0; kk < KW; kk += 4) {
    if (kk + 4 < KW) loadw(kk + 4, nw);
#pragma unroll
    for (int i = 0; i < TB; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(sp + i * KW + kk);
      acc0[i] = fmaf(a.x, w[0].x, acc0[i]); acc1[i] = fmaf(a.x, w[0].y, acc1[i]);
      acc0[i] = fmaf(a.y, w[1].x, acc0[i]); acc1[i] = fmaf(a.y, w[1].y, acc1[i]);
      acc0[i] = fmaf(a.z, w[2].x, acc0[i]); acc1[i] = fmaf(a.z, w[2].y, acc1[i]);
      acc0[i] = fmaf(a.w, w[3].x, acc0[i]); acc1[i] = fmaf(a.w, w[3].y, acc1[i]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = nw[q];
  }
  if (jok) {
#pragma unroll
    for (int i = 0; i < TB; ++i)
      *reinterpret_cast<float2*>(pp + (size_t)i * NCs) = make_float2(acc0[i], acc1[i]);
  }
}

__device__ __noinline__ void gemm_phase(const GemmPlan g, const float* act,
                                        const float* __restrict__ Wt, float* part, int Gcur,
                                        float* stage, int cta, int ncta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwork = g.JW * g.NBT;
  for (int u = cta; u < g.units; u += ncta) {
    const int ks = u / g.njg, jg = u - ks * g.njg;
    const int k0 = ks * g.KW;
    const int kn = min(g.KW, g.K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < g.Gpad * g.KW; i += NT) {
      const int b = i / g.KW, kk = i - b * g.KW;
      float v = 0.0f;
      if (b < Gcur && kk < kn) v = __ldcg(act + (size_t)b * g.lda + k0 + kk);
      stage[i] = v;
    }
    __syncthreads();
    const int jw = warp % g.JW, bt = warp / g.JW;
    const int jbase = (jg * g.JW + jw) * 64;
    if (warp < nwork && jbase < g.NC) {
      const int j = jbase + 2 * lane;
      const bool jok = j < g.NC;          // NC and the row strides are even: a column pair is in or out together
      const float* wp = Wt + (size_t)k0 * g.ldw + (jok ? j : 0);
      const float* sp = stage + bt * g.TB * g.KW;
      float* pp = part + ((size_t)ks * g.Gpad + (size_t)bt * g.TB) * g.NCs + (jok ? j : 0);
      switch (g.TB) {
        case 4: gemm_warp_tile<4>(wp, g.ldw, sp, g.KW, kn, jok, pp, g.NCs); break;
        case 8: gemm_warp_tile<8>(wp, g.ldw, sp, g.KW, kn, jok, pp, g.NCs); break;
        case 12: gemm_warp_tile<12>(wp, g.ldw, sp, g.KW, kn, jok, pp, g.NCs); break;
        default: gemm_warp_tile<16>(wp, g.ldw, sp, g.KW, kn, jok, pp, g.NCs); break;
      }
    }
  }
}

// ------------------------------------- phases A / C on the tensor cores (tcgen05) --
// Same contract as gemm_phase (K-slice partial slabs part[ks][b][j]), but each CTA owns ONE
// unit = (128-column tile, K-slice) whose weights stay RESIDENT IN TENSOR MEMORY for the whole
// kernel as a bf16 "hi" + bf16 "lo" pair (a = hi + lo to ~2^-18): loaded once by
// tc_load_weights, used as the A operand of tcgen05.mma (A from TMEM).  Per timestep only the
// activations move: fp32 [b][k] from L2 -> split into bf16 hi/lo -> K-major SWIZZLE_128B tiles
// in shared memory (B operand, N = sequences); D[128 cols][N] += Whi*Bhi + Whi*Blo + Wlo*Bhi
// accumulates in TMEM (fp32), then goes straight to the partial slab.
__device__ __forceinline__ void tc_load_weights(const GemmPlan& g, const float* __restrict__ Wt, uint32_t tmem,
                                                int cta) {
  using namespace ntm_b200::umma;
  if (!g.tc || cta >= g.units) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = cta / g.njg, tile = cta - ks * g.njg;
  const int k0 = ks * g.KW;
  const int j = tile * 128 + 32 * (warp & 3) + lane;      // weight column = TMEM lane
  const bool jok = j < g.NC;
  const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
  const int kq = g.KW / 16;                               // 16-k groups in the slice
  for (int q = warp >> 2; q < kq; q += NWARP / 4) {       // the 4 warps sharing a lane quarter split k
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + q * 16 + 2 * e;
      const float v0 = (jok && k < g.K) ? __ldg(Wt + (size_t)k * g.ldw + j) : 0.0f;
      const float v1 = (jok && k + 1 < g.K) ? __ldg(Wt + (size_t)(k + 1) * g.ldw + j) : 0.0f;
      split_pack_bf16(v0, v1, hi[e], lo[e]);
    }
    tmem_st_x8(tmem + lane_addr + g.tcol + q * 8, hi);
    tmem_st_x8(tmem + lane_addr + g.tcol + g.KW / 2 + q * 8, lo);
  }
  tmem_wait_st();
}

__device__ __noinline__ void gemm_phase_tc(const GemmPlan g, const float* act, float* part, int Gcur,
                                           uint8_t* stage, uint32_t tmem, uint64_t* mbar, uint32_t& mbar_uses,
                                           int cta) {
  using namespace ntm_b200::umma;
  if (cta >= g.units) return;                              // CTA-uniform
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ks = cta / g.njg, tile = cta - ks * g.njg;
  const int k0 = ks * g.KW;
  const int N = g.Gpad;                                    // MMA N (multiple of 16)
  const int katoms = (g.KW + 63) >> 6;
  uint8_t* sBhi = stage;
  uint8_t* sBlo = stage + (size_t)katoms * N * 128;
  // ---- stage activations: one 16-byte chunk (8 consecutive k) per thread-iteration ----
  const int cpr = katoms * 8;                              // chunks per row
  const bool vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(act) & 15) == 0);
  for (int i = tid; i < N * cpr; i += NT) {
    const int b = i / cpr, c = i - b * cpr;
    const int kk = c * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.0f;
    if (b < Gcur && kk < g.KW) {
      const float* src = act + (size_t)b * g.lda + k0 + kk;
      if (vec && k0 + kk + 8 <= g.K) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(src));
        const float4 c4 = __ldcg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (k0 + kk + e < g.K) v[e] = __ldcg(src + e);
      }
    }
    uint4 h, l;
    split_pack_bf16(v[0], v[1], h.x, l.x);
    split_pack_bf16(v[2], v[3], h.y, l.y);
    split_pack_bf16(v[4], v[5], h.z, l.z);
    split_pack_bf16(v[6], v[7], h.w, l.w);
    const uint32_t off = sw128_offset(b, kk, N);
    *reinterpret_cast<uint4*>(sBhi + off) = h;
    *reinterpret_cast<uint4*>(sBlo + off) = l;
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // ---- MMA issue: one elected thread ----
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(128, N);
      uint32_t accum = 0;
      for (int kk = 0; kk < g.KW; kk += 16) {
        const int atom = kk >> 6, kin = kk & 63;
        const uint64_t dhi = make_sw128_desc(sBhi + (size_t)atom * N * 128 + kin * 2);
        const uint64_t dlo = make_sw128_desc(sBlo + (size_t)atom * N * 128 + kin * 2);
        const uint32_t ahi = tmem + g.tcol + kk / 2, alo = ahi + g.KW / 2;
        mma_ts(tmem, ahi, dhi, idesc, accum);
        accum = 1;
        mma_ts(tmem, ahi, dlo, idesc, accum);
        mma_ts(tmem, alo, dhi, idesc, accum);
      }
      mma_commit(mbar);
    }
    __syncwarp();
  }
  mbar_wait(mbar, mbar_uses & 1u);
  mbar_uses += 1;
  tcgen05_fence_after();
  // ---- epilogue: accumulator rows (weight columns) -> partial slab, coalesced over j ----
  {
    const int j = tile * 128 + 32 * (warp & 3) + lane;
    const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    const int nq = N / (NWARP / 4);                        // accumulator columns (sequences) per warp group
    const int bq0 = (warp >> 2) * nq;                      // (N % 16 == 0, so nq is a multiple of 4)
    float* pp = part + ((size_t)ks * g.Gpad) * g.NCs + j;
    for (int c = 0; c < nq; c += 4) {
      uint32_t v[4];
      tmem_ld_x4(tmem + lane_addr + bq0 + c, v);
      tmem_wait_ld();
      if (j < g.NC) {
#pragma unroll
        for (int e = 0; e < 4; ++e) pp[(size_t)(bq0 + c + e) * g.NCs] = __uint_as_float(v[e]);
      }
    }
  }
  tcgen05_fence_before();   // order the TMEM reads before the next phase's MMAs (after the grid barrier)
}

// ------------------------------------------------------- phase B: LSTM gates --
// BasicLSTMCell (TF 1.0/1.1): i, j, f, o = split4(z); c' = c*sig(f + 0) + sig(i)*tanh(j);
// h' = tanh(c')*sig(o).  z = hoisted x-projection (layer 0, bias folded in) or
// bias (layers > 0) plus the K-slice partials of phase A in slice order.
__device__ __forceinline__ void lstm_phase(const KParams& p, int l, int Gcur, int b0, int t,
                                           int cta, int ncta, float* const* act, float* cst,
                                           const float* partA) {
  const GemmPlan& g = p.gA[l];
  const int C = p.C;
  // one lane per (sequence, unit, gate); the four gates of a unit sit in adjacent lanes
  const int total = Gcur * C * 4;
  const int chunk = ((total + ncta - 1) / ncta + 3) & ~3;
  const int lo = cta * chunk, hi = min(total, lo + chunk);
  const int KS = g.KS;
  for (int base = lo; base < hi; base += NT) {
    const int i = base + (int)threadIdx.x;
    const bool ok = i < hi;
    const int ii = ok ? i : lo;
    const int q = ii & 3, bu = ii >> 2;
    const int b = bu / C, u = bu - b * C;
    const int col = q * C + u;
    float v;
    if (l == 0) {
      const float* xp = p.xw + ((size_t)(b0 + b) * p.T + t) * (size_t)(4 * C) + col;
      v = __ldg(xp);
      // the hoisted projection streams from HBM once: pull the next step's row into L2 now so that its
      // ~1 us DRAM latency is off the next step's critical path
      if (t + 1 < p.T && (col & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + 4 * C));
    } else {
      v = __ldg(p.bA[l] + col);
    }
    v = sum_slabs(partA + (size_t)b * g.NCs + col, (size_t)g.Gpad * g.NCs, KS, v);
    const unsigned lane = threadIdx.x & 31u, gl = lane & ~3u;
    const float zi = __shfl_sync(0xffffffffu, v, gl + 0);
    const float zj = __shfl_sync(0xffffffffu, v, gl + 1);
    const float zf = __shfl_sync(0xffffffffu, v, gl + 2);
    const float zo = __shfl_sync(0xffffffffu, v, gl + 3);
    if (ok && p.hZ != nullptr)   // pre-activation of gate q (training history)
      p.hZ[((((size_t)t * p.B + (b0 + b)) * p.L + l) * 4 + q) * C + u] = v;
    if (ok && q == 0) {
      float* cp = cst + ((size_t)b * p.L + l) * C + u;
      const float c_prev = __ldcg(cp);
      const float c_new = c_prev * sigmoid_f(zf) + sigmoid_f(zi) * tanh_f(zj);
      const float h_new = tanh_f(c_new) * sigmoid_f(zo);
      *cp = c_new;
      if (p.hC != nullptr) p.hC[(((size_t)(t + 1) * p.B + (b0 + b)) * p.L + l) * C + u] = c_new;
      if (p.hH != nullptr) p.hH[(((size_t)(t + 1) * p.B + (b0 + b)) * p.L + l) * C + u] = h_new;
      act[l][(size_t)b * p.actK[l] + (p.actK[l] - C) + u] = h_new;
      if (l + 1 < p.L) act[l + 1][(size_t)b * p.actK[l + 1] + u] = h_new;
    }
  }
}

// -------------------------------------------------- phase D building blocks --
// Partial column sums of squares over this CTA's rows -> xch[0..M4) (used once
// per wave for the initial memory; afterwards pass 2 produces them).
__device__ __forceinline__ void colsq_local(const KParams& p, const float* Ms, int nrows,
                                            float* out) {
  for (int d = threadIdx.x; d < p.M4; d += NT) {
    float s = 0.0f;
    for (int r = 0; r < nrows; ++r) {
      const float v = Ms[r * p.M4 + d];
      s = fmaf(v, v, s);
    }
    out[d] = s;
  }
}

// cn[d] = 1/sqrt(max(sum over the whole cluster of column squares, 1e-12))
// (tf.nn.l2_normalize along N of the transposed memory, ops.py:147-150).
__device__ __forceinline__ void finalize_colnorm(const KParams& p, cg::cluster_group& cluster,
                                                 float* smem, int oXcsq, float* cn) {
  for (int d = threadIdx.x; d < p.M4; d += NT) {
    float s = 0.0f;
    for (int r = 0; r < p.CS; ++r) {
      const float* rem = cluster.map_shared_rank(smem + oXcsq, r);
      s += rem[d];
    }
    cn[d] = 1.0f / sqrtf(fmaxf(s, 1e-12f));
  }
}

template <int R, int W>
__device__ __forceinline__ void phase_d(const KParams& p, cg::cluster_group& cluster, float* smem,
                                        int crank, int gslot, int bglob, int t, int row0, int nrows,
                                        int& wcur, long long* prow, long long& tmark, float* act0,
                                        const float* partC) {
  constexpr int H = R + W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = p.M, M4 = p.M4, MC = p.MC, N = p.N, Npad = p.Npad, S = p.S;
  float* Ms = smem + p.oMs;
  float* wprev = smem + (wcur ? p.oW1 : p.oW0);
  float* wnew = smem + (wcur ? p.oW0 : p.oW1);
  float* cn = smem + p.oCn;
  // [R][M4] read partials, then [M4] column squares.  Aliases the key buffer kS (dead after
  // pass 1; (R+1) <= H rows).  Single-buffered: a peer only reads it in this step's finalize, and
  // three team-wide barriers separate that from the next overwrite.
  float* xch = smem + p.oK;
  float* simA = smem + p.oSim;                   // [H][Npad] private full-length similarities / weights
  float* simL = smem + p.oSl;                    // [H][NR] this CTA's rows, read by the peers
  float* wg = smem + p.oWg;                      // [H][Npad]
  float* kS = smem + p.oK;                       // [H][M4]
  float* eS = smem + p.oE;                       // [W][M4]
  float* aS = smem + p.oA;                       // [W][M4]
  float* sm = smem + p.oSm;                      // beta[H] g[H] gamma[H] rs[H] sw[H][SMAX]
  float* sBeta = sm, *sG = sm + H, *sGam = sm + 2 * H, *sSw = sm + 4 * H;
  float* sPart = sm + 4 * H + H * SMAX;          // [NWARP][H] per-warp partial key norms
  float* sRed = sPart + NWARP * H;               // [H][3][NWARP / H] addressing reductions
  const bool last = (t == p.T - 1);
  float* dbg = (p.dbg != nullptr && last && crank == 0) ? p.dbg + (size_t)bglob * p.dbgStride : nullptr;

  // ---- D0: split-K reduction of phase C + bias, activations (ntm_cell.py:124-196) ----
  const int offBeta = H * M, offG = offBeta + H, offS = offG + H, offGam = offS + S * H,
            offE = offGam + H, offA = offE + M * W;
  // pass a: raw[q] = bias[q] + sum_ks partC[ks][slot][q], float4-vectorised.  `raw` lives in the
  // wg scratch (never written by a peer CTA; simA is, by the pass-1 all-gather).
  float* raw = wg;
  {
    const float4* pc4 = reinterpret_cast<const float4*>(partC + (size_t)gslot * p.gC.NCs);
    const float4* b4 = reinterpret_cast<const float4*>(p.bC);
    const size_t slab4 = (size_t)p.gC.Gpad * p.gC.NCs / 4;
    float4* hp4 = (p.hP != nullptr && crank == 0)
                      ? reinterpret_cast<float4*>(p.hP + ((size_t)t * p.B + bglob) * p.PO4) : nullptr;
    for (int q4 = tid; q4 < p.PO4 / 4; q4 += NT) {
      const float4 v4 = sum_slabs4(pc4 + q4, slab4, p.gC.KS, __ldg(b4 + q4));
      reinterpret_cast<float4*>(raw)[q4] = v4;
      if (hp4) hp4[q4] = v4;
    }
  }
  if (p.hW != nullptr && crank == 0) {   // weightings entering this step
    float* hw = p.hW + ((size_t)t * p.B + bglob) * H * N;
    for (int i = tid; i < H * N; i += NT) {
      const int h = i / N, n = i - h * N;
      hw[i] = wprev[h * Npad + n];
    }
  }
  if (p.hCn != nullptr && crank == 0) {  // inverse column norms of the memory entering this step
    float* hc = p.hCn + ((size_t)t * p.B + bglob) * M;
    for (int d = tid; d < M; d += NT) hc[d] = cn[d];
  }
  if (p.hM != nullptr) {                 // memory entering this step (this CTA's rows)
    float* hm = p.hM + (((size_t)t * p.B + bglob) * N + row0) * M;
    for (int i = tid; i < nrows * M; i += NT) {
      const int r = i / M, d = i - r * M;
      hm[i] = Ms[r * M4 + d];
    }
  }
  __syncthreads();
  mark_slot(prow, tmark, 15);
  // pass b: activations.  Keys: kS[h][d] = tanh(raw) * cn[d]  (the key's own 1/|k| is a per-head
  // scalar and is applied to the similarities later); per-head sum of squares via fixed-order
  // warp partials.  Pad lanes d >= M are written as zeros.
  {
    // loads first, then the H independent activation chains, then the stores: the compiler cannot
    // hoist shared-memory loads over stores that might alias, so the order is spelled out here
    float ss[H];
#pragma unroll
    for (int h = 0; h < H; ++h) ss[h] = 0.0f;
    for (int d = tid; d < M4; d += NT) {
      float rv[H], re[W], ra[W];
      const bool in = d < M;
#pragma unroll
      for (int h = 0; h < H; ++h) rv[h] = in ? raw[h * M + d] : 0.0f;
#pragma unroll
      for (int h = 0; h < W; ++h) {
        re[h] = in ? raw[offE + h * M + d] : 0.0f;
        ra[h] = in ? raw[offA + h * M + d] : 0.0f;
      }
      const float cnd = cn[d];
#pragma unroll
      for (int h = 0; h < H; ++h) rv[h] = in ? tanh_f(rv[h]) : 0.0f;
#pragma unroll
      for (int h = 0; h < W; ++h) {
        re[h] = in ? sigmoid_f(re[h]) : 0.0f;
        ra[h] = in ? tanh_f(ra[h]) : 0.0f;
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        kS[h * M4 + d] = rv[h] * cnd;
        ss[h] = fmaf(rv[h], rv[h], ss[h]);
        if (dbg && in) dbg[h * M + d] = rv[h];
      }
#pragma unroll
      for (int h = 0; h < W; ++h) {
        eS[h * M4 + d] = re[h];
        aS[h * M4 + d] = ra[h];
        if (dbg && in) { dbg[offE + h * M + d] = re[h]; dbg[offA + h * M + d] = ra[h]; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {      // H interleaved warp reductions
#pragma unroll
      for (int h = 0; h < H; ++h) ss[h] += __shfl_xor_sync(0xffffffffu, ss[h], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) sPart[warp * H + h] = ss[h];
    }
  }
  if (tid < H) {   // per-head scalars: beta, g, gamma (ntm_cell.py:140,151,169), shift softmax (:161)
    const float bv = softplus_f(raw[offBeta + tid]);
    const float gv = sigmoid_f(raw[offG + tid]);
    const float gm = 1.0f + softplus_f(raw[offGam + tid]);
    sBeta[tid] = bv; sG[tid] = gv; sGam[tid] = gm;
    float* sp = sSw + tid * SMAX;
    float mx = raw[offS + tid * S];
    for (int i = 1; i < S; ++i) mx = fmaxf(mx, raw[offS + tid * S + i]);
    float sum = 0.0f;
    for (int i = 0; i < S; ++i) { sp[i] = exp_f(raw[offS + tid * S + i] - mx); sum += sp[i]; }
    const float rsum = __frcp_rn(sum);
    for (int i = 0; i < S; ++i) sp[i] = sp[i] * rsum;
    if (dbg) {
      dbg[offBeta + tid] = bv; dbg[offG + tid] = gv; dbg[offGam + tid] = gm;
      for (int i = 0; i < S; ++i) dbg[offS + tid * S + i] = sp[i];
    }
  }
  if (crank == 0 && tid == NT - 1) {   // output projection + softmax (ntm_cell.py:220-221)
    const size_t o = ((size_t)bglob * p.T + t) * p.O;
    const float* lg = raw + p.P;
    float mx = lg[0];
    for (int i = 1; i < p.O; ++i) mx = fmaxf(mx, lg[i]);
    float sum = 0.0f;
    for (int i = 0; i < p.O; ++i) sum += exp_f(lg[i] - mx);
    const float rsum = __frcp_rn(sum);
    for (int i = 0; i < p.O; ++i) {
      p.logits[o + i] = lg[i];
      if (p.outputs) p.outputs[o + i] = exp_f(lg[i] - mx) * rsum;
    }
  }
  __syncthreads();
  mark_slot(prow, tmark, 10);

  // ---- pass 1: sim[h][n] = sum_d kc[h][d] * M[n][d] over this CTA's rows (ops.py:156) ----
  // A warp takes a block of RB1 rows and walks the columns in float4 chunks (one per lane): the
  // H key chunks it loads are reused by all RB1 rows, so the shared-memory traffic of the pass is
  // (1 + H / RB1) x the memory itself.  The pass is shared-memory-bandwidth bound.
  {
    constexpr int RB1 = RB;            // 8-row blocks halve the key traffic but leave half the warps idle: measured slower
    int LPR = 32;                      // lanes cooperating on one block of rows
    while (LPR > 1 && (LPR >> 1) >= MC) LPR >>= 1;
    const int GPW = 32 / LPR;
    const int sg = lane / LPR, lg = lane - sg * LPR;
    const int nRB = (nrows + RB1 - 1) / RB1;
    const int iters = (nRB + NWARP * GPW - 1) / (NWARP * GPW);
    for (int it = 0; it < iters; ++it) {
      const int rb = (it * NWARP + warp) * GPW + sg;
      const bool active = rb < nRB;
      float acc[RB1][H];
#pragma unroll
      for (int i = 0; i < RB1; ++i)
#pragma unroll
        for (int h = 0; h < H; ++h) acc[i][h] = 0.0f;
      int rows[RB1];
#pragma unroll
      for (int i = 0; i < RB1; ++i) rows[i] = min(rb * RB1 + i, nrows - 1);
      if (active) {
        for (int c = lg; c < MC; c += LPR) {
          float4 k4[H];
#pragma unroll
          for (int h = 0; h < H; ++h) k4[h] = *reinterpret_cast<const float4*>(kS + h * M4 + 4 * c);
#pragma unroll
          for (int i = 0; i < RB1; ++i) {
            const float4 m4 = *reinterpret_cast<const float4*>(Ms + rows[i] * M4 + 4 * c);
#pragma unroll
            for (int h = 0; h < H; ++h) {
              acc[i][h] = fmaf(m4.x, k4[h].x, acc[i][h]);
              acc[i][h] = fmaf(m4.y, k4[h].y, acc[i][h]);
              acc[i][h] = fmaf(m4.z, k4[h].z, acc[i][h]);
              acc[i][h] = fmaf(m4.w, k4[h].w, acc[i][h]);
            }
          }
        }
      }
      if (LPR == 32 && RB * H <= 32) {
        // transposing reduction, RB rows at a time: at offset o a lane keeps one half of its value
        // list and receives the partner's sums of that half -- 31 shuffles instead of 5 per value;
        // lane L ends up with the warp total of value L = i * H + h (fixed order, deterministic).
#pragma unroll
        for (int g = 0; g < RB1 / RB; ++g) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.0f;
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int h = 0; h < H; ++h) v[i * H + h] = acc[g * RB + i][h];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < o; ++j) {
              const float send = up ? v[j] : v[j + o];
              const float keep = up ? v[j + o] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          const int vi = lane / H, vh = lane - vi * H;      // value index -> (row in group, head)
          const int rl = rb * RB1 + g * RB + vi;
          if (active && lane < RB * H && rl < nrows) simL[vh * p.NR + rl] = v[0];
        }
      } else {
        for (int o = LPR >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int i = 0; i < RB1; ++i)
#pragma unroll
            for (int h = 0; h < H; ++h) acc[i][h] += __shfl_xor_sync(0xffffffffu, acc[i][h], o);
        }
        if (active && lg == 0) {
#pragma unroll
          for (int i = 0; i < RB1; ++i) {
            const int rl = rb * RB1 + i;
            if (rl < nrows) {
#pragma unroll
              for (int h = 0; h < H; ++h) simL[h * p.NR + rl] = acc[i][h];
            }
          }
        }
      }
    }
  }
  mark_slot(prow, tmark, 6);
  cluster.sync();
  // all-gather over DSMEM: every CTA pulls all slices (its own included) from the owners' simL
  // buffers into its private full-length copy; simL is not touched again before the next step's
  // pass 1, which three team-wide barriers separate from these reads.
  for (int i = tid; i < H * N; i += NT) {
    const int h = i / N, n = i - h * N;
    const int q = n / p.NR;
    const float* rem = cluster.map_shared_rank(simL, q);
    const float sv = rem[h * p.NR + (n - q * p.NR)];
    simA[h * Npad + n] = sv;
    if (p.hSim != nullptr && crank == 0) p.hSim[((size_t)t * p.B + bglob) * H * N + i] = sv;   // un-normalised similarities
  }
  __syncthreads();
  mark_slot(prow, tmark, 11);

  // ---- addressing on the full [H][N] weighting, replicated in every CTA (ntm_cell.py:140-176) ----
  // WPH warps cooperate on one head (elements strided over them); the three N-reductions go
  // through shared memory in fixed (warp-ascending) order behind a named barrier per head.
  if (N == 128 && S <= 7 && p.dbg == nullptr && H <= NWARP && (p.dsw & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0) {
    // N = 128 (the tracker shapes): one warp per head keeps the head's 128 entries in registers -- lane L holds
    // n = 4L .. 4L+3 -- from the similarity to the final weighting; the circular shift takes the neighbour lanes'
    // entries by shuffle.  No shared-memory round trips, no named barriers between the sweeps.
    if (warp < H) {
      const int h = warp;
      const float gate = sG[h], gamma = sGam[h];
      float kn = 0.0f;                                  // |k_h|^2, fixed summation order
      for (int w2 = 0; w2 < NWARP; ++w2) kn += sPart[w2 * H + h];
      const float scale = sBeta[h] * (1.0f / sqrtf(fmaxf(kn, 1e-12f)));   // beta / |k|  (ops.py:152, ntm_cell.py:142)
      const float4 s4 = *reinterpret_cast<const float4*>(simA + h * Npad + 4 * lane);
      const float4 p4 = *reinterpret_cast<const float4*>(wprev + h * Npad + 4 * lane);
      const float x[4] = {s4.x * scale, s4.y * scale, s4.z * scale, s4.w * scale};
      const float mx = warp_max(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
      float e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) e[u] = exp_f(x[u] - mx);
      const float sum = warp_sum((e[0] + e[1]) + (e[2] + e[3]));
      const float gs = gate / sum, g1 = 1.0f - gate;    // w_g = g * softmax + (1 - g) * w_prev
      float win[12];                                     // gated weights of lanes L-1, L, L+1 (circular)
      win[4] = fmaf(e[0], gs, p4.x * g1); win[5] = fmaf(e[1], gs, p4.y * g1);
      win[6] = fmaf(e[2], gs, p4.z * g1); win[7] = fmaf(e[3], gs, p4.w * g1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        win[u] = __shfl_sync(0xffffffffu, win[4 + u], (lane + 31) & 31);
        win[8 + u] = __shfl_sync(0xffffffffu, win[4 + u], (lane + 1) & 31);
      }
      float pw[4];
      auto shift_pow = [&](auto s_tag) {     // circular_shift(x, j)[n] = x[(n + j) mod N], taps j = shift0 .. shift0+S-1
        constexpr int SS = decltype(s_tag)::value, SH0 = -((SS + 1) / 2);
        float swv[SS];
#pragma unroll
        for (int s2 = 0; s2 < SS; ++s2) swv[s2] = sSw[h * SMAX + s2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float conv = 0.0f;
#pragma unroll
          for (int s2 = 0; s2 < SS; ++s2) conv = fmaf(swv[s2], win[4 + u + SH0 + s2], conv);
          pw[u] = exp2f(gamma * log2f(conv));   // conv >= 0, gamma >= 1: == pow(conv, gamma), 0 -> 0
        }
      };
      if (S == 3) shift_pow(std::integral_constant<int, 3>{});
      else if (S == 1) shift_pow(std::integral_constant<int, 1>{});
      else if (S == 5) shift_pow(std::integral_constant<int, 5>{});
      else shift_pow(std::integral_constant<int, 7>{});
      const float psum = warp_sum((pw[0] + pw[1]) + (pw[2] + pw[3]));
      const float rden = 1.0f / (psum + 1e-3f);   // ntm_cell.py:175-176
      const float4 wv = make_float4(pw[0] * rden, pw[1] * rden, pw[2] * rden, pw[3] * rden);
      *reinterpret_cast<float4*>(wnew + h * Npad + 4 * lane) = wv;
      if (last && crank == 0) *reinterpret_cast<float4*>(p.dw + (size_t)bglob * p.dsw + h * N + 4 * lane) = wv;
    }
  } else {
    constexpr int WPH = (NWARP / H) > 0 ? (NWARP / H) : 1;
    const int hgrp = warp / WPH, sub = warp - hgrp * WPH;
    const bool multi = (NWARP / H) > 0;                       // else: one warp walks the heads
    const int hstep = multi ? H : NWARP;
    for (int h = multi ? hgrp : warp; h < H; h += hstep) {
      float* sh = simA + h * Npad;
      float* gh = wg + h * Npad;
      float* red = sRed + h * 3 * WPH;                        // [3][WPH] partial max / sum / psum
      const int nstep = 32 * (multi ? WPH : 1);
      const int n0 = 32 * (multi ? sub : 0) + lane;
      const int nthr = 32 * WPH;
      auto head_bar = [&]() {
        if (multi && WPH > 1) asm volatile("bar.sync %0, %1;" ::"r"(h + 1), "r"(nthr) : "memory");
        else __syncwarp();
      };
      const float gate = sG[h], gamma = sGam[h];
      float kn = 0.0f;                                  // |k_h|^2, fixed summation order
      for (int w2 = 0; w2 < NWARP; ++w2) kn += sPart[w2 * H + h];
      const float rs = 1.0f / sqrtf(fmaxf(kn, 1e-12f));   // ops.py:152
      const float beta = sBeta[h];
      float mx = -INFINITY;
#pragma unroll 2
      for (int n = n0; n < N; n += nstep) {
        const float sv = sh[n] * rs;                    // similarity (ops.py:156)
        const float x = sv * beta;
        if (dbg) dbg[p.P + (0 * H + h) * N + n] = sv;
        sh[n] = x;
        mx = fmaxf(mx, x);
      }
      mx = warp_max(mx);
      if (multi && WPH > 1) {
        if (lane == 0) red[sub] = mx;
        head_bar();
        mx = red[0];
#pragma unroll
        for (int i = 1; i < WPH; ++i) mx = fmaxf(mx, red[i]);
      }
      float sum = 0.0f;
#pragma unroll 2
      for (int n = n0; n < N; n += nstep) {
        const float e = exp_f(sh[n] - mx);
        sh[n] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      if (multi && WPH > 1) {
        if (lane == 0) red[WPH + sub] = sum;
        head_bar();
        sum = red[WPH];
#pragma unroll
        for (int i = 1; i < WPH; ++i) sum += red[WPH + i];
      }
#pragma unroll 2
      for (int n = n0; n < N; n += nstep) {
        const float wc = sh[n] / sum;
        const float v = wc * gate + wprev[h * Npad + n] * (1.0f - gate);
        gh[n] = v;
        if (dbg) {
          dbg[p.P + (1 * H + h) * N + n] = wc;
          dbg[p.P + (2 * H + h) * N + n] = v;
        }
      }
      head_bar();                                       // the shift reads neighbours' gated weights
      float psum = 0.0f;
#pragma unroll 2
      for (int n = n0; n < N; n += nstep) {
        float conv = 0.0f;
        for (int s = 0; s < S; ++s) {
          int idx = n + p.shift0 + s;          // circular_shift(x, j)[n] = x[(n + j) mod N], ops.py:216-242
          idx = idx < 0 ? idx + N : (idx >= N ? idx - N : idx);
          conv = fmaf(sSw[h * SMAX + s], gh[idx], conv);
        }
        const float pw = exp2f(gamma * log2f(conv));   // conv >= 0, gamma >= 1: == pow(conv, gamma), 0 -> 0
        sh[n] = pw;
        psum += pw;
        if (dbg) {
          dbg[p.P + (3 * H + h) * N + n] = conv;
          dbg[p.P + (4 * H + h) * N + n] = pw;
        }
      }
      psum = warp_sum(psum);
      if (multi && WPH > 1) {
        if (lane == 0) red[2 * WPH + sub] = psum;
        head_bar();
        psum = red[2 * WPH];
#pragma unroll
        for (int i = 1; i < WPH; ++i) psum += red[2 * WPH + i];
      }
      const float den = psum + 1e-3f;          // ntm_cell.py:175-176
#pragma unroll 2
      for (int n = n0; n < N; n += nstep) {
        const float wv = sh[n] / den;
        wnew[h * Npad + n] = wv;
        if (last && crank == 0) p.dw[(size_t)bglob * p.dsw + h * N + n] = wv;
      }
    }
  }
  __syncthreads();
  mark_slot(prow, tmark, 12);

  // ---- pass 2: erase/add write, weighted read, next column norms (ntm_cell.py:193-215) ----
  {
    const int cl = lane & 7, rg = lane >> 3;
    const int ncg = (MC + 7) >> 3;
    for (int cgi = warp; cgi < ncg; cgi += NWARP) {
      const int c = cgi * 8 + cl;
      const bool valid = c < MC;
      const int cc = valid ? c : 0;
      float4 e4[W], a4[W];
#pragma unroll
      for (int h = 0; h < W; ++h) {
        e4[h] = *reinterpret_cast<const float4*>(eS + h * M4 + 4 * cc);
        a4[h] = *reinterpret_cast<const float4*>(aS + h * M4 + 4 * cc);
      }
      float4 racc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) racc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 csq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        for (int row = rg; row < nrows; row += 4) {
          const int n = row0 + row;
          float4* mp = reinterpret_cast<float4*>(Ms + row * M4 + 4 * c);
          const float4 m = *mp;
          float4 mn;
          if constexpr (W == 1) {
            // one write head: M' = M (1 - w e) + w a = M + w (a - M e): two FMAs per element
            const float ww = wnew[R * Npad + n];
            mn.x = fmaf(ww, fmaf(-m.x, e4[0].x, a4[0].x), m.x);
            mn.y = fmaf(ww, fmaf(-m.y, e4[0].y, a4[0].y), m.y);
            mn.z = fmaf(ww, fmaf(-m.z, e4[0].z, a4[0].z), m.z);
            mn.w = fmaf(ww, fmaf(-m.w, e4[0].w, a4[0].w), m.w);
          } else {
            float4 E = make_float4(1.f, 1.f, 1.f, 1.f), A = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int h = 0; h < W; ++h) {
              const float ww = wnew[(R + h) * Npad + n];
              E.x *= (1.0f - ww * e4[h].x); E.y *= (1.0f - ww * e4[h].y);
              E.z *= (1.0f - ww * e4[h].z); E.w *= (1.0f - ww * e4[h].w);
              A.x = fmaf(ww, a4[h].x, A.x); A.y = fmaf(ww, a4[h].y, A.y);
              A.z = fmaf(ww, a4[h].z, A.z); A.w = fmaf(ww, a4[h].w, A.w);
            }
            mn.x = fmaf(m.x, E.x, A.x); mn.y = fmaf(m.y, E.y, A.y);
            mn.z = fmaf(m.z, E.z, A.z); mn.w = fmaf(m.w, E.w, A.w);
          }
          const float4 mu = p.write_first ? mn : m;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float wr = wnew[r * Npad + n];
            racc[r].x = fmaf(wr, mu.x, racc[r].x); racc[r].y = fmaf(wr, mu.y, racc[r].y);
            racc[r].z = fmaf(wr, mu.z, racc[r].z); racc[r].w = fmaf(wr, mu.w, racc[r].w);
          }
          csq.x = fmaf(mn.x, mn.x, csq.x); csq.y = fmaf(mn.y, mn.y, csq.y);
          csq.z = fmaf(mn.z, mn.z, csq.z); csq.w = fmaf(mn.w, mn.w, csq.w);
          *mp = mn;
        }
      }
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          racc[r].x += __shfl_xor_sync(0xffffffffu, racc[r].x, o);
          racc[r].y += __shfl_xor_sync(0xffffffffu, racc[r].y, o);
          racc[r].z += __shfl_xor_sync(0xffffffffu, racc[r].z, o);
          racc[r].w += __shfl_xor_sync(0xffffffffu, racc[r].w, o);
        }
        csq.x += __shfl_xor_sync(0xffffffffu, csq.x, o);
        csq.y += __shfl_xor_sync(0xffffffffu, csq.y, o);
        csq.z += __shfl_xor_sync(0xffffffffu, csq.z, o);
        csq.w += __shfl_xor_sync(0xffffffffu, csq.w, o);
      }
      if (valid && rg == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) *reinterpret_cast<float4*>(xch + r * M4 + 4 * c) = racc[r];
        *reinterpret_cast<float4*>(xch + R * M4 + 4 * c) = csq;
      }
    }
  }
  cluster.sync();
  mark_slot(prow, tmark, 13);

  // ---- cluster reduction over DSMEM: column norms (all CTAs), read vector (split by rank) ----
  const int oX = p.oK;
  finalize_colnorm(p, cluster, smem, oX + R * M4, cn);
  for (int i = crank * NT + tid; i < R * M; i += p.CS * NT) {
    const int r = i / M, d = i - r * M;
    float s = 0.0f;
    for (int q = 0; q < p.CS; ++q) {
      const float* rem = cluster.map_shared_rank(smem + oX, q);
      s += rem[r * M4 + d];
    }
    act0[(size_t)gslot * p.actK[0] + i] = s;              // next step's controller input
    if (last) p.dread[(size_t)bglob * p.dsread + i] = s;
    if (p.hRead != nullptr) p.hRead[((size_t)(t + 1) * p.B + bglob) * (R * M) + i] = s;
  }
  wcur ^= 1;
  mark_slot(prow, tmark, 14);
}

// ---------------------------------------------------------- the persistent kernel --
#ifdef NTM_MAXREG
#define NTM_KERNEL_BOUNDS __maxnreg__(NTM_MAXREG)
#else
#define NTM_KERNEL_BOUNDS __launch_bounds__(NT, NTM_MIN_CTAS)
#endif
template <int R, int W>
__global__ void NTM_KERNEL_BOUNDS ntm_seq_kernel(const KParams p) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int H = R + W;
  const int tid = threadIdx.x;
  const int team = blockIdx.x / p.team_ctas;
  const int cta = blockIdx.x - team * p.team_ctas;   // index within the team
  const int ncta = p.team_ctas;
  const int crank = (int)cluster.block_rank();
  const int gslot = cta / p.CS;                    // cluster index in the team = resident-sequence slot
  const int row0 = crank * p.NR;
  const int nrows = max(0, min(p.NR, p.N - row0));
  unsigned epoch = 0;
  float* Ms = smem + p.oMs;
  float* stage = smem + p.oScr;
  // this team's workspace slices and barrier counter
  unsigned* ctr = p.ctr + 32 * team;
  float* act[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) act[l] = (l < p.L) ? p.act[l] + (size_t)team * p.act_ts[l] : nullptr;
  float* cst = p.cst + (size_t)team * p.cst_ts;
  float* partA = p.partA + (size_t)team * p.partA_ts;
  float* partC = p.partC + (size_t)team * p.partC_ts;
  // ---- tensor path: TMEM allocation + one-time load of this CTA's weight tiles ----
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + p.oTc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.oTc + 2);
  uint32_t tmem = 0, mbar_uses = 0;
  uint8_t* stage_tc = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(stage) + 1023) & ~static_cast<uintptr_t>(1023));
  if (p.use_tc) {
    if (tid < 32) ntm_b200::umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    if (tid == 32) ntm_b200::umma::mbar_init(mbar, 1);
    ntm_b200::umma::tcgen05_fence_before();
    __syncthreads();
    ntm_b200::umma::tcgen05_fence_after();
    tmem = *tmem_slot;
    for (int l = 0; l < p.L; ++l) tc_load_weights(p.gA[l], p.wA[l], tmem, cta);
    tc_load_weights(p.gC, p.wC, tmem, cta);
    ntm_b200::umma::tcgen05_fence_before();
    __syncthreads();
    ntm_b200::umma::tcgen05_fence_after();
  }
  long long tmark = clock64();
  // phase-cycle accumulators live in shared memory (a global read-modify-write per mark would sit
  // on the critical path); flushed to p.prof once at the end
  long long* prow = p.prof ? reinterpret_cast<long long*>(smem + p.oTc + 4) : nullptr;
  if (prow != nullptr && tid < PROF_SLOTS) prow[tid] = 0;
  __syncthreads();
  auto mark = [&](int slot) { mark_slot(prow, tmark, slot); };

  for (int b0 = team * p.G; b0 < p.B; b0 += p.nteams * p.G) {
    const int Gcur = min(p.G, p.B - b0);
    const int bglob = b0 + gslot;
    const bool active = gslot < Gcur;
    int wcur = 0;

    // ---- prologue: state -> shared memory / workspace ----
    if (active) {
      const float* srcM = p.sM + (size_t)bglob * p.ssM;
      for (int i = tid; i < nrows * p.M4; i += NT) {
        const int r = i / p.M4, d = i - r * p.M4;
        Ms[i] = (d < p.M) ? __ldg(srcM + (size_t)(row0 + r) * p.M + d) : 0.0f;
      }
      const float* srcw = p.sw + (size_t)bglob * p.ssw;
      float* w0 = smem + p.oW0;
      for (int i = tid; i < H * p.N; i += NT) {
        const int h = i / p.N, n = i - h * p.N;
        w0[h * p.Npad + n] = __ldg(srcw + i);
      }
      if (crank == 0) {
        const float* srcr = p.sread + (size_t)bglob * p.ssread;
        for (int i = tid; i < R * p.M; i += NT) {
          const float v = __ldg(srcr + i);
          act[0][(size_t)gslot * p.actK[0] + i] = v;
          if (p.hRead != nullptr) p.hRead[(size_t)bglob * (R * p.M) + i] = v;
        }
        const float* srcc = p.sctrl + (size_t)bglob * p.ssctrl;
        for (int i = tid; i < p.L * p.C; i += NT) {
          const int l = i / p.C, u = i - l * p.C;
          const float c0 = __ldg(srcc + (size_t)l * 2 * p.C + u), h0 = __ldg(srcc + (size_t)l * 2 * p.C + p.C + u);
          cst[((size_t)gslot * p.L + l) * p.C + u] = c0;
          act[l][(size_t)gslot * p.actK[l] + (p.actK[l] - p.C) + u] = h0;
          if (p.hC != nullptr) p.hC[((size_t)bglob * p.L + l) * p.C + u] = c0;
          if (p.hH != nullptr) p.hH[((size_t)bglob * p.L + l) * p.C + u] = h0;
        }
      }
      __syncthreads();
      colsq_local(p, Ms, nrows, smem + p.oK + R * p.M4);
    }
    cluster.sync();
    if (active) finalize_colnorm(p, cluster, smem, p.oK + R * p.M4, smem + p.oCn);
    cluster.sync();   // peers finished reading the exchange buffer before anything overwrites it
    grid_sync(ctr, p.err, epoch, ncta);
    mark(8);

    for (int t = 0; t < p.T; ++t) {
      for (int l = 0; l < p.L; ++l) {
        if (p.gA[l].tc) gemm_phase_tc(p.gA[l], act[l], partA, Gcur, stage_tc, tmem, mbar, mbar_uses, cta);
        else gemm_phase(p.gA[l], act[l], p.wA[l], partA, Gcur, stage, cta, ncta);
        mark(0);
        grid_sync(ctr, p.err, epoch, ncta);
        mark(1);
        lstm_phase(p, l, Gcur, b0, t, cta, ncta, act, cst, partA);
        mark(2);
        grid_sync(ctr, p.err, epoch, ncta);
        mark(3);
      }
      if (p.gC.tc) gemm_phase_tc(p.gC, act[p.L - 1] + (p.actK[p.L - 1] - p.C), partC, Gcur, stage_tc, tmem, mbar, mbar_uses, cta);
      else gemm_phase(p.gC, act[p.L - 1] + (p.actK[p.L - 1] - p.C), p.wC, partC, Gcur, stage, cta, ncta);
      mark(4);
      grid_sync(ctr, p.err, epoch, ncta);
      mark(5);
      if (active)
        phase_d<R, W>(p, cluster, smem, crank, gslot, bglob, t, row0, nrows, wcur, prow, tmark, act[0], partC);
      grid_sync(ctr, p.err, epoch, ncta);
      mark(7);
    }

    // ---- epilogue: final state (ntm_cell.py:223-228) ----
    // A grid barrier that timed out (error flag) means the steps above ran on unsynchronised data: the
    // barriers pass at once from then on, so the kernel still ends in bounded time, and the results of
    // this call are POISONED (NaN logits / outputs) so that no caller can mistake them for an answer.
    if (active && crank == 0 && *reinterpret_cast<volatile int*>(p.err) != 0) {
      const float qnan = __int_as_float(0x7fc00000);
      for (int i = tid; i < p.T * p.O; i += NT) {
        p.logits[(size_t)bglob * p.T * p.O + i] = qnan;
        if (p.outputs != nullptr) p.outputs[(size_t)bglob * p.T * p.O + i] = qnan;
      }
    }
    if (active) {
      float* dstM = p.dM + (size_t)bglob * p.dsM;
      for (int i = tid; i < nrows * p.M; i += NT) {
        const int r = i / p.M, d = i - r * p.M;
        dstM[(size_t)(row0 + r) * p.M + d] = Ms[r * p.M4 + d];
      }
      if (crank == 0) {
        float* dstc = p.dctrl + (size_t)bglob * p.dsctrl;
        for (int i = tid; i < p.L * p.C; i += NT) {
          const int l = i / p.C, u = i - l * p.C;
          dstc[(size_t)l * 2 * p.C + u] = __ldcg(cst + ((size_t)gslot * p.L + l) * p.C + u);
          dstc[(size_t)l * 2 * p.C + p.C + u] =
              __ldcg(act[l] + (size_t)gslot * p.actK[l] + (p.actK[l] - p.C) + u);
        }
      }
    }
    cluster.sync();   // no CTA re-enters the prologue while a peer still reads its shared memory
    mark(9);
  }
  if (prow != nullptr) {
    __syncthreads();
    if (tid < PROF_SLOTS) p.prof[(size_t)blockIdx.x * PROF_SLOTS + tid] = prow[tid];
  }
  if (p.use_tc) {
    ntm_b200::umma::tcgen05_fence_before();
    __syncthreads();
    if (tid < 32) ntm_b200::umma::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------ host wrappers --
typedef void (*SeqKernel)(const KParams);
#define NTM_K(R, W) ntm_seq_kernel<R, W>
static SeqKernel select_kernel(int R, int W) {
  static const SeqKernel table[NTM_B200_MAX_READ_HEADS][NTM_B200_MAX_WRITE_HEADS] = {
      {NTM_K(1, 1), NTM_K(1, 2), NTM_K(1, 3)},
      {NTM_K(2, 1), NTM_K(2, 2), NTM_K(2, 3)},
      {NTM_K(3, 1), NTM_K(3, 2), NTM_K(3, 3)},
      {NTM_K(4, 1), NTM_K(4, 2), NTM_K(4, 3)}};
  return table[R - 1][W - 1];
}

static void fill_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attrs, int grid_ctas, int cluster_size,
                        int smem_bytes, bool cooperative, cudaStream_t stream) {
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = cluster_size;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(grid_ctas);
  cfg->blockDim = dim3(NT);
  cfg->dynamicSmemBytes = smem_bytes;
  cfg->stream = stream;
  cfg->attrs = attrs;
  cfg->numAttrs = cooperative ? 2 : 1;
}

static cudaError_t set_smem(int R, int W, int smem_bytes) {
  return cudaFuncSetAttribute(select_kernel(R, W), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}

// A kernel that executes tcgen05.alloc is reported by the occupancy API as one CTA per SM whatever
// it allocates, although the hardware does co-schedule two such CTAs (the second allocation simply
// succeeds when the columns are free -- measured with tools/occ_probe.cu: 2 x 256 columns run
// concurrently on all 148 SMs).  For the two-CTAs-per-SM build the cluster capacity is therefore
// taken from a proxy kernel with the same CTA size, shared memory and register budget, and the
// launch is a plain cluster launch (the driver would refuse it as cooperative).
__global__ void __launch_bounds__(NT, NTM_MIN_CTAS) occ_proxy_kernel(float* out) {
  extern __shared__ __align__(16) float smem[];
  if (out != nullptr) out[threadIdx.x] = smem[threadIdx.x];
}

static cudaError_t max_clusters(int R, int W, int cluster_size, int grid_ctas, int smem_bytes, int* out) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  fill_config(&cfg, attrs, grid_ctas, cluster_size, smem_bytes, false, nullptr);
  if (NTM_MIN_CTAS > 1) {
    cudaError_t e = cudaFuncSetAttribute(occ_proxy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveClusters(out, occ_proxy_kernel, &cfg);
  }
  return cudaOccupancyMaxActiveClusters(out, select_kernel(R, W), &cfg);
}

static cudaError_t launch(int R, int W, const KParams& p, int grid_ctas, int cluster_size, int smem_bytes,
                          bool cooperative, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  fill_config(&cfg, attrs, grid_ctas, cluster_size, smem_bytes, cooperative, stream);
  return cudaLaunchKernelEx(&cfg, select_kernel(R, W), p);
}

const KernelVariant& variant() {
  static const KernelVariant v{NT, NTM_MIN_CTAS, NTM_MIN_CTAS == 1, set_smem, max_clusters, launch};
  return v;
}

}  // namespace NTM_KNS
}  // namespace ntm_b200
