// ntm_b200.cu -- B200 (sm_100a) implementation of the NTM-cell hot path behind
// the C ABI of include/ntm_b200.h.
//
// What it replaces (paths relative to the reference root):
//   NTMCell.__call__ .......... ntm_cell.py:53-253
//   batched_smooth_cosine_similarity / batched_circular_convolution ... ops.py:135-242
//   LoopNTMTracker.__call__ ... ntm_tracker_new.py:13-64
//
// Design: ONE persistent kernel per call.  A thread-block
// cluster of CS CTAs owns one sequence; its N x M memory is split by rows over
// the cluster's shared memory and stays resident for all T steps, as do the
// head weightings and the per-column norms.  Each timestep has four phases
// separated by a device-wide barrier:
//   A  recurrent controller projection [read | h] @ W_rh, batched over all
//      resident sequences and split (columns x K-slices) over the whole grid;
//   B  split-K reduction + LSTM gates;
//   C  head-parameter / output projection  h @ [W_addr | W_out], same scheme;
//   D  per-cluster fused addressing: activations, column-normalised cosine
//      similarity (pass 1 over M), softmax, gate, circular shift, sharpening
//      (replicated per CTA after a DSMEM all-gather of the similarities), then
//      erase/add write + weighted read + next-step column norms (pass 2 over M),
//      cluster-reduced through DSMEM.
// The x_t @ W_x part of the controller projection has no state dependence and
// is hoisted out of the recurrence into one GEMM over all B*T rows.
//
// No CPU fallback, no dispatch: every entry point fails loudly without an
// sm_100 device.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ntm_b200.h"
#include "ntm_b200_umma.cuh"
#include "ntm_b200_xproj.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 512;          // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int RB = 4;            // memory rows per warp-group in pass 1
constexpr int TBMAX = 16;        // max sequences per warp tile in the skinny GEMMs
constexpr int MAXL = NTM_B200_MAX_LAYERS;
constexpr int SMAX = 2 * NTM_B200_MAX_SHIFT_RANGE + 1;
constexpr int B200_SMS = 148;
constexpr int B200_SMEM_OPTIN = 232448;   // 227 KiB
constexpr int STAGE_BUDGET_BYTES = 44 * 1024;

struct GemmPlan {
  int K, NC, NCs, ldw, lda;            // NCs: row stride of the partial slabs
  int KS, KW, JW, NBT, TB, Gpad, njg, units;
  int tc;      // 1: tcgen05 path (128-column weight tiles resident in TMEM), 0: SIMT path
  int tcol;    // tc: first TMEM column of this GEMM's weight tile (KW/2 "hi" columns, then KW/2 "lo")
};

struct KParams {
  int D, O, N, M, M4, MC, S, C, L, H, P, PO, PO4, write_first, shift0;
  int B, T, CS, NR, G, Npad;
  GemmPlan gA[MAXL];
  GemmPlan gC;
  const float* wA[MAXL];
  const float* bA[MAXL];
  const float* wC;
  const float* bC;
  const float* xw;
  const float *sM, *sw, *sread, *sctrl;
  long long ssM, ssw, ssread, ssctrl;
  float *dM, *dw, *dread, *dctrl;
  long long dsM, dsw, dsread, dsctrl;
  float* logits;
  float* outputs;
  float* dbg;
  long long dbgStride;
  float* act[MAXL];
  int actK[MAXL];
  float* cst;
  float* partA;
  float* partC;
  unsigned* ctr;
  int* err;
  long long* prof;   // [ncta][16] per-phase cycle counters (measurement hook), or null
  // shared-memory carve-up, offsets in floats
  int oMs, oW0, oW1, oCn, oX0, oX1, oScr;
  int oSim, oWg, oK, oE, oA, oSm, oLog;
  int oTc;       // 4 floats: mbarrier (8 B) + TMEM base address (4 B)
  int use_tc;    // any GEMM on the tensor path -> allocate TMEM
};

// ------------------------------------------------------------------ helpers --
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

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
  int ks = 0;
  for (; ks + 4 <= KS; ks += 4) {
    const float a = __ldcg(p + (size_t)ks * stride), b = __ldcg(p + (size_t)(ks + 1) * stride);
    const float c = __ldcg(p + (size_t)(ks + 2) * stride), d = __ldcg(p + (size_t)(ks + 3) * stride);
    v += a; v += b; v += c; v += d;
  }
  if (ks < KS) {
    const float a = __ldcg(p + (size_t)ks * stride);
    const float b = (ks + 1 < KS) ? __ldcg(p + (size_t)(ks + 1) * stride) : 0.0f;
    const float c = (ks + 2 < KS) ? __ldcg(p + (size_t)(ks + 2) * stride) : 0.0f;
    v += a;
    if (ks + 1 < KS) v += b;
    if (ks + 2 < KS) v += c;
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
    const int nq = N / 4;                                  // accumulator columns (sequences) per warp quarter
    const int bq0 = (warp >> 2) * nq;
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
                                           int cta, int ncta) {
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
    float v = (l == 0) ? __ldg(p.xw + ((size_t)(b0 + b) * p.T + t) * (size_t)(4 * C) + col)
                       : __ldg(p.bA[l] + col);
    v = sum_slabs(p.partA + (size_t)b * g.NCs + col, (size_t)g.Gpad * g.NCs, KS, v);
    const unsigned lane = threadIdx.x & 31u, gl = lane & ~3u;
    const float zi = __shfl_sync(0xffffffffu, v, gl + 0);
    const float zj = __shfl_sync(0xffffffffu, v, gl + 1);
    const float zf = __shfl_sync(0xffffffffu, v, gl + 2);
    const float zo = __shfl_sync(0xffffffffu, v, gl + 3);
    if (ok && q == 0) {
      float* cp = p.cst + ((size_t)b * p.L + l) * C + u;
      const float c_prev = __ldcg(cp);
      const float c_new = c_prev * sigmoid_f(zf) + sigmoid_f(zi) * tanhf(zj);
      const float h_new = tanhf(c_new) * sigmoid_f(zo);
      *cp = c_new;
      p.act[l][(size_t)b * p.actK[l] + (p.actK[l] - C) + u] = h_new;
      if (l + 1 < p.L) p.act[l + 1][(size_t)b * p.actK[l + 1] + u] = h_new;
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
                                        int& wcur, int xpar, long long* prow, long long& tmark) {
  constexpr int H = R + W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = p.M, M4 = p.M4, MC = p.MC, N = p.N, Npad = p.Npad, S = p.S;
  float* Ms = smem + p.oMs;
  float* wprev = smem + (wcur ? p.oW1 : p.oW0);
  float* wnew = smem + (wcur ? p.oW0 : p.oW1);
  float* cn = smem + p.oCn;
  float* xch = smem + (xpar ? p.oX1 : p.oX0);   // [R][M4] read partials, then [M4] column squares
  float* simA = smem + p.oSim;                   // [H][Npad]
  float* wg = smem + p.oWg;                      // [H][Npad]
  float* kS = smem + p.oK;                       // [H][M4]
  float* eS = smem + p.oE;                       // [W][M4]
  float* aS = smem + p.oA;                       // [W][M4]
  float* sm = smem + p.oSm;                      // beta[H] g[H] gamma[H] rs[H] sw[H][SMAX]
  float* sBeta = sm, *sG = sm + H, *sGam = sm + 2 * H, *sSw = sm + 4 * H;
  float* sPart = sm + 4 * H + H * SMAX;          // [NWARP][H] per-warp partial key norms
  const bool last = (t == p.T - 1);
  float* dbg = (p.dbg != nullptr && last && crank == 0) ? p.dbg + (size_t)bglob * p.dbgStride : nullptr;

  // ---- D0: split-K reduction of phase C + bias, activations (ntm_cell.py:124-196) ----
  const int offBeta = H * M, offG = offBeta + H, offS = offG + H, offGam = offS + S * H,
            offE = offGam + H, offA = offE + M * W;
  // pass a: raw[q] = bias[q] + sum_ks partC[ks][slot][q], float4-vectorised.  `raw` lives in the
  // wg scratch (never written by a peer CTA; simA is, by the pass-1 all-gather).
  float* raw = wg;
  {
    const float4* pc4 = reinterpret_cast<const float4*>(p.partC + (size_t)gslot * p.gC.NCs);
    const float4* b4 = reinterpret_cast<const float4*>(p.bC);
    const size_t slab4 = (size_t)p.gC.Gpad * p.gC.NCs / 4;
    for (int q4 = tid; q4 < p.PO4 / 4; q4 += NT)
      reinterpret_cast<float4*>(raw)[q4] = sum_slabs4(pc4 + q4, slab4, p.gC.KS, __ldg(b4 + q4));
  }
  __syncthreads();
  mark_slot(prow, tmark, 15);
  // pass b: activations.  Keys: kS[h][d] = tanh(raw) * cn[d]  (the key's own 1/|k| is a per-head
  // scalar and is applied to the similarities later); per-head sum of squares via fixed-order
  // warp partials.  Pad lanes d >= M are written as zeros.
  {
    float ss[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      ss[h] = 0.0f;
      for (int d = tid; d < M4; d += NT) {
        float kv = 0.0f;
        if (d < M) {
          kv = tanhf(raw[h * M + d]);
          if (dbg) dbg[h * M + d] = kv;
        }
        kS[h * M4 + d] = kv * cn[d];
        ss[h] = fmaf(kv, kv, ss[h]);
      }
      ss[h] = warp_sum(ss[h]);
    }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < H; ++h) sPart[warp * H + h] = ss[h];
    }
#pragma unroll
    for (int h = 0; h < W; ++h) {
      for (int d = tid; d < M4; d += NT) {
        float ev = 0.0f, av = 0.0f;
        if (d < M) {
          ev = sigmoid_f(raw[offE + h * M + d]);
          av = tanhf(raw[offA + h * M + d]);
          if (dbg) { dbg[offE + h * M + d] = ev; dbg[offA + h * M + d] = av; }
        }
        eS[h * M4 + d] = ev;
        aS[h * M4 + d] = av;
      }
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
    for (int i = 0; i < S; ++i) { sp[i] = expf(raw[offS + tid * S + i] - mx); sum += sp[i]; }
    for (int i = 0; i < S; ++i) sp[i] = sp[i] / sum;
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
    for (int i = 0; i < p.O; ++i) sum += expf(lg[i] - mx);
    for (int i = 0; i < p.O; ++i) {
      p.logits[o + i] = lg[i];
      if (p.outputs) p.outputs[o + i] = expf(lg[i] - mx) / sum;
    }
  }
  __syncthreads();
  mark_slot(prow, tmark, 10);

  // ---- pass 1: sim[h][n] = sum_d kc[h][d] * M[n][d] over this CTA's rows (ops.py:156) ----
  {
    int LPR = 32;                      // lanes cooperating on one block of RB rows
    while (LPR > 1 && (LPR >> 1) >= MC) LPR >>= 1;
    const int GPW = 32 / LPR;
    const int sg = lane / LPR, lg = lane - sg * LPR;
    const int nRB = (nrows + RB - 1) / RB;
    const int iters = (nRB + NWARP * GPW - 1) / (NWARP * GPW);
    for (int it = 0; it < iters; ++it) {
      const int rb = (it * NWARP + warp) * GPW + sg;
      const bool active = rb < nRB;
      float acc[RB][H];
#pragma unroll
      for (int i = 0; i < RB; ++i)
#pragma unroll
        for (int h = 0; h < H; ++h) acc[i][h] = 0.0f;
      int rows[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) rows[i] = min(rb * RB + i, nrows - 1);
      if (active) {
        for (int c = lg; c < MC; c += LPR) {
          float4 k4[H];
#pragma unroll
          for (int h = 0; h < H; ++h) k4[h] = *reinterpret_cast<const float4*>(kS + h * M4 + 4 * c);
#pragma unroll
          for (int i = 0; i < RB; ++i) {
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
      for (int o = LPR >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int h = 0; h < H; ++h) acc[i][h] += __shfl_xor_sync(0xffffffffu, acc[i][h], o);
      }
      if (active && lg == 0) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const int rl = rb * RB + i;
          if (rl < nrows) {
            for (int r = 0; r < p.CS; ++r) {
              float* rem = cluster.map_shared_rank(simA, r);
#pragma unroll
              for (int h = 0; h < H; ++h) rem[h * Npad + row0 + rl] = acc[i][h];
            }
          }
        }
      }
    }
  }
  cluster.sync();
  mark_slot(prow, tmark, 11);

  // ---- addressing on the full [H][N] weighting, replicated in every CTA (ntm_cell.py:140-176) ----
  for (int h = warp; h < H; h += NWARP) {
    float* sh = simA + h * Npad;
    float* gh = wg + h * Npad;
    const float gate = sG[h], gamma = sGam[h];
    float kn = 0.0f;                                  // |k_h|^2, fixed summation order
    for (int w2 = 0; w2 < NWARP; ++w2) kn += sPart[w2 * H + h];
    const float rs = 1.0f / sqrtf(fmaxf(kn, 1e-12f));   // ops.py:152
    const float beta = sBeta[h];
    float mx = -INFINITY;
#pragma unroll 4
    for (int n = lane; n < N; n += 32) {
      const float sv = sh[n] * rs;                    // similarity (ops.py:156)
      const float x = sv * beta;
      if (dbg) dbg[p.P + (0 * H + h) * N + n] = sv;
      sh[n] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll 4
    for (int n = lane; n < N; n += 32) {
      const float e = expf(sh[n] - mx);
      sh[n] = e;
      sum += e;
    }
    sum = warp_sum(sum);
#pragma unroll 4
    for (int n = lane; n < N; n += 32) {
      const float wc = sh[n] / sum;
      const float v = wc * gate + wprev[h * Npad + n] * (1.0f - gate);
      gh[n] = v;
      if (dbg) {
        dbg[p.P + (1 * H + h) * N + n] = wc;
        dbg[p.P + (2 * H + h) * N + n] = v;
      }
    }
    __syncwarp();
    float psum = 0.0f;
#pragma unroll 4
    for (int n = lane; n < N; n += 32) {
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
    const float den = psum + 1e-3f;          // ntm_cell.py:175-176
#pragma unroll 4
    for (int n = lane; n < N; n += 32) {
      const float wv = sh[n] / den;
      wnew[h * Npad + n] = wv;
      if (last && crank == 0) p.dw[(size_t)bglob * p.dsw + h * N + n] = wv;
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
          float4 E = make_float4(1.f, 1.f, 1.f, 1.f), A = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int h = 0; h < W; ++h) {
            const float ww = wnew[(R + h) * Npad + n];
            E.x *= (1.0f - ww * e4[h].x); E.y *= (1.0f - ww * e4[h].y);
            E.z *= (1.0f - ww * e4[h].z); E.w *= (1.0f - ww * e4[h].w);
            A.x = fmaf(ww, a4[h].x, A.x); A.y = fmaf(ww, a4[h].y, A.y);
            A.z = fmaf(ww, a4[h].z, A.z); A.w = fmaf(ww, a4[h].w, A.w);
          }
          float4 mn;
          mn.x = fmaf(m.x, E.x, A.x); mn.y = fmaf(m.y, E.y, A.y);
          mn.z = fmaf(m.z, E.z, A.z); mn.w = fmaf(m.w, E.w, A.w);
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
  const int oX = xpar ? p.oX1 : p.oX0;
  finalize_colnorm(p, cluster, smem, oX + R * M4, cn);
  for (int i = crank * NT + tid; i < R * M; i += p.CS * NT) {
    const int r = i / M, d = i - r * M;
    float s = 0.0f;
    for (int q = 0; q < p.CS; ++q) {
      const float* rem = cluster.map_shared_rank(smem + oX, q);
      s += rem[r * M4 + d];
    }
    p.act[0][(size_t)gslot * p.actK[0] + i] = s;          // next step's controller input
    if (last) p.dread[(size_t)bglob * p.dsread + i] = s;
  }
  wcur ^= 1;
  mark_slot(prow, tmark, 14);
}

// ---------------------------------------------------------- the persistent kernel --
template <int R, int W>
__global__ void __launch_bounds__(NT, 1) ntm_seq_kernel(const KParams p) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int H = R + W;
  const int tid = threadIdx.x;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int crank = (int)cluster.block_rank();
  const int gslot = cta / p.CS;                    // cluster index = resident-sequence slot
  const int row0 = crank * p.NR;
  const int nrows = max(0, min(p.NR, p.N - row0));
  unsigned epoch = 0;
  float* Ms = smem + p.oMs;
  float* stage = smem + p.oScr;
  // ---- tensor path: TMEM allocation + one-time load of this CTA's weight tiles ----
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + p.oTc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.oTc + 2);
  uint32_t tmem = 0, mbar_uses = 0;
  uint8_t* stage_tc = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(stage) + 1023) & ~static_cast<uintptr_t>(1023));
  if (p.use_tc) {
    if (tid < 32) ntm_b200::umma::tmem_alloc(tmem_slot, 512);
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
  if (prow != nullptr && tid < 16) prow[tid] = 0;
  __syncthreads();
  auto mark = [&](int slot) { mark_slot(prow, tmark, slot); };

  for (int b0 = 0; b0 < p.B; b0 += p.G) {
    const int Gcur = min(p.G, p.B - b0);
    const int bglob = b0 + gslot;
    const bool active = gslot < Gcur;
    int wcur = 0, xpar = 0;

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
        for (int i = tid; i < R * p.M; i += NT) p.act[0][(size_t)gslot * p.actK[0] + i] = __ldg(srcr + i);
        const float* srcc = p.sctrl + (size_t)bglob * p.ssctrl;
        for (int i = tid; i < p.L * p.C; i += NT) {
          const int l = i / p.C, u = i - l * p.C;
          p.cst[((size_t)gslot * p.L + l) * p.C + u] = __ldg(srcc + (size_t)l * 2 * p.C + u);
          p.act[l][(size_t)gslot * p.actK[l] + (p.actK[l] - p.C) + u] = __ldg(srcc + (size_t)l * 2 * p.C + p.C + u);
        }
      }
      __syncthreads();
      colsq_local(p, Ms, nrows, smem + p.oX1 + R * p.M4);
    }
    cluster.sync();
    if (active) finalize_colnorm(p, cluster, smem, p.oX1 + R * p.M4, smem + p.oCn);
    cluster.sync();   // peers finished reading oX1 before step 1 (xpar == 1) overwrites it
    grid_sync(p.ctr, p.err, epoch, ncta);
    mark(8);

    for (int t = 0; t < p.T; ++t) {
      for (int l = 0; l < p.L; ++l) {
        if (p.gA[l].tc) gemm_phase_tc(p.gA[l], p.act[l], p.partA, Gcur, stage_tc, tmem, mbar, mbar_uses, cta);
        else gemm_phase(p.gA[l], p.act[l], p.wA[l], p.partA, Gcur, stage, cta, ncta);
        mark(0);
        grid_sync(p.ctr, p.err, epoch, ncta);
        mark(1);
        lstm_phase(p, l, Gcur, b0, t, cta, ncta);
        mark(2);
        grid_sync(p.ctr, p.err, epoch, ncta);
        mark(3);
      }
      if (p.gC.tc) gemm_phase_tc(p.gC, p.act[p.L - 1] + (p.actK[p.L - 1] - p.C), p.partC, Gcur, stage_tc, tmem, mbar, mbar_uses, cta);
      else gemm_phase(p.gC, p.act[p.L - 1] + (p.actK[p.L - 1] - p.C), p.wC, p.partC, Gcur, stage, cta, ncta);
      mark(4);
      grid_sync(p.ctr, p.err, epoch, ncta);
      mark(5);
      if (active) {
        phase_d<R, W>(p, cluster, smem, crank, gslot, bglob, t, row0, nrows, wcur, xpar, prow, tmark);
        xpar ^= 1;
      }
      mark(6);
      grid_sync(p.ctr, p.err, epoch, ncta);
      mark(7);
    }

    // ---- epilogue: final state (ntm_cell.py:223-228) ----
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
          dstc[(size_t)l * 2 * p.C + u] = __ldcg(p.cst + ((size_t)gslot * p.L + l) * p.C + u);
          dstc[(size_t)l * 2 * p.C + p.C + u] =
              __ldcg(p.act[l] + (size_t)gslot * p.actK[l] + (p.actK[l] - p.C) + u);
        }
      }
    }
    cluster.sync();   // no CTA re-enters the prologue while a peer still reads its shared memory
    mark(9);
  }
  if (prow != nullptr) {
    __syncthreads();
    if (tid < 16) p.prof[(size_t)cta * 16 + tid] = prow[tid];
  }
  if (p.use_tc) {
    ntm_b200::umma::tcgen05_fence_before();
    __syncthreads();
    if (tid < 32) ntm_b200::umma::tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------ packing --
__global__ void pack_ao_kernel(const float* __restrict__ aw, const float* __restrict__ ab,
                               const float* __restrict__ ow, const float* __restrict__ ob,
                               float* wC, float* bC, int C, int P, int O, int PO4) {
  const int total = (C + 1) * PO4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / PO4, q = i - r * PO4;
    float v = 0.0f;
    if (r < C) {
      if (q < P) v = aw[(size_t)r * P + q];
      else if (q < P + O) v = ow[(size_t)r * O + (q - P)];
      wC[i] = v;
    } else {
      if (q < P) v = ab[q];
      else if (q < P + O) v = ob[q - P];
      bC[q] = v;
    }
  }
}

// --------------------------------------------------------------- host side --
thread_local char g_cuda_err[256] = "";
std::atomic<int> g_profiling{0};
thread_local cudaEvent_t g_ev[3] = {nullptr, nullptr, nullptr};
thread_local bool g_ev_valid = false;
thread_local int g_last_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
std::atomic<long long> g_launches{0};

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
  return NTM_B200_ERR_CUDA;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
inline long long align_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

struct HostPlan {
  int H, S, P, PO, PO4, M4, MC, Npad;
  int CS, NR, Gmax;
  int smem_floats;
  int oMs, oW0, oW1, oCn, oX0, oX1, oScr, oSim, oWg, oK, oE, oA, oSm, oLog, oTc;
  int scr_floats;
  int actK[MAXL];
  long long packed_bytes, debug_floats;
};

int validate_shape(const ntm_b200_shape* s) {
  if (!s) return NTM_B200_ERR_NULL_POINTER;
  if (s->input_dim < 1 || s->output_dim < 1 || s->mem_size < 1 || s->mem_dim < 1 ||
      s->controller_hidden_size < 1 || s->controller_num_layers < 1 ||
      s->controller_num_layers > MAXL)
    return NTM_B200_ERR_BAD_SHAPE;
  if (s->shift_range < 0 || s->shift_range > NTM_B200_MAX_SHIFT_RANGE) return NTM_B200_ERR_BAD_SHIFT;
  // circular_shift asserts 0 <= splitting point < N for every tap (ops.py:229-231)
  const int S = 2 * s->shift_range + 1;
  const int start = -((S + 1) / 2);   // floor(-S / 2) for odd S
  if (-start >= s->mem_size || (S + start - 1) >= s->mem_size) return NTM_B200_ERR_BAD_SHIFT;
  if (s->read_head_size < 1 || s->read_head_size > NTM_B200_MAX_READ_HEADS ||
      s->write_head_size < 1 || s->write_head_size > NTM_B200_MAX_WRITE_HEADS)
    return NTM_B200_ERR_UNSUPPORTED_HEADS;
  return NTM_B200_OK;
}

// Shared-memory carve-up for one CTA of a CS-cluster.  Returns bytes.
long long layout_for(const ntm_b200_shape* s, int CS, HostPlan* hp) {
  const int H = s->read_head_size + s->write_head_size, R = s->read_head_size, W = s->write_head_size;
  hp->H = H;
  hp->S = 2 * s->shift_range + 1;
  hp->M4 = round_up(s->mem_dim, 4);
  hp->MC = hp->M4 / 4;
  hp->Npad = round_up(s->mem_size, 4);
  hp->P = H * s->mem_dim + 3 * H + hp->S * H + 2 * s->mem_dim * W;
  hp->PO = hp->P + s->output_dim;
  hp->PO4 = round_up(hp->PO, 4);
  hp->CS = CS;
  hp->NR = ceil_div(s->mem_size, CS);
  int o = 0;
  auto take = [&](int n) { int r = o; o += round_up(n, 4); return r; };
  hp->oMs = take(hp->NR * hp->M4);
  hp->oW0 = take(H * hp->Npad);
  hp->oW1 = take(H * hp->Npad);
  hp->oCn = take(hp->M4);
  hp->oX0 = take((R + 1) * hp->M4);
  hp->oX1 = take((R + 1) * hp->M4);
  hp->oTc = take(4 + 32);   // mbarrier + TMEM base, then 16 int64 phase-cycle accumulators
  hp->oScr = o;
  // phase-D temporaries inside the scratch union
  int d = o;
  auto taked = [&](int n) { int r = d; d += round_up(n, 4); return r; };
  // [sim | wg] are contiguous and together also hold the raw head-parameter vector (PO4 floats)
  hp->oSim = taked(H * hp->Npad);                      // written by peer CTAs (all-gather)
  hp->oWg = taked(std::max(H * hp->Npad, hp->PO4));    // also holds the raw head-parameter vector
  hp->oK = taked(H * hp->M4);      // kS, eS, aS must stay contiguous (zero-filled together)
  hp->oE = taked(W * hp->M4);
  hp->oA = taked(W * hp->M4);
  hp->oSm = taked(4 * H + H * SMAX + NWARP * H);
  hp->oLog = taked(s->output_dim);
  const int dfl = d - o;
  hp->scr_floats = std::max(dfl, STAGE_BUDGET_BYTES / 4);
  hp->smem_floats = o + hp->scr_floats;
  return 4ll * hp->smem_floats;
}

int make_host_plan(const ntm_b200_shape* s, int nsm, int smem_max, HostPlan* hp) {
  int st = validate_shape(s);
  if (st) return st;
  int CS = 1;
  for (;; CS *= 2) {
    if (CS > 8) return NTM_B200_ERR_TOO_LARGE;
    if (layout_for(s, CS, hp) <= smem_max) break;
  }
  hp->Gmax = std::max(1, nsm / CS);
  const int C = s->controller_hidden_size;
  for (int l = 0; l < s->controller_num_layers; ++l)
    hp->actK[l] = (l == 0) ? s->read_head_size * s->mem_dim + C : 2 * C;
  hp->packed_bytes = 4ll * (long long)(C + 1) * hp->PO4;
  hp->debug_floats = (long long)hp->P + 5ll * hp->H * s->mem_size;
  return NTM_B200_OK;
}

GemmPlan plan_gemm(int K, int NC, int NCs, int ldw, int lda, int G, int ncta) {
  GemmPlan g{};
  g.K = K; g.NC = NC; g.NCs = NCs; g.ldw = ldw; g.lda = lda;
  const int G4 = round_up(std::max(G, 1), 4);
  int best_nbt = NWARP, best_pad = 1 << 30;
  for (int nbt = ceil_div(G4, TBMAX); nbt <= NWARP; ++nbt) {
    const int tb = round_up(ceil_div(G4, nbt), 4);
    if (tb > TBMAX) continue;
    const int pad = nbt * tb;
    if (pad < best_pad) { best_pad = pad; best_nbt = nbt; }
  }
  g.NBT = best_nbt;
  g.TB = round_up(ceil_div(G4, g.NBT), 4);
  g.Gpad = g.NBT * g.TB;
  g.JW = std::max(1, NWARP / g.NBT);
  const int nj64 = ceil_div(NC, 64);
  g.JW = std::min(g.JW, nj64);
  g.njg = ceil_div(nj64, g.JW);
  int KS = std::max(1, std::min(ncta / std::max(1, g.njg), ceil_div(K, 16)));
  int KW = round_up(ceil_div(K, KS), 4);
  const int kw_cap = std::max(4, (STAGE_BUDGET_BYTES / 4 / g.Gpad) / 4 * 4);
  KW = std::min(KW, kw_cap);
  g.KW = KW;
  g.KS = ceil_div(K, KW);
  g.units = g.KS * g.njg;
  return g;
}

// Tensor-path plan: one unit (128-column tile x K-slice) per CTA, weights resident in TMEM.
bool plan_gemm_tc(int K, int NC, int NCs, int ldw, int lda, int G, int ncta, int tcol, int scr_bytes,
                  GemmPlan* out) {
  GemmPlan g{};
  g.K = K; g.NC = NC; g.NCs = NCs; g.ldw = ldw; g.lda = lda;
  g.tc = 1;
  g.tcol = tcol;
  g.Gpad = round_up(std::max(G, 1), 16);
  if (g.Gpad > 256) return false;
  const int tiles = ceil_div(NC, 128);
  if (tiles > ncta) return false;
  int KS = std::max(1, std::min(ncta / tiles, ceil_div(K, 16)));
  if (const char* e = getenv("NTM_B200_TC_MAX_KS")) KS = std::max(1, std::min(KS, atoi(e)));
  const int KW = round_up(ceil_div(K, KS), 16);
  KS = ceil_div(K, KW);
  const int katoms = ceil_div(KW, 64);
  if (2 * g.Gpad * katoms * 128 + 1024 > scr_bytes) return false;
  g.KW = KW; g.KS = KS; g.njg = tiles; g.units = KS * tiles;
  *out = g;
  return true;
}

// All GEMM plans of one launch.  The tensor path is used when every GEMM's weight tile fits the
// 512 TMEM columns next to the accumulator (else the SIMT path, e.g. for small grids).
void choose_plans(const ntm_b200_shape* s, const HostPlan& hp, int G, int ncta, bool allow_tc,
                  GemmPlan* gA, GemmPlan* gC, int* use_tc) {
  const int C = s->controller_hidden_size, L = s->controller_num_layers;
  const int scr_bytes = 4 * hp.scr_floats;
  bool tc = allow_tc && getenv("NTM_B200_DISABLE_TC") == nullptr;
  if (tc) {
    int col = round_up(round_up(std::max(G, 1), 16), 32);     // accumulator columns first
    for (int l = 0; l < L && tc; ++l) {
      tc = plan_gemm_tc(hp.actK[l], 4 * C, 4 * C, 4 * C, hp.actK[l], G, ncta, col, scr_bytes, &gA[l]);
      if (tc) col += gA[l].KW;
    }
    if (tc) tc = plan_gemm_tc(C, hp.PO, hp.PO4, hp.PO4, hp.actK[L - 1], G, ncta, col, scr_bytes, gC);
    if (tc) col += gC->KW;
    if (col > 512) tc = false;
  }
  if (!tc) {
    for (int l = 0; l < L; ++l) gA[l] = plan_gemm(hp.actK[l], 4 * C, 4 * C, 4 * C, hp.actK[l], G, ncta);
    *gC = plan_gemm(C, hp.PO, hp.PO4, hp.PO4, hp.actK[L - 1], G, ncta);
  }
  *use_tc = tc ? 1 : 0;
}

struct Workspace {
  long long off_ctr, off_err, off_prof, off_act[MAXL], off_cst, off_partA, off_partC, off_xw, total;
};

// Workspace sized for the planner's upper bounds (Gmax resident sequences, a
// full grid), so it does not depend on what the occupancy query returns later.
void layout_workspace(const ntm_b200_shape* s, const HostPlan& hp, long long B, long long T,
                      Workspace* ws) {
  const int C = s->controller_hidden_size, L = s->controller_num_layers;
  const int Gm = hp.Gmax;
  long long o = 0;
  auto take = [&](long long bytes) { long long r = o; o = align_up_ll(o + bytes, 256); return r; };
  ws->off_ctr = take(256);
  ws->off_err = take(256);
  ws->off_prof = take(8ll * 16 * 1024);   // directly after ctr/err: zeroed by the same memset
  for (int l = 0; l < L; ++l) ws->off_act[l] = take(4ll * Gm * hp.actK[l]);
  ws->off_cst = take(4ll * Gm * L * C);
  // partial-slab sizes: maximum over every resident-sequence count the launch may end up with,
  // on either GEMM path
  long long pa = 0, pc = 0;
  for (int G = 1; G <= Gm; ++G) {
    const int ncta = G * hp.CS;
    for (int variant = 0; variant < 2; ++variant) {
      GemmPlan gA[MAXL], gC;
      int use_tc = 0;
      choose_plans(s, hp, G, ncta, variant == 1, gA, &gC, &use_tc);
      for (int l = 0; l < L; ++l) pa = std::max(pa, 4ll * gA[l].KS * gA[l].Gpad * gA[l].NCs);
      pc = std::max(pc, 4ll * gC.KS * gC.Gpad * gC.NCs);
    }
  }
  ws->off_partA = take(pa);
  ws->off_partC = take(pc);
  ws->off_xw = take(4ll * B * T * 4 * C);
  ws->total = o;
}

struct DeviceInfo { int ok; int nsm; int smem_optin; int cc_major; };

DeviceInfo device_info() {
  DeviceInfo d{0, B200_SMS, B200_SMEM_OPTIN, 0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return d; }
  int major = 0, nsm = 0, smem = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
    cudaGetLastError();
    return d;
  }
  d.ok = (major == 10);
  d.nsm = nsm; d.smem_optin = smem; d.cc_major = major;
  return d;
}

typedef void (*SeqKernel)(const KParams);
#define NTM_K(R, W) ntm_seq_kernel<R, W>
SeqKernel select_kernel(int R, int W) {
  static const SeqKernel table[NTM_B200_MAX_READ_HEADS][NTM_B200_MAX_WRITE_HEADS] = {
      {NTM_K(1, 1), NTM_K(1, 2), NTM_K(1, 3)},
      {NTM_K(2, 1), NTM_K(2, 2), NTM_K(2, 3)},
      {NTM_K(3, 1), NTM_K(3, 2), NTM_K(3, 3)},
      {NTM_K(4, 1), NTM_K(4, 2), NTM_K(4, 3)}};
  return table[R - 1][W - 1];
}

int check_state(const ntm_b200_state* st) {
  if (!st || !st->M || !st->w || !st->read || !st->controller_state) return NTM_B200_ERR_NULL_POINTER;
  return NTM_B200_OK;
}

}  // namespace

// ------------------------------------------------------------------- C ABI --
extern "C" {

int32_t ntm_b200_abi_version(void) { return NTM_B200_ABI_VERSION; }

const char* ntm_b200_status_string(int32_t status) {
  switch (status) {
    case NTM_B200_OK: return "ok";
    case NTM_B200_ERR_BAD_SHAPE: return "bad shape (non-positive dimension or too many controller layers)";
    case NTM_B200_ERR_BAD_SHIFT: return "shift_range out of range for mem_size";
    case NTM_B200_ERR_NULL_POINTER: return "null pointer argument";
    case NTM_B200_ERR_UNSUPPORTED_HEADS: return "head count outside 1..4 read / 1..3 write";
    case NTM_B200_ERR_TOO_LARGE: return "per-sequence state does not fit an 8-CTA cluster's shared memory";
    case NTM_B200_ERR_WORKSPACE: return "workspace or packed-weight buffer too small";
    case NTM_B200_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
    case NTM_B200_ERR_CUDA: return "CUDA runtime error";
    case NTM_B200_ERR_DEVICE_TIMEOUT: return "device-side grid barrier timed out";
    default: return "unknown status";
  }
}

const char* ntm_b200_last_cuda_error(void) { return g_cuda_err; }

int64_t ntm_b200_launch_count(void) { return g_launches.load(); }

int32_t ntm_b200_query(const ntm_b200_shape* shape, int64_t batch, int64_t steps,
                       ntm_b200_plan* plan_out) {
  if (!shape || !plan_out) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || steps < 1) return NTM_B200_ERR_BAD_SHAPE;
  DeviceInfo di = device_info();
  HostPlan hp{};
  int st = make_host_plan(shape, di.nsm, di.smem_optin, &hp);
  if (st) return st;
  Workspace ws{};
  layout_workspace(shape, hp, batch, steps, &ws);
  plan_out->cluster_size = hp.CS;
  plan_out->rows_per_cta = hp.NR;
  plan_out->sequences_resident = (int32_t)std::min<long long>(hp.Gmax, batch);
  plan_out->threads_per_cta = NT;
  plan_out->smem_bytes_per_cta = 4ll * hp.smem_floats;
  plan_out->workspace_bytes = ws.total;
  plan_out->packed_bytes = hp.packed_bytes;
  plan_out->debug_floats_per_sequence = hp.debug_floats;
  return NTM_B200_OK;
}

int32_t ntm_b200_pack_weights(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                              void* packed, int64_t packed_bytes, void* stream) {
  if (!shape || !weights || !packed) return NTM_B200_ERR_NULL_POINTER;
  DeviceInfo di = device_info();
  HostPlan hp{};
  int st = make_host_plan(shape, di.nsm, di.smem_optin, &hp);
  if (st) return st;
  if (!di.ok) return NTM_B200_ERR_NO_DEVICE;
  if (packed_bytes < hp.packed_bytes) return NTM_B200_ERR_WORKSPACE;
  if (!weights->addr_w || !weights->addr_b || !weights->out_w || !weights->out_b)
    return NTM_B200_ERR_NULL_POINTER;
  const int C = shape->controller_hidden_size;
  float* wC = static_cast<float*>(packed);
  float* bC = wC + (size_t)C * hp.PO4;
  pack_ao_kernel<<<B200_SMS, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      weights->addr_w, weights->addr_b, weights->out_w, weights->out_b, wC, bC, C, hp.P,
      shape->output_dim, hp.PO4);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "pack_ao_kernel");
  return NTM_B200_OK;
}

int32_t ntm_b200_forward_seq(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                             const void* packed, int64_t batch, int64_t steps,
                             const float* inputs, const ntm_b200_state* state_in,
                             const ntm_b200_state* state_out, float* logits, float* outputs,
                             float* debug_taps, void* workspace, int64_t workspace_bytes,
                             void* stream_v) {
  if (!shape || !weights || !packed || !inputs || !logits || !workspace) return NTM_B200_ERR_NULL_POINTER;
  int st = check_state(state_in);
  if (st) return st;
  st = check_state(state_out);
  if (st) return st;
  if (batch < 1 || steps < 1 || batch > (1 << 24) || steps > (1 << 24)) return NTM_B200_ERR_BAD_SHAPE;
  DeviceInfo di = device_info();
  HostPlan hp{};
  st = make_host_plan(shape, di.nsm, di.smem_optin, &hp);
  if (st) return st;
  if (!di.ok) return NTM_B200_ERR_NO_DEVICE;
  const int L = shape->controller_num_layers, C = shape->controller_hidden_size;
  for (int l = 0; l < L; ++l)
    if (!weights->lstm_w[l] || !weights->lstm_b[l]) return NTM_B200_ERR_NULL_POINTER;
  Workspace ws{};
  layout_workspace(shape, hp, batch, steps, &ws);
  if (workspace_bytes < ws.total) return NTM_B200_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* wsb = static_cast<char*>(workspace);
  cudaError_t e;

  SeqKernel kern = select_kernel(shape->read_head_size, shape->write_head_size);
  const int smem_bytes = 4 * hp.smem_floats;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(smem)");

  // how many clusters are co-resident (1 CTA per SM by shared-memory footprint)
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = hp.CS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(hp.Gmax * hp.CS);
  int max_clusters = 0;
  e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveClusters");
  if (max_clusters < 1) return NTM_B200_ERR_TOO_LARGE;
  const int G = (int)std::min<long long>(std::min(max_clusters, hp.Gmax), batch);
  const int ncta = G * hp.CS;

  const bool prof = g_profiling.load() != 0;
  if (prof) {
    for (int i = 0; i < 3; ++i)
      if (!g_ev[i] && (e = cudaEventCreate(&g_ev[i])) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
    cudaEventRecord(g_ev[0], stream);
  }
  // hoisted x-projection: xw[b,t,:] = x[b,t,:] @ W_lstm0[0:D,:] + b_lstm0
  float* xw = reinterpret_cast<float*>(wsb + ws.off_xw);
  st = ntm_b200::launch_xproj(inputs, weights->lstm_w[0], weights->lstm_b[0], xw,
                              (long long)batch * steps, shape->input_dim, 4 * C, stream);
  g_launches++;
  if (st) return set_cuda_error(cudaGetLastError(), "xproj");

  KParams p{};
  p.D = shape->input_dim; p.O = shape->output_dim; p.N = shape->mem_size; p.M = shape->mem_dim;
  p.M4 = hp.M4; p.MC = hp.MC; p.S = hp.S; p.C = C; p.L = L; p.H = hp.H; p.P = hp.P; p.PO = hp.PO;
  p.PO4 = hp.PO4; p.write_first = shape->write_first ? 1 : 0;
  p.shift0 = -((hp.S + 1) / 2);   // Python-2 floor(-S/2), ops.py:204
  p.B = (int)batch; p.T = (int)steps; p.CS = hp.CS; p.NR = hp.NR; p.G = G; p.Npad = hp.Npad;
  for (int l = 0; l < L; ++l) {
    p.actK[l] = hp.actK[l];
    p.act[l] = reinterpret_cast<float*>(wsb + ws.off_act[l]);
    p.wA[l] = weights->lstm_w[l] + (l == 0 ? (size_t)shape->input_dim * 4 * C : 0);
    p.bA[l] = weights->lstm_b[l];
  }
  choose_plans(shape, hp, G, ncta, true, p.gA, &p.gC, &p.use_tc);
  g_last_info[0] = p.use_tc; g_last_info[1] = G; g_last_info[2] = ncta; g_last_info[3] = hp.CS;
  g_last_info[4] = p.gA[0].KS; g_last_info[5] = p.gA[0].KW; g_last_info[6] = p.gC.KS; g_last_info[7] = p.gC.KW;
  p.wC = static_cast<const float*>(packed);
  p.bC = p.wC + (size_t)C * hp.PO4;
  p.xw = xw;
  p.sM = state_in->M; p.sw = state_in->w; p.sread = state_in->read; p.sctrl = state_in->controller_state;
  p.ssM = state_in->stride_M; p.ssw = state_in->stride_w; p.ssread = state_in->stride_read;
  p.ssctrl = state_in->stride_controller_state;
  p.dM = state_out->M; p.dw = state_out->w; p.dread = state_out->read; p.dctrl = state_out->controller_state;
  p.dsM = state_out->stride_M; p.dsw = state_out->stride_w; p.dsread = state_out->stride_read;
  p.dsctrl = state_out->stride_controller_state;
  p.logits = logits; p.outputs = outputs; p.dbg = debug_taps; p.dbgStride = hp.debug_floats;
  p.cst = reinterpret_cast<float*>(wsb + ws.off_cst);
  p.partA = reinterpret_cast<float*>(wsb + ws.off_partA);
  p.partC = reinterpret_cast<float*>(wsb + ws.off_partC);
  p.ctr = reinterpret_cast<unsigned*>(wsb + ws.off_ctr);
  p.err = reinterpret_cast<int*>(wsb + ws.off_err);
  p.prof = prof ? reinterpret_cast<long long*>(wsb + ws.off_prof) : nullptr;
  p.oMs = hp.oMs; p.oW0 = hp.oW0; p.oW1 = hp.oW1; p.oCn = hp.oCn; p.oX0 = hp.oX0; p.oX1 = hp.oX1;
  p.oScr = hp.oScr; p.oSim = hp.oSim; p.oWg = hp.oWg; p.oK = hp.oK; p.oE = hp.oE; p.oA = hp.oA;
  p.oSm = hp.oSm; p.oLog = hp.oLog; p.oTc = hp.oTc;

  e = cudaMemsetAsync(wsb + ws.off_ctr, 0, 512 + 8 * 16 * 1024, stream);   // barrier counter, error flag, phase counters
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");

  if (prof) cudaEventRecord(g_ev[1], stream);
  cfg.gridDim = dim3(ncta);
  // cluster + cooperative (co-residency enforced by the driver).  NTM_B200_NO_COOP=1 drops the
  // cooperative attribute (the grid is sized from the occupancy query, so the CTAs are still
  // co-resident); needed under Nsight Compute, whose kernel replay rejects cooperative+cluster launches.
  cfg.numAttrs = getenv("NTM_B200_NO_COOP") ? 1 : 2;
  e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) {
    // some driver/toolkit combinations reject cooperative+cluster; the grid is
    // sized from the occupancy query, so co-residency still holds.
    cudaGetLastError();
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
  }
  g_launches++;
  if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchKernelEx(ntm_seq_kernel)");
  if (prof) {
    cudaEventRecord(g_ev[2], stream);
    g_ev_valid = true;
  }
  return NTM_B200_OK;
}

int32_t ntm_b200_step(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                      const void* packed, int64_t batch, const float* inputs,
                      const ntm_b200_state* state_in, const ntm_b200_state* state_out,
                      float* logits, float* outputs, float* debug_taps, void* workspace,
                      int64_t workspace_bytes, void* stream) {
  return ntm_b200_forward_seq(shape, weights, packed, batch, 1, inputs, state_in, state_out, logits,
                              outputs, debug_taps, workspace, workspace_bytes, stream);
}

int32_t ntm_b200_last_launch_info(int32_t* out8) {
  if (!out8) return NTM_B200_ERR_NULL_POINTER;
  for (int i = 0; i < 8; ++i) out8[i] = g_last_info[i];
  return NTM_B200_OK;
}

int32_t ntm_b200_set_profiling(int32_t enable) {
  g_profiling.store(enable ? 1 : 0);
  return NTM_B200_OK;
}

int32_t ntm_b200_last_kernel_ms(float* xproj_ms, float* seq_kernel_ms) {
  if (!xproj_ms || !seq_kernel_ms) return NTM_B200_ERR_NULL_POINTER;
  if (!g_ev_valid) return NTM_B200_ERR_BAD_SHAPE;
  cudaError_t e = cudaEventElapsedTime(xproj_ms, g_ev[0], g_ev[1]);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaEventElapsedTime");
  e = cudaEventElapsedTime(seq_kernel_ms, g_ev[1], g_ev[2]);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaEventElapsedTime");
  return NTM_B200_OK;
}

int32_t ntm_b200_phase_cycles(const void* workspace, int64_t* out, int32_t max_ctas) {
  if (!workspace || !out) return NTM_B200_ERR_NULL_POINTER;
  if (max_ctas < 1 || max_ctas > 1024) return NTM_B200_ERR_BAD_SHAPE;
  cudaError_t e = cudaMemcpy(out, static_cast<const char*>(workspace) + 512, 8ll * 16 * max_ctas,
                             cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemcpy(phase cycles)");
  return NTM_B200_OK;
}

int32_t ntm_b200_finish(void* workspace, void* stream) {
  if (!workspace) return NTM_B200_ERR_NULL_POINTER;
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamSynchronize");
  int flag = 0;
  e = cudaMemcpy(&flag, static_cast<char*>(workspace) + 256, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemcpy(err flag)");
  return flag ? NTM_B200_ERR_DEVICE_TIMEOUT : NTM_B200_OK;
}

}  // extern "C"
