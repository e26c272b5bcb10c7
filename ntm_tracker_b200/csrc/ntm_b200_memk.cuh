// ntm_b200_memk.cuh -- the TMA-ring memory kernel of the streaming mode (mem_step_tma_kernel) with its launch
// templates, shared by the translation units that instantiate it: ntm_b200_stream.cu (1-2 read heads) and
// ntm_b200_memk_r34.cu (3-4 read heads) -- one TU with all 168 variants was 2.5 minutes of a clean build.
// See ntm_b200_stream.cu / DESIGN.md s4.3 for what the kernel does.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <type_traits>

#include "ntm_b200.h"
#include "ntm_b200_gemm_ws.cuh"
#include "ntm_b200_params.h"

namespace ntm_b200 {
namespace memk {

__device__ __forceinline__ float exp_f(float x) { return exp2f(x * 1.4426950408889634f); }
__device__ __forceinline__ float sigmoid_f(float x) { return __frcp_rn(1.0f + exp_f(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - 2.0f * __frcp_rn(1.0f + exp_f(2.0f * x)); }
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.0f) + __logf(1.0f + exp_f(-fabsf(x))); }
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

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
inline long long align_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

// (Measured negative result: rewriting the two passes with the packed fp32x2 FMA of sm_100, __ffma2_rn, made
// the kernel slower -- 460 -> 543 us per step of 4096 sequences -- so the FMAs below stay scalar.)

struct MemArgs {
  int N, M, M4, MC, Npad, S, shift0, P, PO4, O, write_first, T, t;
  int nslab; long long slab;                 // raw[q] = bias[q] + sum_s mc[s*slab + b*PO4 + q]
  const float* mc; const float* bias;
  const float* Min; long long sMin;          // memory entering the step [B][N][M]
  float* Mout; long long sMout;              // memory leaving it (may be the same buffer)
  const float* w_in; long long sw_in;        // weightings entering the step [B][H][N]
  float* w_out; long long sw_out;
  float* cn;                                 // [B][M4] inverse column norms, in: of Min, out: of Mout
  float* act_read; long long s_act;          // read vectors -> next step's controller input rows
  float* read_out; long long s_read;         // state / history copy of the read vectors (may be null)
  float* logits; float* outputs;             // [B][T][O]
  int oK, oE, oA, oSim, oWg, oWn, oSm, oX;   // shared-memory carve-up (floats)
  int WPC;                                   // warps sharing one 8-chunk column group in pass 2
  // TMA-ring kernel: stages of RPS memory rows, NS stages, NCH chunks per pass
  int RPS, NS, NCH, RP, qps_shift;           // RP: quads of 4 rows per pass-2 iteration (threads = RP * MC) = two stages;
                                             // qps_shift: log2(stage uses per sequence) when that is a power of two, else -1
  int NR;                                    // stages of pass 1 that stay in the ring for pass 2 (NS, or 0 = none)
  int exp;                                   // experiment switches (EnvSwitches::exp)
  int P2S;                                   // pass 2: stage pairs (iterations) per CTA barrier / re-issue round, 1 or 2
  int rot;                                   // compact shared-memory plan (large N): three [H][N] buffers rotate between
                                             // the roles w_prev / gated / final weighting, quad-slot partials live in the
                                             // similarity buffer; goes with the 4-stage ring (template flag NS4)
  unsigned qps_magic;                        // ceil(2^32 / uses per sequence): division by multiplication (0 = divide)
  int oWp, oRaw, oCn, oBar, oRing;           // w_prev copy, raw parameter row, column norms, mbarriers, ring (floats)
  int vec_out;                               // read-vector rows are 16-byte aligned
  uint8_t* tilesA; int KAtotA;               // controller-GEMM operand tiles (read vectors at k = r*M + d), or null
  long long B;
  long long* prof;                           // [B][8] phase timestamps (globaltimer ns) of the last launch, or null
  float* sim_hist; float* cn_hist;           // training history of this step: [B][H][N] un-normalised similarities, [B][M] cn
};

// ------------------------------------------------------------------------------------------------
// TMA-ring variant of the memory kernel (the fast path: M <= 512, H * ceil(M / 128) <= 20, N a multiple of
// the rows of a pass-2 iteration).  Same arithmetic as mem_step_kernel; what changes is how the memory
// moves.  The sequence's N x M rows are contiguous in HBM, so they stream through a ring of NS = 8
// shared-memory stages (RPS rows = 8 KiB each) with 1-D bulk copies (cp.async.bulk + mbarrier
// complete_tx).  CTAs are persistent (2 per SM) and the ring never drains: stage uses are numbered over
// all the sequences of a CTA, whoever releases use Q issues the load of use Q + NS, so the tail of pass 1
// prefetches the head of pass 2 (out of L2) behind the addressing phase and the tail of pass 2 prefetches
// the next sequence's pass 1 (from HBM) -- together with its head parameters, weightings and column norms
// -- behind finalize / activations.  Pass 1 is consumed by four two-warp teams that own two slots each
// (keys in registers); pass 2 reads the ring and stores M' straight to HBM.  Bytes in flight per CTA = the
// ring, independent of registers and occupancy -- which is what an HBM-latency-bound stream needs.
constexpr int TMA_NT = 256;
constexpr int TMA_NS = 8;     // ring stages (one stage = half the rows of a pass-2 iteration, 8 KiB at M*RP = 1024)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init_(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = s_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol, bool hint) {
  if (hint)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_global_hint(float* p, const float4 v, uint64_t pol, bool hint) {
  if (hint)
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
  else
    __stcg(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define MEM_PROF(slot) do { if (a.prof != nullptr && tid == 0) a.prof[(size_t)b * 16 + (slot)] = gtimer(); } while (0)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}

// FULLM: M == 128 * CPL exactly (the tracker and large-memory shapes).  Row stride, stage size and the thread
// mapping of pass 2 are then compile-time constants: the inner loops lose their address arithmetic and their
// column-bound checks (the kernel is issue-bound; a third of its instructions were integer / control).
// N128: N == 128 and S <= 7 (every BASELINE tracker shape): the addressing phase keeps a head's weighting in registers.
// A template parameter rather than a run-time branch: with both addressing variants in one kernel the code grew from
// 83 KB to 113 KB and the once-per-sequence phases of the OTHER shapes paid for it in instruction-cache misses
// (C4, N = 1024: addressing 12 -> 25 us per sequence).
// NS4: a ring of 4 stages (and pass-1 teams of four warps) instead of 8 -- with MemArgs::rot the kernel then fits two
// CTAs per SM at N*M*4 = 1 MiB (C4: 110 KB per CTA instead of 174 KB), where a second CTA covers the 12 us addressing
// phase of the first and 512 sequences take two rounds of 296 CTAs instead of four of 148.
template <int R, int W, int CPL, bool FULLM, bool N128, bool NS4>
__global__ void __launch_bounds__(TMA_NT, 2) mem_step_tma_kernel(const MemArgs a) {
  constexpr int H = R + W, NT = TMA_NT, NWARP = NT / 32, NS = NS4 ? 4 : TMA_NS;
  extern __shared__ float4 mem_smem4[];
  float* smem = reinterpret_cast<float*>(mem_smem4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, Npad = a.Npad, S = a.S;   // here N % 4 == 0: Npad == N
  const int M = FULLM ? 128 * CPL : a.M, M4 = M, MC = FULLM ? 32 * CPL : a.MC;
  const int RPS = FULLM ? 2 * (NT / (32 * CPL)) : a.RPS, NCH = a.NCH;   // RPS: power of two >= 4 that divides N
  float* kS = smem + a.oK;      // [H][M4]; after pass 1: partials of the quad slots rp >= 1
  float* eS = smem + a.oE;
  float* aS = smem + a.oA;
  float* simS = smem + a.oSim;
  float* wg = smem + a.oWg;          // (with MemArgs::rot these three change roles from sequence to sequence)
  float* wnew = smem + a.oWn;
  float* wprevS = smem + a.oWp;
  float* raw = smem + a.oRaw;
  float* cnS = smem + a.oCn;
  float* sm = smem + a.oSm;
  float* ring = smem + a.oRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.oBar);   // full[NS], parbar
  uint64_t* parbar = bars + NS;
  float* sBeta = sm, *sG = sm + H, *sGam = sm + 2 * H, *sSw = sm + 4 * H;
  float* sPart = sm + 4 * H + H * SMAX;
  float* sRed = sPart + NWARP * H;
  const int stage_floats = RPS * M;
  const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;
  // L2 eviction hints (measured on the final kernel, C3, same build otherwise): none 494 us per launch and 1.75 GB
  // of DRAM reads; evict-first on pass-2 loads / stores / parameters only 431 us, 1.31 GB; plus evict-last on the
  // pass-1 loads (below) 422 us, 1.43 GB.  -DNTM_NO_L2_HINTS rebuilds the first variant.
#ifdef NTM_NO_L2_HINTS
  constexpr bool hint = false;
#else
  constexpr bool hint = true;
#endif
  // pass 1 brings the rows in and wants them to survive in L2 until pass 2 re-reads them; after that
  // re-read, and for the rewritten rows, the next use is a whole timestep (the other sequences) away
  // (round 2, after a quarter of the re-read moved into the ring: the re-read stages are best loaded with the
  // NORMAL policy -- 409-413 us per launch against 415-416 with evict-last, 427 with evict-last on half the lines;
  // NTM_B200_EXP bit 256 brings evict-last back)
  uint64_t pol_keep;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(pol_keep));
  const uint64_t pol_drop = l2_policy_evict_first();
  if (a.exp & 256) pol_keep = l2_policy_evict_last();

  // Persistent CTA: sequences b = blockIdx.x + si * gridDim.x.  The ring never drains between sequences:
  // stage use Qg (global over this CTA's sequences) lives in slot Qg % NS with mbarrier parity (Qg / NS) & 1,
  // and whoever releases use Qg issues the load of use Qg + NS into the same slot -- so the head of the next
  // sequence streams in behind pass 2.  Per sequence: NCH pass-1 uses (row stages 0 .. NCH-1, from HBM), then
  // the pass-2 uses.  The last NR (= NS) stages of pass 1 are NOT released: pass 2 starts with them where they
  // lie (rows of the stages NCH-NR .. NCH-1) and goes on with the row stages 0 .. NCH-NR-1, re-read through L2
  // -- a quarter of the re-read traffic and of the lines that must survive in L2 gone at N*M*4 = 256 KiB.  So
  // pass-2 position p (0 .. NCH-1) is use Qb + NCH - NR + p, and a sequence takes QPS = 2*NCH - NR uses.
  const int G = gridDim.x;
  const int nseq = ((int)a.B - (int)blockIdx.x + G - 1) / G;
  const int NR = a.NR;
  const int QPS = 2 * NCH - NR, QT = nseq * QPS;
  // (the division runs on the issuing lane only, ~28 times per sequence; carrying (sequence, use) pairs
  // instead costs registers this kernel does not have: it spilled and ran 11 % slower)
  auto issue_load = [&](int Qg) {
    const int si = a.qps_shift >= 0 ? (Qg >> a.qps_shift)
                                    : (a.qps_magic != 0u ? (int)__umulhi((unsigned)Qg, a.qps_magic) : Qg / QPS);
    const int Q = Qg - si * QPS;
    const int j = Q < NCH ? Q : Q - NCH;      // row stage: pass 1 in order; pass-2 loads start over at stage 0
    const float* src = a.Min + (size_t)(blockIdx.x + si * G) * a.sMin + (size_t)j * stage_floats;
    uint64_t* fb = bars + (Qg & (NS - 1));
    mbar_expect_tx(fb, stage_bytes);
    // only the pass-1 stages that pass 2 re-reads through L2 are worth keeping there
    bulk_load(ring + (Qg & (NS - 1)) * stage_floats, src, stage_bytes, fb, Q < NCH - NR ? pol_keep : pol_drop, hint);
  };
  // head parameters, entering weightings and inverse column norms of sequence si -> shared memory
  const uint32_t par_bytes = (uint32_t)(a.PO4 + H * N + M4) * 4u;
  auto issue_params = [&](int si, float* wdst) {
    const size_t bb = (size_t)(blockIdx.x + si * G);
    mbar_expect_tx(parbar, par_bytes);
    bulk_load(raw, a.mc + bb * a.PO4, (uint32_t)a.PO4 * 4u, parbar, pol_drop, hint);
    bulk_load(wdst, a.w_in + bb * a.sw_in, (uint32_t)(H * N) * 4u, parbar, pol_drop, hint);
    bulk_load(cnS, a.cn + bb * M4, (uint32_t)M4 * 4u, parbar, pol_drop, hint);
  };

  if (tid == 0) {
    for (int s2 = 0; s2 <= NS; ++s2) mbar_init_(bars + s2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();            // barrier inits visible to every thread before anyone polls
  pdl_trigger();
  pdl_wait();                 // head parameters, memories, weightings: all written by earlier kernels of the chain
  if (tid == 0 && nseq > 0) {
    issue_params(0, wprevS);
    for (int Qg = 0; Qg < NS && Qg < QT; ++Qg) issue_load(Qg);
  }

  const int offBeta = H * M, offG = offBeta + H, offS = offG + H, offGam = offS + S * H,
            offE = offGam + H, offA = offE + M * W;
  const int RP = FULLM ? NT / (32 * CPL) : a.RP;   // pass 2: RP quads of 4 rows per iteration = two stages
  const bool worker = tid < RP * MC;
  const int rp = worker ? tid / MC : 0, c = worker ? tid - rp * MC : 0;

  for (int si = 0; si < nseq; ++si) {
    const int b = blockIdx.x + si * G;
    const int Qb = si * QPS;
    float* Mo = a.Mout + (size_t)b * a.sMout;
    if (a.rot) {      // sequence si: w_prev arrived in X (even) / Z (odd); gated weights go to the other, the final
      float* bufX = smem + a.oWp, *bufZ = smem + a.oWg;   // weighting overwrites w_prev head by head
      wprevS = (si & 1) ? bufZ : bufX;
      wg = (si & 1) ? bufX : bufZ;
      wnew = wprevS;
    }
    MEM_PROF(0);
    mbar_wait_(parbar, (uint32_t)si & 1u);
    MEM_PROF(1);

    // ---- activations (ntm_cell.py:133-196) ----
    {
      float ss[H];
#pragma unroll
      for (int h = 0; h < H; ++h) ss[h] = 0.0f;
      for (int d = tid; d < M; d += NT) {
        const float cnd = cnS[d];
        if (a.cn_hist != nullptr) a.cn_hist[(size_t)b * M + d] = cnd;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float kv = tanh_f(raw[h * M + d]);
          kS[h * M4 + d] = kv * cnd;
          ss[h] = fmaf(kv, kv, ss[h]);
        }
#pragma unroll
        for (int h = 0; h < W; ++h) {
          eS[h * M4 + d] = sigmoid_f(raw[offE + h * M + d]);
          aS[h * M4 + d] = tanh_f(raw[offA + h * M + d]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int h = 0; h < H; ++h) ss[h] += __shfl_xor_sync(0xffffffffu, ss[h], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < H; ++h) sPart[warp * H + h] = ss[h];
      }
    }
    if (warp < 4 && lane < H) {   // per-head scalars (ntm_cell.py:140,151,161,169): one WARP per quantity (divergent
      const int h = lane, job = warp;  // branches inside a warp would run one after the other), one lane per head
      if (job == 0) sBeta[h] = softplus_f(raw[offBeta + h]);
      else if (job == 1) sG[h] = sigmoid_f(raw[offG + h]);
      else if (job == 2) sGam[h] = 1.0f + softplus_f(raw[offGam + h]);
      else {
        float* sp = sSw + h * SMAX;
        float mx = -INFINITY;
        for (int i = 0; i < S; ++i) { sp[i] = raw[offS + h * S + i]; mx = fmaxf(mx, sp[i]); }
        float sum = 0.0f;
        for (int i = 0; i < S; ++i) { sp[i] = exp_f(sp[i] - mx); sum += sp[i]; }
        const float rsum = __frcp_rn(sum);
        for (int i = 0; i < S; ++i) sp[i] = sp[i] * rsum;
      }
    }
    if (tid == NT - 1) {   // output projection + softmax (ntm_cell.py:220-221)
      const size_t o = ((size_t)b * a.T + a.t) * a.O;
      float mx = -INFINITY;
      for (int i = 0; i < a.O; ++i) mx = fmaxf(mx, raw[a.P + i]);
      float sum = 0.0f;
      for (int i = 0; i < a.O; ++i) sum += exp_f(raw[a.P + i] - mx);
      const float rsum = __frcp_rn(sum);
      for (int i = 0; i < a.O; ++i) {
        const float lg = raw[a.P + i];
        a.logits[o + i] = lg;
        if (a.outputs) a.outputs[o + i] = exp_f(lg - mx) * rsum;
      }
    }
    __syncthreads();
    MEM_PROF(2);

    // ---- pass 1: a warp owns stage q = warp, warp + NWARP, ...; keys in registers; four rows at a time ----
    {
      float4 kr[H][CPL];
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int cc = lane + 32 * j;
          kr[h][j] = cc < MC ? *reinterpret_cast<const float4*>(kS + h * M4 + 4 * cc) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      // NTEAM teams of two warps; team t consumes the stages q = t, t + NTEAM, ... which alternate between
      // the slots t and t + NTEAM of the ring, so the team's next stage is already in flight while it works
      // on the current one (a warp only ever waits on the use that directly follows the one it consumed in
      // that slot -- an mbarrier cannot be waited on two phases ahead).  Within a quad of four rows the two
      // members take two rows each.
      constexpr int NTEAM = NS / 2, WPT = NWARP / NTEAM;      // 4 teams of 2 warps, or (NS4) 2 teams of 4
      static_assert(WPT == 2 || WPT == 4, "two or four warps per team");
      constexpr int RBT = 2 * WPT;                            // rows a team takes per sweep step, two per member
      const int team = warp % NTEAM, member = warp / NTEAM;
      const int rv = (lane & 15) / H, hv = (lane & 15) - rv * H;      // reduced value index -> (row of the pair, head)
      for (int q = team; q < NCH; q += NTEAM) {
        const int Qg = Qb + q;
        mbar_wait_(bars + (Qg & (NS - 1)), (uint32_t)(Qg / NS) & 1u);
        const float* sp = ring + (Qg & (NS - 1)) * stage_floats + 4 * lane + 2 * member * M;
        float* simq = simS + hv * Npad + q * RPS + 2 * member + rv;
        for (int g0 = 0; g0 < RPS; g0 += RBT, sp += RBT * M) {
          float acc[2][H];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int h = 0; h < H; ++h) acc[i][h] = 0.0f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            if (lane + 32 * j < MC) {
              float4 m4[2];
#pragma unroll
              for (int i = 0; i < 2; ++i) m4[i] = *reinterpret_cast<const float4*>(sp + i * M + 128 * j);
#pragma unroll
              for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int h = 0; h < H; ++h) {
                  acc[i][h] = fmaf(m4[i].x, kr[h][j].x, acc[i][h]);
                  acc[i][h] = fmaf(m4[i].y, kr[h][j].y, acc[i][h]);
                  acc[i][h] = fmaf(m4[i].z, kr[h][j].z, acc[i][h]);
                  acc[i][h] = fmaf(m4[i].w, kr[h][j].w, acc[i][h]);
                }
            }
          }
          // reduction of the 2*H <= 16 values over the warp: one butterfly step over the half-warps, then
          // a transposing reduction over 16 lanes (at offset o a lane keeps one half of its value list and
          // receives the partner's sums of that half): lane L ends with the total of value L & 15
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.0f;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int h = 0; h < H; ++h) v[i * H + h] = acc[i][h];
#pragma unroll
          for (int j = 0; j < 2 * H; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < o; ++j) {
              const float send = up ? v[j] : v[j + o];
              const float keep = up ? v[j + o] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          if (lane < 2 * H) simq[g0] = v[0];
        }
        // the team were the stage's only readers: once both members are done, refill the slot with use Qg + NS
        // (the last NR stages stay where they are for pass 2)
        if (q < NCH - NR) {
          asm volatile("bar.sync %0, %1;" ::"r"(8 + team), "r"(32 * WPT) : "memory");
          // (rotating the issuing member over the team's warps changes nothing: measured)
          if (member == 0 && lane == 0 && Qg + NS < QT) issue_load(Qg + NS);
        }
      }
    }
    __syncthreads();
    MEM_PROF(3);
    if (a.sim_hist != nullptr)     // training history: un-normalised similarities (here Npad == N)
      for (int i = tid; i < H * N; i += NT) a.sim_hist[(size_t)b * H * N + i] = simS[i];

    // ---- addressing (ntm_cell.py:140-176): one warp (WPH warps when there are spares) per head.  Every
    //      sweep over the head's N entries handles four entries per lane at a time -- loads, then math,
    //      then stores -- so the four dependent chains (shared-memory load -> SFU -> store) overlap ----
    if constexpr (N128) {
      // N = 128 (every BASELINE tracker shape): one warp per head, the head's 128 entries live in REGISTERS -- lane L
      // holds n = 4L .. 4L+3 -- from the similarity to the final weighting: no shared-memory round trip and no
      // barrier between the five sweeps of the general path below, the circular shift takes the neighbour lanes'
      // entries by shuffle (3.4 -> ~2 us of a CTA's 29 us per sequence; nothing streams during this phase)
      if (warp < H) {
        const int h = warp;
        const float gate = sG[h], gamma = sGam[h];
        float kn = 0.0f;
#pragma unroll
        for (int w2 = 0; w2 < NWARP; ++w2) kn += sPart[w2 * H + h];
        const float scale = sBeta[h] / sqrtf(fmaxf(kn, 1e-12f));   // beta / |k|  (ops.py:152, ntm_cell.py:142)
        const float4 s4 = *reinterpret_cast<const float4*>(simS + h * Npad + 4 * lane);
        const float4 p4 = *reinterpret_cast<const float4*>(wprevS + h * N + 4 * lane);
        float x[4] = {s4.x * scale, s4.y * scale, s4.z * scale, s4.w * scale};
        const float mx = warp_max(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
        float e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = exp_f(x[u] - mx);
        const float sum = warp_sum((e[0] + e[1]) + (e[2] + e[3]));
        const float gs = gate / sum, g1 = 1.0f - gate;          // w_g = g * softmax + (1 - g) * w_prev
        float win[12];                                           // gated weights of lanes L-1, L, L+1 (circular)
        win[4] = fmaf(e[0], gs, p4.x * g1); win[5] = fmaf(e[1], gs, p4.y * g1);
        win[6] = fmaf(e[2], gs, p4.z * g1); win[7] = fmaf(e[3], gs, p4.w * g1);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          win[u] = __shfl_sync(0xffffffffu, win[4 + u], (lane + 31) & 31);
          win[8 + u] = __shfl_sync(0xffffffffu, win[4 + u], (lane + 1) & 31);
        }
        float pw[4];
        auto shift_pow = [&](auto s_tag) {       // circular_shift(x, j)[n] = x[(n + j) mod N], taps j = shift0 .. shift0+S-1
          constexpr int SS = decltype(s_tag)::value, SH0 = -((SS + 1) / 2);
          float swv[SS];
#pragma unroll
          for (int s2 = 0; s2 < SS; ++s2) swv[s2] = sSw[h * SMAX + s2];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float conv = 0.0f;
#pragma unroll
            for (int s2 = 0; s2 < SS; ++s2) conv = fmaf(swv[s2], win[4 + u + SH0 + s2], conv);
            pw[u] = exp2f(gamma * __log2f(conv));   // conv >= 0, gamma >= 1: pow on the SFU; 0 -> 0
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
        *reinterpret_cast<float4*>(a.w_out + (size_t)b * a.sw_out + h * N + 4 * lane) = wv;
      }
    } else {
      constexpr int WPH = (NWARP / H) > 0 ? (NWARP / H) : 1;
      constexpr bool multi = (NWARP / H) > 0;
      constexpr int U = 4;
      const int hgrp = warp / WPH, sub = warp - hgrp * WPH;
      const int hstep = multi ? H : NWARP;
      float* wout = a.w_out + (size_t)b * a.sw_out;
      for (int h = multi ? hgrp : warp; h < H; h += hstep) {
        float* sh = simS + h * Npad;
        float* gh = wg + h * Npad;
        const float* wp = wprevS + h * N;
        float* red = sRed + h * 3 * WPH;
        const int nstep = 32 * WPH;
        const int n0 = 32 * sub + lane;
        auto head_bar = [&]() {
          if (WPH > 1) asm volatile("bar.sync %0, %1;" ::"r"(h + 1), "r"(32 * WPH) : "memory");
          else __syncwarp();
        };
        const float gate = sG[h], gamma = sGam[h];
        float kn = 0.0f;
#pragma unroll
        for (int w2 = 0; w2 < NWARP; ++w2) kn += sPart[w2 * H + h];
        const float scale = sBeta[h] / sqrtf(fmaxf(kn, 1e-12f));   // beta / |k|  (ops.py:152, ntm_cell.py:142)
        float mx = -INFINITY;
        for (int nb = n0; nb < N; nb += U * nstep) {
          float x[U];
#pragma unroll
          for (int u = 0; u < U; ++u) x[u] = (nb + u * nstep < N) ? sh[nb + u * nstep] * scale : -INFINITY;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (nb + u * nstep < N) sh[nb + u * nstep] = x[u];
            mx = fmaxf(mx, x[u]);
          }
        }
        mx = warp_max(mx);
        if (WPH > 1) {
          if (lane == 0) red[sub] = mx;
          head_bar();
          mx = red[0];
#pragma unroll
          for (int i = 1; i < WPH; ++i) mx = fmaxf(mx, red[i]);
        }
        float sum = 0.0f;
        for (int nb = n0; nb < N; nb += U * nstep) {
          float e[U];
#pragma unroll
          for (int u = 0; u < U; ++u) e[u] = (nb + u * nstep < N) ? exp_f(sh[nb + u * nstep] - mx) : 0.0f;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (nb + u * nstep < N) sh[nb + u * nstep] = e[u];
            sum += e[u];
          }
        }
        sum = warp_sum(sum);
        if (WPH > 1) {
          if (lane == 0) red[WPH + sub] = sum;
          head_bar();
          sum = red[WPH];
#pragma unroll
          for (int i = 1; i < WPH; ++i) sum += red[WPH + i];
        }
        const float gs = gate / sum, g1 = 1.0f - gate;          // w_g = g * softmax + (1 - g) * w_prev
        for (int nb = n0; nb < N; nb += U * nstep) {
          float g[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
            g[u] = (nb + u * nstep < N) ? fmaf(sh[nb + u * nstep], gs, wp[nb + u * nstep] * g1) : 0.0f;
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (nb + u * nstep < N) gh[nb + u * nstep] = g[u];
        }
        head_bar();   // the shift reads neighbours' gated weights
        float psum = 0.0f;
        for (int nb = n0; nb < N; nb += U * nstep) {
          float pw[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int n = nb + u * nstep;
            float conv = 0.0f;
            if (n < N) {
              for (int s2 = 0; s2 < S; ++s2) {
                int idx = n + a.shift0 + s2;   // circular_shift(x, j)[n] = x[(n + j) mod N], ops.py:216-242
                idx = idx < 0 ? idx + N : (idx >= N ? idx - N : idx);
                conv = fmaf(sSw[h * SMAX + s2], gh[idx], conv);
              }
            }
            // conv >= 0, gamma >= 1: pow(conv, gamma) on the SFU (lg2 + ex2, ~1e-6 relative); 0 -> 0
            pw[u] = (n < N) ? exp2f(gamma * __log2f(conv)) : 0.0f;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (nb + u * nstep < N) sh[nb + u * nstep] = pw[u];
            psum += pw[u];
          }
        }
        psum = warp_sum(psum);
        if (WPH > 1) {
          if (lane == 0) red[2 * WPH + sub] = psum;
          head_bar();
          psum = red[2 * WPH];
#pragma unroll
          for (int i = 1; i < WPH; ++i) psum += red[2 * WPH + i];
        }
        const float rden = 1.0f / (psum + 1e-3f);   // ntm_cell.py:175-176
        for (int nb = n0; nb < N; nb += U * nstep) {
          float wv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) wv[u] = (nb + u * nstep < N) ? sh[nb + u * nstep] * rden : 0.0f;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (nb + u * nstep < N) {
              wnew[h * Npad + nb + u * nstep] = wv[u];
              wout[h * N + nb + u * nstep] = wv[u];
            }
          }
        }
      }
    }
    __syncthreads();
    MEM_PROF(4);
    // raw / wprevS / cnS are dead now: the next sequence's parameters stream in behind pass 2
    // (rot: the next sequence's weightings land in the gated-weights buffer, dead since the barrier above; the
    // entering weightings' own buffer holds the final weighting now and is read all through pass 2)
    if (tid == 0 && si + 1 < nseq) issue_params(si + 1, a.rot ? wg : wprevS);

    // ---- pass 2: thread -> (16-byte column chunk c, quad slot rp); per iteration the CTA reads RP quads
    //      of four consecutive rows (= two whole stages) from the ring and writes M' straight to HBM ----
    float4 racc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) racc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 csq = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      float4 e4[W], a4[W];
#pragma unroll
      for (int h = 0; h < W; ++h) {
        e4[h] = *reinterpret_cast<const float4*>(eS + h * M4 + 4 * c);
        a4[h] = *reinterpret_cast<const float4*>(aS + h * M4 + 4 * c);
      }
      const int RS = 4 * RP;                       // rows per iteration = two stages
      const int qps = RP >> 1;                     // quads per stage
      const int sk = rp >= qps ? 1 : 0;            // which of the iteration's two stages this thread's quad is in
      const int NIT = N / RS;                      // iterations = NCH / 2
      // P2S iterations (stage pairs) share one CTA barrier and one round of re-issues: every bulk copy costs its
      // issuing warp ~240 ns right behind the barrier, i.e. on the whole CTA's critical path
      const int P2S = a.P2S;
      for (int it0 = 0; it0 < NIT; it0 += P2S) {
       for (int it = it0; it < it0 + P2S; ++it) {
        // position p = 2 * it of pass 2: the stages retained from pass 1 first (rows from stage NCH - NR on), then
        // the re-read stages 0 .. NCH-NR-1
        const int nb = (2 * it < NR ? (NCH - NR + 2 * it) : (2 * it - NR)) * RPS;
        const int n0 = nb + 4 * rp;                // this thread's quad
        if (worker && n0 < N) {
          const int Qg = Qb + NCH - NR + 2 * it + sk;
          mbar_wait_(bars + (Qg & (NS - 1)), (uint32_t)(Qg / NS) & 1u);
          const float* mp = ring + (Qg & (NS - 1)) * stage_floats + 4 * (rp - sk * qps) * M + 4 * c;
          float* gp = Mo + (size_t)n0 * M + 4 * c;
          float4 wv[H];                            // the four rows' weights of every head
#pragma unroll
          for (int h = 0; h < H; ++h) wv[h] = *reinterpret_cast<const float4*>(wnew + h * Npad + n0);
          float4 m[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) m[i] = *reinterpret_cast<const float4*>(mp + i * M);
          // (write_first is uniform: the branch is taken once per quad, outside the unrolled row loop, so
          // the four rows' FMA chains still interleave)
          auto quad = [&](auto wf_tag) {
            constexpr bool WF = decltype(wf_tag)::value;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              auto wsel = [&](int h) -> float { return i == 0 ? wv[h].x : (i == 1 ? wv[h].y : (i == 2 ? wv[h].z : wv[h].w)); };
              float4 mn;
              if constexpr (W == 1) {
                // one write head: M' = M (1 - w e) + w a = M + w (a - M e)
                const float ww = wsel(R);
                mn.x = fmaf(ww, fmaf(-m[i].x, e4[0].x, a4[0].x), m[i].x);
                mn.y = fmaf(ww, fmaf(-m[i].y, e4[0].y, a4[0].y), m[i].y);
                mn.z = fmaf(ww, fmaf(-m[i].z, e4[0].z, a4[0].z), m[i].z);
                mn.w = fmaf(ww, fmaf(-m[i].w, e4[0].w, a4[0].w), m[i].w);
              } else {
                float4 E = make_float4(1.f, 1.f, 1.f, 1.f), A = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int h = 0; h < W; ++h) {
                  const float ww = wsel(R + h);
                  E.x *= (1.0f - ww * e4[h].x); E.y *= (1.0f - ww * e4[h].y);
                  E.z *= (1.0f - ww * e4[h].z); E.w *= (1.0f - ww * e4[h].w);
                  A.x = fmaf(ww, a4[h].x, A.x); A.y = fmaf(ww, a4[h].y, A.y);
                  A.z = fmaf(ww, a4[h].z, A.z); A.w = fmaf(ww, a4[h].w, A.w);
                }
                mn.x = fmaf(m[i].x, E.x, A.x); mn.y = fmaf(m[i].y, E.y, A.y);
                mn.z = fmaf(m[i].z, E.z, A.z); mn.w = fmaf(m[i].w, E.w, A.w);
              }
              const float4 mu = WF ? mn : m[i];   // read from the updated memory when write_first (ntm_cell.py:212-215)
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const float wr = wsel(r);
                racc[r].x = fmaf(wr, mu.x, racc[r].x); racc[r].y = fmaf(wr, mu.y, racc[r].y);
                racc[r].z = fmaf(wr, mu.z, racc[r].z); racc[r].w = fmaf(wr, mu.w, racc[r].w);
              }
              csq.x = fmaf(mn.x, mn.x, csq.x); csq.y = fmaf(mn.y, mn.y, csq.y);
              csq.z = fmaf(mn.z, mn.z, csq.z); csq.w = fmaf(mn.w, mn.w, csq.w);
              st_global_hint(gp + (size_t)i * M, mn, pol_drop, hint);   // evict-first: M' is not needed for a whole step
            }
          };
          if (a.write_first) quad(std::true_type{}); else quad(std::false_type{});
        }
       }
        __syncthreads();         // every reader of these stages is done
        // one bulk copy costs its issuing thread ~240 ns (tools/tma_probe.cu): one issuer per freed slot, in
        // different warps, rotating over the warps so that no warp pays it twice in a row
        const int ik = (warp - 2 * it0) & (NWARP - 1);
        if ((tid & 31) == 0 && ik < 2 * P2S) {
          const int Qn = Qb + NCH - NR + 2 * it0 + ik + NS;
          if (Qn < QT) issue_load(Qn);
        }
      }
    }
    MEM_PROF(5);

    // ---- finalize: quad slots rp >= 1 park their partials in the (dead) key buffer; slot 0 adds them in
    //      fixed order and writes the read vector and the new inverse column norms ----
    {
      float* xch = a.rot ? simS : kS;     // [RP-1][R+1][M4] (rot: the similarity buffer, dead since the addressing)
      if (worker && rp > 0) {
        float* xw = xch + (rp - 1) * (R + 1) * M4 + 4 * c;
#pragma unroll
        for (int r = 0; r < R; ++r) *reinterpret_cast<float4*>(xw + r * M4) = racc[r];
        *reinterpret_cast<float4*>(xw + R * M4) = csq;
      }
      __syncthreads();
      MEM_PROF(6);
      if (worker && rp == 0) {
        for (int q = 1; q < RP; ++q) {
          const float* xr = xch + (q - 1) * (R + 1) * M4 + 4 * c;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 t = *reinterpret_cast<const float4*>(xr + r * M4);
            racc[r].x += t.x; racc[r].y += t.y; racc[r].z += t.z; racc[r].w += t.w;
          }
          const float4 t = *reinterpret_cast<const float4*>(xr + R * M4);
          csq.x += t.x; csq.y += t.y; csq.z += t.z; csq.w += t.w;
        }
        // fp32 rows for the fallback GEMM (not needed when the next GEMM reads the operand tiles) and for the state
        float* ar = a.tilesA == nullptr ? a.act_read + (size_t)b * a.s_act + 4 * c : nullptr;
        float* ro = a.read_out != nullptr ? a.read_out + (size_t)b * a.s_read + 4 * c : nullptr;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (a.vec_out) {
            if (ar) *reinterpret_cast<float4*>(ar + r * M) = racc[r];
            if (ro) *reinterpret_cast<float4*>(ro + r * M) = racc[r];
          } else {
            if (ar) { ar[r * M] = racc[r].x; ar[r * M + 1] = racc[r].y; ar[r * M + 2] = racc[r].z; ar[r * M + 3] = racc[r].w; }
            if (ro) { ro[r * M] = racc[r].x; ro[r * M + 1] = racc[r].y; ro[r * M + 2] = racc[r].z; ro[r * M + 3] = racc[r].w; }
          }
        }
      }
      if (a.tilesA != nullptr) {
        // next step's controller-GEMM operand: bf16 hi/lo of the read vectors, straight into the tile
        // records (lane pairs assemble 8 consecutive columns; every thread takes part in the shuffles)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float v[8];
          v[0] = racc[r].x; v[1] = racc[r].y; v[2] = racc[r].z; v[3] = racc[r].w;
          v[4] = __shfl_xor_sync(0xffffffffu, racc[r].x, 1); v[5] = __shfl_xor_sync(0xffffffffu, racc[r].y, 1);
          v[6] = __shfl_xor_sync(0xffffffffu, racc[r].z, 1); v[7] = __shfl_xor_sync(0xffffffffu, racc[r].w, 1);
          if (worker && rp == 0 && (c & 1) == 0) gemmws::store_split8(a.tilesA, a.KAtotA, b, r * M + 4 * c, v);
        }
      }
      if (worker && rp == 0) {
        float4 cn4;   // tf.nn.l2_normalize over N, ops.py:147-150
        cn4.x = 1.0f / sqrtf(fmaxf(csq.x, 1e-12f)); cn4.y = 1.0f / sqrtf(fmaxf(csq.y, 1e-12f));
        cn4.z = 1.0f / sqrtf(fmaxf(csq.z, 1e-12f)); cn4.w = 1.0f / sqrtf(fmaxf(csq.w, 1e-12f));
        *reinterpret_cast<float4*>(a.cn + (size_t)b * M4 + 4 * c) = cn4;
      }
      __syncthreads();           // kS is rewritten by the next sequence's activations
    }
    MEM_PROF(7);
  }
}


// what the caller's environment switches ask of a launch, and what the launch reports back
struct TmaCtl {
  int ctas_per_sm_cap;   // EnvSwitches::mem_ctas_per_sm (0 = occupancy)
  int grid_cap;          // EnvSwitches::mem_grid (0 = none)
  bool pdl;              // programmatic-dependent-launch attribute
  int occ;               // out: co-resident CTAs per SM of the variant launched
};

// ---- TMA-ring kernel dispatch ----
template <int R, int W, int CPL, bool FULLM, bool N128, bool NS4>
inline cudaError_t launch_tma_v2(const MemArgs& a, long long B, int smem, cudaStream_t stream, TmaCtl& ctl) {
  if constexpr ((R + W) * CPL > 20) {
    return cudaErrorInvalidValue;
  } else {
    // per instantiation AND device: configured size, co-resident CTAs per SM at that size, SM count
    static int configured[MAX_DEVICES] = {0}, occs[MAX_DEVICES] = {0}, sms[MAX_DEVICES] = {0};
    int occ = 1, nsm = B200_SMS;
    {
      std::lock_guard<std::mutex> lk(config_mutex());
      const int dev = current_device_slot();
      if (configured[dev] != smem) {
        cudaError_t e = cudaFuncSetAttribute(mem_step_tma_kernel<R, W, CPL, FULLM, N128, NS4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occs[dev], mem_step_tma_kernel<R, W, CPL, FULLM, N128, NS4>, TMA_NT, smem);
        if (occs[dev] < 1) occs[dev] = 1;
        int rdev = 0;
        sms[dev] = B200_SMS;
        if (cudaGetDevice(&rdev) == cudaSuccess) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, rdev);
      }
      occ = occs[dev]; nsm = sms[dev];
    }
    ctl.occ = occ;
    int per_sm = occ;
    if (ctl.ctas_per_sm_cap > 0) per_sm = std::max(1, std::min(per_sm, ctl.ctas_per_sm_cap));
    long long grid = std::min<long long>(B, (long long)per_sm * nsm);   // persistent CTAs
    if (ctl.grid_cap > 0) grid = std::min<long long>(grid, ctl.grid_cap);
    return launch_chain(mem_step_tma_kernel<R, W, CPL, FULLM, N128, NS4>, (unsigned)grid, TMA_NT, (size_t)smem, stream, ctl.pdl, a);
  }
}
template <int R, int W, int CPL, bool FULLM>
inline cudaError_t launch_tma_v(const MemArgs& a, long long B, int smem, cudaStream_t stream, TmaCtl& ctl) {
  const bool n128 = a.N == 128 && a.S <= 7 && (R + W) <= TMA_NT / 32 && (a.sw_out & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.w_out) & 15) == 0 && a.NS == TMA_NS;
  if constexpr (FULLM) {      // (instantiated for M = 128 / 256 / 512 only: every extra variant costs build time)
    if (n128) return launch_tma_v2<R, W, CPL, FULLM, true, false>(a, B, smem, stream, ctl);
  }
  if constexpr (CPL <= 2 && FULLM) {      // the 4-stage ring exists for M = 128 / 256 only (stages of >= 8 rows)
    if (a.NS == 4) return launch_tma_v2<R, W, CPL, FULLM, false, true>(a, B, smem, stream, ctl);
  }
  if (a.NS != TMA_NS) return cudaErrorInvalidValue;
  return launch_tma_v2<R, W, CPL, FULLM, false, false>(a, B, smem, stream, ctl);
}
template <int R, int W>
inline cudaError_t launch_tma_rw(int CPL, const MemArgs& a, long long B, int smem, cudaStream_t stream, TmaCtl& ctl) {
  switch (CPL) {
    case 1: return a.M == 128 ? launch_tma_v<R, W, 1, true>(a, B, smem, stream, ctl) : launch_tma_v<R, W, 1, false>(a, B, smem, stream, ctl);
    case 2: return a.M == 256 ? launch_tma_v<R, W, 2, true>(a, B, smem, stream, ctl) : launch_tma_v<R, W, 2, false>(a, B, smem, stream, ctl);
    case 4: return a.M == 512 ? launch_tma_v<R, W, 4, true>(a, B, smem, stream, ctl) : launch_tma_v<R, W, 4, false>(a, B, smem, stream, ctl);
  }
  return cudaErrorInvalidValue;
}
template <int R>
inline cudaError_t launch_tma_r(int W, int CPL, const MemArgs& a, long long B, int smem, cudaStream_t stream, TmaCtl& ctl) {
  switch (W) {
    case 1: return launch_tma_rw<R, 1>(CPL, a, B, smem, stream, ctl);
    case 2: return launch_tma_rw<R, 2>(CPL, a, B, smem, stream, ctl);
    case 3: return launch_tma_rw<R, 3>(CPL, a, B, smem, stream, ctl);
  }
  return cudaErrorInvalidValue;
}

// 3 and 4 read heads: instantiated in ntm_b200_memk_r34.cu
cudaError_t launch_tma_r34(int R, int W, int CPL, const MemArgs& a, long long B, int smem, cudaStream_t stream, TmaCtl& ctl);

}  // namespace memk
}  // namespace ntm_b200
