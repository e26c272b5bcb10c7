// ntm_b200_stream.cu -- streaming (throughput) mode of the NTM sequence path for batches far larger
// than what fits the SMs' shared memory.
//
// What it replaces (paths relative to the reference root), per timestep and for ALL sequences of the
// shard in lockstep:
//   controller projection + BasicLSTMCell ... ntm_cell.py:101-105     -> gemm_tc (tcgen05) + lstm_stream_kernel
//   _linear head parameters + logits ........ ntm_cell.py:113-130,220 -> gemm_tc (tcgen05)
//   activations, batched_smooth_cosine_similarity, content focus, gate, batched_circular_convolution,
//   sharpening, erase/add write, read ....... ntm_cell.py:133-215, ops.py:135-242 -> mem_step_kernel
//
// The HBM-bound kernel of this mode is the fused addressing / memory kernel, in two variants:
//   mem_step_tma_kernel  persistent CTAs, the sequence's rows stream through a shared-memory ring of bulk
//                        TMA copies (the fast path: M <= 512, N a multiple of the pass-2 iteration);
//   mem_step_kernel      one CTA per sequence-step, register-streamed ld.global (any M % 4 == 0).
// Both make two passes over the N x M memory -- pass 1 from HBM (similarities), the addressing in shared
// memory, pass 2 out of L2, where pass 1 just put the rows, writing M' back: one HBM read and one HBM
// write of M per sequence-step instead of the three passes of the algorithmic count.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "ntm_b200_stream.h"
#include "ntm_b200_gemm_ws.cuh"
#include "ntm_b200_memk.cuh"

namespace ntm_b200 {
namespace {

using namespace memk;      // device helpers, MemArgs, the TMA-ring kernel and its launch templates (ntm_b200_memk.cuh)

// generic memory kernel below: (RB = 4 rows per transposing reduction group comes from ntm_b200_params.h)
constexpr int RB1 = 8;       // rows a warp streams at once in pass 1 (two groups)
constexpr int U2 = 8;        // rows a thread keeps in flight in pass 2


// ------------------------------------------------------------------------------------------------
// One CTA = one sequence, one timestep.
template <int R, int W, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) mem_step_kernel(const MemArgs a) {
  pdl_trigger();
  pdl_wait();
  constexpr int H = R + W, NWARP = NT / 32;
  extern __shared__ float4 mem_smem4[];
  float* smem = reinterpret_cast<float*>(mem_smem4);
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, M = a.M, M4 = a.M4, MC = a.MC, Npad = a.Npad, S = a.S;
  float* kS = smem + a.oK;      // [H][M4]  tanh(k) * cn
  float* eS = smem + a.oE;      // [W][M4]
  float* aS = smem + a.oA;      // [W][M4]
  float* simS = smem + a.oSim;  // [H][Npad]
  float* wg = smem + a.oWg;     // [H][Npad]
  float* wnew = smem + a.oWn;   // [H][Npad]
  float* sm = smem + a.oSm;     // beta[H] g[H] gamma[H] pad[H] sw[H][SMAX] sPart[NWARP][H] sRed[H][3][WPH]
  float* xch = smem + a.oX;     // [WPC][R+1][M4]
  float* sBeta = sm, *sG = sm + H, *sGam = sm + 2 * H, *sSw = sm + 4 * H;
  float* sPart = sm + 4 * H + H * SMAX;
  float* sRed = sPart + NWARP * H;
  const float* Mi = a.Min + (size_t)b * a.sMin;
  float* Mo = a.Mout + (size_t)b * a.sMout;
  float* cnb = a.cn + (size_t)b * M4;

  const int offBeta = H * M, offG = offBeta + H, offS = offG + H, offGam = offS + S * H,
            offE = offGam + H, offA = offE + M * W;
  const float* mcb = a.mc + (size_t)b * a.PO4;
  auto rawv = [&](int q) -> float {
    float v = a.bias != nullptr ? __ldg(a.bias + q) : 0.0f;
    for (int s = 0; s < a.nslab; ++s) v += __ldcg(mcb + (size_t)s * a.slab + q);
    return v;
  };

  // ---- activations (ntm_cell.py:133-196).  kS[h][d] = tanh(k) * cn[d]: the key's own 1/|k| is a
  //      per-head scalar applied to the similarities later; pad lanes d >= M are zeros ----
  {
    float ss[H];
#pragma unroll
    for (int h = 0; h < H; ++h) ss[h] = 0.0f;
    for (int d = tid; d < M4; d += NT) {
      float rv[H], re[W], ra[W];
      const bool in = d < M;
#pragma unroll
      for (int h = 0; h < H; ++h) rv[h] = in ? rawv(h * M + d) : 0.0f;
#pragma unroll
      for (int h = 0; h < W; ++h) {
        re[h] = in ? rawv(offE + h * M + d) : 0.0f;
        ra[h] = in ? rawv(offA + h * M + d) : 0.0f;
      }
      const float cnd = in ? __ldcg(cnb + d) : 0.0f;
      if (in && a.cn_hist != nullptr) a.cn_hist[(size_t)b * M + d] = cnd;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        rv[h] = in ? tanh_f(rv[h]) : 0.0f;
        kS[h * M4 + d] = rv[h] * cnd;
        ss[h] = fmaf(rv[h], rv[h], ss[h]);
      }
#pragma unroll
      for (int h = 0; h < W; ++h) {
        eS[h * M4 + d] = in ? sigmoid_f(re[h]) : 0.0f;
        aS[h * M4 + d] = in ? tanh_f(ra[h]) : 0.0f;
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
  if (tid < H) {   // per-head scalars: beta, g, gamma (ntm_cell.py:140,151,169), shift softmax (:161)
    sBeta[tid] = softplus_f(rawv(offBeta + tid));
    sG[tid] = sigmoid_f(rawv(offG + tid));
    sGam[tid] = 1.0f + softplus_f(rawv(offGam + tid));
    float* sp = sSw + tid * SMAX;
    float mx = -INFINITY;
    for (int i = 0; i < S; ++i) { sp[i] = rawv(offS + tid * S + i); mx = fmaxf(mx, sp[i]); }
    float sum = 0.0f;
    for (int i = 0; i < S; ++i) { sp[i] = exp_f(sp[i] - mx); sum += sp[i]; }
    const float rsum = __frcp_rn(sum);
    for (int i = 0; i < S; ++i) sp[i] = sp[i] * rsum;
  }
  if (tid == NT - 1) {   // output projection + softmax (ntm_cell.py:220-221)
    const size_t o = ((size_t)b * a.T + a.t) * a.O;
    float mx = -INFINITY;
    for (int i = 0; i < a.O; ++i) mx = fmaxf(mx, rawv(a.P + i));
    float sum = 0.0f;
    for (int i = 0; i < a.O; ++i) sum += exp_f(rawv(a.P + i) - mx);
    const float rsum = __frcp_rn(sum);
    for (int i = 0; i < a.O; ++i) {
      const float lg = rawv(a.P + i);
      a.logits[o + i] = lg;
      if (a.outputs) a.outputs[o + i] = exp_f(lg - mx) * rsum;
    }
  }
  __syncthreads();

  // ---- pass 1 (HBM): sim[h][n] = sum_d kS[h][d] * M[n][d]  (ops.py:156).  A warp streams RB1 rows at
  //      once (RB1 independent 16-byte loads per lane in flight); the H key chunks it reads from shared
  //      memory are reused by all RB1 rows ----
  {
    const int nRB = (N + RB1 - 1) / RB1;
    for (int rb = warp; rb < nRB; rb += NWARP) {
      float acc[RB1][H];
#pragma unroll
      for (int i = 0; i < RB1; ++i)
#pragma unroll
        for (int h = 0; h < H; ++h) acc[i][h] = 0.0f;
      const float4* rp[RB1];
#pragma unroll
      for (int i = 0; i < RB1; ++i)
        rp[i] = reinterpret_cast<const float4*>(Mi + (size_t)min(rb * RB1 + i, N - 1) * M);
      for (int c = lane; c < MC; c += 32) {
        float4 m4[RB1];
#pragma unroll
        for (int i = 0; i < RB1; ++i) m4[i] = __ldcg(rp[i] + c);
        float4 k4[H];
#pragma unroll
        for (int h = 0; h < H; ++h) k4[h] = *reinterpret_cast<const float4*>(kS + h * M4 + 4 * c);
#pragma unroll
        for (int i = 0; i < RB1; ++i)
#pragma unroll
          for (int h = 0; h < H; ++h) {
            acc[i][h] = fmaf(m4[i].x, k4[h].x, acc[i][h]);
            acc[i][h] = fmaf(m4[i].y, k4[h].y, acc[i][h]);
            acc[i][h] = fmaf(m4[i].z, k4[h].z, acc[i][h]);
            acc[i][h] = fmaf(m4[i].w, k4[h].w, acc[i][h]);
          }
      }
      // transposing reduction, RB rows (RB*H <= 32 values) at a time: at offset o a lane keeps one half
      // of its value list and receives the partner's sums of that half; lane L ends with the warp
      // total of value L = i*H + h (fixed order: deterministic)
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
        const int vi = lane / H, vh = lane - vi * H;
        const int n = rb * RB1 + g * RB + vi;
        if (lane < RB * H && n < N) simS[vh * Npad + n] = v[0];
      }
    }
  }
  __syncthreads();
  if (a.sim_hist != nullptr)
    for (int i = tid; i < H * N; i += NT) a.sim_hist[(size_t)b * H * N + i] = simS[(i / N) * Npad + (i % N)];

  // ---- addressing on the [H][N] weightings (ntm_cell.py:140-176): WPH warps per head ----
  {
    constexpr int WPH = (NWARP / H) > 0 ? (NWARP / H) : 1;
    constexpr bool multi = (NWARP / H) > 0;
    const int hgrp = warp / WPH, sub = warp - hgrp * WPH;
    const int hstep = multi ? H : NWARP;
    const float* wprev = a.w_in + (size_t)b * a.sw_in;
    float* wout = a.w_out + (size_t)b * a.sw_out;
    for (int h = multi ? hgrp : warp; h < H; h += hstep) {
      float* sh = simS + h * Npad;
      float* gh = wg + h * Npad;
      float* red = sRed + h * 3 * WPH;
      const int nstep = 32 * WPH;
      const int n0 = 32 * sub + lane;
      auto head_bar = [&]() {
        if (WPH > 1) asm volatile("bar.sync %0, %1;" ::"r"(h + 1), "r"(32 * WPH) : "memory");
        else __syncwarp();
      };
      const float gate = sG[h], gamma = sGam[h];
      float kn = 0.0f;
      for (int w2 = 0; w2 < NWARP; ++w2) kn += sPart[w2 * H + h];
      const float rs = 1.0f / sqrtf(fmaxf(kn, 1e-12f));   // ops.py:152
      const float beta = sBeta[h];
      float mx = -INFINITY;
      for (int n = n0; n < N; n += nstep) {
        const float x = sh[n] * rs * beta;
        sh[n] = x;
        mx = fmaxf(mx, x);
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
      for (int n = n0; n < N; n += nstep) {
        const float e = exp_f(sh[n] - mx);
        sh[n] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      if (WPH > 1) {
        if (lane == 0) red[WPH + sub] = sum;
        head_bar();
        sum = red[WPH];
#pragma unroll
        for (int i = 1; i < WPH; ++i) sum += red[WPH + i];
      }
      for (int n = n0; n < N; n += nstep) {
        const float wc = sh[n] / sum;
        gh[n] = wc * gate + __ldcg(wprev + h * N + n) * (1.0f - gate);
      }
      head_bar();   // the shift reads neighbours' gated weights
      float psum = 0.0f;
      for (int n = n0; n < N; n += nstep) {
        float conv = 0.0f;
        for (int s = 0; s < S; ++s) {
          int idx = n + a.shift0 + s;   // circular_shift(x, j)[n] = x[(n + j) mod N], ops.py:216-242
          idx = idx < 0 ? idx + N : (idx >= N ? idx - N : idx);
          conv = fmaf(sSw[h * SMAX + s], gh[idx], conv);
        }
        const float pw = exp2f(gamma * log2f(conv));   // conv >= 0, gamma >= 1: pow(conv, gamma), 0 -> 0
        sh[n] = pw;
        psum += pw;
      }
      psum = warp_sum(psum);
      if (WPH > 1) {
        if (lane == 0) red[2 * WPH + sub] = psum;
        head_bar();
        psum = red[2 * WPH];
#pragma unroll
        for (int i = 1; i < WPH; ++i) psum += red[2 * WPH + i];
      }
      const float den = psum + 1e-3f;   // ntm_cell.py:175-176
      for (int n = n0; n < N; n += nstep) {
        const float wv = sh[n] / den;
        wnew[h * Npad + n] = wv;
        wout[h * N + n] = wv;
      }
    }
  }
  __syncthreads();

  // ---- pass 2 (L2): erase/add write, weighted read, next column norms (ntm_cell.py:193-215).
  //      lane -> (16-byte column chunk cl of an 8-chunk group, row phase rg); WPC warps share a group ----
  {
    const int cl = lane & 7, rg = lane >> 3;
    const int ncg = (MC + 7) >> 3, WPC = a.WPC;
    const int rstep = 4 * WPC;
    for (int si = warp; si < ncg * WPC; si += NWARP) {
      const int cgi = si / WPC, wsub = si - cgi * WPC;
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
        for (int nb = wsub * 4 + rg; nb < N; nb += rstep * U2) {
          float4 mv[U2];
#pragma unroll
          for (int u = 0; u < U2; ++u) {
            const int n = nb + u * rstep;
            if (n < N) mv[u] = __ldcg(reinterpret_cast<const float4*>(Mi + (size_t)n * M) + c);
          }
#pragma unroll
          for (int u = 0; u < U2; ++u) {
            const int n = nb + u * rstep;
            if (n < N) {
              const float4 m = mv[u];
              float4 mn;
              if constexpr (W == 1) {
                // one write head: M' = M (1 - w e) + w a = M + w (a - M e)
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
              const float4 mu = a.write_first ? mn : m;
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const float wr = wnew[r * Npad + n];
                racc[r].x = fmaf(wr, mu.x, racc[r].x); racc[r].y = fmaf(wr, mu.y, racc[r].y);
                racc[r].z = fmaf(wr, mu.z, racc[r].z); racc[r].w = fmaf(wr, mu.w, racc[r].w);
              }
              csq.x = fmaf(mn.x, mn.x, csq.x); csq.y = fmaf(mn.y, mn.y, csq.y);
              csq.z = fmaf(mn.z, mn.z, csq.z); csq.w = fmaf(mn.w, mn.w, csq.w);
              __stcg(reinterpret_cast<float4*>(Mo + (size_t)n * M) + c, mn);
            }
          }
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
        float* xw = xch + (size_t)wsub * (R + 1) * M4;
#pragma unroll
        for (int r = 0; r < R; ++r) *reinterpret_cast<float4*>(xw + r * M4 + 4 * c) = racc[r];
        *reinterpret_cast<float4*>(xw + R * M4 + 4 * c) = csq;
      }
    }
  }
  __syncthreads();

  // ---- finalize: fixed-order sum over the WPC row slices; read vector + inverse column norms ----
  {
    float* ar = a.act_read + (size_t)b * a.s_act;
    float* ro = a.read_out != nullptr ? a.read_out + (size_t)b * a.s_read : nullptr;
    for (int i = tid; i < (R + 1) * M; i += NT) {
      const int r = i / M, d = i - r * M;
      float s = 0.0f;
      for (int q = 0; q < a.WPC; ++q) s += xch[((size_t)q * (R + 1) + r) * M4 + d];
      if (r < R) {
        ar[i] = s;
        if (ro) ro[i] = s;
      } else {
        cnb[d] = 1.0f / sqrtf(fmaxf(s, 1e-12f));   // tf.nn.l2_normalize over N, ops.py:147-150
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Once per call: M_in (batch stride may be 0) -> working memory, and its inverse column norms.
constexpr int INIT_NT = 256;
__global__ void __launch_bounds__(INIT_NT) init_mem_kernel(const float* __restrict__ Min, long long sMin,
                                                           float* Mout, long long sMout, float* cn, int N,
                                                           int M, int M4, int MC, int WPC) {
  extern __shared__ float4 init_smem4[];
  float* xch = reinterpret_cast<float*>(init_smem4);   // [WPC][M4]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Mi = Min + (size_t)b * sMin;
  float* Mo = Mout + (size_t)b * sMout;
  const bool copy = (Mi != Mo);
  const int cl = lane & 7, rg = lane >> 3;
  const int ncg = (MC + 7) >> 3;
  for (int si = warp; si < ncg * WPC; si += INIT_NT / 32) {
    const int cgi = si / WPC, wsub = si - cgi * WPC;
    const int c = cgi * 8 + cl;
    const bool valid = c < MC;
    float4 csq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      for (int n = wsub * 4 + rg; n < N; n += 4 * WPC) {
        const float4 m = __ldcg(reinterpret_cast<const float4*>(Mi + (size_t)n * M) + c);
        csq.x = fmaf(m.x, m.x, csq.x); csq.y = fmaf(m.y, m.y, csq.y);
        csq.z = fmaf(m.z, m.z, csq.z); csq.w = fmaf(m.w, m.w, csq.w);
        if (copy) __stcg(reinterpret_cast<float4*>(Mo + (size_t)n * M) + c, m);
      }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      csq.x += __shfl_xor_sync(0xffffffffu, csq.x, o);
      csq.y += __shfl_xor_sync(0xffffffffu, csq.y, o);
      csq.z += __shfl_xor_sync(0xffffffffu, csq.z, o);
      csq.w += __shfl_xor_sync(0xffffffffu, csq.w, o);
    }
    if (valid && rg == 0) *reinterpret_cast<float4*>(xch + (size_t)wsub * M4 + 4 * c) = csq;
  }
  __syncthreads();
  for (int d = tid; d < M; d += INIT_NT) {
    float s = 0.0f;
    for (int q = 0; q < WPC; ++q) s += xch[(size_t)q * M4 + d];
    cn[(size_t)b * M4 + d] = 1.0f / sqrtf(fmaxf(s, 1e-12f));
  }
}

// rows 1 .. B-1 of a [B][n] array := row 0 (the initial memory is one broadcast copy: its column norms are computed once)
__global__ void bcast_row0_kernel(float* a, int n, long long B) {
  const long long total = (B - 1) * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    a[n + i] = a[i % n];
}

struct SmallInitArgs {
  int H, N, R, M, C, L;
  const float *w_in, *read_in, *ctrl_in;
  long long sw_in, sread_in, sctrl_in;
  float *w_out, *ctrl_out;
  long long sw_out, sctrl_out;
  float* act[MAXL];
  int actK[MAXL];
  float *hC, *hH, *hRead;     // slot 0 of the training history (or null)
  long long B;
};
// weightings, read vectors and controller state -> working buffers (one CTA per sequence)
__global__ void init_small_kernel(const SmallInitArgs a) {
  const long long b = blockIdx.x;
  const int tid = threadIdx.x, NT = blockDim.x;
  const float* wi = a.w_in + b * a.sw_in;
  float* wo = a.w_out + b * a.sw_out;
  if (wi != wo)
    for (int i = tid; i < a.H * a.N; i += NT) wo[i] = wi[i];
  for (int i = tid; i < a.R * a.M; i += NT) {
    const float v = a.read_in[b * a.sread_in + i];
    a.act[0][b * a.actK[0] + i] = v;
    if (a.hRead) a.hRead[b * (long long)(a.R * a.M) + i] = v;
  }
  for (int i = tid; i < a.L * a.C; i += NT) {
    const int l = i / a.C, u = i - l * a.C;
    const float c = a.ctrl_in[b * a.sctrl_in + (2 * l) * a.C + u];
    const float h = a.ctrl_in[b * a.sctrl_in + (2 * l + 1) * a.C + u];
    a.ctrl_out[b * a.sctrl_out + (2 * l) * a.C + u] = c;
    a.ctrl_out[b * a.sctrl_out + (2 * l + 1) * a.C + u] = h;
    a.act[l][b * a.actK[l] + (a.actK[l] - a.C) + u] = h;
    if (a.hC) a.hC[(b * a.L + l) * a.C + u] = c;
    if (a.hH) a.hH[(b * a.L + l) * a.C + u] = h;
  }
}

// BasicLSTMCell gates for all sequences (TF 1.0/1.1: i, j, f, o = split4(z); c' = c*sig(f) +
// sig(i)*tanh(j); h' = tanh(c')*sig(o)); z = hoisted x-projection (layer 0, bias folded in) or bias
// plus the K-slice partials of the controller GEMM in slice order.
struct LstmArgs {
  long long B; int C, L, l, T, t, KS; long long slab;
  const float* xw; const float* bias; const float* part;
  float* ctrl; long long sctrl;          // state_out.controller_state, updated in place
  float* act_self; int actK_self;        // h -> recurrent input of this layer (last C columns)
  float* act_next; int actK_next;        // h -> first C columns of the next layer's input (or null)
  float *hZ, *hC, *hH;                   // training history (or null)
  uint8_t *tilesA, *tilesC;              // GEMM operand tiles receiving h (C % 8 == 0), or null
  int KAtotA, koffA, KAtotC;
  const float* xr; const float* wrem; int nrem;   // layer 0: input columns the hoisted projection left out
};                                                // ([B*T][nrem]) and their weight rows ([nrem][4C]); nrem 0 = none
__global__ void lstm_stream_kernel(const LstmArgs a) {
  pdl_trigger();
  pdl_wait();
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < a.B * a.C;
  const long long i = live ? i0 : 0;      // tail lanes recompute element 0 and store nothing
  const long long b = i / a.C;
  const int u = (int)(i - b * a.C), C = a.C;
  float z[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = q * C + u;
    float v = (a.l == 0) ? __ldg(a.xw + (b * a.T + a.t) * (long long)(4 * C) + col) : __ldg(a.bias + col);
    z[q] = v;
  }
  if (a.l == 0 && a.nrem > 0) {
    const float* xr = a.xr + (b * a.T + a.t) * (long long)a.nrem;
    for (int c2 = 0; c2 < a.nrem; ++c2) {
      const float xv = __ldg(xr + c2);
#pragma unroll
      for (int q = 0; q < 4; ++q) z[q] = fmaf(xv, __ldg(a.wrem + (size_t)c2 * 4 * C + q * C + u), z[q]);
    }
  }
  // K-slice partials of the controller GEMM, summed in slice order (deterministic); the loads of a slice
  // for all four gates are issued together
  {
    const float* pp = a.part + b * (long long)(4 * C) + u;
#pragma unroll 4
    for (int ks = 0; ks < a.KS; ++ks) {
      float sl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) sl[q] = __ldcg(pp + (size_t)ks * a.slab + q * C);
#pragma unroll
      for (int q = 0; q < 4; ++q) z[q] += sl[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (a.hZ && live) a.hZ[((((size_t)a.t * a.B + b) * a.L + a.l) * 4 + q) * C + u] = z[q];
  }
  float* cp = a.ctrl + b * a.sctrl + (2 * a.l) * C + u;
  const float c_prev = *cp;
  const float c_new = c_prev * sigmoid_f(z[2]) + sigmoid_f(z[0]) * tanh_f(z[1]);
  const float h_new = tanh_f(c_new) * sigmoid_f(z[3]);
  if (live) {
    cp[0] = c_new;
    cp[C] = h_new;
    if (a.hC) a.hC[(((size_t)(a.t + 1) * a.B + b) * a.L + a.l) * C + u] = c_new;
    if (a.hH) a.hH[(((size_t)(a.t + 1) * a.B + b) * a.L + a.l) * C + u] = h_new;
    a.act_self[b * a.actK_self + (a.actK_self - C) + u] = h_new;
    if (a.act_next) a.act_next[b * a.actK_next + u] = h_new;
  }
  if (a.tilesA != nullptr) {   // C % 8 == 0: aligned groups of 8 lanes hold 8 consecutive units of one sequence
    const int lane = threadIdx.x & 31, g0 = lane & ~7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __shfl_sync(0xffffffffu, h_new, g0 + e);
    if (live && (lane & 7) == 0) gemmws::store_split8(a.tilesA, a.KAtotA, b, a.koffA + u, v);
    if (live && (lane & 7) == 1) gemmws::store_split8(a.tilesC, a.KAtotC, b, u - 1, v);
  }
}

// Same step, four consecutive units per thread (C % 4 == 0, 16-byte aligned rows everywhere): every access is a
// 16-byte vector -- 4x the bytes in flight per thread of the scalar kernel above, which was latency-bound on the
// K-slice slabs (33 us per step of 4096 sequences; this one: see DESIGN.md s4.3).  Same summation order.
__global__ void __launch_bounds__(256) lstm_stream_kernel_v4(const LstmArgs a) {
  pdl_trigger();
  pdl_wait();
  const int C = a.C, C4 = C >> 2;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < a.B * C4;
  const unsigned ii = live ? (unsigned)i0 : 0u;      // host guarantees B * C4 < 2^31
  const unsigned bu = ii / (unsigned)C4;
  const long long b = bu;
  const int u = 4 * (int)(ii - bu * (unsigned)C4);
  float4 z[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = q * C + u;
    z[q] = (a.l == 0) ? __ldg(reinterpret_cast<const float4*>(a.xw + (b * a.T + a.t) * (long long)(4 * C) + col))
                      : __ldg(reinterpret_cast<const float4*>(a.bias + col));
  }
  if (a.l == 0 && a.nrem > 0) {
    const float* xr = a.xr + (b * a.T + a.t) * (long long)a.nrem;
    for (int c2 = 0; c2 < a.nrem; ++c2) {
      const float xv = __ldg(xr + c2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(a.wrem + (size_t)c2 * 4 * C + q * C + u));
        z[q].x = fmaf(xv, wv.x, z[q].x); z[q].y = fmaf(xv, wv.y, z[q].y);
        z[q].z = fmaf(xv, wv.z, z[q].z); z[q].w = fmaf(xv, wv.w, z[q].w);
      }
    }
  }
  {
    const float* pp = a.part + b * (long long)(4 * C) + u;
#pragma unroll 5
    for (int ks = 0; ks < a.KS; ++ks) {
      float4 sl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) sl[q] = __ldcg(reinterpret_cast<const float4*>(pp + (size_t)ks * a.slab + q * C));
#pragma unroll
      for (int q = 0; q < 4; ++q) { z[q].x += sl[q].x; z[q].y += sl[q].y; z[q].z += sl[q].z; z[q].w += sl[q].w; }
    }
  }
  float* cp = a.ctrl + b * a.sctrl + (2 * a.l) * C + u;
  const float4 c_prev = *reinterpret_cast<const float4*>(cp);
  float4 c_new, h_new;
  c_new.x = c_prev.x * sigmoid_f(z[2].x) + sigmoid_f(z[0].x) * tanh_f(z[1].x);
  c_new.y = c_prev.y * sigmoid_f(z[2].y) + sigmoid_f(z[0].y) * tanh_f(z[1].y);
  c_new.z = c_prev.z * sigmoid_f(z[2].z) + sigmoid_f(z[0].z) * tanh_f(z[1].z);
  c_new.w = c_prev.w * sigmoid_f(z[2].w) + sigmoid_f(z[0].w) * tanh_f(z[1].w);
  h_new.x = tanh_f(c_new.x) * sigmoid_f(z[3].x);
  h_new.y = tanh_f(c_new.y) * sigmoid_f(z[3].y);
  h_new.z = tanh_f(c_new.z) * sigmoid_f(z[3].z);
  h_new.w = tanh_f(c_new.w) * sigmoid_f(z[3].w);
  if (live) {
    if (a.hZ) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(a.hZ + ((((size_t)a.t * a.B + b) * a.L + a.l) * 4 + q) * C + u) = z[q];
    }
    *reinterpret_cast<float4*>(cp) = c_new;
    *reinterpret_cast<float4*>(cp + C) = h_new;
    if (a.hC) *reinterpret_cast<float4*>(a.hC + (((size_t)(a.t + 1) * a.B + b) * a.L + a.l) * C + u) = c_new;
    if (a.hH) *reinterpret_cast<float4*>(a.hH + (((size_t)(a.t + 1) * a.B + b) * a.L + a.l) * C + u) = h_new;
    *reinterpret_cast<float4*>(a.act_self + b * a.actK_self + (a.actK_self - C) + u) = h_new;
    if (a.act_next) *reinterpret_cast<float4*>(a.act_next + b * a.actK_next + u) = h_new;
  }
  if (a.tilesA != nullptr) {   // C % 8 == 0: an (even, odd) lane pair holds 8 consecutive units of one sequence
    const bool odd = (threadIdx.x & 1) != 0;
    const float px = __shfl_xor_sync(0xffffffffu, h_new.x, 1), py = __shfl_xor_sync(0xffffffffu, h_new.y, 1);
    const float pz = __shfl_xor_sync(0xffffffffu, h_new.z, 1), pw = __shfl_xor_sync(0xffffffffu, h_new.w, 1);
    float v[8];
    v[0] = odd ? px : h_new.x; v[1] = odd ? py : h_new.y; v[2] = odd ? pz : h_new.z; v[3] = odd ? pw : h_new.w;
    v[4] = odd ? h_new.x : px; v[5] = odd ? h_new.y : py; v[6] = odd ? h_new.z : pz; v[7] = odd ? h_new.w : pw;
    if (live && !odd) gemmws::store_split8(a.tilesA, a.KAtotA, b, a.koffA + u, v);
    if (live && odd) gemmws::store_split8(a.tilesC, a.KAtotC, b, u - 4, v);
  }
}

// ------------------------------------------------------------------------------------- host side --
constexpr int MEM_NT = 256;
thread_local int g_mem_occ = 0;
thread_local int g_env_mem_ctas_per_sm = 0;   // EnvSwitches::mem_ctas_per_sm of the call in progress
thread_local int g_env_mem_grid = 0;          // EnvSwitches::mem_grid of the call in progress
thread_local bool g_chain_pdl = false;        // launch the memory kernel with the programmatic-dependent-launch attribute

template <int R, int W, int NT, int MINB>
cudaError_t launch_mem_v(const MemArgs& a, long long B, int smem, cudaStream_t stream) {
  static int configured[MAX_DEVICES] = {0}, occs[MAX_DEVICES] = {0};   // per instantiation AND device
  {
    std::lock_guard<std::mutex> lk(config_mutex());
    const int dev = current_device_slot();
    if (configured[dev] < smem) {
      cudaError_t e = cudaFuncSetAttribute(mem_step_kernel<R, W, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
      configured[dev] = smem;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occs[dev], mem_step_kernel<R, W, NT, MINB>, NT, smem);
    }
    g_mem_occ = occs[dev];
  }
  return launch_chain(mem_step_kernel<R, W, NT, MINB>, (unsigned)B, NT, (size_t)smem, stream, g_chain_pdl, a);
}
template <int R, int W>
cudaError_t launch_mem_rw(const MemArgs& a, long long B, int smem, cudaStream_t stream) {
  return launch_mem_v<R, W, MEM_NT, 2>(a, B, smem, stream);   // 2 CTAs/SM: 720 us per C3 step vs 1073 (1/SM) and 796 (3/SM)
}
template <int R>
cudaError_t launch_mem_r(int W, const MemArgs& a, long long B, int smem, cudaStream_t stream) {
  switch (W) {
    case 1: return launch_mem_rw<R, 1>(a, B, smem, stream);
    case 2: return launch_mem_rw<R, 2>(a, B, smem, stream);
    case 3: return launch_mem_rw<R, 3>(a, B, smem, stream);
  }
  return cudaErrorInvalidValue;
}
cudaError_t launch_mem(int R, int W, const MemArgs& a, long long B, int smem, cudaStream_t stream) {
  switch (R) {
    case 1: return launch_mem_r<1>(W, a, B, smem, stream);
    case 2: return launch_mem_r<2>(W, a, B, smem, stream);
    case 3: return launch_mem_r<3>(W, a, B, smem, stream);
    case 4: return launch_mem_r<4>(W, a, B, smem, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_tma(int R, int W, int CPL, const MemArgs& a, long long B, int smem, cudaStream_t stream) {
  TmaCtl ctl{g_env_mem_ctas_per_sm, g_env_mem_grid, g_chain_pdl, g_mem_occ};
  cudaError_t e = cudaErrorInvalidValue;
  switch (R) {
    case 1: e = launch_tma_r<1>(W, CPL, a, B, smem, stream, ctl); break;
    case 2: e = launch_tma_r<2>(W, CPL, a, B, smem, stream, ctl); break;
    case 3:
    case 4: e = launch_tma_r34(R, W, CPL, a, B, smem, stream, ctl); break;
  }
  g_mem_occ = ctl.occ;
  return e;
}
// Rows per ring stage = half the rows one pass-2 iteration covers (2 * RP; 8 KiB when MC divides 256);
// the iteration's rows must divide N.  0 = shape not covered (the generic kernel runs).
int tma_rp(int MC) {   // quads per pass-2 iteration: the largest power of two <= NT / MC
  int rp = 1;
  while (2 * rp * MC <= TMA_NT) rp *= 2;
  return rp;
}
int tma_rps(int N, int M) {
  if (M % 4 != 0 || M > 512) return 0;
  const int rp = tma_rp(M / 4);
  if (rp < 2) return 0;
  const int rs = 4 * rp;                       // rows of one pass-2 iteration = two stages
  return (N % rs == 0) ? rs / 2 : 0;
}
// 0 when the TMA-ring kernel does not cover the shape (then the generic register-streaming kernel runs)
int tma_cpl(int H, int MC) {
  const int cpl = MC <= 32 ? 1 : (MC <= 64 ? 2 : (MC <= 128 ? 4 : 0));
  if (cpl == 0 || H * cpl > 20) return 0;
  return cpl;
}

// Development only (NTM_B200_EXP bit 8): per-CTA timestamps of the last controller / head-parameter GEMM launch of a
// call, printed to stderr at the end of the call (synchronises the stream).
long long* g_gemm_prof_dev[MAX_DEVICES] = {nullptr};     // one buffer per device, allocated on first use, never freed
long long* gemm_prof_buffer(int which) {
  std::lock_guard<std::mutex> lk(config_mutex());
  long long*& buf = g_gemm_prof_dev[current_device_slot()];
  if (buf == nullptr && (cudaMalloc(&buf, 2 * 256 * 8 * sizeof(long long)) != cudaSuccess ||
                         cudaMemset(buf, 0, 2 * 256 * 8 * sizeof(long long)) != cudaSuccess)) {
    buf = nullptr;
    return nullptr;
  }
  return buf + (size_t)which * 256 * 8;
}
void gemm_prof_dump(cudaStream_t stream) {
  long long* g_gemm_prof = gemm_prof_buffer(0);
  if (g_gemm_prof == nullptr || cudaStreamSynchronize(stream) != cudaSuccess) return;
  std::vector<long long> h(2 * 256 * 8);
  if (cudaMemcpy(h.data(), g_gemm_prof, h.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  for (int w = 0; w < 2; ++w) {
    double pro = 0, loop = 0, tail = 0, wempty = 0, wfull = 0, wacc = 0; long long first = 0, last = 0; int n = 0;
    for (int c = 0; c < 256; ++c) {
      const long long* r = &h[((size_t)w * 256 + c) * 8];
      if (r[0] == 0 || r[3] == 0) continue;
      pro += (double)(r[1] - r[0]); loop += (double)(r[2] - r[1]); tail += (double)(r[3] - r[2]);
      wempty += (double)r[4]; wfull += (double)r[5]; wacc += (double)r[6];
      first = (n == 0) ? r[0] : std::min(first, r[0]); last = (n == 0) ? r[3] : std::max(last, r[3]);
      ++n;
    }
    if (n > 0)
      fprintf(stderr, "[gemm_ws prof %s] ctas %d span %.1f us | per CTA: prologue %.1f, MMA issue loop %.1f (waiting full %.1f, "
              "acc_empty %.1f), tail %.1f us; producer waiting empty %.1f us\n", w == 0 ? "controller" : "head", n,
              (last - first) / 1e3, pro / n / 1e3, loop / n / 1e3, wfull / n / 1e3, wacc / n / 1e3, tail / n / 1e3, wempty / n / 1e3);
  }
  cudaMemset(g_gemm_prof, 0, 2 * 256 * 8 * sizeof(long long));
}

thread_local std::vector<cudaEvent_t> g_sev;
thread_local long long* g_prof_ptr = nullptr;
thread_local long long g_prof_B = 0;
thread_local int g_sev_steps = 0;

}  // namespace

bool stream_supported(const ntm_b200_shape* s, int nsm) {
  if (s->mem_dim % 4 != 0) return false;
  const int C = s->controller_hidden_size, H = s->read_head_size + s->write_head_size;
  const int S = 2 * s->shift_range + 1;
  const int P = H * s->mem_dim + 3 * H + S * H + 2 * s->mem_dim * s->write_head_size;
  const int PO4 = round_up(P + s->output_dim, 4);
  // the tile kernel needs (column tiles x K-slices) <= SMs for every GEMM
  for (int l = 0; l < s->controller_num_layers; ++l) {
    const int K = (l == 0) ? s->read_head_size * s->mem_dim + C : 2 * C;
    if ((K & 1) != 0) return false;                       // 8-byte aligned activation rows
    if (ceil_div(4 * C, 128) * gemm_tc_slices(K) > nsm) return false;
  }
  if (gemm_tc_slices(C) != 1) return false;               // head parameters: one slice (bias in the GEMM)
  if (ceil_div(PO4, 128) > nsm) return false;
  // shared memory of the memory kernel
  const int M4 = s->mem_dim, Npad = round_up(s->mem_size, 4);
  const long long fl = (long long)(H + 2 * s->write_head_size) * M4 + 3ll * H * Npad + 256 +
                       (long long)(MEM_NT / 32) * (s->read_head_size + 1) * M4;
  return 4 * fl <= 200 * 1024;
}

// The hoisted input projection runs on the warp-specialised GEMM when the controller is single-layer with the
// operand-tile path (as stream_forward decides), and the input width is a whole number of 64-wide K atoms (at most
// KA_MAX of them) plus at most 8 columns (the tracker's 512 features + delimiter + target channels).
// true when the per-timestep chain runs on the warp-specialised GEMMs and the TMA-ring memory kernel (what
// stream_forward's `use_ws` decides): the fast streaming path the mode choice's crossover was measured on
bool stream_ws_path(const ntm_b200_shape* s) {
  const int C = s->controller_hidden_size, H = s->read_head_size + s->write_head_size, M = s->mem_dim;
  return s->controller_num_layers == 1 && M % 8 == 0 && C % 8 == 0 && tma_rps(s->mem_size, M) > 0 && tma_cpl(H, M / 4) > 0;
}

static bool xproj_ws_shape(const ntm_b200_shape* s, int* xK, int* xrem) {
  const int C = s->controller_hidden_size, H = s->read_head_size + s->write_head_size, M = s->mem_dim, D = s->input_dim;
  if (s->controller_num_layers != 1 || M % 8 != 0 || C % 8 != 0 || M % 4 != 0) return false;
  if (tma_rps(s->mem_size, M) <= 0 || tma_cpl(H, M / 4) <= 0) return false;
  const int K = D / 64 * 64;
  if (K < 64 || K / 64 > gemmws::KA_MAX || D - K > 8) return false;
  if ((4 * C + 127) / 128 > B200_SMS) return false;
  *xK = K; *xrem = D - K;
  return true;
}

namespace {
// the xrem last columns of the frames, compacted: xr[r][c] = x[r][xK + c]
__global__ void extract_cols_kernel(const float* __restrict__ x, long long rows, int D, int xK, int nrem, float* __restrict__ xr) {
  const long long total = rows * nrem;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nrem;
    const int c = (int)(i - r * nrem);
    xr[i] = __ldg(x + r * D + xK + c);
  }
}
// Feature-layout source (FeatureSource): row r = (b, t), t = l * (F+1) + pos; the delimiter row of a frame is pos == F
// (training layout) or pos == 0 (serve layout); feature f of frame l sits at pos f (resp. f + 1).  Same values, bit
// for bit, as serialize_kernel (ntm_b200_io.cu) followed by pack_act_tiles_kernel / extract_cols_kernel.
__global__ void pack_feature_tiles_kernel(const float* __restrict__ feat, long long rows, int L, int F, int Cch, int dfirst,
                                          uint8_t* tiles, int KAtot) {
  const long long nrb = (rows + 127) / 128;
  const int T = L * (F + 1);
  const long long total = nrb * 128 * (long long)KAtot * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(i & 7);
    long long t2 = i >> 3;
    const int row = (int)(t2 & 127); t2 >>= 7;
    const int ka = (int)(t2 % KAtot);
    const long long rb = t2 / KAtot;
    const long long r = rb * 128 + row;
    const int k = ka * 64 + chunk * 8;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (r < rows) {
      const long long b = r / T;
      const int t = (int)(r - b * T);
      const int l = t / (F + 1), pos = t - l * (F + 1);
      const bool is_delim = dfirst ? (pos == 0) : (pos == F);
      if (!is_delim) {
        const int f = dfirst ? pos - 1 : pos;
        const float4* src = reinterpret_cast<const float4*>(feat + (((b * L + l) * F + f) * (long long)Cch + k));
        v0 = __ldg(src); v1 = __ldg(src + 1);
      }
    }
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    gemmws::store_split8(tiles, KAtot, r, k, v);
  }
}
// the two synthesised channels: xr[r] = {delimiter flag, first-frame target on the feature rows}
__global__ void synth_cols_kernel(const float* __restrict__ target, long long rows, int L, int F, int dfirst, float* __restrict__ xr) {
  const int T = L * (F + 1);
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const long long b = r / T;
    const int t = (int)(r - b * T);
    const int l = t / (F + 1), pos = t - l * (F + 1);
    const bool is_delim = dfirst ? (pos == 0) : (pos == F);
    const int f = dfirst ? pos - 1 : pos;
    xr[2 * r] = is_delim ? 1.0f : 0.0f;
    xr[2 * r + 1] = (l == 0 && !is_delim) ? __ldg(target + b * F + f) : 0.0f;
  }
}
}  // namespace

int stream_xproj(const ntm_b200_shape* s, const ntm_b200_weights* w, long long B, long long T, const float* x, float* xw,
                 char* wsb, const StreamWorkspace& ws, int nsm, cudaStream_t stream, bool cont, const EnvSwitches& env,
                 const FeatureSource* fs) {
  if (ws.xK <= 0 || env.no_tma_ring || env.old_gemm || (env.exp & 32)) return -1;
  if (fs != nullptr && (fs->Cch != ws.xK || ws.xrem != 2 || (long long)fs->L * (fs->F + 1) != T ||
                        (reinterpret_cast<uintptr_t>(fs->features) & 15) != 0))
    return -1;
  const int C = s->controller_hidden_size, D = s->input_dim;
  const long long rows = B * T;
  const gemmws::Plan px = gemmws::make_plan(ws.xK, 4 * C, rows, nsm);
  if (!gemmws::plan_ok(px, nsm) || px.kslices != 1) return -1;
  uint32_t* whiX = reinterpret_cast<uint32_t*>(wsb + ws.off_whiX);
  uint8_t* wloX = reinterpret_cast<uint8_t*>(wsb + ws.off_wloX);
  uint8_t* xt = reinterpret_cast<uint8_t*>(wsb + ws.off_xtiles);
  cudaError_t e;
  if (!cont) {
    gemmws::pack_weight_tiles_kernel<<<2 * nsm, 256, 0, stream>>>(w->lstm_w[0], ws.xK, 4 * C, 4 * C, whiX, wloX, px.ntiles,
                                                                px.kslices, px.KA, px.wlo_tmem);
    count_launch();
  }
  // frames -> operand tiles (rows past B*T of the last block and nothing else are zero-filled by the kernel itself)
  if (fs != nullptr) {
    pack_feature_tiles_kernel<<<8 * nsm, 256, 0, stream>>>(fs->features, rows, fs->L, fs->F, fs->Cch, fs->delimiter_first ? 1 : 0,
                                                          xt, px.KAtot);
    synth_cols_kernel<<<2 * nsm, 256, 0, stream>>>(fs->target, rows, fs->L, fs->F, fs->delimiter_first ? 1 : 0,
                                                  reinterpret_cast<float*>(wsb + ws.off_xr));
    count_launch(); count_launch();
  } else {
    gemmws::pack_act_tiles_kernel<<<8 * nsm, 256, 0, stream>>>(x, rows, ws.xK, D, xt, px.KAtot, 0);
    count_launch();
    if (ws.xrem > 0) {
      extract_cols_kernel<<<2 * nsm, 256, 0, stream>>>(x, rows, D, ws.xK, ws.xrem, reinterpret_cast<float*>(wsb + ws.off_xr));
      count_launch();
    }
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error_ext(e, "x-projection pack kernels");
  e = gemmws::launch(px, xt, whiX, wloX, w->lstm_b[0], xw, 4 * C, 0, rows, stream, env.exp);
  count_launch();
  if (e != cudaSuccess) return set_cuda_error_ext(e, "gemm_ws(x-projection)");
  return 0;
}

void stream_layout(const ntm_b200_shape* s, long long B, long long T, StreamWorkspace* ws) {
  const int C = s->controller_hidden_size, L = s->controller_num_layers;
  const int H = s->read_head_size + s->write_head_size, S = 2 * s->shift_range + 1;
  const int P = H * s->mem_dim + 3 * H + S * H + 2 * s->mem_dim * s->write_head_size;
  const int PO4 = round_up(P + s->output_dim, 4);
  long long o = 0;
  auto take = [&](long long bytes) { long long r = o; o = align_up_ll(o + bytes, 256); return r; };
  int ksmax = 1;
  for (int l = 0; l < L; ++l) {
    ws->actK[l] = (l == 0) ? s->read_head_size * s->mem_dim + C : 2 * C;
    ws->ksA[l] = gemm_tc_slices(ws->actK[l]);
    ksmax = std::max(ksmax, ws->ksA[l]);
    ws->off_act[l] = take(4ll * B * ws->actK[l]);
  }
  ws->ksC = 1;
  ws->slabA = B * 4ll * C;
  ws->slabC = B * (long long)PO4;
  {   // operand tiles + packed weights of the warp-specialised GEMMs (layer-0 controller, head parameters)
    const gemmws::Plan pa = gemmws::make_plan(ws->actK[0], 4 * C, B, B200_SMS);
    const gemmws::Plan pc = gemmws::make_plan(C, PO4, B, B200_SMS);
    ksmax = std::max(ksmax, pa.kslices);
    ws->off_tilesA = take((long long)pa.act_bytes);
    ws->off_tilesC = take((long long)pc.act_bytes);
    ws->off_whiA = take((long long)pa.whi_bytes);
    ws->off_wloA = take((long long)pa.wlo_bytes);
    ws->off_whiC = take((long long)pc.whi_bytes);
    ws->off_wloC = take((long long)pc.wlo_bytes);
  }
  ws->xK = 0; ws->xrem = 0;
  ws->off_whiX = ws->off_wloX = ws->off_xtiles = ws->off_xr = 0;
  gemmws::Plan px{};
  if (xproj_ws_shape(s, &ws->xK, &ws->xrem)) {
    px = gemmws::make_plan(ws->xK, 4 * C, B * T, B200_SMS);
    ws->off_whiX = take((long long)px.whi_bytes);
    ws->off_wloX = take((long long)px.wlo_bytes);
  }
  ws->off_partA = take(4ll * ksmax * ws->slabA);
  ws->off_mc = take(4ll * ws->slabC);
  ws->off_cn = take(4ll * B * round_up(s->mem_dim, 4));
  ws->off_prof = take(8ll * 16 * B);
  ws->off_xw = take(4ll * B * T * 4 * C);
  if (ws->xK > 0) {      // (everything that depends on T stays at the end: see ntm_b200_forward_seq_continue)
    ws->off_xr = take(4ll * B * T * std::max(1, ws->xrem));
    ws->off_xtiles = take((long long)px.act_bytes);
  }
  ws->total = o;
}

int stream_forward(const ntm_b200_shape* s, const ntm_b200_weights* w, const float* wC, const float* bC,
                   long long B, long long T, const float* xw, const ntm_b200_state* in,
                   const ntm_b200_state* out, float* logits, float* outputs, const ntm_b200_history* hist,
                   char* wsb, const StreamWorkspace& ws, int nsm, cudaStream_t stream, bool prof, bool cont,
                   const EnvSwitches& env, bool xw_partial) {
  g_env_mem_ctas_per_sm = env.mem_ctas_per_sm;
  g_env_mem_grid = env.mem_grid;
  const int C = s->controller_hidden_size, L = s->controller_num_layers;
  const int R = s->read_head_size, W = s->write_head_size, H = R + W, S = 2 * s->shift_range + 1;
  const int N = s->mem_size, M = s->mem_dim, M4 = round_up(M, 4), MC = M4 / 4, Npad = round_up(N, 4);
  const int P = H * M + 3 * H + S * H + 2 * M * W, PO = P + s->output_dim, PO4 = round_up(PO, 4);
  cudaError_t e;
  float* act[MAXL];
  for (int l = 0; l < L; ++l) act[l] = reinterpret_cast<float*>(wsb + ws.off_act[l]);
  float* partA = reinterpret_cast<float*>(wsb + ws.off_partA);
  float* mcbuf = reinterpret_cast<float*>(wsb + ws.off_mc);
  float* cn = reinterpret_cast<float*>(wsb + ws.off_cn);
  const bool hM = hist && hist->M_prev, hW = hist && hist->w_prev, hP = hist && hist->params;

  if (prof) {
    const size_t need = 4 * (size_t)T + 2;
    while (g_sev.size() < need) {
      cudaEvent_t ev;
      if ((e = cudaEventCreate(&ev)) != cudaSuccess) return set_cuda_error_ext(e, "cudaEventCreate");
      g_sev.push_back(ev);
    }
    g_sev_steps = 0;
    cudaEventRecord(g_sev[0], stream);
  }

  // ---- init: working copies of the state, inverse column norms of the initial memory ----
  const int ncg = (MC + 7) / 8;
  const int WPCi = std::max(1, (INIT_NT / 32) / ncg);
  float* M0 = hM ? hist->M_prev : out->M;                      // memory entering step 0
  const long long sM0 = hM ? (long long)N * M : out->stride_M;
  // Without a history the initial memory is not copied at all: step 0 reads it where it is (one shared copy when
  // zero_state broadcasts it, stride 0) and writes the working memory; only its column norms are derived here.
  const bool lazy_M0 = !hM;
  if (!cont) {
    // zero_state hands ONE copy of the initial memory to all sequences (batch stride 0): without a history nothing is
    // copied, so one CTA derives the column norms and a second kernel replicates the row (4096 CTAs re-reading the
    // same 256 KiB cost 0.09-0.18 ms per call)
    const bool one_copy = lazy_M0 && in->stride_M == 0 && B > 1;
    init_mem_kernel<<<one_copy ? 1u : (unsigned)B, INIT_NT, WPCi * M4 * 4, stream>>>(
        in->M, in->stride_M, lazy_M0 ? const_cast<float*>(in->M) : M0, lazy_M0 ? in->stride_M : sM0, cn, N, M, M4, MC, WPCi);
    count_launch();
    if (one_copy) {
      bcast_row0_kernel<<<nsm, 256, 0, stream>>>(cn, M4, B);
      count_launch();
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error_ext(e, "init_mem_kernel");
  }
  if (!cont) {
    SmallInitArgs ia{};
    ia.H = H; ia.N = N; ia.R = R; ia.M = M; ia.C = C; ia.L = L; ia.B = B;
    ia.w_in = in->w; ia.read_in = in->read; ia.ctrl_in = in->controller_state;
    ia.sw_in = in->stride_w; ia.sread_in = in->stride_read; ia.sctrl_in = in->stride_controller_state;
    ia.w_out = hW ? hist->w_prev : out->w; ia.sw_out = hW ? (long long)H * N : out->stride_w;
    ia.ctrl_out = out->controller_state; ia.sctrl_out = out->stride_controller_state;
    for (int l = 0; l < L; ++l) { ia.act[l] = act[l]; ia.actK[l] = ws.actK[l]; }
    ia.hC = hist ? hist->c : nullptr; ia.hH = hist ? hist->h : nullptr; ia.hRead = hist ? hist->read : nullptr;
    init_small_kernel<<<(unsigned)B, 256, 0, stream>>>(ia);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error_ext(e, "init_small_kernel");
  }
  // ---- warp-specialised tensor-core GEMMs (single-layer controller, M and C multiples of 8, TMA memory
  //      kernel): operands pre-split into bf16 hi/lo tile records by their producers ----
  const bool use_ws = (L == 1) && (M % 8 == 0) && (C % 8 == 0) && tma_rps(N, M) > 0 && tma_cpl(H, MC) > 0 &&
                      !env.no_tma_ring && !env.old_gemm;
  gemmws::Plan planA{}, planC{};
  uint8_t* tilesA = reinterpret_cast<uint8_t*>(wsb + ws.off_tilesA);
  uint8_t* tilesC = reinterpret_cast<uint8_t*>(wsb + ws.off_tilesC);
  uint32_t* whiA = reinterpret_cast<uint32_t*>(wsb + ws.off_whiA);
  uint8_t* wloA = reinterpret_cast<uint8_t*>(wsb + ws.off_wloA);
  uint32_t* whiC = reinterpret_cast<uint32_t*>(wsb + ws.off_whiC);
  uint8_t* wloC = reinterpret_cast<uint8_t*>(wsb + ws.off_wloC);
  bool ws_ok = false;
  if (use_ws) {
    planA = gemmws::make_plan(ws.actK[0], 4 * C, B, nsm);
    planC = gemmws::make_plan(C, PO4, B, nsm);
    ws_ok = gemmws::plan_ok(planA, nsm) && gemmws::plan_ok(planC, nsm);
  }
  if (ws_ok && !cont) {
    if ((e = cudaMemsetAsync(tilesA, 0, planA.act_bytes, stream)) != cudaSuccess) return set_cuda_error_ext(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(tilesC, 0, planC.act_bytes, stream)) != cudaSuccess) return set_cuda_error_ext(e, "cudaMemsetAsync");
    const int RM = R * M;
    gemmws::pack_act_tiles_kernel<<<2 * nsm, 256, 0, stream>>>(act[0], B, RM, ws.actK[0], tilesA, planA.KAtot, 0);
    gemmws::pack_act_tiles_kernel<<<2 * nsm, 256, 0, stream>>>(act[0] + RM, B, C, ws.actK[0], tilesA, planA.KAtot, RM);
    gemmws::pack_act_tiles_kernel<<<2 * nsm, 256, 0, stream>>>(act[0] + RM, B, C, ws.actK[0], tilesC, planC.KAtot, 0);
    gemmws::pack_weight_tiles_kernel<<<2 * nsm, 256, 0, stream>>>(w->lstm_w[0] + (size_t)s->input_dim * 4 * C, ws.actK[0],
                                                                4 * C, 4 * C, whiA, wloA, planA.ntiles, planA.kslices, planA.KA, planA.wlo_tmem);
    gemmws::pack_weight_tiles_kernel<<<2 * nsm, 256, 0, stream>>>(wC, C, PO4, PO4, whiC, wloC, planC.ntiles, planC.kslices,
                                                                planC.KA, planC.wlo_tmem);
    for (int i = 0; i < 5; ++i) count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return set_cuda_error_ext(e, "gemmws pack kernels");
  }
  if (prof) cudaEventRecord(g_sev[1], stream);

  // ---- shared-memory carve-up of the memory kernel ----
  MemArgs ma{};
  {
    const int nwarp = MEM_NT / 32;
    ma.WPC = std::max(1, nwarp / ncg);
    int o = 0;
    auto take = [&](int n) { int r = o; o += round_up(n, 4); return r; };
    ma.oK = take(H * M4); ma.oE = take(W * M4); ma.oA = take(W * M4);
    ma.oSim = take(H * Npad); ma.oWg = take(H * Npad); ma.oWn = take(H * Npad);
    ma.oSm = take(4 * H + H * SMAX + nwarp * H + 3 * H * std::max(1, nwarp / H) + 8);
    ma.oX = take(ma.WPC * (R + 1) * M4);
    const int smem = 4 * o;
    // TMA-ring kernel (when it covers the shape): its own carve-up replaces the generic one
    int smem_tma = 0;
    const int cpl = (!env.no_tma_ring && tma_rps(N, M) > 0) ? tma_cpl(H, MC) : 0;
    if (cpl) {
      int o2 = 0;
      auto take2 = [&](int n) { int r = o2; o2 += round_up(n, 4); return r; };
      ma.RP = tma_rp(MC);
      ma.oK = take2(std::max(H * M4, (ma.RP - 1) * (R + 1) * M4));   // keys, later the quad-slot partials
      ma.oE = take2(W * M4); ma.oA = take2(W * M4);
      ma.oSim = take2(H * Npad); ma.oWg = take2(H * Npad); ma.oWn = take2(H * Npad); ma.oWp = take2(H * Npad);
      ma.oRaw = take2(PO4); ma.oCn = take2(M4);
      ma.oSm = take2(4 * H + H * SMAX + (TMA_NT / 32) * H + 3 * H * std::max(1, (TMA_NT / 32) / H) + 8);
      ma.NS = TMA_NS;
      ma.RPS = tma_rps(N, M);
      ma.NCH = N / ma.RPS;
      // stages of pass 1 kept in the ring for pass 2 (needs a whole ring of them; NTM_B200_EXP bit 1 = off)
      ma.NR = (ma.NCH >= ma.NS && !(env.exp & 2)) ? ma.NS : 0;
      ma.exp = env.exp;
      const int qps = 2 * ma.NCH - ma.NR;          // stage uses per sequence
      ma.qps_shift = -1;
      for (int sh = 0; sh < 30; ++sh)
        if ((1 << sh) == qps) ma.qps_shift = sh;
      ma.qps_magic = 0u;                           // exact for Qg * qps < 2^32
      if ((B + 1) * (long long)qps * qps < (1ll << 32)) ma.qps_magic = (unsigned)(((1ull << 32) + qps - 1) / qps);
      ma.RP = tma_rp(MC);
      ma.oBar = take2(2 * (ma.NS + 1) + 2);
      o2 = round_up(o2, 32);                      // 128-byte aligned ring
      ma.oRing = o2;
      o2 += ma.NS * ma.RPS * M;
      smem_tma = 4 * o2;
      // Large N (C4: N = 1024): that plan allows one CTA per SM only.  The compact plan -- 4 ring stages, three
      // rotating [H][N] buffers instead of four, quad-slot partials in the similarity buffer -- fits two.
      const int per_cta_2 = (B200_SMEM_SM - 2 * 1024) / 2;
      const bool full_m = (M == 128 * cpl);
      if (smem_tma > per_cta_2 && cpl <= 2 && full_m && ma.RPS >= 8 && ma.NCH >= 4 && N != 128 &&
          (ma.RP - 1) * (R + 1) * M4 <= H * Npad && !(env.exp & 2048)) {
        int o3 = 0;
        auto take3 = [&](int n) { int r = o3; o3 += round_up(n, 4); return r; };
        MemArgs mb = ma;
        mb.oK = take3(H * M4);
        mb.oE = take3(W * M4); mb.oA = take3(W * M4);
        mb.oSim = take3(H * Npad); mb.oWg = take3(H * Npad); mb.oWp = take3(H * Npad); mb.oWn = mb.oWp;
        mb.oRaw = take3(PO4); mb.oCn = take3(M4);
        mb.oSm = take3(4 * H + H * SMAX + (TMA_NT / 32) * H + 3 * H * std::max(1, (TMA_NT / 32) / H) + 8);
        mb.NS = 4; mb.rot = 1;
        mb.NR = (mb.NCH >= mb.NS && !(env.exp & 2)) ? mb.NS : 0;
        const int qps2 = 2 * mb.NCH - mb.NR;
        mb.qps_shift = -1;
        for (int sh = 0; sh < 30; ++sh)
          if ((1 << sh) == qps2) mb.qps_shift = sh;
        mb.qps_magic = 0u;
        if ((B + 1) * (long long)qps2 * qps2 < (1ll << 32)) mb.qps_magic = (unsigned)(((1ull << 32) + qps2 - 1) / qps2);
        mb.oBar = take3(2 * (mb.NS + 1) + 2);
        o3 = round_up(o3, 32);
        mb.oRing = o3;
        o3 += mb.NS * mb.RPS * M;
        if (4 * o3 <= per_cta_2) { ma = mb; smem_tma = 4 * o3; }
      }
    }
    ma.P2S = 1;
    if (cpl && ma.NS == TMA_NS && ma.RP > 0 && (N / (4 * ma.RP)) % 2 == 0 && ma.NR % 4 == 0 && !(env.exp & 4096)) ma.P2S = 2;
    ma.N = N; ma.M = M; ma.M4 = M4; ma.MC = MC; ma.Npad = Npad; ma.S = S;
    ma.shift0 = -((S + 1) / 2);   // Python-2 floor(-S/2), ops.py:204
    ma.P = P; ma.PO4 = PO4; ma.O = s->output_dim; ma.write_first = s->write_first ? 1 : 0; ma.T = (int)T;
    ma.nslab = 1; ma.slab = 0; ma.bias = nullptr;
    ma.cn = cn; ma.act_read = act[0]; ma.s_act = ws.actK[0];
    ma.logits = logits; ma.outputs = outputs;
    ma.B = B;
    ma.tilesA = ws_ok ? tilesA : nullptr; ma.KAtotA = planA.KAtot;
    ma.vec_out = (ws.actK[0] % 4 == 0 && out->stride_read % 4 == 0 && (R * M) % 4 == 0) ? 1 : 0;
    ma.prof = prof ? reinterpret_cast<long long*>(wsb + ws.off_prof) : nullptr;
    if (prof && (e = cudaMemsetAsync(wsb + ws.off_prof, 0, 8ll * 16 * B, stream)) != cudaSuccess)
      return set_cuda_error_ext(e, "cudaMemsetAsync(prof)");
    g_prof_ptr = ma.prof; g_prof_B = B;

    // programmatic dependent launch along the chain GEMM -> gates -> GEMM -> memory kernel (NTM_B200_EXP bit 16 = off;
    // with per-kernel profiling events between the launches there is nothing to overlap)
    const bool pdl = !(env.exp & 16) && !prof;
    for (long long t = 0; t < T; ++t) {
      const bool last = (t == T - 1);
      if (prof) cudaEventRecord(g_sev[2 + 4 * t + 0], stream);
      // ---- controller: per layer GEMM (tensor cores) + gates ----
      for (int l = 0; l < L; ++l) {
        const float* wl = w->lstm_w[l] + (l == 0 ? (size_t)s->input_dim * 4 * C : 0);
        if (ws_ok) {
          // (the first launch of the call reads weights packed by the kernels just before it: no early start)
          e = gemmws::launch(planA, tilesA, whiA, wloA, nullptr, partA, 4 * C, ws.slabA, B, stream, env.exp,
                             (env.exp & 8) ? gemm_prof_buffer(0) : nullptr, pdl && t > 0);
          count_launch();
          if (e != cudaSuccess) return set_cuda_error_ext(e, "gemm_ws(controller)");
        } else {
          int st = gemm_tc(act[l], ws.actK[l], wl, 4 * C, nullptr, partA, 4 * C, ws.slabA, B, ws.actK[l], 4 * C,
                           ws.ksA[l], nsm, stream);
          count_launch();
          if (st != 0) return set_cuda_error_ext(cudaGetLastError(), "gemm_tc(controller)");
        }
        LstmArgs la{};
        la.B = B; la.C = C; la.L = L; la.l = l; la.T = (int)T; la.t = (int)t; la.KS = ws_ok ? planA.kslices : ws.ksA[l]; la.slab = ws.slabA;
        if (ws_ok) {
          la.tilesA = tilesA; la.tilesC = tilesC; la.KAtotA = planA.KAtot; la.koffA = R * M; la.KAtotC = planC.KAtot;
        }
        la.xw = xw; la.bias = w->lstm_b[l]; la.part = partA;
        if (l == 0 && xw_partial && ws.xrem > 0) {
          la.xr = reinterpret_cast<const float*>(wsb + ws.off_xr); la.nrem = ws.xrem;
          la.wrem = w->lstm_w[0] + (size_t)ws.xK * 4 * C;
        }
        la.ctrl = out->controller_state; la.sctrl = out->stride_controller_state;
        la.act_self = act[l]; la.actK_self = ws.actK[l];
        la.act_next = (l + 1 < L) ? act[l + 1] : nullptr; la.actK_next = (l + 1 < L) ? ws.actK[l + 1] : 0;
        la.hZ = hist ? hist->z : nullptr; la.hC = hist ? hist->c : nullptr; la.hH = hist ? hist->h : nullptr;
        const long long tot = B * C;
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
        const bool v4 = !(env.exp & 4) && C % 4 == 0 && la.actK_self % 4 == 0 && (la.act_next == nullptr || la.actK_next % 4 == 0) &&
                        la.sctrl % 4 == 0 && al16(la.ctrl) && al16(la.xw) && al16(la.bias) && al16(la.part) && ws.slabA % 4 == 0 &&
                        al16(la.act_self) && al16(la.act_next) && al16(la.hZ) && al16(la.hC) && al16(la.hH) && tot / 4 < (1ll << 31) &&
                        al16(la.wrem);
        const bool lpdl = pdl && ws_ok;     // (the fallback GEMM before it is not part of the chain protocol)
        if (v4) e = launch_chain(lstm_stream_kernel_v4, (unsigned)((tot / 4 + 255) / 256), 256u, (size_t)0, stream, lpdl, la);
        else e = launch_chain(lstm_stream_kernel, (unsigned)((tot + 255) / 256), 256u, (size_t)0, stream, lpdl, la);
        count_launch();
        if (e != cudaSuccess) return set_cuda_error_ext(e, "lstm_stream_kernel");
      }
      if (prof) cudaEventRecord(g_sev[2 + 4 * t + 1], stream);
      // ---- head parameters + logits: one GEMM, bias folded in ----
      float* mc_t = hP ? hist->params + (size_t)t * B * PO4 : mcbuf;
      if (ws_ok) {
        e = gemmws::launch(planC, tilesC, whiC, wloC, bC, mc_t, PO4, 0, B, stream, env.exp,
                           (env.exp & 8) ? gemm_prof_buffer(1) : nullptr, pdl);
        count_launch();
        if (e != cudaSuccess) return set_cuda_error_ext(e, "gemm_ws(head parameters)");
      } else {
        int st = gemm_tc(act[L - 1] + (ws.actK[L - 1] - C), ws.actK[L - 1], wC, PO4, bC, mc_t, PO4, 0, B, C, PO4, 1,
                         nsm, stream);
        count_launch();
        if (st != 0) return set_cuda_error_ext(cudaGetLastError(), "gemm_tc(head parameters)");
      }
      if (prof) cudaEventRecord(g_sev[2 + 4 * t + 2], stream);
      // ---- fused addressing + memory update ----
      ma.t = (int)t; ma.mc = mc_t;
      ma.Min = hM ? hist->M_prev + (size_t)t * B * N * M : ((t == 0 && !cont) ? in->M : out->M);
      ma.sMin = hM ? (long long)N * M : ((t == 0 && !cont) ? in->stride_M : out->stride_M);
      ma.Mout = (hM && !last) ? hist->M_prev + (size_t)(t + 1) * B * N * M : out->M;
      ma.sMout = (hM && !last) ? (long long)N * M : out->stride_M;
      if (hW) {
        ma.w_in = hist->w_prev + (size_t)t * B * H * N; ma.sw_in = (long long)H * N;
        ma.w_out = last ? out->w : hist->w_prev + (size_t)(t + 1) * B * H * N;
        ma.sw_out = last ? out->stride_w : (long long)H * N;
      } else {
        ma.w_in = out->w; ma.sw_in = out->stride_w; ma.w_out = out->w; ma.sw_out = out->stride_w;
      }
      ma.sim_hist = (hist && hist->sim) ? hist->sim + (size_t)t * B * H * N : nullptr;
      ma.cn_hist = (hist && hist->cn) ? hist->cn + (size_t)t * B * M : nullptr;
      if (hist && hist->read) {
        ma.read_out = hist->read + (size_t)(t + 1) * B * R * M; ma.s_read = (long long)R * M;
      } else {
        ma.read_out = last ? out->read : nullptr; ma.s_read = out->stride_read;
      }
      g_chain_pdl = pdl && ws_ok;
      e = cpl ? launch_tma(R, W, cpl, ma, B, smem_tma, stream) : launch_mem(R, W, ma, B, smem, stream);
      count_launch();
      if (e != cudaSuccess) return set_cuda_error_ext(e, "mem_step_kernel");
      if (prof) cudaEventRecord(g_sev[2 + 4 * t + 3], stream);
    }
    if (hist && hist->read) {   // the state copy of the last read vectors
      e = cudaMemcpy2DAsync(out->read, 4 * out->stride_read, hist->read + (size_t)T * B * R * M, 4ll * R * M,
                            4ll * R * M, B, cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return set_cuda_error_ext(e, "cudaMemcpy2DAsync(read)");
    }
  }
  if (prof) g_sev_steps = (int)T;
  if (env.exp & 8) gemm_prof_dump(stream);
  return NTM_B200_OK;
}

int stream_mem_occupancy() { return g_mem_occ; }

// Mean duration (ns) of the memory kernel's phases over the CTAs of the last launch: {wait for the head
// parameters, activations, pass 1, addressing, pass 2, store drain, finalize, whole CTA}; out[8] = span of
// the launch (first CTA start to last CTA end).  Synchronous; call after the stream was synchronised.
int stream_phase_ns(double* out9) {
  for (int i = 0; i < 12; ++i) out9[i] = 0.0;
  if (g_prof_ptr == nullptr || g_prof_B <= 0) return 0;
  std::vector<long long> h((size_t)g_prof_B * 16);
  if (cudaMemcpy(h.data(), g_prof_ptr, h.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  long long first = h[0], last = h[7];
  for (long long b = 0; b < g_prof_B; ++b) {
    const long long* r = &h[(size_t)b * 16];
    for (int i = 0; i < 7; ++i) out9[i] += (double)(r[i + 1] - r[i]);
    out9[7] += (double)(r[7] - r[0]);
    first = std::min(first, r[0]); last = std::max(last, r[7]);
  }
  for (int i = 0; i < 8; ++i) out9[i] /= (double)g_prof_B;
  out9[8] = (double)(last - first);
  for (long long b = 0; b < g_prof_B; ++b)
    for (int i = 0; i < 3; ++i) out9[9 + i] += (double)h[(size_t)b * 16 + 8 + i] / (double)g_prof_B;
  return (int)g_prof_B;
}

int stream_last_ms(float* out4) {
  out4[0] = out4[1] = out4[2] = out4[3] = 0.0f;
  if (g_sev_steps <= 0) return 0;
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, g_sev[0], g_sev[1]);
  out4[3] = ms;
  for (int t = 0; t < g_sev_steps; ++t) {
    const cudaEvent_t* ev = &g_sev[2 + 4 * t];
    if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) out4[0] += ms;
    if (cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) out4[1] += ms;
    if (cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) out4[2] += ms;
  }
  return g_sev_steps;
}

}  // namespace ntm_b200
