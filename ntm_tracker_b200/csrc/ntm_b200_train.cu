// ntm_b200_train.cu -- reverse-time step of the NTM memory / addressing backward pass.
//
// The reference trains through tf.gradients of the unrolled while_loop
// (direct_offset_output.py:611-626); TensorFlow derives the backward graph op by op.  Here the
// backward of everything between the head-parameter projection and the memory (K3-K9 of
// column-normalised similarity, beta-softmax, gate, circular shift, sharpening,
// erase/add write, weighted read -- ntm_cell.py:133-215, ops.py:135-242) is ONE kernel per
// timestep, one CTA per sequence.  The dense projections' gradients are plain GEMMs and are left
// to the caller (ntm_tracker_b200/training.py).
//
// Per sequence and step, given  dL/dM_t, dL/dw_t, dL/dread_t  it recomputes the forward
// quantities from the recorded history (M_{t-1}, w_{t-1}, raw head parameters) and returns
// dL/dM_{t-1} (in place), dL/dw_{t-1} and dL/d(raw head parameters).
//
// Mapping: threads form a (column-chunk x row-group) tile over the N x M memory, which streams
// from HBM/L2 as coalesced float4 rows; row reductions go warp-shuffle -> shared memory, column
// reductions go registers -> shared memory, both summed in fixed order (deterministic).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>

#include "ntm_b200.h"
#include "ntm_b200_params.h"
#include "ntm_b200_train.h"
#include "ntm_b200_gemm_ws.cuh"

namespace ntm_b200 {
namespace train {

constexpr int NT = 256;
constexpr int NWARP = NT / 32;

struct BwdParams {
  int N, M, M4, MC, Np, S, shift0, R, W, H, P, PO4, write_first;
  int TPR, RG, TW;            // threads per row (multiple of 32), row groups, warps per row
  const float* sim_hist;      // [B, H, N] un-normalised similarities the forward pass recorded, or null
  const float* cn_hist;       // [B, M] inverse column norms the forward pass recorded, or null
  uint8_t* tiles_raw;         // operand tiles receiving d_raw (row operand of the d_h GEMM), or null
  int KAtot_raw;
  const float* M_prev;        // [B, N, M]
  const float* w_prev;        // [B, H, N]
  const float* raw;           // [B, PO4]
  const float* d_read;        // [B, R, M], sequences sdr floats apart
  long long sdr;
  const float* dlogits;       // [B, T, O] gradient w.r.t. the logits (fills the logit slots of d_raw), or null
  int T, t, O;
  const float* d_w;           // [B, H, N]
  float* dM;                  // [B, N, M] in: dL/dM_t, out: dL/dM_{t-1}
  float* d_w_prev;            // [B, H, N]
  float* d_raw;               // [B, PO4]
  // shared-memory offsets (floats)
  int oK, oKhat, oE, oA, oCn, oCok, oDkh, oDe, oDa, oCt;
  int oSim, oWc, oWg, oWt, oPw, oW, oDw, oDsim, oWp;
  int oRow, oCol, oSc;
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
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
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
// VEC: M % 4 == 0 and every row base 16-byte aligned (checked once on the host): plain vector accesses
template <bool VEC>
__device__ __forceinline__ float4 ld4t(const float* base, int M, int d0);
template <>
__device__ __forceinline__ float4 ld4t<true>(const float* base, int, int d0) {
  return __ldcg(reinterpret_cast<const float4*>(base + d0));
}
template <bool VEC>
__device__ __forceinline__ void st4t(float* base, int M, int d0, const float4& v);
template <>
__device__ __forceinline__ void st4t<true>(float* base, int, int d0, const float4& v) {
  __stcg(reinterpret_cast<float4*>(base + d0), v);
}
__device__ __forceinline__ float4 ld4(const float* base, int M, int d0) {
  // row pointer `base`, columns d0..d0+3 of a row of M valid floats (vector load when aligned)
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (((M & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) return *reinterpret_cast<const float4*>(base + d0);
  if (d0 + 0 < M) v.x = base[d0 + 0];
  if (d0 + 1 < M) v.y = base[d0 + 1];
  if (d0 + 2 < M) v.z = base[d0 + 2];
  if (d0 + 3 < M) v.w = base[d0 + 3];
  return v;
}
__device__ __forceinline__ void st4(float* base, int M, int d0, const float4& v) {
  if (((M & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) { *reinterpret_cast<float4*>(base + d0) = v; return; }
  if (d0 + 0 < M) base[d0 + 0] = v.x;
  if (d0 + 1 < M) base[d0 + 1] = v.y;
  if (d0 + 2 < M) base[d0 + 2] = v.z;
  if (d0 + 3 < M) base[d0 + 3] = v.w;
}

template <>
__device__ __forceinline__ float4 ld4t<false>(const float* base, int M, int d0) { return ld4(base, M, d0); }
template <>
__device__ __forceinline__ void st4t<false>(float* base, int M, int d0, const float4& v) { st4(base, M, d0, v); }

// Transposing warp reduction of NV <= 32 per-lane partial sums: at offset o a lane keeps one half of its value
// list and receives the partner's sums of that half; lane L ends with the warp total of value L (fixed order).
// 31 shuffles for up to 32 values instead of 5 per value.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
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
  return v[0];
}

// Thread mapping: thread = (16-byte column chunk c = tid % TPR, row group rg = tid / TPR); a warp covers 32
// consecutive chunks of one row group (TW = TPR / 32 warps per row).  Every sweep over the memory handles a
// QUAD of four consecutive rows per iteration (four independent loads in flight; 4 * H <= 28 row partials per
// lane, combined by one transposing reduction per quad).  Two CTAs per SM.
template <int R, int W, bool VEC>
__global__ void __launch_bounds__(NT, 2) mem_backward_kernel(const BwdParams q) {
  pdl_trigger();      // reverse-time loop chained by programmatic dependent launch: the launch latency of the next
  pdl_wait();         // kernel hides behind this one; nothing is touched before the predecessor has completed
  constexpr int H = R + W;
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int N = q.N, M = q.M, M4 = q.M4, MC = q.MC, Np = q.Np, S = q.S;
  const int c = tid % q.TPR, rg = tid / q.TPR;      // column chunk, row group
  const bool cvalid = c < MC && rg < q.RG;
  const int wrow = (tid % q.TPR) >> 5;              // warp index within the row
  const int NQ = (N + 3) >> 2;                      // quads of rows
  float* kS = sm + q.oK;   float* khat = sm + q.oKhat;
  float* eS = sm + q.oE;   float* aS = sm + q.oA;      float* cn = sm + q.oCn;  float* cok = sm + q.oCok;
  float* dkh = sm + q.oDkh; float* deS = sm + q.oDe;   float* daS = sm + q.oDa; float* ct = sm + q.oCt;
  float* sim = sm + q.oSim; float* wc = sm + q.oWc; float* wg = sm + q.oWg; float* wt = sm + q.oWt;
  float* pw = sm + q.oPw;   float* wv = sm + q.oW;  float* dw = sm + q.oDw; float* dsim = sm + q.oDsim;
  float* wp = sm + q.oWp;
  float* rowbuf = sm + q.oRow;                       // [N][TW][H]
  float* colbuf = sm + q.oCol;                       // [RG][NQC][M4]
  float* sc = sm + q.oSc;                            // scalars
  float* sBeta = sc, *sG = sc + H, *sGam = sc + 2 * H, *sRs = sc + 3 * H, *sKok = sc + 4 * H,
        *sSw = sc + 5 * H, *sDsw = sc + 5 * H + H * SMAX, *sDbeta = sc + 5 * H + 2 * H * SMAX,
        *sDg = sDbeta + H, *sDgam = sDg + H, *sSum = sDgam + H;
  const float* Mb = q.M_prev + (size_t)b * N * M;
  float* dMb = q.dM + (size_t)b * N * M;
  const float* raw = q.raw + (size_t)b * q.PO4;
  const int offBeta = H * M, offG = offBeta + H, offS = offG + H, offGam = offS + S * H,
            offE = offGam + H, offA = offE + M * W;
  const bool recorded = q.sim_hist != nullptr && q.cn_hist != nullptr;

  // ---- (1) activations of the recorded raw head parameters ----
  for (int i = tid; i < H * M4; i += NT) {
    const int h = i / M4, d = i - h * M4;
    kS[i] = (d < M) ? tanhf(raw[h * M + d]) : 0.0f;
  }
  for (int i = tid; i < W * M4; i += NT) {
    const int h = i / M4, d = i - h * M4;
    eS[i] = (d < M) ? sigmoid_f(raw[offE + h * M + d]) : 0.0f;
    aS[i] = (d < M) ? tanhf(raw[offA + h * M + d]) : 0.0f;
  }
  if (tid < H) {
    sBeta[tid] = softplus_f(raw[offBeta + tid]);
    sG[tid] = sigmoid_f(raw[offG + tid]);
    sGam[tid] = 1.0f + softplus_f(raw[offGam + tid]);
    float mx = raw[offS + tid * S];
    for (int i = 1; i < S; ++i) mx = fmaxf(mx, raw[offS + tid * S + i]);
    float sum = 0.0f;
    for (int i = 0; i < S; ++i) { sSw[tid * SMAX + i] = expf(raw[offS + tid * S + i] - mx); sum += sSw[tid * SMAX + i]; }
    for (int i = 0; i < S; ++i) sSw[tid * SMAX + i] /= sum;
  }
  for (int i = tid; i < H * N; i += NT) {
    const int h = i / N, n = i - h * N;
    wp[h * Np + n] = q.w_prev[(size_t)b * H * N + i];
    dw[h * Np + n] = q.d_w[(size_t)b * H * N + i];
    if (recorded) sim[h * Np + n] = q.sim_hist[(size_t)b * H * N + i];   // un-normalised: times 1/|k| below
  }
  if (recorded) {
    // inverse column norms as the forward pass computed them; the clamp of tf.nn.l2_normalize
    // (sum of squares <= 1e-12, ops.py:147-150) shows as cn >= 1e6 and cuts the gradient
    for (int d = tid; d < M4; d += NT) {
      const float v = d < M ? q.cn_hist[(size_t)b * M + d] : 0.0f;
      cn[d] = v;
      cok[d] = (d < M && v < 0.999e6f) ? 1.0f : 0.0f;
    }
  }
  __syncthreads();

  // ---- (2) F1: column sums of squares -> cn (only when the forward pass did not record them) ----
  if (!recorded) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cvalid)
      for (int n = rg; n < N; n += q.RG) {
        const float4 m = ld4(Mb + (size_t)n * M, M, 4 * c);
        acc.x = fmaf(m.x, m.x, acc.x); acc.y = fmaf(m.y, m.y, acc.y);
        acc.z = fmaf(m.z, m.z, acc.z); acc.w = fmaf(m.w, m.w, acc.w);
      }
    if (cvalid) *reinterpret_cast<float4*>(colbuf + (size_t)rg * M4 + 4 * c) = acc;
    __syncthreads();
    for (int d = tid; d < M4; d += NT) {
      float s = 0.0f;
      for (int r2 = 0; r2 < q.RG; ++r2) s += colbuf[(size_t)r2 * M4 + d];
      cn[d] = 1.0f / sqrtf(fmaxf(s, 1e-12f));
      cok[d] = (s > 1e-12f) ? 1.0f : 0.0f;
    }
  }
  for (int h = warp; h < H; h += NWARP) {
    float s = 0.0f;
    for (int d = lane; d < M; d += 32) s = fmaf(kS[h * M4 + d], kS[h * M4 + d], s);
    s = warp_sum(s);
    if (lane == 0) { sRs[h] = 1.0f / sqrtf(fmaxf(s, 1e-12f)); sKok[h] = (s > 1e-12f) ? 1.0f : 0.0f; }
  }
  __syncthreads();
  for (int i = tid; i < H * M4; i += NT) khat[i] = kS[i] * sRs[i / M4];
  if (recorded)
    for (int i = tid; i < H * N; i += NT) {
      const int h = i / N, n = i - h * N;
      sim[h * Np + n] *= sRs[h];
    }
  __syncthreads();

  // ---- (3) F2: similarities (row reductions), only when not recorded ----
  if (!recorded) {
    float4 kc4[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      kc4[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cvalid) {
        const float4 kh = *reinterpret_cast<const float4*>(khat + h * M4 + 4 * c);
        const float4 c4 = *reinterpret_cast<const float4*>(cn + 4 * c);
        kc4[h] = make_float4(kh.x * c4.x, kh.y * c4.y, kh.z * c4.z, kh.w * c4.w);
      }
    }
    for (int q0 = 0; q0 < NQ; q0 += q.RG) {
      const int n0 = 4 * (q0 + rg);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.0f;
      if (cvalid && n0 < N) {
        float4 m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = ld4t<VEC>(Mb + (size_t)min(n0 + i, N - 1) * M, M, 4 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int h = 0; h < H; ++h) v[i * H + h] = dot4(m[i], kc4[h]);
      }
      const float tot = warp_transpose_sum(v, lane);
      const int vi = lane / H, vh = lane - vi * H;
      if (lane < 4 * H && rg < q.RG && n0 + vi < N) rowbuf[((size_t)(n0 + vi) * q.TW + wrow) * H + vh] = tot;
    }
    __syncthreads();
    for (int i = tid; i < H * N; i += NT) {
      const int h = i / N, n = i - h * N;
      float s = 0.0f;
      for (int w2 = 0; w2 < q.TW; ++w2) s += rowbuf[((size_t)n * q.TW + w2) * H + h];
      sim[h * Np + n] = s;
    }
    __syncthreads();
  }

  // ---- (4) forward weightings, one warp per head (ntm_cell.py:140-176) ----
  for (int h = warp; h < H; h += NWARP) {
    const float beta = sBeta[h], g = sG[h], gamma = sGam[h];
    float mx = -INFINITY;
    for (int n = lane; n < N; n += 32) mx = fmaxf(mx, sim[h * Np + n] * beta);
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int n = lane; n < N; n += 32) { const float e = expf(sim[h * Np + n] * beta - mx); wc[h * Np + n] = e; sum += e; }
    sum = warp_sum(sum);
    for (int n = lane; n < N; n += 32) {
      const float v = wc[h * Np + n] / sum;
      wc[h * Np + n] = v;
      wg[h * Np + n] = v * g + wp[h * Np + n] * (1.0f - g);
    }
    __syncwarp();
    float psum = 0.0f;
    for (int n = lane; n < N; n += 32) {
      float conv = 0.0f;
      for (int s = 0; s < S; ++s) {
        int idx = n + q.shift0 + s;
        idx = idx < 0 ? idx + N : (idx >= N ? idx - N : idx);
        conv = fmaf(sSw[h * SMAX + s], wg[h * Np + idx], conv);
      }
      const float pv = exp2f(gamma * log2f(conv));
      wt[h * Np + n] = conv;
      pw[h * Np + n] = pv;
      psum += pv;
    }
    psum = warp_sum(psum);
    const float den = psum + 1e-3f;
    if (lane == 0) sSum[h] = den;
    for (int n = lane; n < N; n += 32) wv[h * Np + n] = pw[h * Np + n] / den;
  }
  __syncthreads();

  // ---- (5) B1: read + write backward; row sums d_w(read|write), column sums d_e, d_a ----
  {
    float4 dr4[R], e4[W], a4[W], dea[W], daa[W];
#pragma unroll
    for (int r = 0; r < R; ++r)
      dr4[r] = cvalid ? ld4(q.d_read + (size_t)b * q.sdr + (size_t)r * M, M, 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < W; ++h) {
      e4[h] = cvalid ? *reinterpret_cast<const float4*>(eS + h * M4 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      a4[h] = cvalid ? *reinterpret_cast<const float4*>(aS + h * M4 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      dea[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      daa[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // Software pipeline over PAIRS of rows: while a pair is being worked on, the loads of the next pair
    // (second half of the quad, then the first half of the next quad) are already in flight.
    float v[32];
    auto load_pair = [&](int n0p, float4 (&m)[2], float4 (&d)[2]) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const size_t ro = (size_t)min(n0p + i, N - 1) * M;
        m[i] = ld4t<VEC>(Mb + ro, M, 4 * c);
        d[i] = ld4t<VEC>(dMb + ro, M, 4 * c);
      }
    };
    auto compute_pair = [&](int n0p, const float4 (&m2)[2], const float4 (&d2)[2], int vbase) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n = n0p + i;
        if (n < N) {
          const float4 m = m2[i], dmn = d2[i];
          float ww[W];
          float4 F[W];
          float4 E = make_float4(1.f, 1.f, 1.f, 1.f), A = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int h = 0; h < W; ++h) {
            ww[h] = wv[(R + h) * Np + n];
            F[h] = make_float4(1.0f - ww[h] * e4[h].x, 1.0f - ww[h] * e4[h].y, 1.0f - ww[h] * e4[h].z, 1.0f - ww[h] * e4[h].w);
            E.x *= F[h].x; E.y *= F[h].y; E.z *= F[h].z; E.w *= F[h].w;
            A.x = fmaf(ww[h], a4[h].x, A.x); A.y = fmaf(ww[h], a4[h].y, A.y);
            A.z = fmaf(ww[h], a4[h].z, A.z); A.w = fmaf(ww[h], a4[h].w, A.w);
          }
          const float4 mn = make_float4(fmaf(m.x, E.x, A.x), fmaf(m.y, E.y, A.y), fmaf(m.z, E.z, A.z), fmaf(m.w, E.w, A.w));
          const float4 mu = q.write_first ? mn : m;
          float4 dmu = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float wr = wv[r * Np + n];
            dmu.x = fmaf(wr, dr4[r].x, dmu.x); dmu.y = fmaf(wr, dr4[r].y, dmu.y);
            dmu.z = fmaf(wr, dr4[r].z, dmu.z); dmu.w = fmaf(wr, dr4[r].w, dmu.w);
            v[vbase + i * H + r] = dot4(dr4[r], mu);
          }
          float4 dt = dmn;                                // dL/dM_t including the read path if write_first
          if (q.write_first) { dt.x += dmu.x; dt.y += dmu.y; dt.z += dmu.z; dt.w += dmu.w; }
          const float4 dE = make_float4(dt.x * m.x, dt.y * m.y, dt.z * m.z, dt.w * m.w);
#pragma unroll
          for (int h = 0; h < W; ++h) {
            float4 pex = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
            for (int o = 0; o < W; ++o)
              if (o != h) { pex.x *= F[o].x; pex.y *= F[o].y; pex.z *= F[o].z; pex.w *= F[o].w; }
            const float4 dF = make_float4(dE.x * pex.x, dE.y * pex.y, dE.z * pex.z, dE.w * pex.w);
            v[vbase + i * H + R + h] = dot4(dt, a4[h]) - dot4(dF, e4[h]);
            dea[h].x -= dF.x * ww[h]; dea[h].y -= dF.y * ww[h]; dea[h].z -= dF.z * ww[h]; dea[h].w -= dF.w * ww[h];
            daa[h].x = fmaf(dt.x, ww[h], daa[h].x); daa[h].y = fmaf(dt.y, ww[h], daa[h].y);
            daa[h].z = fmaf(dt.z, ww[h], daa[h].z); daa[h].w = fmaf(dt.w, ww[h], daa[h].w);
          }
          float4 dout = make_float4(dt.x * E.x, dt.y * E.y, dt.z * E.z, dt.w * E.w);
          if (!q.write_first) { dout.x += dmu.x; dout.y += dmu.y; dout.z += dmu.z; dout.w += dmu.w; }
          st4t<VEC>(dMb + (size_t)n * M, M, 4 * c, dout);
        }
      }
    };
    float4 mA[2], dA[2], mB[2], dB[2];
    if (cvalid && 4 * rg < N) load_pair(4 * rg, mA, dA);
    for (int q0 = 0; q0 < NQ; q0 += q.RG) {
      const int n0 = 4 * (q0 + rg);
      const bool active = cvalid && n0 < N;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.0f;
      if (active) {
        load_pair(n0 + 2, mB, dB);
        compute_pair(n0, mA, dA, 0);
        const int nn0 = n0 + 4 * q.RG;
        if (nn0 < N) load_pair(nn0, mA, dA);
        compute_pair(n0 + 2, mB, dB, 2 * H);
      }
      const float tot = warp_transpose_sum(v, lane);
      const int vi = lane / H, vh = lane - vi * H;
      if (lane < 4 * H && rg < q.RG && n0 + vi < N) rowbuf[((size_t)(n0 + vi) * q.TW + wrow) * H + vh] = tot;
    }
    if (cvalid) {
#pragma unroll
      for (int h = 0; h < W; ++h) {
        *reinterpret_cast<float4*>(colbuf + ((size_t)rg * 2 * W + h) * M4 + 4 * c) = dea[h];
        *reinterpret_cast<float4*>(colbuf + ((size_t)rg * 2 * W + W + h) * M4 + 4 * c) = daa[h];
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < H * N; i += NT) {               // total dL/dw_t
    const int h = i / N, n = i - h * N;
    float s = dw[h * Np + n];
    for (int w2 = 0; w2 < q.TW; ++w2) s += rowbuf[((size_t)n * q.TW + w2) * H + h];
    dw[h * Np + n] = s;
  }
  for (int i = tid; i < W * M4; i += NT) {
    const int h = i / M4, d = i - h * M4;
    float se = 0.0f, sa = 0.0f;
    for (int r2 = 0; r2 < q.RG; ++r2) {
      se += colbuf[((size_t)r2 * 2 * W + h) * M4 + d];
      sa += colbuf[((size_t)r2 * 2 * W + W + h) * M4 + d];
    }
    deS[i] = se;
    daS[i] = sa;
  }
  __syncthreads();

  // ---- (6) weightings backward, one warp per head ----
  for (int h = warp; h < H; h += NWARP) {
    const float beta = sBeta[h], g = sG[h], gamma = sGam[h], den = sSum[h];
    float* dwt = dsim + h * Np;                          // scratch: dL/dw~, later dL/dsim
    float t1 = 0.0f;
    for (int n = lane; n < N; n += 32) t1 = fmaf(dw[h * Np + n], pw[h * Np + n], t1);
    t1 = warp_sum(t1);
    float dgam = 0.0f;
    for (int n = lane; n < N; n += 32) {
      const float dp = dw[h * Np + n] / den - t1 / (den * den);
      const float x = wt[h * Np + n], pv = pw[h * Np + n];
      float d = 0.0f;
      if (x > 0.0f) {
        d = dp * gamma * pv / x;
        dgam = fmaf(dp * pv, logf(x), dgam);
      }
      dwt[n] = d;
    }
    dgam = warp_sum(dgam);
    __syncwarp();
    float dgate = 0.0f, t2 = 0.0f;
    float dsw[SMAX];
#pragma unroll
    for (int s = 0; s < SMAX; ++s) dsw[s] = 0.0f;
    for (int n = lane; n < N; n += 32) {
      float dwg = 0.0f;
#pragma unroll
      for (int s = 0; s < SMAX; ++s) {
        if (s < S) {
          const int off = q.shift0 + s;
          int src = n - off;                             // w~[src] reads wg[(src + off) mod N] = wg[n]
          src = src < 0 ? src + N : (src >= N ? src - N : src);
          dwg = fmaf(sSw[h * SMAX + s], dwt[src], dwg);
          int idx = n + off;
          idx = idx < 0 ? idx + N : (idx >= N ? idx - N : idx);
          dsw[s] = fmaf(dwt[n], wg[h * Np + idx], dsw[s]);
        }
      }
      const float wcv = wc[h * Np + n], wpv = wp[h * Np + n];
      dgate = fmaf(dwg, wcv - wpv, dgate);
      q.d_w_prev[(size_t)b * H * N + h * N + n] = (1.0f - g) * dwg;
      const float dwc = g * dwg;
      pw[h * Np + n] = dwc;                              // pw is dead: reuse for dL/dwc
      t2 = fmaf(dwc, wcv, t2);
    }
    dgate = warp_sum(dgate);
    t2 = warp_sum(t2);
#pragma unroll
    for (int s = 0; s < SMAX; ++s) dsw[s] = warp_sum(dsw[s]);
    __syncwarp();
    float dbeta = 0.0f;
    for (int n = lane; n < N; n += 32) {
      const float dx = wc[h * Np + n] * (pw[h * Np + n] - t2);
      dbeta = fmaf(dx, sim[h * Np + n], dbeta);
      dwt[n] = beta * dx;                                // dL/dsim
    }
    dbeta = warp_sum(dbeta);
    if (lane == 0) {
      sDbeta[h] = dbeta; sDg[h] = dgate; sDgam[h] = dgam;
      float dot = 0.0f;
      for (int s = 0; s < S; ++s) dot = fmaf(dsw[s], sSw[h * SMAX + s], dot);
      for (int s = 0; s < S; ++s) sDsw[h * SMAX + s] = sSw[h * SMAX + s] * (dsw[s] - dot);   // through the softmax
    }
  }
  __syncthreads();

  // ---- (7) B2: column sums  A[h][d] = sum_n dsim[h][n] M[n][d]  ->  dL/dkhat = A * cn,  ct = sum_h khat * A ----
  {
    float4 acc[H];
#pragma unroll
    for (int i = 0; i < H; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cvalid) {
      for (int q0 = rg; q0 < NQ; q0 += q.RG) {
        const int n0 = 4 * q0;
        float4 m4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) m4[i] = ld4t<VEC>(Mb + (size_t)min(n0 + i, N - 1) * M, M, 4 * c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (n0 + i < N) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
              const float ds = dsim[h * Np + n0 + i];
              acc[h].x = fmaf(ds, m4[i].x, acc[h].x); acc[h].y = fmaf(ds, m4[i].y, acc[h].y);
              acc[h].z = fmaf(ds, m4[i].z, acc[h].z); acc[h].w = fmaf(ds, m4[i].w, acc[h].w);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < H; ++i) *reinterpret_cast<float4*>(colbuf + ((size_t)rg * H + i) * M4 + 4 * c) = acc[i];
    }
    __syncthreads();
    for (int d = tid; d < M4; d += NT) {
      float ctd = 0.0f;
      const float cnd = cn[d];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        float s = 0.0f;
        for (int r2 = 0; r2 < q.RG; ++r2) s += colbuf[((size_t)r2 * H + h) * M4 + d];
        const float sc2 = s * cnd;                        // sum_n dsim[h][n] * M[n][d] * cn[d]
        dkh[h * M4 + d] = sc2;
        ctd = fmaf(khat[h * M4 + d], sc2, ctd);
      }
      ct[d] = ctd;                                        // sum_n (sum_h dsim khat) * M * cn  (the cn factor folded in)
    }
    __syncthreads();
  }

  // ---- (8) B3: dL/dM_{t-1} += similarity path (through the column normalisation) ----
  if (cvalid) {
    const float4 cn4 = *reinterpret_cast<const float4*>(cn + 4 * c);
    const float4 ok4 = *reinterpret_cast<const float4*>(cok + 4 * c);
    const float4 ct4 = *reinterpret_cast<const float4*>(ct + 4 * c);
    float4 kh4[H];
#pragma unroll
    for (int h = 0; h < H; ++h) kh4[h] = *reinterpret_cast<const float4*>(khat + h * M4 + 4 * c);
    // d(M cn)/dM: cn * dmh - M * cn^3 * sum_n(dmh M); ct already carries one factor cn
    const float4 c3 = make_float4(ok4.x * cn4.x * cn4.x * ct4.x, ok4.y * cn4.y * cn4.y * ct4.y,
                                  ok4.z * cn4.z * cn4.z * ct4.z, ok4.w * cn4.w * cn4.w * ct4.w);
    for (int q0 = rg; q0 < NQ; q0 += q.RG) {
      const int n0 = 4 * q0;
      float4 m4[4], d4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = min(n0 + i, N - 1);
        m4[i] = ld4t<VEC>(Mb + (size_t)n * M, M, 4 * c);
        d4[i] = ld4t<VEC>(dMb + (size_t)n * M, M, 4 * c);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + i;
        if (n < N) {
          float4 dmh = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const float ds = dsim[h * Np + n];
            dmh.x = fmaf(ds, kh4[h].x, dmh.x); dmh.y = fmaf(ds, kh4[h].y, dmh.y);
            dmh.z = fmaf(ds, kh4[h].z, dmh.z); dmh.w = fmaf(ds, kh4[h].w, dmh.w);
          }
          float4 d = d4[i];
          d.x += cn4.x * dmh.x - m4[i].x * c3.x; d.y += cn4.y * dmh.y - m4[i].y * c3.y;
          d.z += cn4.z * dmh.z - m4[i].z * c3.z; d.w += cn4.w * dmh.w - m4[i].w * c3.w;
          st4t<VEC>(dMb + (size_t)n * M, M, 4 * c, d);
        }
      }
    }
  }

  // ---- (9) back through the activations -> dL/d(raw head parameters) ----
  float* draw = q.d_raw + (size_t)b * q.PO4;
  for (int h = warp; h < H; h += NWARP) {             // key: through the l2 normalisation, then tanh
    float dot = 0.0f;
    for (int d = lane; d < M; d += 32) dot = fmaf(khat[h * M4 + d], dkh[h * M4 + d], dot);
    dot = warp_sum(dot) * sKok[h];
    for (int d = lane; d < M; d += 32) {
      const float k = kS[h * M4 + d];
      const float dk = sRs[h] * (dkh[h * M4 + d] - khat[h * M4 + d] * dot);
      draw[h * M + d] = dk * (1.0f - k * k);
    }
  }
  for (int i = tid; i < W * M; i += NT) {
    const int h = i / M, d = i - h * M;
    const float e = eS[h * M4 + d], a = aS[h * M4 + d];
    draw[offE + i] = deS[h * M4 + d] * e * (1.0f - e);
    draw[offA + i] = daS[h * M4 + d] * (1.0f - a * a);
  }
  if (tid < H) {
    draw[offBeta + tid] = sDbeta[tid] * sigmoid_f(raw[offBeta + tid]);
    draw[offG + tid] = sDg[tid] * sG[tid] * (1.0f - sG[tid]);
    draw[offGam + tid] = sDgam[tid] * sigmoid_f(raw[offGam + tid]);
    for (int s = 0; s < S; ++s) draw[offS + tid * S + s] = sDsw[tid * SMAX + s];
  }
  // logit slots: the loss gradient w.r.t. this step's logits when the caller hands it over (the projection
  // h @ [W_addr | W_out] is ONE GEMM, so its backward wants them in the same row), else zero
  for (int i = q.P + tid; i < q.PO4; i += NT) {
    const int o = i - q.P;
    draw[i] = (q.dlogits != nullptr && o < q.O) ? q.dlogits[((size_t)b * q.T + q.t) * q.O + o] : 0.0f;
  }
  if (q.tiles_raw != nullptr) {
    // the row operand of d_h = d_raw @ [W_addr | W_out]^T: this sequence's d_raw row as bf16 hi/lo tile records
    // (the CTA re-reads the row it has just written; 8 consecutive entries per 16-byte chunk)
    __syncthreads();
    for (int ch = tid; 8 * ch < q.PO4; ch += NT) {
      float v8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v8[e] = (8 * ch + e < q.PO4) ? draw[8 * ch + e] : 0.0f;
      gemmws::store_split8(q.tiles_raw, q.KAtot_raw, b, 8 * ch, v8);
    }
  }
}

// LSTM cell backward for one layer and one timestep, all sequences: elementwise over [B, C].
// BasicLSTMCell forward (TF 1.0/1.1): c' = c*sig(f) + sig(i)*tanh(j); h' = tanh(c')*sig(o).
__global__ void __launch_bounds__(256) lstm_backward_kernel(int B, int C, const float* __restrict__ dh_a, long long lda,
                                     const float* __restrict__ dh_b, long long ldb, int nslab, long long slab,
                                     const float* __restrict__ z, long long z_stride,
                                     const float* __restrict__ c_prev, const float* __restrict__ c_new,
                                     long long c_stride, float* __restrict__ dc, float* __restrict__ dz,
                                     long long dz_stride, uint8_t* tiles, int KAtot) {
  pdl_trigger();
  pdl_wait();
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < (long long)B * C;
  const long long i = live ? i0 : 0;            // tail lanes recompute element 0 and store nothing
  const int b = (int)(i / C), u = (int)(i - (long long)b * C);
  const float* zb = z + (size_t)b * z_stride;
  const float gi = sigmoid_f(zb[u]), gj = tanhf(zb[C + u]), gf = sigmoid_f(zb[2 * C + u]), go = sigmoid_f(zb[3 * C + u]);
  const float cp = c_prev[(size_t)b * c_stride + u], tc = tanhf(c_new[(size_t)b * c_stride + u]);
  float dh = dh_a != nullptr ? dh_a[(size_t)b * lda + u] : 0.0f;
  if (dh_b != nullptr)
    for (int s2 = 0; s2 < nslab; ++s2) dh += dh_b[(size_t)s2 * slab + (size_t)b * ldb + u];   // K-slice slabs, slice order
  const float dct = dc[i] + dh * go * (1.0f - tc * tc);
  float g4[4];
  g4[0] = dct * gj * gi * (1.0f - gi);
  g4[1] = dct * gi * (1.0f - gj * gj);
  g4[2] = dct * cp * gf * (1.0f - gf);
  g4[3] = dh * tc * go * (1.0f - go);
  if (live) {
    float* dzb = dz + (size_t)b * dz_stride;
#pragma unroll
    for (int g = 0; g < 4; ++g) dzb[g * C + u] = g4[g];
    dc[i] = dct * gf;
  }
  if (tiles != nullptr) {
    // row operand of d_cat = d_z @ W^T as bf16 hi/lo tile records (C % 8 == 0: aligned groups of 8 lanes hold 8
    // consecutive units of one sequence; lane g of a group stores gate g's chunk, k = g * C + u)
    const int lane = threadIdx.x & 31, l0 = lane & ~7;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v8[e] = __shfl_sync(0xffffffffu, g4[g], l0 + e);
      if (live && (lane & 7) == g) gemmws::store_split8(tiles, KAtot, b, g * C + (u - g), v8);
    }
  }
}

typedef void (*BwdKernel)(const BwdParams);
#define NTM_BK(R, W) {mem_backward_kernel<R, W, false>, mem_backward_kernel<R, W, true>}
static BwdKernel select_bwd(int R, int W, bool vec) {
  static const BwdKernel table[NTM_B200_MAX_READ_HEADS][NTM_B200_MAX_WRITE_HEADS][2] = {
      {NTM_BK(1, 1), NTM_BK(1, 2), NTM_BK(1, 3)},
      {NTM_BK(2, 1), NTM_BK(2, 2), NTM_BK(2, 3)},
      {NTM_BK(3, 1), NTM_BK(3, 2), NTM_BK(3, 3)},
      {NTM_BK(4, 1), NTM_BK(4, 2), NTM_BK(4, 3)}};
  return table[R - 1][W - 1][vec ? 1 : 0];
}


int launch_memory_backward(const ntm_b200_shape* s, long long batch, const float* M_prev, const float* w_prev,
                           const float* raw_params, const float* d_read, long long sdr, const float* d_w, float* dM,
                           float* d_w_prev, float* d_raw_params, const float* dlogits, int T, int t,
                           const float* sim_hist, const float* cn_hist, uint8_t* tiles_raw, int KAtot_raw,
                           cudaStream_t stream, bool pdl) {
  BwdParams q{};
  q.sim_hist = sim_hist; q.cn_hist = cn_hist; q.tiles_raw = tiles_raw; q.KAtot_raw = KAtot_raw;
  const int R = s->read_head_size, W = s->write_head_size, H = R + W;
  q.N = s->mem_size; q.M = s->mem_dim; q.M4 = (q.M + 3) / 4 * 4; q.MC = q.M4 / 4; q.Np = (q.N + 3) / 4 * 4;
  q.S = 2 * s->shift_range + 1; q.shift0 = -((q.S + 1) / 2); q.R = R; q.W = W; q.H = H;
  q.P = H * q.M + 3 * H + q.S * H + 2 * q.M * W;
  q.PO4 = (q.P + s->output_dim + 3) / 4 * 4;
  q.write_first = s->write_first ? 1 : 0;
  q.TPR = std::min(NT, (q.MC + 31) / 32 * 32);
  if (q.MC > NT) return NTM_B200_ERR_TOO_LARGE;
  q.RG = NT / q.TPR; q.TW = q.TPR / 32;
  q.M_prev = M_prev; q.w_prev = w_prev; q.raw = raw_params; q.d_read = d_read; q.sdr = sdr; q.d_w = d_w; q.dM = dM;
  q.d_w_prev = d_w_prev; q.d_raw = d_raw_params;
  q.dlogits = dlogits; q.T = T; q.t = t; q.O = s->output_dim;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
  q.oK = take(H * q.M4); q.oKhat = take(H * q.M4); q.oE = take(W * q.M4); q.oA = take(W * q.M4);
  q.oCn = take(q.M4); q.oCok = take(q.M4); q.oDkh = take(H * q.M4); q.oDe = take(W * q.M4); q.oDa = take(W * q.M4);
  q.oCt = take(q.M4);
  q.oSim = take(H * q.Np); q.oWc = take(H * q.Np); q.oWg = take(H * q.Np); q.oWt = take(H * q.Np); q.oPw = take(H * q.Np);
  q.oW = take(H * q.Np); q.oDw = take(H * q.Np); q.oDsim = take(H * q.Np); q.oWp = take(H * q.Np);
  q.oRow = take(q.N * q.TW * H);
  q.oCol = take(q.RG * std::max(2 * W, H) * q.M4);
  q.oSc = take(5 * H + 2 * H * SMAX + 4 * H);
  const int smem = 4 * o;
  if (smem > B200_SMEM_OPTIN) return NTM_B200_ERR_TOO_LARGE;
  const bool vec = (q.M % 4 == 0) && (sdr % 4 == 0) && ((reinterpret_cast<uintptr_t>(M_prev) | reinterpret_cast<uintptr_t>(dM) |
                                                         reinterpret_cast<uintptr_t>(d_read)) & 15) == 0;
  BwdKernel k = select_bwd(R, W, vec);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    cudaGetLastError();
    return NTM_B200_ERR_CUDA;
  }
  const cudaError_t le = launch_chain(k, (unsigned)batch, (unsigned)NT, (size_t)smem, stream, pdl, q);
  count_launch();
  return le == cudaSuccess ? NTM_B200_OK : NTM_B200_ERR_CUDA;
}

int launch_lstm_backward(long long batch, int hidden, const float* dh_a, long long lda, const float* dh_b, long long ldb,
                         int nslab, long long slab, const float* z, long long z_stride, const float* c_prev,
                         const float* c_new, long long c_stride, float* dc, float* dz, long long dz_stride,
                         uint8_t* tiles, int KAtot, cudaStream_t stream, bool pdl) {
  const long long total = batch * hidden;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (hidden % 8 != 0) tiles = nullptr;      // the caller packs d_z with the generic kernel instead
  const cudaError_t le = launch_chain(lstm_backward_kernel, blocks, 256u, (size_t)0, stream, pdl, (int)batch, hidden, dh_a, lda, dh_b,
                                      ldb, nslab, slab, z, z_stride, c_prev, c_new, c_stride, dc, dz, dz_stride, tiles, KAtot);
  count_launch();
  return le == cudaSuccess ? NTM_B200_OK : NTM_B200_ERR_CUDA;
}

}  // namespace train
}  // namespace ntm_b200

namespace {
bool on_sm100() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      major != 10) {
    cudaGetLastError();
    return false;
  }
  return true;
}
}  // namespace

extern "C" int32_t ntm_b200_memory_backward_step(const ntm_b200_shape* s, int64_t batch, const float* M_prev,
                                                 const float* w_prev, const float* raw_params,
                                                 const float* d_read, const float* d_w, float* dM,
                                                 float* d_w_prev, float* d_raw_params, void* stream) {
  if (!s || !M_prev || !w_prev || !raw_params || !d_read || !d_w || !dM || !d_w_prev || !d_raw_params)
    return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || s->mem_size < 1 || s->mem_dim < 1) return NTM_B200_ERR_BAD_SHAPE;
  if (s->read_head_size < 1 || s->read_head_size > NTM_B200_MAX_READ_HEADS || s->write_head_size < 1 ||
      s->write_head_size > NTM_B200_MAX_WRITE_HEADS)
    return NTM_B200_ERR_UNSUPPORTED_HEADS;
  if (s->shift_range < 0 || s->shift_range > NTM_B200_MAX_SHIFT_RANGE) return NTM_B200_ERR_BAD_SHIFT;
  if (!on_sm100()) return NTM_B200_ERR_NO_DEVICE;
  return ntm_b200::train::launch_memory_backward(s, batch, M_prev, w_prev, raw_params, d_read,
                                                 (long long)s->read_head_size * s->mem_dim, d_w, dM, d_w_prev,
                                                 d_raw_params, nullptr, 1, 0, nullptr, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int32_t ntm_b200_lstm_backward_step(int64_t batch, int32_t hidden, const float* dh_a, const float* dh_b,
                                               const float* z, int64_t z_stride, const float* c_prev,
                                               const float* c_new, int64_t c_stride, float* dc, float* dz,
                                               int64_t dz_stride, void* stream) {
  if (!dh_a || !z || !c_prev || !c_new || !dc || !dz) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || hidden < 1 || batch * (int64_t)hidden > (1ll << 30)) return NTM_B200_ERR_BAD_SHAPE;
  if (!on_sm100()) return NTM_B200_ERR_NO_DEVICE;
  return ntm_b200::train::launch_lstm_backward(batch, hidden, dh_a, hidden, dh_b, hidden, 1, 0, z, z_stride, c_prev, c_new,
                                               c_stride, dc, dz, dz_stride, nullptr, 0, static_cast<cudaStream_t>(stream));
}
