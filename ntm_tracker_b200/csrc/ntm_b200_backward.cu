// ntm_b200_backward.cu -- the training step's device side around the forward pass (BASELINE configs[4]):
//
//   ntm_b200_offset_loss      tanh / l2_loss of the delimiter-step logits ...... direct_offset_output.py:581-606
//   ntm_b200_backward_seq     tf.gradients through the unrolled while_loop ..... direct_offset_output.py:611-613
//   ntm_b200_rmsprop_step     clip_by_global_norm + RMSProp.apply_gradients .... direct_offset_output.py:606-626
//
// The reverse-time loop runs here, not in Python: per timestep the fused memory/addressing backward kernel
// (ntm_b200_train.cu), the LSTM gate backward, and the two data-gradient contractions on the tensor cores
// (ntm_b200_gemm_tiles.cuh: d_h = d_raw @ [W_addr|W_out]^T, d_[read|h] = d_z @ W_lstm^T); after the loop the
// weight gradients X^T @ dZ as large-K tcgen05 GEMMs over all (t, b), the bias gradients (column sums) and
// the init_state gradients (batch sums through tanh / sigmoid).  Every reduction has a fixed order.
#include <vector>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "ntm_b200.h"
#include "ntm_b200_gemm_tiles.cuh"
#include "ntm_b200_params.h"
#include "ntm_b200_stream.h"
#include "ntm_b200_train.h"

namespace ntm_b200 {
namespace bwd {

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline long long align_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------ loss --
constexpr int MAX_GATHER = 64;
struct LossArgs {
  const float* logits; const float* targets; float* loss; float* dlogits;
  long long B; int T, O, G;
  int steps[MAX_GATHER];
};
// One CTA: zero dlogits, then y = tanh(logit), diff = y - target, dlogit = diff * (1 - y^2), loss = 0.5 sum diff^2.
__global__ void __launch_bounds__(1024) offset_loss_kernel(const LossArgs a) {
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const long long all = a.B * a.T * a.O;
  for (long long i = tid; i < all; i += 1024) a.dlogits[i] = 0.0f;
  __syncthreads();
  const long long n = a.B * a.G * a.O;
  float acc = 0.0f;
  for (long long i = tid; i < n; i += 1024) {
    const int o = (int)(i % a.O);
    const long long bg = i / a.O;
    const int g = (int)(bg % a.G);
    const long long b = bg / a.G;
    const long long li = (b * a.T + a.steps[g]) * a.O + o;
    const float y = tanhf(a.logits[li]);
    const float diff = y - a.targets[i];
    a.dlogits[li] = diff * (1.0f - y * y);
    acc = fmaf(diff, diff, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.0f;
    for (int i = 0; i < 32; ++i) s += red[i];
    a.loss[0] = 0.5f * s;
  }
}

// ------------------------------------------------------------------------------- column sums --
constexpr int COLSUM_MAX_CHUNKS = 16;
// dst[j] = f(sum_r src[r * ld + j]), j < ncols: bias gradients (sum over all (t, b)) and the init_state
// gradients (sum over the batch, times the derivative of the activation the variable goes through:
// mode 1: s0 = tanh(var) -> (1 - s0^2); mode 2: s0 = sigmoid(var) -> s0 (1 - s0)).  A CTA owns 32 columns;
// 8 row phases per column, combined in phase order.
// Tall matrices (8192 rows x 800 columns would be 25 CTAs): gridDim.y row chunks write partial sums to `scratch`
// [chunks][ncols] and colsum_final_kernel adds them in chunk order; a single chunk writes dst directly.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ src, long long ld, long long nrows_all, int ncols,
                                                     float* __restrict__ dst, const float* __restrict__ s0, int mode,
                                                     float* __restrict__ scratch, long long rows_per_chunk) {
  __shared__ float part[8][33];
  const int cl = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + cl;
  const long long r_begin = (long long)blockIdx.y * rows_per_chunk;
  const long long nrows = r_begin + rows_per_chunk < nrows_all ? r_begin + rows_per_chunk : nrows_all;
  float acc = 0.0f;
  if (j < ncols) {
    long long r = r_begin + ph;
    for (; r + 24 < nrows; r += 32) {      // four independent loads in flight
      const float v0 = __ldg(src + r * ld + j), v1 = __ldg(src + (r + 8) * ld + j);
      const float v2 = __ldg(src + (r + 16) * ld + j), v3 = __ldg(src + (r + 24) * ld + j);
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; r < nrows; r += 8) acc += __ldg(src + r * ld + j);
  }
  part[ph][cl] = acc;
  __syncthreads();
  if (ph == 0 && j < ncols) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][cl];
    if (gridDim.y > 1) { scratch[(size_t)blockIdx.y * ncols + j] = s; return; }
    if (mode == 1) { const float a0 = s0[j]; s *= (1.0f - a0 * a0); }
    else if (mode == 2) { const float a0 = s0[j]; s *= a0 * (1.0f - a0); }
    dst[j] = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ scratch, int chunks, int ncols, float* __restrict__ dst,
                                    const float* __restrict__ s0, int mode) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  float s = 0.0f;
  for (int c = 0; c < chunks; ++c) s += scratch[(size_t)c * ncols + j];
  if (mode == 1) { const float a0 = s0[j]; s *= (1.0f - a0 * a0); }
  else if (mode == 2) { const float a0 = s0[j]; s *= a0 * (1.0f - a0); }
  dst[j] = s;
}

// dst[i] = sum over slabs (slice order) of src[s * slab + i]   (split-K partials of a weight-gradient GEMM)
__global__ void sum_slabs_kernel(const float* __restrict__ src, long long slab, int nslab, long long n, float* __restrict__ dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < nslab; ++k) s += src[(size_t)k * slab + i];
    dst[i] = s;
  }
}
// [C, PO4] slabs of d[W_addr | W_out] -> d W_addr [C, P], d W_out [C, O]; tmp_b [PO4] -> d b_addr [P], d b_out [O]
__global__ void unpack_ao_grads_kernel(const float* __restrict__ src, long long slab, int nslab, const float* __restrict__ tmp_b,
                                       int C, int P, int O, int PO4, float* aw, float* ab, float* ow, float* ob) {
  const long long total = (long long)(C + 1) * PO4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / PO4), q = (int)(i - (long long)r * PO4);
    if (q >= P + O) continue;
    if (r < C) {
      float s = 0.0f;
      for (int k = 0; k < nslab; ++k) s += src[(size_t)k * slab + i];
      if (q < P) aw[(size_t)r * P + q] = s; else ow[(size_t)r * O + (q - P)] = s;
    } else {
      const float s = tmp_b[q];
      if (q < P) ab[q] = s; else ob[q - P] = s;
    }
  }
}

// ------------------------------------------------------------------------------- optimizer --
constexpr int NORM_BLOCKS = 512;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ part) {
  __shared__ float red[8];
  float acc = 0.0f;
  // contiguous chunk per CTA, fixed assignment: the partials (and their order) do not depend on timing
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = (long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) { const float v = g[i]; acc = fmaf(v, v, acc); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += red[i];
    part[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(32) norm_final_kernel(const float* __restrict__ part, int nparts, float* gnorm) {
  // fixed order: lane l sums parts l, l+32, ...; then a butterfly
  float s = 0.0f;
  for (int i = threadIdx.x; i < nparts; i += 32) s += part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) gnorm[0] = sqrtf(s);
}
__global__ void __launch_bounds__(256) rmsprop_update_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                             float* __restrict__ rms, float* __restrict__ mom, long long n,
                                                             float lr, float decay, float momentum, float eps, float clip,
                                                             const float* __restrict__ gnorm) {
  const float gn = gnorm[0];
  const float scale = clip / fmaxf(gn, clip);            // tf.clip_by_global_norm
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * scale;
    const float r = decay * rms[i] + (1.0f - decay) * gi * gi;
    const float m = momentum * mom[i] + lr * gi / sqrtf(r + eps);
    rms[i] = r;
    mom[i] = m;
    p[i] -= m;
  }
}

// ------------------------------------------------------------------------------- workspace --
struct Layout {
  int C, L, R, W, H, N, M, D, O, S, P, PO4, RM;
  int ncat[MAXL], inK[MAXL];                  // columns of d_cat_l = d[inputs after x | h]; rows of W_l
  gemmt::Plan pA;                             // d_h      = d_raw [B, PO4]  x  Wao^T
  gemmt::Plan pB[MAXL];                       // d_cat_l  = d_z_l [B, 4C]   x  W_l[off:]^T
  gemmt::Plan pGao;                           // d Wao    = hc^T [C, TB]    x  DMC
  gemmt::Plan pGw[MAXL];                      // d W_l    = inp_l^T         x  DZ_l
  long long off_dM, off_dw[2], off_dc, off_dcat[MAXL], off_dh, off_DMC, off_DZ, off_tmpb, off_slabs, off_colsum;
  long long colsum_floats;
  long long off_rowDraw, off_colWao, off_rowDz, off_colWl[MAXL], off_rowG, off_colG;
  long long tiles_begin, tiles_end, total;
  long long slabs_floats;
};

void make_layout(const ntm_b200_shape* s, long long B, long long T, Layout* y) {
  y->C = s->controller_hidden_size; y->L = s->controller_num_layers; y->R = s->read_head_size;
  y->W = s->write_head_size; y->H = y->R + y->W; y->N = s->mem_size; y->M = s->mem_dim; y->D = s->input_dim;
  y->O = s->output_dim; y->S = 2 * s->shift_range + 1;
  y->P = y->H * y->M + 3 * y->H + y->S * y->H + 2 * y->M * y->W;
  y->PO4 = round_up(y->P + y->O, 4);
  y->RM = y->R * y->M;
  const int C = y->C, L = y->L;
  const long long TB = T * B;
  const int nsm = B200_SMS;
  for (int l = 0; l < L; ++l) {
    y->ncat[l] = (l == 0) ? y->RM + C : 2 * C;
    y->inK[l] = (l == 0) ? y->D + y->RM + C : 2 * C;
  }
  auto slices_for = [&](int nrows, int ncols, int K) {       // enough K slices to put ~one wave of CTAs to work
    const int tiles = ((nrows + 127) / 128) * ((ncols + 127) / 128);
    int ks = nsm / std::max(1, tiles);
    return std::max(1, std::min(ks, (K + 63) / 64));
  };
  y->pA = gemmt::make_plan((int)B, C, y->PO4, std::min(16, slices_for((int)B, C, y->PO4)));
  for (int l = 0; l < L; ++l) y->pB[l] = gemmt::make_plan((int)B, y->ncat[l], 4 * C, 1);
  y->pGao = gemmt::make_plan(C, y->PO4, (int)TB, std::min(8, slices_for(C, y->PO4, (int)TB)));
  for (int l = 0; l < L; ++l) y->pGw[l] = gemmt::make_plan(y->inK[l], 4 * C, (int)TB, std::min(8, slices_for(y->inK[l], 4 * C, (int)TB)));
  long long o = 0;
  auto take = [&](long long bytes) { long long r = o; o = align_up_ll(o + bytes, 1024); return r; };
  y->off_dM = take(4ll * B * y->N * y->M);
  y->off_dw[0] = take(4ll * B * y->H * y->N);
  y->off_dw[1] = take(4ll * B * y->H * y->N);
  y->off_dc = take(4ll * L * B * C);
  for (int l = 0; l < L; ++l) y->off_dcat[l] = take(4ll * B * y->ncat[l]);
  y->off_dh = take(4ll * y->pA.kslices * B * C);
  y->off_DMC = take(4ll * TB * y->PO4);
  y->off_DZ = take(4ll * TB * L * 4 * C);
  y->off_tmpb = take(4ll * y->PO4);
  y->colsum_floats = (long long)bwd::COLSUM_MAX_CHUNKS * std::max(y->PO4, 4 * C);     // row-chunk partials of the bias sums
  y->off_colsum = take(4ll * y->colsum_floats);
  long long slabs = (long long)y->pGao.kslices * C * y->PO4;
  for (int l = 0; l < L; ++l) slabs = std::max(slabs, (long long)y->pGw[l].kslices * y->inK[l] * 4 * C);
  y->slabs_floats = slabs;
  y->off_slabs = take(4ll * slabs);
  y->tiles_begin = o;
  y->off_rowDraw = take((long long)y->pA.row_bytes);
  y->off_colWao = take((long long)y->pA.col_bytes);
  long long rowdz = 0;
  for (int l = 0; l < L; ++l) rowdz = std::max(rowdz, (long long)y->pB[l].row_bytes);
  y->off_rowDz = take(rowdz);
  for (int l = 0; l < L; ++l) y->off_colWl[l] = take((long long)y->pB[l].col_bytes);
  long long rowG = (long long)y->pGao.row_bytes, colG = (long long)y->pGao.col_bytes;
  for (int l = 0; l < L; ++l) {
    rowG = std::max(rowG, (long long)y->pGw[l].row_bytes);
    colG = std::max(colG, (long long)y->pGw[l].col_bytes);
  }
  y->off_rowG = take(rowG);
  y->off_colG = take(colG);
  y->tiles_end = o;
  y->total = o;
}

#define BWD_CK(call, what)                                            \
  do {                                                                \
    cudaError_t e_ = (call);                                          \
    if (e_ != cudaSuccess) return set_cuda_error_ext(e_, what);       \
  } while (0)

// scratch (may be null -> one chunk): room for COLSUM_MAX_CHUNKS * ncols floats
int launch_colsum(const float* src, long long ld, long long nrows, int ncols, float* dst, const float* s0, int mode,
                  cudaStream_t stream, float* scratch = nullptr, long long scratch_floats = 0) {
  const int cblocks = (ncols + 31) / 32;
  int chunks = 1;
  if (scratch != nullptr && cblocks < 2 * B200_SMS) {      // not enough column blocks to fill the device: split the rows
    chunks = (int)std::min<long long>(COLSUM_MAX_CHUNKS, std::max<long long>(1, nrows / 256));
    chunks = (int)std::min<long long>(chunks, scratch_floats / std::max(1, ncols));
    if (chunks < 1) chunks = 1;
  }
  const long long rpc = (nrows + chunks - 1) / chunks;
  colsum_kernel<<<dim3(cblocks, chunks), 256, 0, stream>>>(src, ld, nrows, ncols, dst, s0, mode, scratch, rpc);
  count_launch();
  if (chunks > 1) {
    colsum_final_kernel<<<(ncols + 255) / 256, 256, 0, stream>>>(scratch, chunks, ncols, dst, s0, mode);
    count_launch();
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace bwd
}  // namespace ntm_b200

using namespace ntm_b200;

namespace {
// profiling (ntm_b200_set_profiling): events on the launching stream -- [0] start, per reverse step (before,
// after) the memory-backward kernel, then end of the reverse loop, end of the call
thread_local std::vector<cudaEvent_t> g_bev;
thread_local int g_bev_steps = 0;

bool device_is_sm100() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      major != 10) {
    cudaGetLastError();
    return false;
  }
  return true;
}
int check_train_shape(const ntm_b200_shape* s) {
  if (s->input_dim < 1 || s->output_dim < 1 || s->mem_size < 1 || s->mem_dim < 1 || s->controller_hidden_size < 1 ||
      s->controller_num_layers < 1 || s->controller_num_layers > MAXL)
    return NTM_B200_ERR_BAD_SHAPE;
  if (s->shift_range < 0 || s->shift_range > NTM_B200_MAX_SHIFT_RANGE) return NTM_B200_ERR_BAD_SHIFT;
  if (s->read_head_size < 1 || s->read_head_size > NTM_B200_MAX_READ_HEADS || s->write_head_size < 1 ||
      s->write_head_size > NTM_B200_MAX_WRITE_HEADS)
    return NTM_B200_ERR_UNSUPPORTED_HEADS;
  return NTM_B200_OK;
}
}  // namespace

extern "C" int32_t ntm_b200_offset_loss(const float* logits, const float* targets, const int32_t* gather_steps,
                                        int32_t num_gather, int64_t batch, int64_t steps, int32_t output_dim,
                                        float* loss_out, float* dlogits, void* stream) {
  if (!logits || !targets || !gather_steps || !loss_out || !dlogits) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || steps < 1 || output_dim < 1 || num_gather < 1 || num_gather > bwd::MAX_GATHER) return NTM_B200_ERR_BAD_SHAPE;
  for (int g = 0; g < num_gather; ++g)
    if (gather_steps[g] < 0 || gather_steps[g] >= steps) return NTM_B200_ERR_BAD_SHAPE;
  if (!device_is_sm100()) return NTM_B200_ERR_NO_DEVICE;
  bwd::LossArgs a{};
  a.logits = logits; a.targets = targets; a.loss = loss_out; a.dlogits = dlogits;
  a.B = batch; a.T = (int)steps; a.O = output_dim; a.G = num_gather;
  for (int g = 0; g < num_gather; ++g) a.steps[g] = gather_steps[g];
  bwd::offset_loss_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? NTM_B200_OK : NTM_B200_ERR_CUDA;
}

extern "C" int64_t ntm_b200_backward_workspace_bytes(const ntm_b200_shape* shape, int64_t batch, int64_t steps) {
  if (!shape || batch < 1 || steps < 1 || check_train_shape(shape) != NTM_B200_OK) return 0;
  bwd::Layout y{};
  bwd::make_layout(shape, batch, steps, &y);
  return y.total;
}

extern "C" int32_t ntm_b200_backward_seq(const ntm_b200_shape* s, const ntm_b200_weights* w, const void* packed,
                                         int64_t B, int64_t T, const float* inputs, const ntm_b200_history* hist,
                                         const float* dlogits, const ntm_b200_state* state0, const ntm_b200_grads* g,
                                         void* workspace, int64_t workspace_bytes, void* stream_v) {
  if (!s || !w || !packed || !inputs || !hist || !dlogits || !g || !workspace) return NTM_B200_ERR_NULL_POINTER;
  if (!hist->M_prev || !hist->w_prev || !hist->params || !hist->z || !hist->c || !hist->h || !hist->read)
    return NTM_B200_ERR_NULL_POINTER;
  int st = check_train_shape(s);
  if (st) return st;
  if (B < 1 || T < 1 || B > (1 << 22) || T > (1 << 16) || B * T > (1ll << 26)) return NTM_B200_ERR_BAD_SHAPE;
  if (!g->addr_w || !g->addr_b || !g->out_w || !g->out_b) return NTM_B200_ERR_NULL_POINTER;
  for (int l = 0; l < s->controller_num_layers; ++l)
    if (!w->lstm_w[l] || !g->lstm_w[l] || !g->lstm_b[l]) return NTM_B200_ERR_NULL_POINTER;
  if (!device_is_sm100()) return NTM_B200_ERR_NO_DEVICE;
  bwd::Layout y{};
  bwd::make_layout(s, B, T, &y);
  if (workspace_bytes < y.total) return NTM_B200_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* ws = static_cast<char*>(workspace);
  const int C = y.C, L = y.L, H = y.H, N = y.N, M = y.M, D = y.D, PO4 = y.PO4, RM = y.RM, P = y.P, O = y.O;
  const long long TB = T * B;
  int nsm = B200_SMS;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  float* dM = reinterpret_cast<float*>(ws + y.off_dM);
  float* dwbuf[2] = {reinterpret_cast<float*>(ws + y.off_dw[0]), reinterpret_cast<float*>(ws + y.off_dw[1])};
  float* dc = reinterpret_cast<float*>(ws + y.off_dc);
  float* dcat[MAXL];
  for (int l = 0; l < L; ++l) dcat[l] = reinterpret_cast<float*>(ws + y.off_dcat[l]);
  float* dh = reinterpret_cast<float*>(ws + y.off_dh);
  float* DMC = reinterpret_cast<float*>(ws + y.off_DMC);
  float* DZ = reinterpret_cast<float*>(ws + y.off_DZ);
  float* tmpb = reinterpret_cast<float*>(ws + y.off_tmpb);
  float* slabs = reinterpret_cast<float*>(ws + y.off_slabs);
  uint8_t* rowDraw = reinterpret_cast<uint8_t*>(ws + y.off_rowDraw);
  uint8_t* colWao = reinterpret_cast<uint8_t*>(ws + y.off_colWao);
  uint8_t* rowDz = reinterpret_cast<uint8_t*>(ws + y.off_rowDz);
  uint8_t* rowG = reinterpret_cast<uint8_t*>(ws + y.off_rowG);
  uint8_t* colG = reinterpret_cast<uint8_t*>(ws + y.off_colG);
  const float* Wao = static_cast<const float*>(packed);                 // [C, PO4]: [W_addr | W_out | 0]

  // ---- zero state of the recursion: dM, dw, dc, d_cat (= d_read and d_h entering step T-1), operand tiles ----
  BWD_CK(cudaMemsetAsync(ws + y.off_dM, 0, (size_t)(y.off_dh - y.off_dM), stream), "cudaMemsetAsync(backward state)");
  BWD_CK(cudaMemsetAsync(ws + y.tiles_begin, 0, (size_t)(y.tiles_end - y.tiles_begin), stream), "cudaMemsetAsync(operand tiles)");

  // ---- weights as column operands (once per call): element (j, k) = W[j][k], K contiguous ----
  BWD_CK(gemmt::pack(Wao, PO4, 1, 0, 0, C, PO4, colWao, y.pA.KAtot, 0, 0, false, nsm, stream), "pack(Wao)");
  count_launch();
  for (int l = 0; l < L; ++l) {
    const float* wl = w->lstm_w[l] + (l == 0 ? (size_t)D * 4 * C : 0);   // rows after the x rows: no gradient flows into x
    BWD_CK(gemmt::pack(wl, 4 * C, 1, 0, 0, y.ncat[l], 4 * C, reinterpret_cast<uint8_t*>(ws + y.off_colWl[l]), y.pB[l].KAtot, 0, 0,
                       false, nsm, stream), "pack(W_lstm)");
    count_launch();
  }

  const bool prof = profiling_enabled();
  g_bev_steps = 0;
  if (prof) {
    const size_t need = 2 * (size_t)T + 3;
    while (g_bev.size() < need) {
      cudaEvent_t ev;
      BWD_CK(cudaEventCreate(&ev), "cudaEventCreate");
      g_bev.push_back(ev);
    }
    cudaEventRecord(g_bev[0], stream);
  }

  // ---- reverse-time loop: four launches per step, chained by programmatic dependent launch (every kernel waits for
  //      its predecessor before touching memory; what hides is the launch latency).  Off with per-kernel profiling
  //      events between the launches and with NTM_B200_EXP bit 16. ----
  const bool pdl = !prof && !(read_env().exp & 16);
  int cur = 0;
  for (long long t = T - 1; t >= 0; --t) {
    float* draw = DMC + (size_t)t * B * PO4;
    if (prof) cudaEventRecord(g_bev[1 + 2 * (T - 1 - t)], stream);
    st = train::launch_memory_backward(s, B, hist->M_prev + (size_t)t * B * N * M, hist->w_prev + (size_t)t * B * H * N,
                                       hist->params + (size_t)t * B * PO4, dcat[0], y.ncat[0], dwbuf[cur], dM,
                                       dwbuf[cur ^ 1], draw, dlogits, (int)T, (int)t,
                                       (hist->sim && hist->cn) ? hist->sim + (size_t)t * B * H * N : nullptr,
                                       (hist->sim && hist->cn) ? hist->cn + (size_t)t * B * M : nullptr, rowDraw, y.pA.KAtot,
                                       stream, pdl && t < T - 1);
    if (st) return st == NTM_B200_ERR_CUDA ? set_cuda_error_ext(cudaGetLastError(), "mem_backward_kernel") : st;
    if (prof) cudaEventRecord(g_bev[2 + 2 * (T - 1 - t)], stream);
    cur ^= 1;
    // d_h(top) = d_raw @ [W_addr | W_out]^T, K slices into slabs
    // (the memory-backward kernel has already written d_raw into the row-operand tiles)
    BWD_CK(gemmt::launch(y.pA, rowDraw, colWao, dh, C, B * (long long)C, stream, pdl), "gemm_tiles(d_h)");
    count_launch();
    for (int l = L - 1; l >= 0; --l) {
      const float* dh_b = (l == L - 1) ? dh : dcat[l + 1];
      const long long ldb = (l == L - 1) ? C : y.ncat[l + 1];
      const int nslab = (l == L - 1) ? y.pA.kslices : 1;
      float* dz = DZ + ((size_t)t * B * L + l) * 4 * C;
      st = train::launch_lstm_backward(B, C, dcat[l] + (y.ncat[l] - C), y.ncat[l], dh_b, ldb, nslab, B * (long long)C,
                                       hist->z + ((size_t)t * B * L + l) * 4 * C, (long long)L * 4 * C,
                                       hist->c + ((size_t)t * B * L + l) * C, hist->c + ((size_t)(t + 1) * B * L + l) * C,
                                       (long long)L * C, dc + (size_t)l * B * C, dz, (long long)L * 4 * C, rowDz, y.pB[l].KAtot,
                                       stream, pdl);
      if (st) return set_cuda_error_ext(cudaGetLastError(), "lstm_backward_kernel");
      // d_cat_l = d_z_l @ W_l[off:]^T  -> [d_read | d_h_l] (layer 0) or [d_h_{l-1} (this step) | d_h_l (previous step)]
      if (C % 8 != 0) {     // (else the LSTM backward kernel has written d_z into the row-operand tiles itself)
        BWD_CK(gemmt::pack(dz, (long long)L * 4 * C, 1, 0, 0, (int)B, 4 * C, rowDz, y.pB[l].KAtot, 0, 0, false, nsm, stream), "pack(d_z)");
        count_launch();
      }
      BWD_CK(gemmt::launch(y.pB[l], rowDz, reinterpret_cast<uint8_t*>(ws + y.off_colWl[l]), dcat[l], y.ncat[l], 0, stream,
                           pdl && C % 8 == 0),
             "gemm_tiles(d_cat)");
      count_launch();
    }
  }
  float* dw_final = dwbuf[cur];          // gradient w.r.t. the weightings entering step 0
  if (prof) cudaEventRecord(g_bev[1 + 2 * T], stream);

  // ---- weight gradients: one large-K GEMM per variable over all (t, b) ----
  const dim3 eg(2 * nsm), eb(256);
  {   // d [W_addr | W_out] [C, PO4] = hc^T @ DMC, hc = top-layer h after each step (history slots 1..T)
    BWD_CK(cudaMemsetAsync(rowG, 0, y.pGao.row_bytes, stream), "cudaMemsetAsync");
    BWD_CK(cudaMemsetAsync(colG, 0, y.pGao.col_bytes, stream), "cudaMemsetAsync");
    BWD_CK(gemmt::pack(hist->h + (size_t)B * L * C + (size_t)(L - 1) * C, 1, (long long)L * C, 0, 0, C, (int)TB, rowG,
                       y.pGao.KAtot, 0, 0, true, nsm, stream), "pack(h^T)");
    BWD_CK(gemmt::pack(DMC, 1, PO4, 0, 0, PO4, (int)TB, colG, y.pGao.KAtot, 0, 0, true, nsm, stream), "pack(d_raw^T)");
    BWD_CK(gemmt::launch(y.pGao, rowG, colG, slabs, PO4, (long long)C * PO4, stream), "gemm_tiles(d Wao)");
    for (int i = 0; i < 3; ++i) count_launch();
    if (bwd::launch_colsum(DMC, PO4, TB, PO4, tmpb, nullptr, 0, stream, reinterpret_cast<float*>(ws + y.off_colsum), y.colsum_floats)) return set_cuda_error_ext(cudaGetLastError(), "colsum(d_raw)");
    bwd::unpack_ao_grads_kernel<<<eg, eb, 0, stream>>>(slabs, (long long)C * PO4, y.pGao.kslices, tmpb, C, P, O, PO4,
                                                      g->addr_w, g->addr_b, g->out_w, g->out_b);
    count_launch();
    BWD_CK(cudaGetLastError(), "unpack_ao_grads_kernel");
  }
  for (int l = 0; l < L; ++l) {   // d W_l [in_l + C, 4C] = [inputs_l | h_l entering]^T @ DZ_l
    const gemmt::Plan& pg = y.pGw[l];
    BWD_CK(cudaMemsetAsync(rowG, 0, pg.row_bytes, stream), "cudaMemsetAsync");
    BWD_CK(cudaMemsetAsync(colG, 0, pg.col_bytes, stream), "cudaMemsetAsync");
    if (l == 0) {
      // x is batch-major [B, T, D]; the contraction index is k = t * B + b like everything recorded
      BWD_CK(gemmt::pack(inputs, 1, D, (int)B, (int)T, D, (int)TB, rowG, pg.KAtot, 0, 0, true, nsm, stream), "pack(x^T)");
      BWD_CK(gemmt::pack(hist->read, 1, RM, 0, 0, RM, (int)TB, rowG, pg.KAtot, D, 0, true, nsm, stream), "pack(read^T)");
      BWD_CK(gemmt::pack(hist->h, 1, (long long)L * C, 0, 0, C, (int)TB, rowG, pg.KAtot, D + RM, 0, true, nsm, stream), "pack(h^T)");
      for (int i = 0; i < 3; ++i) count_launch();
    } else {
      BWD_CK(gemmt::pack(hist->h + (size_t)B * L * C + (size_t)(l - 1) * C, 1, (long long)L * C, 0, 0, C, (int)TB, rowG, pg.KAtot,
                         0, 0, true, nsm, stream), "pack(h_below^T)");
      BWD_CK(gemmt::pack(hist->h + (size_t)l * C, 1, (long long)L * C, 0, 0, C, (int)TB, rowG, pg.KAtot, C, 0, true, nsm, stream),
             "pack(h^T)");
      for (int i = 0; i < 2; ++i) count_launch();
    }
    BWD_CK(gemmt::pack(DZ + (size_t)l * 4 * C, 1, (long long)L * 4 * C, 0, 0, 4 * C, (int)TB, colG, pg.KAtot, 0, 0, true, nsm, stream),
           "pack(d_z^T)");
    count_launch();
    const long long nW = (long long)y.inK[l] * 4 * C;
    if (pg.kslices == 1) {
      BWD_CK(gemmt::launch(pg, rowG, colG, g->lstm_w[l], 4 * C, 0, stream), "gemm_tiles(d W_lstm)");
      count_launch();
    } else {
      BWD_CK(gemmt::launch(pg, rowG, colG, slabs, 4 * C, nW, stream), "gemm_tiles(d W_lstm)");
      bwd::sum_slabs_kernel<<<eg, eb, 0, stream>>>(slabs, nW, pg.kslices, nW, g->lstm_w[l]);
      count_launch(); count_launch();
      BWD_CK(cudaGetLastError(), "sum_slabs_kernel");
    }
    if (bwd::launch_colsum(DZ + (size_t)l * 4 * C, (long long)L * 4 * C, TB, 4 * C, g->lstm_b[l], nullptr, 0, stream,
                           reinterpret_cast<float*>(ws + y.off_colsum), y.colsum_floats))
      return set_cuda_error_ext(cudaGetLastError(), "colsum(d_z)");
  }
  // ---- init_state variables: tiled over the batch, so their gradients are batch sums (ntm_cell.py:292-306) ----
  if (state0 != nullptr && g->init_M && g->init_w && g->init_read && state0->M && state0->w && state0->read &&
      state0->stride_M == 0 && state0->stride_w == 0 && state0->stride_read == 0) {
    if (bwd::launch_colsum(dM, (long long)N * M, B, N * M, g->init_M, state0->M, 1, stream) ||
        bwd::launch_colsum(dw_final, (long long)H * N, B, H * N, g->init_w, state0->w, 2, stream) ||
        bwd::launch_colsum(dcat[0], y.ncat[0], B, RM, g->init_read, state0->read, 1, stream))
      return set_cuda_error_ext(cudaGetLastError(), "colsum(init_state)");
  }
  if (prof) {
    cudaEventRecord(g_bev[2 + 2 * T], stream);
    g_bev_steps = (int)T;
  }
  return NTM_B200_OK;
}

// Profiling enabled (ntm_b200_set_profiling) and the stream synchronised: device ms of this thread's last
// ntm_b200_backward_seq -- {memory-backward kernel summed over the steps, rest of the reverse loop (GEMMs +
// LSTM gate backward), weight gradients + init_state sums, whole call}; *steps = 0 if there is nothing to report.
extern "C" int32_t ntm_b200_last_backward_ms(float* out4, int32_t* steps) {
  if (!out4 || !steps) return NTM_B200_ERR_NULL_POINTER;
  out4[0] = out4[1] = out4[2] = out4[3] = 0.0f;
  *steps = g_bev_steps;
  if (g_bev_steps <= 0) return NTM_B200_OK;
  const int T = g_bev_steps;
  float ms = 0.0f, loop = 0.0f;
  for (int i = 0; i < T; ++i)
    if (cudaEventElapsedTime(&ms, g_bev[1 + 2 * i], g_bev[2 + 2 * i]) == cudaSuccess) out4[0] += ms;
  if (cudaEventElapsedTime(&loop, g_bev[0], g_bev[1 + 2 * T]) == cudaSuccess) out4[1] = loop - out4[0];
  if (cudaEventElapsedTime(&ms, g_bev[1 + 2 * T], g_bev[2 + 2 * T]) == cudaSuccess) out4[2] = ms;
  if (cudaEventElapsedTime(&ms, g_bev[0], g_bev[2 + 2 * T]) == cudaSuccess) out4[3] = ms;
  cudaGetLastError();
  return NTM_B200_OK;
}

extern "C" int32_t ntm_b200_rmsprop_step(float* params, const float* grads, float* rms, float* mom, int64_t n,
                                         float learning_rate, float decay, float momentum, float epsilon, float clip_norm,
                                         float* gnorm_out, void* scratch, void* stream_v) {
  if (!params || !grads || !rms || !mom || !gnorm_out || !scratch) return NTM_B200_ERR_NULL_POINTER;
  if (n < 1) return NTM_B200_ERR_BAD_SHAPE;
  if (!device_is_sm100()) return NTM_B200_ERR_NO_DEVICE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  float* part = static_cast<float*>(scratch);
  const int nb = (int)std::min<long long>(bwd::NORM_BLOCKS, (n + 255) / 256);
  bwd::sumsq_partial_kernel<<<nb, 256, 0, stream>>>(grads, n, part);
  bwd::norm_final_kernel<<<1, 32, 0, stream>>>(part, nb, gnorm_out);
  const int ub = (int)std::min<long long>(4ll * B200_SMS, (n + 255) / 256);
  bwd::rmsprop_update_kernel<<<ub, 256, 0, stream>>>(params, grads, rms, mom, n, learning_rate, decay, momentum, epsilon,
                                                    clip_norm, gnorm_out);
  for (int i = 0; i < 3; ++i) count_launch();
  return cudaGetLastError() == cudaSuccess ? NTM_B200_OK : NTM_B200_ERR_CUDA;
}

// out[r, j] = sum_k a[r * lda + k] * b[j * ldb + k]  (A @ B^T, fp32 in / fp32 out, 3-term bf16 split on the tensor
// cores): the tile-record GEMM of the training path on plain row-major operands -- packs both, runs
// gemm_tiles_kernel with `kslices` K slices, sums the slabs.  workspace >= ntm_b200_gemm_nt_workspace_bytes.
extern "C" int64_t ntm_b200_gemm_nt_workspace_bytes(int64_t nrows, int64_t ncols, int64_t K, int32_t kslices) {
  if (nrows < 1 || ncols < 1 || K < 1 || nrows > (1 << 24) || ncols > (1 << 24) || K > (1 << 26)) return 0;
  const gemmt::Plan p = gemmt::make_plan((int)nrows, (int)ncols, (int)K, kslices);
  return (int64_t)(p.row_bytes + p.col_bytes + 2048 + (p.kslices > 1 ? 4ll * p.kslices * nrows * ncols : 0));
}
extern "C" int32_t ntm_b200_gemm_nt(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
                                    int64_t nrows, int64_t ncols, int64_t K, int32_t kslices, void* workspace,
                                    int64_t workspace_bytes, void* stream_v) {
  if (!a || !b || !out || !workspace) return NTM_B200_ERR_NULL_POINTER;
  const int64_t need = ntm_b200_gemm_nt_workspace_bytes(nrows, ncols, K, kslices);
  if (need == 0 || lda < K || ldb < K || ldo < ncols) return NTM_B200_ERR_BAD_SHAPE;
  if (workspace_bytes < need) return NTM_B200_ERR_WORKSPACE;
  if (!device_is_sm100()) return NTM_B200_ERR_NO_DEVICE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const gemmt::Plan p = gemmt::make_plan((int)nrows, (int)ncols, (int)K, kslices);
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  uint8_t* rowT = reinterpret_cast<uint8_t*>(ws);
  uint8_t* colT = rowT + p.row_bytes;
  float* slabs = reinterpret_cast<float*>(colT + p.col_bytes);
  BWD_CK(cudaMemsetAsync(rowT, 0, p.row_bytes + p.col_bytes, stream), "cudaMemsetAsync(operand tiles)");
  BWD_CK(gemmt::pack(a, lda, 1, 0, 0, (int)nrows, (int)K, rowT, p.KAtot, 0, 0, false, B200_SMS, stream), "pack(a)");
  BWD_CK(gemmt::pack(b, ldb, 1, 0, 0, (int)ncols, (int)K, colT, p.KAtot, 0, 0, false, B200_SMS, stream), "pack(b)");
  if (p.kslices == 1) {
    BWD_CK(gemmt::launch(p, rowT, colT, out, (int)ldo, 0, stream), "gemm_tiles");
    for (int i = 0; i < 3; ++i) count_launch();
  } else {
    BWD_CK(gemmt::launch(p, rowT, colT, slabs, (int)ncols, nrows * ncols, stream), "gemm_tiles");
    if (ldo == ncols) {
      bwd::sum_slabs_kernel<<<2 * B200_SMS, 256, 0, stream>>>(slabs, nrows * ncols, p.kslices, nrows * ncols, out);
    } else {
      for (int64_t r = 0; r < nrows; ++r)   // strided destination: row by row (debug / small shapes only)
        bwd::sum_slabs_kernel<<<1, 256, 0, stream>>>(slabs + r * ncols, nrows * ncols, p.kslices, ncols, out + r * ldo);
    }
    for (int i = 0; i < 4; ++i) count_launch();
    BWD_CK(cudaGetLastError(), "sum_slabs_kernel");
  }
  return NTM_B200_OK;
}
