// ntm_b200_io.cu -- the two data-format steps either side of the NTM path.
//
//  * serialise:  conv features [B, L, F, Cch] + first-frame target map [B, F]
//                -> tracker inputs [B, L*(F+1), Cch+2]      (direct_offset_output.py:439-500)
//       row (l, f<F) = [features[b,l,f,:], 0, target]; row (l, F) = frame delimiter = [0...0, 1, 0];
//       the target channel carries target[b, f] on the F FEATURE rows of the first frame and 0 everywhere
//       else (training layout: those are steps 0..F-1, direct_offset_output.py:490-494).
//       `delimiter_first` puts the delimiter row at the START of every frame instead, which is the
//       serve path's layout (test_tracker.py:385-404: rows [feat_f, 0, gt_f] with the delimiter
//       [0...0, 1, 0] prepended) -- there feature f of the first frame is step f + 1.
//  * gather:     logits [B, L*(F+1), O] -> tanh(logits at every frame's delimiter row, first frame
//                dropped) [B, L-1, O]                        (direct_offset_output.py:581-593)
// Pure HBM-bound copies: one pass, coalesced (float2 where the row strides allow).
#include <cuda_runtime.h>

#include <algorithm>

#include "ntm_b200.h"
#include "ntm_b200_params.h"

namespace ntm_b200 {
namespace io {

__global__ void serialize_kernel(const float* __restrict__ feat, const float* __restrict__ target,
                                 float* __restrict__ out, long long rows, int L, int F, int Cch,
                                 int delimiter_first) {
  const int D = Cch + 2, T = L * (F + 1);
  const long long total = rows * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / D;
    const int col = (int)(i - row * D);
    const long long b = row / T;
    const int t = (int)(row - b * T);
    const int l = t / (F + 1), r = t - l * (F + 1);
    const bool is_delim = delimiter_first ? (r == 0) : (r == F);
    const int f = delimiter_first ? r - 1 : r;
    float v;
    if (col < Cch) v = is_delim ? 0.0f : __ldg(feat + (((b * L + l) * F + f) * (long long)Cch + col));
    else if (col == Cch) v = is_delim ? 1.0f : 0.0f;
    else v = (l == 0 && !is_delim) ? __ldg(target + b * F + f) : 0.0f;
    out[i] = v;
  }
}

__global__ void gather_kernel(const float* __restrict__ logits, float* __restrict__ out, long long B,
                              int L, int F, int O) {
  const long long total = B * (L - 1) * O;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const long long bl = i / O;
    const int l = (int)(bl % (L - 1));
    const long long b = bl / (L - 1);
    const long long t = (long long)(l + 1) * (F + 1) + F;          // delimiter row of frame l+1
    out[i] = tanhf(__ldg(logits + (b * (long long)L * (F + 1) + t) * O + o));
  }
}

// zero_state (ntm_cell.py:284-315): M = tanh(var_M), w = sigmoid(var_w), read = tanh(var_read), one launch
__global__ void zero_state_kernel(const float* __restrict__ vM, long long nM, const float* __restrict__ vw, long long nw,
                                  const float* __restrict__ vr, long long nr, float* __restrict__ M, float* __restrict__ w,
                                  float* __restrict__ r) {
  const long long total = nM + nw + nr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < nM) M[i] = tanhf(vM[i]);
    else if (i < nM + nw) w[i - nM] = 1.0f / (1.0f + expf(-vw[i - nM]));
    else r[i - nM - nw] = tanhf(vr[i - nM - nw]);
  }
}

}  // namespace io
}  // namespace ntm_b200

extern "C" int32_t ntm_b200_serialize_tracker_inputs(const float* features, const float* target, float* inputs,
                                                     int64_t batch, int32_t frames, int32_t num_features,
                                                     int32_t channels, int32_t delimiter_first, void* stream) {
  if (!features || !target || !inputs) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || frames < 1 || num_features < 1 || channels < 1) return NTM_B200_ERR_BAD_SHAPE;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
    cudaGetLastError();
    return NTM_B200_ERR_NO_DEVICE;
  }
  const long long rows = (long long)batch * frames * (num_features + 1);
  ntm_b200::io::serialize_kernel<<<ntm_b200::B200_SMS * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      features, target, inputs, rows, frames, num_features, channels, delimiter_first ? 1 : 0);
  ntm_b200::count_launch();
  return cudaGetLastError() == cudaSuccess ? NTM_B200_OK : NTM_B200_ERR_CUDA;
}

extern "C" int32_t ntm_b200_gather_offsets(const float* logits, float* offsets, int64_t batch, int32_t frames,
                                           int32_t num_features, int32_t output_dim, void* stream) {
  if (!logits || !offsets) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || frames < 2 || num_features < 1 || output_dim < 1) return NTM_B200_ERR_BAD_SHAPE;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
    cudaGetLastError();
    return NTM_B200_ERR_NO_DEVICE;
  }
  ntm_b200::io::gather_kernel<<<ntm_b200::B200_SMS, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, offsets, batch, frames, num_features, output_dim);
  ntm_b200::count_launch();
  return cudaGetLastError() == cudaSuccess ? NTM_B200_OK : NTM_B200_ERR_CUDA;
}

extern "C" int32_t ntm_b200_zero_state(const float* var_M, int64_t n_M, const float* var_w, int64_t n_w, const float* var_read,
                                       int64_t n_read, float* M, float* w, float* read, void* stream) {
  if (!var_M || !var_w || !var_read || !M || !w || !read) return NTM_B200_ERR_NULL_POINTER;
  if (n_M < 1 || n_w < 1 || n_read < 1) return NTM_B200_ERR_BAD_SHAPE;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
    cudaGetLastError();
    return NTM_B200_ERR_NO_DEVICE;
  }
  const long long total = n_M + n_w + n_read;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 4ll * ntm_b200::B200_SMS);
  ntm_b200::io::zero_state_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(var_M, n_M, var_w, n_w, var_read, n_read,
                                                                                         M, w, read);
  ntm_b200::count_launch();
  return cudaGetLastError() == cudaSuccess ? NTM_B200_OK : NTM_B200_ERR_CUDA;
}
