// ntm_b200_train.h -- internal launchers of the backward kernels (ntm_b200_train.cu), shared by the
// single-step C-ABI entry points and the in-library reverse-time loop (ntm_b200_backward.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ntm_b200.h"

namespace ntm_b200 {
namespace train {

// One reverse-time step of the memory / addressing backward for `batch` sequences (see
// ntm_b200_memory_backward_step).  d_read: sequences `sdr` floats apart; dlogits (may be null) [B,T,O]
// fills the logit slots of d_raw for step t; sim_hist [B,H,N] / cn_hist [B,M] (both or neither): what the
// forward pass recorded for this step (ntm_b200_history::sim / ::cn); tiles_raw (may be null): operand tile records
// (ntm_b200_gemm_tiles.cuh) that receive the d_raw rows as well.  Returns an ntm_b200_status.
int launch_memory_backward(const ntm_b200_shape* s, long long batch, const float* M_prev, const float* w_prev,
                           const float* raw_params, const float* d_read, long long sdr, const float* d_w, float* dM,
                           float* d_w_prev, float* d_raw_params, const float* dlogits, int T, int t,
                           const float* sim_hist, const float* cn_hist, uint8_t* tiles_raw, int KAtot_raw,
                           cudaStream_t stream, bool pdl = false);

// BasicLSTMCell backward, elementwise part.  dh = dh_a[b*lda + u] (null = 0) + sum over `nslab` K-slice slabs
// of dh_b[s*slab + b*ldb + u] (null = 0).  tiles (may be null; used when hidden % 8 == 0): operand tile records that
// receive d_z rows as well.
int launch_lstm_backward(long long batch, int hidden, const float* dh_a, long long lda, const float* dh_b, long long ldb,
                         int nslab, long long slab, const float* z, long long z_stride, const float* c_prev,
                         const float* c_new, long long c_stride, float* dc, float* dz, long long dz_stride,
                         uint8_t* tiles, int KAtot, cudaStream_t stream, bool pdl = false);

}  // namespace train
}  // namespace ntm_b200
