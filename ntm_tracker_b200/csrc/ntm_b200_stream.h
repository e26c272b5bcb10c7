// ntm_b200_stream.h -- host interface of the streaming (throughput) mode of the NTM sequence path.
//
// The persistent kernel (ntm_b200_seq_kernel.cuh) keeps each sequence's memory in shared memory, which
// caps the sequences in flight at ~74 (C2 shapes) however large the batch is.  For batches far beyond
// that the same step is run over ALL sequences of the shard in lockstep: the dense projections become
// large tensor-core GEMMs and the memory is streamed from HBM by one fused addressing kernel per
// timestep (one read + one write of M per sequence-step; the second read hits L2).  See DESIGN.md s4.3.
#pragma once
#include <cuda_runtime.h>

#include "ntm_b200.h"
#include "ntm_b200_params.h"

namespace ntm_b200 {

struct StreamWorkspace {
  long long off_act[MAXL], off_partA, off_mc, off_cn, off_prof, off_xw, total;
  long long off_tilesA, off_tilesC, off_whiA, off_wloA, off_whiC, off_wloC;   // warp-specialised GEMM operands
  // hoisted x-projection through the same GEMM: packed W_x, operand tiles of the frames' first xK columns, and the
  // remaining xrem (< 64, <= 8) columns as a compact side array [B*T][xrem] the gate kernel folds in
  long long off_whiX, off_wloX, off_xtiles, off_xr;
  int xK, xrem;                 // 0 / 0 when the shape does not take this path
  long long slabA, slabC;       // floats per K-slice slab
  int ksA[MAXL], ksC;           // K-slices of each controller GEMM / of the head-parameter GEMM
  int actK[MAXL];
};

// true when the shape is one the streaming kernels cover (M % 4 == 0, GEMM widths within the tile kernel)
bool stream_supported(const ntm_b200_shape* s, int nsm);
bool stream_ws_path(const ntm_b200_shape* s);   // the chain runs on gemm_ws_kernel + mem_step_tma_kernel
void stream_layout(const ntm_b200_shape* s, long long B, long long T, StreamWorkspace* ws);

// Runs T steps for B sequences.  xw = hoisted x-projection [B,T,4C] (already computed on `stream`),
// wC/bC = packed [C,PO4] head-parameter + output projection.  Returns an ntm_b200_status.
// `cont`: continuation of the previous call on this workspace with the state updated in place (in == out):
// column norms, activation rows, operand tiles and packed weights in the workspace are still valid, so the
// per-call initialisation is skipped.
// `hist` (may be null) = training history buffers (ntm_b200.h); in this mode M_prev[t] IS the working
// memory of step t (pass 2 writes slot t+1), so recording it costs no extra traffic.
int stream_forward(const ntm_b200_shape* s, const ntm_b200_weights* w, const float* wC, const float* bC,
                   long long B, long long T, const float* xw, const ntm_b200_state* in,
                   const ntm_b200_state* out, float* logits, float* outputs, const ntm_b200_history* hist,
                   char* wsb, const StreamWorkspace& ws, int nsm, cudaStream_t stream, bool prof, bool cont,
                   const EnvSwitches& env, bool xw_partial = false);   // xw_partial: xw came from stream_xproj

// Hoisted input projection of the streaming mode on the warp-specialised GEMM: xw[b,t,:] = x[b,t,0:xK] @ W_x[0:xK] + b
// (pack pass -> tile records, then gemm_ws); the last xrem input columns are copied to ws.off_xr and added by the
// gate kernel of layer 0.  Returns 0 when done, -1 when the shape does not take this path (the caller then runs
// the tile kernel of ntm_b200_xproj_tc.cuh over all D columns), > 0 = ntm_b200_status.
// `fs` (may be null): the frames arrive as conv features [B, L, F, Cch] + first-frame target map [B, F] instead of
// serialised rows x [B, T, Cch+2] (x is ignored then): the pack pass reads the feature rows where they lie (16-byte
// loads: Cch-wide rows are aligned, the 514-wide serialised ones are not) and synthesises the delimiter / target
// channels into the side array -- the tf.concat / tile / reshape chain of direct_offset_output.py:439-500 never
// materialises.  Needs Cch == xK (else -1).
struct FeatureSource { const float* features; const float* target; int L, F, Cch, delimiter_first; };
int stream_xproj(const ntm_b200_shape* s, const ntm_b200_weights* w, long long B, long long T, const float* x, float* xw,
                 char* wsb, const StreamWorkspace& ws, int nsm, cudaStream_t stream, bool cont, const EnvSwitches& env,
                 const FeatureSource* fs = nullptr);

// profiling (after the stream was synchronised): {controller GEMM + LSTM, head-parameter GEMM, memory
// kernel, init} summed over the T steps, in ms; returns the number of memory-kernel launches (0 = none)
int stream_last_ms(float* out4);
// co-resident CTAs per SM of the memory kernel last configured (occupancy query)
int stream_mem_occupancy();
int stream_phase_ns(double* out9);

// defined in ntm_b200.cu (the tcgen05 tile kernel lives in a header with internal linkage state)
int gemm_tc(const float* x, int ldx, const float* w, int ldw, const float* bias, float* out, int ldo,
            long long slab, long long rows, int K, int ncols, int kslices, int nsm, cudaStream_t stream);
int gemm_tc_slices(int K);
int set_cuda_error_ext(cudaError_t e, const char* where);

}  // namespace ntm_b200
