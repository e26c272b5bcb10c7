/* ntm_b200.h -- C ABI of the B200-native NTM-cell hot path.
 *
 * Drop-in boundary for ONE path of JeffOwOSun/ntm-tracker: the NTMCell step
 * unrolled over T timesteps for B independent sequences.  The reference has no
 * FFI of its own (it is TensorFlow-1 graph code); each entry point below cites
 * the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes only; every buffer is caller-owned
 * DEVICE memory (fp32, contiguous, row-major) unless marked host; every call
 * returns an ntm_b200_status and never throws; `stream` is a cudaStream_t passed
 * as void*; calls are asynchronous on that stream and re-entrant (no global
 * mutable state).  There is NO CPU fallback: on a machine without an sm_100
 * device every compute entry point returns NTM_B200_ERR_NO_DEVICE.
 */
#ifndef NTM_B200_H_
#define NTM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTM_B200_ABI_VERSION 2
#define NTM_B200_MAX_LAYERS 16
#define NTM_B200_MAX_READ_HEADS 4
#define NTM_B200_MAX_WRITE_HEADS 3
#define NTM_B200_MAX_SHIFT_RANGE 4

typedef enum ntm_b200_status {
  NTM_B200_OK = 0,
  NTM_B200_ERR_BAD_SHAPE = 1,      /* reference: ValueError in _linear, ntm_cell.py:334-347 */
  NTM_B200_ERR_BAD_SHIFT = 2,      /* reference: assert in circular_shift, ops.py:231 */
  NTM_B200_ERR_NULL_POINTER = 3,
  NTM_B200_ERR_UNSUPPORTED_HEADS = 4,
  NTM_B200_ERR_TOO_LARGE = 5,      /* per-sequence state does not fit an 8-CTA cluster */
  NTM_B200_ERR_WORKSPACE = 6,      /* workspace / packed buffer too small */
  NTM_B200_ERR_NO_DEVICE = 7,      /* no CUDA device of compute capability 10.x */
  NTM_B200_ERR_CUDA = 8,           /* a CUDA runtime call failed; see ntm_b200_last_cuda_error */
  NTM_B200_ERR_DEVICE_TIMEOUT = 9  /* device-side grid barrier timed out (reported by ntm_b200_finish) */
} ntm_b200_status;

/* Constructor arguments of NTMCell (ntm_cell.py:18-20) plus the input width
 * (the reference infers it from the `inputs` tensor, ntm_cell.py:103). */
typedef struct ntm_b200_shape {
  int32_t input_dim;             /* D: width of `inputs` */
  int32_t output_dim;            /* O */
  int32_t mem_size;              /* N */
  int32_t mem_dim;               /* M */
  int32_t shift_range;           /* S = 2*shift_range + 1 taps */
  int32_t controller_hidden_size;/* C */
  int32_t controller_num_layers; /* L */
  int32_t write_head_size;       /* W */
  int32_t read_head_size;        /* R */
  int32_t write_first;           /* ntm_cell.py:212-215 */
} ntm_b200_shape;

/* Trainable variables in the reference's TensorFlow layout:
 *   lstm_w[l]  ntm-cell/lstm-controller/cell_l/basic_lstm_cell/weights  [in_l + C, 4C]
 *              in_0 = D + R*M (rows ordered x | read head-major | h), in_l = C
 *   lstm_b[l]  .../biases [4C]            gate order i, j, f, o
 *   addr_w/b   ntm-cell/addressing/{weights [C,P], biases [P]}  (_linear, ntm_cell.py:124)
 *   out_w/b    ntm-cell/{weights [C,O], biases [O]}             (_linear, ntm_cell.py:220)  */
typedef struct ntm_b200_weights {
  const float* lstm_w[NTM_B200_MAX_LAYERS];
  const float* lstm_b[NTM_B200_MAX_LAYERS];
  const float* addr_w;
  const float* addr_b;
  const float* out_w;
  const float* out_b;
} ntm_b200_weights;

/* The state dict {'M','w','read','controller_state'} (ntm_cell.py:223-228,
 * 255-315).  Element [b] of each tensor starts at ptr + b*stride (in floats);
 * a stride of 0 broadcasts one copy to every sequence (what zero_state's
 * tf.stack([M]*batch) expresses, ntm_cell.py:296,301,306). */
typedef struct ntm_b200_state {
  float* M;                 /* [B, N, M] */
  float* w;                 /* [B, R+W, N]  (read heads first) */
  float* read;              /* [B, R, M] */
  float* controller_state;  /* [B, 2*C*L]  per layer: c then h */
  int64_t stride_M, stride_w, stride_read, stride_controller_state;
} ntm_b200_state;

/* Launch geometry the library will use for a shape (all fields outputs). */
typedef struct ntm_b200_plan {
  int32_t cluster_size;        /* CTAs sharing one sequence's memory rows (DSMEM) */
  int32_t rows_per_cta;        /* memory rows resident in each CTA's shared memory */
  int32_t sequences_resident;  /* sequences advanced concurrently per wave */
  int32_t threads_per_cta;
  int32_t ctas_per_sm;         /* co-resident CTAs per SM the chosen kernel shape is built for */
  int32_t teams;               /* decoupled groups of CTAs, each advancing its own sequences */
  int64_t smem_bytes_per_cta;
  int64_t workspace_bytes;     /* for ntm_b200_forward_seq with this (B, T) */
  int64_t packed_bytes;        /* for ntm_b200_pack_weights */
  int64_t debug_floats_per_sequence; /* record size of the debug-tap buffer */
} ntm_b200_plan;

int32_t ntm_b200_abi_version(void);
const char* ntm_b200_status_string(int32_t status);
/* Text of the last CUDA error seen by this thread inside the library. */
const char* ntm_b200_last_cuda_error(void);

/* Shape validation + launch plan.  Pure host arithmetic: works without a GPU
 * (assumes a B200: 148 SMs, 227 KiB shared memory per CTA) so the planner is
 * unit-testable on CPU.  Replaces nothing in the reference (TF sizes its own
 * kernels); it is the "smem/cluster plan" entry SURVEY.md s8b asks for. */
int32_t ntm_b200_query(const ntm_b200_shape* shape, int64_t batch, int64_t steps,
                       ntm_b200_plan* plan_out);

/* Which execution mode ntm_b200_forward_seq will pick for `batch` sequences of this shape:
 * *mode_out = 0: the persistent kernel with the memories resident in shared memory (batches of up to a few
 * waves of resident sequences), 1: the streaming kernels (larger batches; memories streamed from HBM once per
 * step).  Pure host arithmetic like ntm_b200_query; honours NTM_B200_MODE / NTM_B200_STREAM_MIN_BATCH.  Calls that
 * request debug taps always run resident. */
int32_t ntm_b200_query_mode(const ntm_b200_shape* shape, int64_t batch, int32_t* mode_out);

/* One-time repacking of the variables into the kernel's layout (the
 * [C, P+O] concatenation of the two _linear projections of ntm_cell.py:124,220).
 * `packed` is a device buffer of plan.packed_bytes.  Must be re-run when the
 * variables change.  Reference counterpart: tf.train.Saver.restore +
 * variable creation in _linear (ntm_cell.py:354-369). */
int32_t ntm_b200_pack_weights(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                              void* packed, int64_t packed_bytes, void* stream);

/* LoopNTMTracker.__call__ (ntm_tracker_new.py:13-64): T cell steps for B
 * sequences.  inputs [B,T,D] batch-major; logits/outputs [B,T,O] batch-major
 * (`outputs` = softmax of logits, ntm_cell.py:221; may be NULL).
 * state_in may alias state_out.  debug_taps (may be NULL) receives, for the LAST
 * step, plan.debug_floats_per_sequence floats per sequence laid out as
 * [k H*M | beta H | g H | sw H*S | gamma H | erase W*M | add W*M |
 *  similarity H*N | w_content_focused H*N | w_gated H*N | w_conv H*N |
 *  w_conv_powed H*N]  -- the taps of the reference's `debug` dict
 * (ntm_cell.py:230-250) that are not derivable from the state. */
int32_t ntm_b200_forward_seq(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                             const void* packed, int64_t batch, int64_t steps,
                             const float* inputs, const ntm_b200_state* state_in,
                             const ntm_b200_state* state_out, float* logits, float* outputs,
                             float* debug_taps, void* workspace, int64_t workspace_bytes,
                             void* stream);

/* Next block of steps of the sequences the previous ntm_b200_forward_seq call on this thread advanced: same
 * workspace, batch, shape and packed weights, `state` = that call's state_out, updated in place.  In streaming
 * mode the per-call initialisation (column norms of the memory, operand packing) is skipped because the
 * workspace still holds it; in every other case this is an ordinary in-place ntm_b200_forward_seq.  The
 * reference's counterpart is feeding the fetched state back into the next sess.run (test_tracker.py:284-299). */
int32_t ntm_b200_forward_seq_continue(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                                      const void* packed, int64_t batch, int64_t steps, const float* inputs,
                                      const ntm_b200_state* state, float* logits, float* outputs,
                                      void* workspace, int64_t workspace_bytes, void* stream);

/* Training-mode history: what the backward pass needs from the forward pass (the reference gets
 * it from TensorFlow's while_loop gradient machinery, `swap_memory=True`,
 * ntm_tracker_new.py:34-40).  Every pointer is an optional caller-owned device buffer:
 *   M_prev   [T, B, N, M]    memory entering step t
 *   w_prev   [T, B, R+W, N]  weightings entering step t
 *   params   [T, B, PO4]     raw (pre-activation) head parameters + logits of step t,
 *                            PO4 = round_up(P + O, 4), layout of ntm_cell.py:126-130 then logits
 *   z        [T, B, L, 4, C] LSTM gate pre-activations i, j, f, o of step t
 *   c, h     [T+1, B, L, C]  LSTM cell / hidden state: slot 0 = initial, slot t+1 = after step t
 *   read     [T+1, B, R*M]   read vectors: slot 0 = initial, slot t+1 = produced by step t
 *   sim      [T, B, R+W, N]  (optional) un-normalised similarities of step t: sum_d tanh(k)[h,d] * cn[d] *
 *                            M_prev[n,d]  (the key's own 1/|k| not yet applied, ops.py:152-156)
 *   cn       [T, B, M]       (optional) inverse column norms of M_prev[t] (tf.nn.l2_normalize over N,
 *                            ops.py:147-150)
 * sim and cn are by-products of the forward pass; with both recorded the backward pass skips two of its five
 * sweeps over the memory (it re-derives them otherwise). */
typedef struct ntm_b200_history {
  float* M_prev;
  float* w_prev;
  float* params;
  float* z;
  float* c;
  float* h;
  float* read;
  float* sim;
  float* cn;
} ntm_b200_history;

/* ntm_b200_forward_seq that also records `history` (may be NULL = plain forward). */
int32_t ntm_b200_forward_seq_train(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                                   const void* packed, int64_t batch, int64_t steps,
                                   const float* inputs, const ntm_b200_state* state_in,
                                   const ntm_b200_state* state_out, float* logits, float* outputs,
                                   float* debug_taps, const ntm_b200_history* history,
                                   void* workspace, int64_t workspace_bytes, void* stream);

/* One reverse-time step of the backward pass through the memory / addressing part of the cell
 * (everything between the head-parameter projection and the memory: ntm_cell.py:133-215,
 * ops.py:135-242) for `batch` sequences.  The reference obtains this from tf.gradients of the
 * unrolled graph (direct_offset_output.py:611-613).  Inputs: the recorded history of that step
 * (M_prev [B,N,M], w_prev [B,H,N], raw_params [B,PO4]) and the upstream gradients d_read [B,R,M]
 * (w.r.t. the read vectors the step produced), d_w [B,H,N] (w.r.t. the weightings it produced)
 * and dM [B,N,M] (w.r.t. the memory it produced).  Outputs: dM is overwritten with the gradient
 * w.r.t. M_prev; d_w_prev [B,H,N]; d_raw_params [B,PO4] (logit slots zeroed).  The dense
 * projections' gradients are plain GEMMs and stay with the caller. */
int32_t ntm_b200_memory_backward_step(const ntm_b200_shape* shape, int64_t batch, const float* M_prev,
                                      const float* w_prev, const float* raw_params,
                                      const float* d_read, const float* d_w, float* dM,
                                      float* d_w_prev, float* d_raw_params, void* stream);

/* The data formats either side of the path.
 * serialize: conv features [B, L, F, Cch] + first-frame target map [B, F] -> tracker inputs
 * [B, L*(F+1), Cch+2] with a delimiter row per frame (last row of the frame as in training,
 * direct_offset_output.py:439-500; first row when delimiter_first != 0 as in the serve path,
 * test_tracker.py:385-404), channel Cch = delimiter bit, channel Cch+1 = target[b, f] on the F feature
 * rows of the FIRST frame only (training layout: steps 0..F-1; serve layout: steps 1..F), 0 on every
 * delimiter row and on all later frames.
 * gather: logits [B, L*(F+1), O] -> tanh(logits at each frame's delimiter row, first frame
 * dropped) [B, L-1, O]  (direct_offset_output.py:581-593). */
int32_t ntm_b200_serialize_tracker_inputs(const float* features, const float* target, float* inputs,
                                          int64_t batch, int32_t frames, int32_t num_features,
                                          int32_t channels, int32_t delimiter_first, void* stream);
int32_t ntm_b200_gather_offsets(const float* logits, float* offsets, int64_t batch, int32_t frames,
                                int32_t num_features, int32_t output_dim, void* stream);

/* NTMCell.zero_state (ntm_cell.py:284-315): the initial memory / weightings / read vectors of every sequence are
 * tanh / sigmoid / tanh of the three init_state variables (the weighting is NOT normalised: reference quirk 4).
 * One launch over the three arrays ([N,M], [H,N], [R,M] floats); the batch dimension is a stride-0 broadcast on
 * the caller's side. */
int32_t ntm_b200_zero_state(const float* var_M, int64_t n_M, const float* var_w, int64_t n_w, const float* var_read,
                            int64_t n_read, float* M, float* w, float* read, void* stream);

/* ntm_b200_forward_seq for frames in FEATURE layout: conv features [B, L, F, Cch] + first-frame target map [B, F]
 * (what direct_offset_output.py:439-500 concatenates / tiles / reshapes into the tracker inputs, and what
 * test_tracker.py:385-404 builds per frame on the serve path), with shape->input_dim == Cch + 2 and T = L*(F+1)
 * steps.  In streaming mode at tracker shapes (Cch a whole number of 64-wide K atoms) the serialised rows are never
 * materialised: the input projection's pack pass reads the 16-byte aligned feature rows where they lie and
 * synthesises the delimiter / target channels (identical values to ntm_b200_serialize_tracker_inputs followed by
 * ntm_b200_forward_seq, bit for bit); otherwise the rows are materialised behind the workspace and the call proceeds
 * as ntm_b200_forward_seq.  Workspace: ntm_b200_features_workspace_bytes (-1 on a bad shape). */
int64_t ntm_b200_features_workspace_bytes(const ntm_b200_shape* shape, int64_t batch, int32_t frames, int32_t num_features);
int32_t ntm_b200_forward_seq_features(const ntm_b200_shape* shape, const ntm_b200_weights* weights, const void* packed,
                                      int64_t batch, int32_t frames, int32_t num_features, const float* features,
                                      const float* target, int32_t delimiter_first, const ntm_b200_state* state_in,
                                      const ntm_b200_state* state_out, float* logits, float* outputs, void* workspace,
                                      int64_t workspace_bytes, void* stream);

/* Host -> device copy of the frames [t0, t1) of every sequence: inputs_host [B, T, D] (page-locked for a
 * truly asynchronous copy) -> frames_dev [B, t1 - t0, D] contiguous, as one strided DMA on `stream`.  Lets a
 * caller that holds the frames on the host (the reference feeds them through feed_dict on every sess.run,
 * direct_offset_output.py:300-330, test_tracker.py:284-299) overlap the upload of the next block of steps
 * with ntm_b200_forward_seq on the current one (state_in of a block = state_out of the previous block). */
int32_t ntm_b200_copy_frames_h2d(float* frames_dev, const float* inputs_host, int64_t batch, int64_t steps,
                                 int64_t input_dim, int64_t t0, int64_t t1, void* stream);

/* Backward of one BasicLSTMCell layer at one timestep for `batch` sequences (elementwise part; the
 * two GEMMs around it stay with the caller).  dh = dh_a + dh_b (dh_b may be NULL) is the gradient
 * w.r.t. the layer's new hidden state, dc [B,C] carries the cell-state gradient in and out, z holds
 * the recorded gate pre-activations i|j|f|o of sequence b at z + b*z_stride, c_prev / c_new the cell
 * state before / after the step at + b*c_stride, dz receives the pre-activation gradients at
 * dz + b*dz_stride (4*C floats). */
int32_t ntm_b200_lstm_backward_step(int64_t batch, int32_t hidden, const float* dh_a, const float* dh_b,
                                    const float* z, int64_t z_stride, const float* c_prev,
                                    const float* c_new, int64_t c_stride, float* dc, float* dz,
                                    int64_t dz_stride, void* stream);

/* ---- training step (BASELINE configs[4]; direct_offset_output.py:581-626) -------------------------------- */

/* Gradients w.r.t. the trainable variables, same layout as ntm_b200_weights plus the three init_state
 * variables (ntm_cell.py:292-306).  Caller-owned device buffers (typically views into ONE flat buffer, so that
 * a single NCCL all-reduce and a single optimizer pass cover everything).  init_* may be NULL (skipped). */
typedef struct ntm_b200_grads {
  float* lstm_w[NTM_B200_MAX_LAYERS];
  float* lstm_b[NTM_B200_MAX_LAYERS];
  float* addr_w;
  float* addr_b;
  float* out_w;
  float* out_b;
  float* init_M;     /* [N, M]   through tanh    */
  float* init_w;     /* [R+W, N] through sigmoid */
  float* init_read;  /* [R, M]   through tanh    */
} ntm_b200_grads;

/* Loss of the tracker and its gradient w.r.t. the logits (direct_offset_output.py:581-606):
 * y = tanh(logits[:, steps[g], :]); loss = tf.nn.l2_loss(y - targets) = 0.5 * sum (y - targets)^2.
 * logits [B,T,O]; targets [B,G,O]; gather_steps: HOST array of G <= 64 timesteps; loss_out: device scalar;
 * dlogits [B,T,O] receives dLoss/dlogits (zero at the steps that do not enter the loss).  One kernel,
 * fixed-order reduction. */
int32_t ntm_b200_offset_loss(const float* logits, const float* targets, const int32_t* gather_steps,
                             int32_t num_gather, int64_t batch, int64_t steps, int32_t output_dim,
                             float* loss_out, float* dlogits, void* stream);

/* Workspace the backward pass needs for this geometry (device bytes). */
int64_t ntm_b200_backward_workspace_bytes(const ntm_b200_shape* shape, int64_t batch, int64_t steps);

/* The whole backward pass through the unrolled loop -- what tf.gradients(loss, tf.trainable_variables())
 * derives op by op in the reference (direct_offset_output.py:611-613) -- as ONE call: the reverse-time loop
 * runs inside the library (per step: the fused memory/addressing backward kernel, the LSTM gate backward and
 * the two data-gradient GEMMs on the tensor cores), then the weight gradients as large-K tcgen05 GEMMs over
 * all (t, b), the bias gradients and the init_state gradients.  `history` = what ntm_b200_forward_seq_train
 * recorded for the same inputs (all seven buffers); `dlogits` [B,T,O] from ntm_b200_offset_loss (or any other
 * loss); `state0` = the state the forward started from (zero_state: its stride-0 tensors hold tanh / sigmoid /
 * tanh of the init_state variables, which is what their gradient needs; per-sequence initial states get no
 * init_state gradient).  No cuBLAS, no host synchronisation; gradients are written, not accumulated. */
int32_t ntm_b200_backward_seq(const ntm_b200_shape* shape, const ntm_b200_weights* weights, const void* packed,
                              int64_t batch, int64_t steps, const float* inputs,
                              const ntm_b200_history* history, const float* dlogits,
                              const ntm_b200_state* state0, const ntm_b200_grads* grads, void* workspace,
                              int64_t workspace_bytes, void* stream);

/* tf.clip_by_global_norm(grads, clip_norm) + RMSPropOptimizer(lr, decay, momentum, epsilon).apply_gradients
 * (direct_offset_output.py:606-626) fused over flat buffers of n floats: global norm (two-stage fixed-order
 * reduction), then ONE pass  g *= clip_norm / max(norm, clip_norm);  rms = decay*rms + (1-decay)*g*g;
 * mom = momentum*mom + lr*g/sqrt(rms + epsilon);  param -= mom.  `grads` holds the (all-reduced) gradient,
 * `scratch` >= 4 KiB of device memory, gnorm_out a device scalar (the unclipped global norm).  TF slot
 * initialisation is the caller's: rms = 1, mom = 0. */
int32_t ntm_b200_rmsprop_step(float* params, const float* grads, float* rms, float* mom, int64_t n,
                              float learning_rate, float decay, float momentum, float epsilon, float clip_norm,
                              float* gnorm_out, void* scratch, void* stream);

/* out[r, j] = sum_k a[r*lda + k] * b[j*ldb + k]  (A @ B^T; fp32 in, fp32 out, ~fp32 accuracy through the
 * 3-term bf16 operand split on the tensor cores): the tile-record tcgen05 GEMM the backward pass is built from
 * (csrc/ntm_b200_gemm_tiles.cuh), exposed on plain row-major operands so that it can be tested and reused on
 * its own.  kslices >= 1 = K slices summed in slice order (deterministic).  Replaces the MatMul gradient ops
 * TensorFlow derives for ntm_cell.py:101-105,124-130,220. */
int64_t ntm_b200_gemm_nt_workspace_bytes(int64_t nrows, int64_t ncols, int64_t K, int32_t kslices);
int32_t ntm_b200_gemm_nt(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
                         int64_t nrows, int64_t ncols, int64_t K, int32_t kslices, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* NTMCell.__call__ (ntm_cell.py:53-253): one step, inputs [B,D], logits/outputs
 * [B,O].  The serve path's unit of work (test_tracker.py:284-299). */
int32_t ntm_b200_step(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                      const void* packed, int64_t batch, const float* inputs,
                      const ntm_b200_state* state_in, const ntm_b200_state* state_out,
                      float* logits, float* outputs, float* debug_taps, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* Synchronise `stream` and report device-side failures of earlier calls that
 * used `workspace` (grid-barrier timeout).  Optional: a caller that never
 * checks still gets correct results when nothing failed. */
int32_t ntm_b200_finish(void* workspace, void* stream);

/* Measurement hooks for the bench harness: when enabled, ntm_b200_forward_seq
 * brackets its two kernels (hoisted x-projection, persistent sequence kernel)
 * with CUDA events on the caller's stream; after the stream has been
 * synchronised ntm_b200_last_kernel_ms returns their device durations. */
int32_t ntm_b200_set_profiling(int32_t enable);
int32_t ntm_b200_last_kernel_ms(float* xproj_ms, float* seq_kernel_ms);
/* Streaming mode (large batches): with profiling enabled, device time of the last
 * call summed over its timesteps, out4 = {controller GEMM + LSTM gates, head-parameter GEMM, fused
 * addressing/memory kernel, state initialisation} in ms; *steps = number of timesteps (0 when the last
 * call on this thread did not run in streaming mode). */
int32_t ntm_b200_last_stream_ms(float* out4, int32_t* steps);
/* Streaming mode, profiling enabled: mean duration in ns, over the CTAs (= sequences) of the last
 * memory-kernel launch, of {wait for the head parameters, activations, pass 1, addressing, pass 2, store
 * drain, finalize, whole CTA}; out12[8] = first CTA start to last CTA end; out12[9..11] = SM cycles warp 0
 * spent in pass 1 {waiting for ring stages, computing, team barrier + bulk-copy issue}.  Synchronous. */
int32_t ntm_b200_stream_phase_ns(double* out12, int32_t* ctas);
/* Training (ntm_b200_backward_seq), profiling enabled, stream synchronised: device ms of this thread's last
 * backward call, out4 = {memory/addressing backward kernel summed over the steps, rest of the reverse-time
 * loop (data-gradient GEMMs + LSTM gate backward), weight-gradient GEMMs + init_state sums, whole call};
 * *steps = number of timesteps (0 = nothing recorded).  Measurement hook for bench.py's training roofline
 * (reference counterpart: none -- tf.gradients has no per-op timing, direct_offset_output.py:611-613). */
int32_t ntm_b200_last_backward_ms(float* out4, int32_t* steps);
/* With profiling enabled the persistent kernel also accumulates, per CTA, SM-clock
 * cycles spent in each phase of the timestep (16 int64 slots per CTA: 0/2/4/6 =
 * phases A/B/C/D compute, 1/3/5/7 = the device-wide barrier after each, 8 =
 * per-wave prologue, 9 = epilogue).  Copies max_ctas*16 counters to host `out`
 * (synchronous; call after the stream has been synchronised). */
int32_t ntm_b200_phase_cycles(const void* workspace, int64_t* out, int32_t max_ctas);

/* Geometry of this thread's last ntm_b200_forward_seq launch, 16 ints: {tensor path used (0/1),
 * resident sequences, CTAs, cluster size, K-slices and K-slice width of the layer-0 controller
 * GEMM, K-slices and K-slice width of the head-parameter GEMM, teams, threads per CTA, CTAs per
 * SM, shared-memory bytes per CTA, x-projection on the tensor path (0/1), execution mode (0 = persistent
 * shared-memory-resident kernel, 1 = streaming), continuation taken (0/1), 0}. */
int32_t ntm_b200_last_launch_info(int32_t* out16);

/* Number of kernel launches the library has issued in this process (for the
 * bench harness' `gpu_launches` claim). */
int64_t ntm_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NTM_B200_H_ */
