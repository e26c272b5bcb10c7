"""LoopNTMTracker -- drop-in for the reference's batched sequence driver
(ntm_tracker_new.py:4-64): unrolls the NTM cell over the T frames-steps of B
independent sequences.  The reference does it with tf.while_loop + TensorArrays
(one graph dispatch per op per step); here the whole loop is ONE call through the
C ABI (ntm_b200_forward_seq), which runs it in one of two ways:
small batches as ONE persistent kernel that keeps every sequence's memory in shared
memory for all T steps; large batches (BASELINE's 4096 sequences) in streaming mode,
all sequences in lockstep with a few fused kernels per timestep and the memories in HBM.
"""
import numpy as np
import torch

from .ntm_cell import NTMCell, random_uniform_initializer


class LoopNTMTracker(object):
    def __init__(self, sequence_length, output_dim,
                 initializer=random_uniform_initializer(-.1, .1), **kwargs):
        self.cell = NTMCell(output_dim, **kwargs)
        self.initializer = initializer
        self.sequence_length = sequence_length
        self.final_state = None
        self.history = None

    # The reference's loop also fills TensorArrays Ms / ws / reads with the memory, weightings and read
    # vectors AFTER every step (ntm_tracker_new.py:22-26,57-61; their stacking into the return value is
    # commented out there, :46-49, as are the summaries that used them, direct_offset_output.py:549-575).
    # Set record_history = True to get them: after a call, ``tracker.history`` holds time-major
    # {'Ms': [T,B,N,M], 'ws': [T,B,H,N], 'reads': [T,B,R,M]} device tensors (the reference would stack
    # them on the LAST axis: ``Ms.permute(1, 2, 3, 0)``).  Costs T copies of the state in HBM, like the
    # reference's swap_memory TensorArrays.
    record_history = False

    def _history_buffers(self, B, T):
        cell, dev = self.cell, self.cell.device
        f32 = dict(dtype=torch.float32, device=dev)
        N, M, H, R = cell.mem_size, cell.mem_dim, cell.num_heads, cell.read_head_size
        return {"M_prev": torch.empty(T + 1, B, N, M, **f32), "w_prev": torch.empty(T + 1, B, H, N, **f32),
                "read": torch.empty(T + 1, B, R * M, **f32)}

    def _publish_history(self, bufs, new_state, B, T):
        bufs["M_prev"][T].copy_(new_state["M"])       # slot t of the recording = state ENTERING step t; the
        bufs["w_prev"][T].copy_(new_state["w"])       # last step's result is the returned state
        self.history = {"Ms": bufs["M_prev"][1:], "ws": bufs["w_prev"][1:],
                        "reads": bufs["read"][1:].view(T, B, self.cell.read_head_size, self.cell.mem_dim)}

    def __call__(self, inputs, state=None, scope=None):
        """inputs [B, T, D] batch-major -> (outputs [B,T,O], output_logits [B,T,O])
        (ntm_tracker_new.py:42-49).  T must equal ``sequence_length``.  Host
        inputs (NumPy / CPU tensors, ideally pinned) are copied to the device and
        the results are returned on the host, like a ``sess.run`` of the reference
        graph; CUDA inputs stay on the device.  The final state dict is kept in
        ``self.final_state`` (the reference's loop_vars M, w, read, controller_state)."""
        host_in = isinstance(inputs, np.ndarray) or not inputs.is_cuda
        as_numpy = isinstance(inputs, np.ndarray)
        if host_in and inputs.ndim == 3 and not self.record_history and self._time_blocks(inputs) > 1:
            if inputs.shape[1] != self.sequence_length:
                raise ValueError("inputs have %d steps but sequence_length is %d" % (inputs.shape[1], self.sequence_length))
            if self.cell.input_dim is None:
                self.cell.build(inputs.shape[2], self.initializer)
            state = state or self.cell.zero_state(inputs.shape[0], self.initializer)
            return self._call_host_time_pipelined(inputs, state, self._time_blocks(inputs))
        if host_in and inputs.ndim == 3 and not self.record_history and \
                self._pipeline_chunks(inputs.shape[0], inputs.shape[1]) > 1:
            x = None
            B, T, D = inputs.shape
        else:
            x = self.cell._prepare_inputs(inputs, 3)
            B, T, D = x.shape
        if T != self.sequence_length:
            raise ValueError("inputs have %d steps but sequence_length is %d" % (T, self.sequence_length))
        if self.cell.input_dim is None:
            self.cell.build(D, self.initializer)
        state = state or self.cell.zero_state(B, self.initializer)
        if x is None:
            return self._call_host_pipelined(inputs, state, as_numpy)
        bufs = self._history_buffers(B, T) if self.record_history else None
        logits, outputs, new_state, _ = self.cell._run(x, state, T, history=bufs)
        self.final_state = new_state
        if bufs is not None:
            self._publish_history(bufs, new_state, B, T)
        if host_in:
            self.cell.finish()          # the call synchronises here anyway: surface device-side failures
            outputs, logits = outputs.cpu(), logits.cpu()
            if as_numpy:
                outputs, logits = outputs.numpy(), logits.numpy()
        return (outputs, logits)

    def call_features(self, features, target, delimiter_first=False, state=None):
        """The same loop for frames in FEATURE layout: features [B, L, F, Cch] (conv4_3 vectors at the F sampled points
        of each of L frames, CUDA) and target [B, F] (first-frame ground-truth map) -- what the reference's trainer
        concatenates / tiles / reshapes into ``inputs`` before calling the tracker (direct_offset_output.py:439-500).
        Here the delimiter and target channels are synthesised inside the library (streaming mode: by the input
        projection's pack pass, the serialised [B, T, Cch+2] rows never exist).  T = L*(F+1) must equal
        ``sequence_length``.  Returns (outputs, output_logits) on the device, bit-identical to
        ``self(serialize.tracker_inputs(features, target, delimiter_first))``."""
        B, L, F, Cch = features.shape
        if L * (F + 1) != self.sequence_length:
            raise ValueError("features hold %d steps but sequence_length is %d" % (L * (F + 1), self.sequence_length))
        if self.cell.input_dim is None:
            self.cell.build(Cch + 2, self.initializer)
        state = state or self.cell.zero_state(B, self.initializer)
        logits, outputs, new_state = self.cell._run_features(features, target, state, delimiter_first)
        self.final_state = new_state
        return (outputs, logits)

    # Host-resident frames (page-locked torch tensor): the call is cut into blocks of timesteps and the
    # upload of block i+1 (one strided DMA, ntm_b200_copy_frames_h2d) overlaps the kernels of block i; the
    # state is carried from block to block on the device, every sequence stays in every launch, so the
    # kernels run at full batch.  None = automatic (once the frames exceed 64 MB: a short first block, so
    # that little of the upload is exposed, then blocks of ~16 steps), 1 = off, n = n equal blocks.
    time_blocks = None

    def _time_blocks(self, inputs):
        if isinstance(inputs, np.ndarray) or inputs.dtype != torch.float32 or not inputs.is_contiguous() \
                or not inputs.is_pinned():
            return 1
        B, T, D = inputs.shape
        n = self.time_blocks
        if n is None:
            n = len(self._auto_bounds(T)) if (B * T * D * 4 >= (64 << 20) and T >= 16) else 1
        return max(1, min(int(n), T))

    first_cuts = (2, 8, 16)      # class attribute: ends of the first three blocks (experiments: LoopNTMTracker.first_cuts = ...)

    @classmethod
    def _auto_bounds(cls, T):
        """[0,2) [2,8) [8,16) [16,32) ... : only the first 2 steps' upload is not hidden behind kernels
        (blocks after the first are continuations, so short blocks cost next to nothing)."""
        c0, c1, c2 = cls.first_cuts
        cuts = [0, min(c0, T), min(c1, T)]
        if T > c1:
            cuts.append(min(c2, T))
        while cuts[-1] < T:
            cuts.append(min(cuts[-1] + 16, T))
        cuts = sorted(set(cuts))
        return list(zip(cuts[:-1], cuts[1:]))

    def _call_host_time_pipelined(self, xh, state, n):
        import ctypes as C
        from . import _cabi
        cell, dev = self.cell, self.cell.device
        lib = _cabi.load()
        B, T, D = xh.shape
        bounds = self._auto_bounds(T) if self.time_blocks is None else \
            [(i * T // n, (i + 1) * T // n) for i in range(n)]
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        copy.wait_stream(compute)
        staged = []
        for t0, t1 in bounds:                      # all uploads are queued on the copy stream up front
            xd = torch.empty(B, t1 - t0, D, dtype=torch.float32, device=dev)
            with torch.cuda.stream(copy):
                _cabi.check(lib.ntm_b200_copy_frames_h2d(xd.data_ptr(), xh.data_ptr(), B, T, D, t0, t1,
                                                         C.c_void_p(copy.cuda_stream)), "copy_frames_h2d")
                ev = torch.cuda.Event()
                ev.record(copy)
            staged.append((xd, ev))
        # one workspace for all blocks (sized for the longest), so that every block after the first is a
        # continuation: state updated in place, column norms / operand tiles / packed weights reused
        ws_bytes = max(cell.plan(B, t1 - t0)["workspace_bytes"] for t0, t1 in bounds)
        ws = getattr(self, "_block_ws", None)
        if ws is None or ws.numel() < ws_bytes or ws.device != dev:
            ws = self._block_ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=dev)
        outs, logs = [], []
        for i, ((t0, t1), (xd, ev)) in enumerate(zip(bounds, staged)):
            compute.wait_event(ev)
            xd.record_stream(compute)
            lg, out, state, _ = cell._run(xd, state, t1 - t0, workspace=ws,
                                          continuation=(i > 0 and not cell.debug))
            outs.append(out); logs.append(lg)
        self.final_state = state
        out_d, log_d = torch.cat(outs, 1), torch.cat(logs, 1)
        # results leave through page-locked buffers (torch's caching host allocator: no cudaHostAlloc per
        # call) as two asynchronous copies and ONE synchronisation, which also checks the error flag
        out_h = torch.empty(out_d.shape, dtype=torch.float32, pin_memory=True)
        log_h = torch.empty(log_d.shape, dtype=torch.float32, pin_memory=True)
        out_h.copy_(out_d, non_blocking=True)
        log_h.copy_(log_d, non_blocking=True)
        cell.finish()
        return out_h, log_h

    # Number of batch chunks a host-resident call is split into so that the host->device copy of
    # chunk i+1 overlaps the kernels of chunk i (sequences are independent, so any split is exact).
    # Default 1 (off): measured on B200 at C3 (B=4096, T=64, 539 MB of frames) the single copy costs
    # 9.7 ms of a 128 ms call, while 4 chunks take 226 ms -- the persistent cooperative kernels do not
    # overlap usefully with the queued DMA -- so splitting is left as an opt-in.
    host_chunks = 1

    def _pipeline_chunks(self, B, T):
        """Chunks for a host-resident call: only when every chunk still spans several waves of
        resident sequences (otherwise splitting would just idle SMs)."""
        if self.host_chunks <= 1 or self.cell.input_dim is None:
            return 1
        resident = self.cell.plan(B, T)["sequences_resident"]
        return self.host_chunks if B >= 4 * self.host_chunks * resident else 1

    def _call_host_pipelined(self, inputs, state, as_numpy):
        cell, dev = self.cell, self.cell.device
        xh = torch.from_numpy(inputs) if isinstance(inputs, np.ndarray) else inputs
        xh = xh.float()
        B, T, D = xh.shape
        if cell.input_dim is None:
            cell.build(D, self.initializer)
        n = self._pipeline_chunks(B, T)
        bounds = [(i * B // n, (i + 1) * B // n) for i in range(n)]
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        copy.wait_stream(compute)
        staged = []
        for lo, hi in bounds:                      # all copies are queued on the copy stream up front
            with torch.cuda.stream(copy):
                xd = xh[lo:hi].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            staged.append((xd, ev))
        outs, logs, states = [], [], []
        for (lo, hi), (xd, ev) in zip(bounds, staged):
            compute.wait_event(ev)
            xd.record_stream(compute)
            st = {k: v[lo:hi] for k, v in state.items()}
            lg, out, ns, _ = cell._run(xd.contiguous(), st, T)
            outs.append(out); logs.append(lg); states.append(ns)
        cell.finish()
        outputs, logits = torch.cat(outs, 0).cpu(), torch.cat(logs, 0).cpu()
        self.final_state = {k: torch.cat([s_[k] for s_ in states], 0) for k in states[0]}
        if as_numpy:
            outputs, logits = outputs.numpy(), logits.numpy()
        return (outputs, logits)
