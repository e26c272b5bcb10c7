"""Training step for the NTM tracker path (BASELINE config 5).

What the reference does (direct_offset_output.py:581-626): gather the logits at every frame's
delimiter step (first frame dropped), ``tanh``, ``tf.nn.l2_loss`` against the offsets,
``tf.gradients`` through the unrolled while_loop, ``tf.clip_by_global_norm(5)``,
``RMSPropOptimizer(1e-4, decay=0.95, momentum=0.9)``.  It is single-device.

Here every stage is a call into libntm_b200.so (no torch arithmetic, no cuBLAS):
  * forward  -- ``ntm_b200_forward_seq_train``: the forward kernels, recording the history the
                backward needs (memories / weightings entering each step, raw head parameters, LSTM
                gate pre-activations, c / h, read vectors);
  * loss     -- ``ntm_b200_offset_loss``: tanh + l2_loss of the gathered logits and dLoss/dlogits;
  * backward -- ``ntm_b200_backward_seq``: the reverse-time loop inside the library (fused
                memory/addressing backward kernel, LSTM gate backward, data-gradient GEMMs on the tensor
                cores), then the weight gradients as large-K tcgen05 GEMMs over all (t, b); gradients land
                in ONE flat buffer;
  * multi-GPU -- sequences are sharded over ranks (no collective in forward or backward); the ONE
                collective of the path is an all-reduce (sum) of that flat gradient over NCCL;
  * update   -- ``ntm_b200_rmsprop_step``: global norm, clip and RMSProp + momentum fused over the
                flat buffers, identical on every rank, so the replicas stay identical.
The variables of the cell are re-homed as views into one flat parameter buffer (sorted by TF variable
name), which is what makes the single all-reduce / single optimizer pass possible.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _cabi
from .ntm_tracker_new import LoopNTMTracker


def delimiter_steps(T, frame):
    """Timesteps whose logits enter the loss: the last row of every frame but the first
    (direct_offset_output.py:581-588 with frame = num_features + 1 = 65)."""
    return [f * frame + frame - 1 for f in range(1, T // frame)]


ALIGN = 64      # floats: every variable starts on a 256-byte boundary of the flat buffers


class NTMTrainer(object):
    def __init__(self, tracker: LoopNTMTracker, learning_rate=1e-4, decay=0.95, momentum=0.9,
                 max_gradient_norm=5.0, epsilon=1e-10, frame=65):
        self.tracker = tracker
        self.cell = tracker.cell
        self.lr, self.decay, self.momentum = learning_rate, decay, momentum
        self.clip, self.eps, self.frame = max_gradient_norm, epsilon, frame
        self._flat = None     # parameters; cell.variables are views into it
        self._grad = None     # gradient, same offsets
        self._rms = None      # TF's RMSProp slots: 'rms' starts at ones, 'momentum' at zeros
        self._mom = None
        self._offsets = None
        self._bufs = {}
        self.global_step = 0
        self.last_gnorm = None   # device scalar: unclipped global norm of the last update

    # ------------------------------------------------------------------ flat buffers --
    def _ensure_flat(self):
        """Re-home the cell's variables as views into one flat buffer (name-sorted, 256-byte aligned)."""
        V = self.cell.variables
        names = sorted(V)
        if self._offsets is not None and [n for n, _, _ in self._offsets] == names and \
                all(V[n].data_ptr() == self._flat.data_ptr() + 4 * off for n, off, _ in self._offsets):
            return
        dev = self.cell.device
        offs, o = [], 0
        for n in names:
            offs.append((n, o, V[n].numel()))
            o += (V[n].numel() + ALIGN - 1) // ALIGN * ALIGN
        flat = torch.zeros(o, dtype=torch.float32, device=dev)
        for n, off, k in offs:
            flat[off:off + k].copy_(V[n].reshape(-1))
            V[n] = flat[off:off + k].view(V[n].shape)
        self._flat, self._offsets = flat, offs
        self._grad = torch.zeros_like(flat)
        self._rms = torch.ones_like(flat)
        self._mom = torch.zeros_like(flat)
        self._scratch = torch.zeros(4096, dtype=torch.float32, device=dev)
        self.last_gnorm = self._scratch[2048:2049]
        self.cell.mark_weights_dirty()

    def _grad_view(self, name):
        for n, off, k in self._offsets:
            if n == name:
                return self._grad[off:off + k].view(self.cell.variables[n].shape)
        raise KeyError(name)

    def _buffer(self, key, shape):
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float32, device=self.cell.device)
            self._bufs[key] = t
        return t

    # ------------------------------------------------------------------ forward + backward --
    def loss_and_grads(self, inputs, targets, gather=None):
        """inputs [B,T,D] (CUDA), targets [B, len(gather), O].  Returns (loss, {variable name: grad}).
        loss = tf.nn.l2_loss(tanh(logits[:, gather]) - targets) = 0.5 * sum(diff^2) (a device scalar); the
        gradients are views into the trainer's flat gradient buffer (valid until the next call)."""
        cell, lib = self.cell, _cabi.load()
        dev = cell.device
        with torch.cuda.device(dev):
            x = cell._prepare_inputs(inputs, 3)
            B, T, D = x.shape
            if cell.input_dim is None:
                cell.build(D, self.tracker.initializer)
            cell.ensure_init_state(self.tracker.initializer)          # before the variables are re-homed
            self._ensure_flat()
            gather = list(gather) if gather is not None else delimiter_steps(T, self.frame)
            N, M, R, W = cell.mem_size, cell.mem_dim, cell.read_head_size, cell.write_head_size
            H, Cc, L, O = R + W, cell.controller_hidden_size, cell.controller_num_layers, cell.output_dim
            PO4 = (cell.param_size + O + 3) // 4 * 4
            hist = {
                "M_prev": self._buffer("M_prev", (T, B, N, M)), "w_prev": self._buffer("w_prev", (T, B, H, N)),
                "params": self._buffer("params", (T, B, PO4)), "z": self._buffer("z", (T, B, L, 4, Cc)),
                "c": self._buffer("c", (T + 1, B, L, Cc)), "h": self._buffer("h", (T + 1, B, L, Cc)),
                "read": self._buffer("read", (T + 1, B, R * M)),
                # by-products of the forward pass that spare the backward two sweeps over the memory
                "sim": self._buffer("sim", (T, B, H, N)), "cn": self._buffer("cn", (T, B, M)),
            }
            state = cell.zero_state(B, self.tracker.initializer)     # views of the (re-homed) variables
            logits, _, final_state, _ = cell._run(x, state, T, history=hist)
            self.tracker.final_state = final_state
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

            # ---- loss + dLoss/dlogits (direct_offset_output.py:581-606) ----
            tg = targets.to(dev, torch.float32).contiguous()
            if tuple(tg.shape) != (B, len(gather), O):
                raise ValueError("targets have shape %s, expected %s" % (tuple(tg.shape), (B, len(gather), O)))
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            dlogits = self._buffer("dlogits", (B, T, O))
            steps = (C.c_int32 * len(gather))(*[int(g_) for g_ in gather])
            _cabi.check(lib.ntm_b200_offset_loss(logits.data_ptr(), tg.data_ptr(), steps, len(gather), B, T, O,
                                                 loss.data_ptr(), dlogits.data_ptr(), stream), "offset_loss")

            # ---- backward: ONE library call (reverse-time loop + weight gradients) ----
            shp = cell._shape_struct(D)
            wts = cell._weights_struct()
            need = int(lib.ntm_b200_backward_workspace_bytes(C.byref(shp), B, T))
            ws = self._bufs.get("bwd_ws")
            if ws is None or ws.numel() < need:
                ws = self._bufs["bwd_ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
            hstruct = _cabi.History(*[hist[k].data_ptr() for k, _ in _cabi.History._fields_])
            inner = {"M": N * M, "w": H * N, "read": R * M, "controller_state": 2 * Cc * L}
            s0, keep = cell._state_struct(state, inner)
            gs = _cabi.Grads()
            for l in range(L):
                gs.lstm_w[l] = self._grad_view(cell._lstm(l, "weights")).data_ptr()
                gs.lstm_b[l] = self._grad_view(cell._lstm(l, "biases")).data_ptr()
            gs.addr_w = self._grad_view(cell._cell("addressing/weights")).data_ptr()
            gs.addr_b = self._grad_view(cell._cell("addressing/biases")).data_ptr()
            gs.out_w = self._grad_view(cell._cell("weights")).data_ptr()
            gs.out_b = self._grad_view(cell._cell("biases")).data_ptr()
            gs.init_M = self._grad_view(cell.scope + "/init_state/M").data_ptr()
            gs.init_w = self._grad_view(cell.scope + "/init_state/w").data_ptr()
            gs.init_read = self._grad_view(cell.scope + "/init_state/read").data_ptr()
            _cabi.check(lib.ntm_b200_backward_seq(
                C.byref(shp), C.byref(wts), cell._packed.data_ptr(), B, T, x.data_ptr(), C.byref(hstruct),
                dlogits.data_ptr(), C.byref(s0), C.byref(gs), ws.data_ptr(), ws.numel(), stream), "backward_seq")
            grads = {n: self._grad[off:off + k].view(cell.variables[n].shape) for n, off, k in self._offsets}
            return loss[0], grads

    # ------------------------------------------------------------------ optimizer --
    def apply_gradients(self, grads=None, sync=True):
        """All-reduce (sum) the flat gradient over the ranks, then clip by global norm + RMSProp in one fused
        pass (``ntm_b200_rmsprop_step``).  `grads` (optional): a {name: tensor} dict to load into the flat
        buffer first (the dict ``loss_and_grads`` returns already lives there).  Returns the unclipped global
        norm -- as a Python float when `sync`, else as a device scalar (no host synchronisation)."""
        self._ensure_flat()
        if grads is not None:
            for n, off, k in self._offsets:
                gv = grads[n]
                if gv.data_ptr() != self._grad.data_ptr() + 4 * off:
                    self._grad[off:off + k].copy_(gv.reshape(-1))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self._grad, op=dist.ReduceOp.SUM)          # the path's only collective (NCCL)
        self._fused_update()
        self.cell.mark_weights_dirty()
        self.global_step += 1
        return float(self.last_gnorm) if sync else self.last_gnorm

    def _fused_update(self):
        """clip_by_global_norm + RMSProp + momentum + parameter update over the flat buffers: one library call
        (three kernels, no host synchronisation).  There is no CPU version of this in the product; the
        world-size-2 gloo test substitutes a NumPy restatement for this one method to exercise the sharding /
        all-reduce logic around it on CPU."""
        lib = _cabi.load()
        dev = self.cell.device
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(lib.ntm_b200_rmsprop_step(
                self._flat.data_ptr(), self._grad.data_ptr(), self._rms.data_ptr(), self._mom.data_ptr(),
                self._flat.numel(), self.lr, self.decay, self.momentum, self.eps, self.clip,
                self.last_gnorm.data_ptr(), self._scratch.data_ptr(), stream), "rmsprop_step")

    def train_step(self, inputs, targets, gather=None, sync=True):
        loss, _ = self.loss_and_grads(inputs, targets, gather)
        gnorm = self.apply_gradients(sync=sync)
        if sync:
            self.cell.finish()                  # surface device-side failures of the forward
        return loss, gnorm
