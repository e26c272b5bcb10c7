"""Training step for the NTM tracker path (BASELINE config 5).

What the reference does (direct_offset_output.py:581-626): gather the logits at every frame's
delimiter step (first frame dropped), ``tanh``, ``tf.nn.l2_loss`` against the offsets,
``tf.gradients`` through the unrolled while_loop, ``tf.clip_by_global_norm(5)``,
``RMSPropOptimizer(1e-4, decay=0.95, momentum=0.9)``.  It is single-device.

Here:
  * forward  -- the persistent CUDA kernel, recording the history the backward needs
                (``ntm_b200_forward_seq_train``);
  * backward -- reverse-time loop; the memory / addressing part of every step is ONE hand-written
                kernel (``ntm_b200_memory_backward_step``, csrc/ntm_b200_train.cu); the dense
                projections' data- and weight-gradients are plain GEMMs (``torch.matmul`` = cuBLAS);
  * the LSTM gate algebra's backward is one elementwise kernel per layer and step
                (``ntm_b200_lstm_backward_step``);
  * multi-GPU -- sequences are sharded over ranks (no collective in forward or backward); the ONE
                collective of the path is an all-reduce (sum) of the flat gradient over NCCL, followed
                by the same clip + RMSProp on every rank, so the replicas stay identical.
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


class NTMTrainer(object):
    def __init__(self, tracker: LoopNTMTracker, learning_rate=1e-4, decay=0.95, momentum=0.9,
                 max_gradient_norm=5.0, epsilon=1e-10, frame=65):
        self.tracker = tracker
        self.cell = tracker.cell
        self.lr, self.decay, self.momentum = learning_rate, decay, momentum
        self.clip, self.eps, self.frame = max_gradient_norm, epsilon, frame
        self._rms = None      # TF's RMSProp slots: 'rms' starts at ones, 'momentum' at zeros
        self._mom = None
        self.global_step = 0

    # ------------------------------------------------------------------ forward + backward --
    def loss_and_grads(self, inputs, targets, gather=None):
        """inputs [B,T,D] (CUDA), targets [B, len(gather), O].  Returns (loss, {variable name: grad}).
        loss = tf.nn.l2_loss(tanh(logits[:, gather]) - targets) = 0.5 * sum(diff^2)."""
        cell, lib = self.cell, _cabi.load()
        dev = cell.device
        x = cell._prepare_inputs(inputs, 3)
        B, T, D = x.shape
        if cell.input_dim is None:
            cell.build(D, self.tracker.initializer)
        gather = list(gather) if gather is not None else delimiter_steps(T, self.frame)
        N, M, R, W = cell.mem_size, cell.mem_dim, cell.read_head_size, cell.write_head_size
        H, Cc, L, O = R + W, cell.controller_hidden_size, cell.controller_num_layers, cell.output_dim
        P = cell.param_size
        PO4 = (P + O + 3) // 4 * 4
        f32 = dict(dtype=torch.float32, device=dev)
        hist = {
            "M_prev": torch.empty(T, B, N, M, **f32), "w_prev": torch.empty(T, B, H, N, **f32),
            "params": torch.empty(T, B, PO4, **f32), "z": torch.empty(T, B, L, 4, Cc, **f32),
            "c": torch.empty(T + 1, B, L, Cc, **f32), "h": torch.empty(T + 1, B, L, Cc, **f32),
            "read": torch.empty(T + 1, B, R * M, **f32),
        }
        state = cell.zero_state(B, self.tracker.initializer)
        logits, _, final_state, _ = cell._run(x, state, T, history=hist)
        self.tracker.final_state = final_state

        # ---- loss (direct_offset_output.py:581-606) ----
        gi = torch.as_tensor(gather, device=dev, dtype=torch.long)
        y = torch.tanh(logits.index_select(1, gi))
        diff = y - targets.to(dev, torch.float32)
        loss = 0.5 * torch.sum(diff * diff)
        dlogits = torch.zeros_like(logits)
        dlogits.index_copy_(1, gi, diff * (1.0 - y * y))

        # ---- reverse-time loop ----
        V = cell.variables
        Wl = [V[cell._lstm(l, "weights")] for l in range(L)]
        Wao = cell._packed.view(torch.float32)[: Cc * PO4].view(Cc, PO4)       # [C, P+O (padded)]
        shp = cell._shape_struct(D)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        dM = torch.zeros(B, N, M, **f32)
        dw = torch.zeros(B, H, N, **f32)
        dw_prev = torch.empty(B, H, N, **f32)
        dread = torch.zeros(B, R, M, **f32)
        dh = [torch.zeros(B, Cc, **f32) for _ in range(L)]
        dc = [torch.zeros(B, Cc, **f32) for _ in range(L)]
        DMC = torch.empty(T, B, PO4, **f32)
        DZ = torch.empty(T, B, L, 4 * Cc, **f32)
        for t in range(T - 1, -1, -1):
            draw = DMC[t]
            _cabi.check(lib.ntm_b200_memory_backward_step(
                C.byref(shp), B, hist["M_prev"][t].data_ptr(), hist["w_prev"][t].data_ptr(),
                hist["params"][t].data_ptr(), dread.data_ptr(), dw.data_ptr(), dM.data_ptr(),
                dw_prev.data_ptr(), draw.data_ptr(), stream), "memory_backward_step")
            dw, dw_prev = dw_prev, dw
            draw[:, P:P + O] = dlogits[:, t]
            d_in = torch.matmul(draw, Wao.t()).contiguous()                     # dL/dh_top via both projections
            for l in range(L - 1, -1, -1):
                # elementwise LSTM backward (one kernel): dz -> DZ[t, :, l], dc[l] updated in place
                _cabi.check(lib.ntm_b200_lstm_backward_step(
                    B, Cc, dh[l].data_ptr(), d_in.data_ptr(), hist["z"][t, :, l].data_ptr(), L * 4 * Cc,
                    hist["c"][t, :, l].data_ptr(), hist["c"][t + 1, :, l].data_ptr(), L * Cc,
                    dc[l].data_ptr(), DZ[t, :, l].data_ptr(), L * 4 * Cc, stream), "lstm_backward_step")
                d_cat = torch.matmul(DZ[t, :, l], Wl[l].t())                    # [B, in_l + C]
                dh[l] = d_cat[:, -Cc:].contiguous()
                d_in = d_cat[:, :-Cc].contiguous() if l > 0 else d_cat[:, :-Cc]
            dread = d_in[:, D:].reshape(B, R, M).contiguous()

        # ---- weight gradients: one GEMM per variable over all (t, b) ----
        grads = {}
        hc = hist["h"][1:, :, L - 1].reshape(T * B, Cc)
        gao = torch.matmul(hc.t(), DMC.reshape(T * B, PO4))
        bao = DMC.reshape(T * B, PO4).sum(0)
        grads[cell._cell("addressing/weights")] = gao[:, :P].contiguous()
        grads[cell._cell("addressing/biases")] = bao[:P].contiguous()
        grads[cell._cell("weights")] = gao[:, P:P + O].contiguous()
        grads[cell._cell("biases")] = bao[P:P + O].contiguous()
        for l in range(L):
            if l == 0:
                inp = torch.cat([x.transpose(0, 1), hist["read"][:T], hist["h"][:T, :, 0]], dim=2)
            else:
                inp = torch.cat([hist["h"][1:, :, l - 1], hist["h"][:T, :, l]], dim=2)
            dzl = DZ[:, :, l].reshape(T * B, 4 * Cc)
            grads[cell._lstm(l, "weights")] = torch.matmul(inp.reshape(T * B, -1).t(), dzl)
            grads[cell._lstm(l, "biases")] = dzl.sum(0)
        # initial-state variables (tiled over the batch: gradients are batch sums, ntm_cell.py:292-306)
        M0 = torch.tanh(V[cell.scope + "/init_state/M"])
        w0 = torch.sigmoid(V[cell.scope + "/init_state/w"])
        r0 = torch.tanh(V[cell.scope + "/init_state/read"])
        grads[cell.scope + "/init_state/M"] = dM.sum(0) * (1.0 - M0 * M0)
        grads[cell.scope + "/init_state/w"] = dw.sum(0) * w0 * (1.0 - w0)
        grads[cell.scope + "/init_state/read"] = dread.sum(0) * (1.0 - r0 * r0)
        return loss, grads

    # ------------------------------------------------------------------ optimizer --
    def apply_gradients(self, grads):
        """All-reduce (sum) the flat gradient over the ranks, clip by global norm, RMSProp."""
        V = self.cell.variables
        names = sorted(V)
        flat = torch.cat([grads[n].reshape(-1) for n in names])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)          # the path's only collective (NCCL)
        gnorm = torch.linalg.vector_norm(flat)
        flat = flat * (self.clip / torch.clamp(gnorm, min=self.clip))   # tf.clip_by_global_norm
        if self._rms is None:
            self._rms = torch.ones_like(flat)
            self._mom = torch.zeros_like(flat)
        self._rms.mul_(self.decay).addcmul_(flat, flat, value=1.0 - self.decay)
        self._mom.mul_(self.momentum).add_(self.lr * flat / torch.sqrt(self._rms + self.eps))
        off = 0
        for n in names:
            k = V[n].numel()
            V[n].sub_(self._mom[off:off + k].view_as(V[n]))
            off += k
        self.cell.mark_weights_dirty()
        self.global_step += 1
        return float(gnorm)

    def train_step(self, inputs, targets, gather=None):
        loss, grads = self.loss_and_grads(inputs, targets, gather)
        gnorm = self.apply_gradients(grads)     # synchronises (returns the host value of the global norm)
        self.cell.finish()                      # ... so device-side failures of the forward surface here
        return loss, gnorm
