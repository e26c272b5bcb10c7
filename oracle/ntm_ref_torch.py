"""fp32 PyTorch-CPU, op-for-op restatement of the reference TF graph (TEST /
BASELINE INFRASTRUCTURE -- never imported by the product package).

Same op granularity as the TensorFlow graph the reference builds, so that timing
it on the host cores is a fair stand-in for "the reference TF CPU path" (which
cannot be installed here: no TensorFlow, no Python 2):
  * similarity: materialised transpose + two l2_normalize chains + batched matmul
    (ops.py:147-156);
  * circular convolution: S x (2 slices + concat), stack to [B,H,N,S], batched
    matmul with the [B,H,S,1] kernel, squeeze (ops.py:201-214, 216-242);
  * erase/add: two outer-product batched matmuls materialising [B,W,N,M], prod /
    sum over W (ntm_cell.py:202-208);
  * controller: concat + matmul + bias + split + gates per layer (TF 1.0/1.1
    BasicLSTMCell / MultiRNNCell);
  * driver: per-step slices of a time-major copy of the inputs, per-step writes of
    outputs, logits, M, w and read histories (ntm_tracker_new.py:17-40, 53-61).
Used by bench.py (`cpu_baseline`, `--impl reference`) and by tests as a second
CPU implementation checked against the NumPy oracle.
"""
import torch
import torch.nn.functional as F

from . import ntm_oracle as O


def _l2_normalize(x, dim, eps=1e-12):
    ss = torch.sum(torch.square(x), dim=dim, keepdim=True)
    return x * torch.rsqrt(torch.clamp_min(ss, eps))


def batched_smooth_cosine_similarity(memory, keys):
    memory = memory.permute(0, 2, 1).contiguous()      # tf.transpose materialises
    memory = _l2_normalize(memory, 2)
    keys = _l2_normalize(keys, 2)
    return torch.matmul(keys, memory)


def circular_shift(t, shift):
    n = t.shape[-1]
    sp = n + shift if shift < 0 else shift
    return torch.cat([t[..., sp:], t[..., :sp]], dim=-1)


def batched_circular_convolution(t, kernel):
    S = kernel.shape[-1]
    aug = torch.stack([circular_shift(t, j) for j in O.shift_offsets(S)], dim=-1)
    return torch.matmul(aug, kernel.unsqueeze(-1)).squeeze(-1)


class TorchRefNTM(object):
    def __init__(self, shape: O.NTMShape, params, dtype=torch.float32, requires_grad=False):
        self.s = shape
        self.dtype = dtype
        self.p = {k: torch.as_tensor(v).to(dtype).clone().requires_grad_(requires_grad) for k, v in params.items()}

    def zero_state(self, B):
        s, p = self.s, self.p
        M = torch.tanh(p[O.SCOPE + "/init_state/M"])
        w = torch.sigmoid(p[O.SCOPE + "/init_state/w"])
        r = torch.tanh(p[O.SCOPE + "/init_state/read"])
        return {"M": torch.stack([M] * B, 0), "w": torch.stack([w] * B, 0),
                "read": torch.stack([r] * B, 0),
                "controller_state": torch.zeros(B, 2 * s.controller_hidden_size * s.controller_num_layers,
                                                dtype=self.dtype)}

    def controller(self, inp, state):
        s, C = self.s, self.s.controller_hidden_size
        cur, new_states = inp, []
        for l in range(s.controller_num_layers):
            st = state[:, 2 * C * l: 2 * C * (l + 1)]
            c, h = st[:, :C], st[:, C:]
            wn, bn = O.lstm_names(l)
            z = torch.matmul(torch.cat([cur, h], 1), self.p[wn]) + self.p[bn]
            i, j, f, o = torch.split(z, C, dim=1)
            new_c = c * torch.sigmoid(f + 0.0) + torch.sigmoid(i) * torch.tanh(j)
            new_h = torch.tanh(new_c) * torch.sigmoid(o)
            new_states += [new_c, new_h]
            cur = new_h
        return cur, torch.cat(new_states, 1)

    def step(self, x, M_prev, w_prev, read_prev, ctrl):
        s, p = self.s, self.p
        B = x.shape[0]
        H, R, W, Md, S = s.num_heads, s.read_head_size, s.write_head_size, s.mem_dim, s.shift_space
        hc, ctrl = self.controller(torch.cat([x, read_prev.reshape(B, R * Md)], 1), ctrl)
        mc = torch.matmul(hc, p[O.CELL + "/addressing/weights"]) + p[O.CELL + "/addressing/biases"]
        k, beta, g, sw, gamma, erase, add = torch.split(
            mc, [Md * H, H, H, S * H, H, Md * W, Md * W], dim=1)
        k = torch.tanh(k.reshape(B, H, Md))
        sim = batched_smooth_cosine_similarity(M_prev, k)
        beta = F.softplus(beta).unsqueeze(-1)
        w_c = torch.softmax(sim * beta, dim=-1)
        g = torch.sigmoid(g).unsqueeze(-1)
        w_g = w_c * g + w_prev * (1.0 - g)
        sw = torch.softmax(sw.reshape(B, H, S), dim=-1)
        w_conv = batched_circular_convolution(w_g, sw)
        gamma = (F.softplus(gamma) + 1.0).unsqueeze(-1)
        powed = torch.pow(w_conv, gamma)
        w = powed / (torch.sum(powed, dim=2, keepdim=True) + 1e-3)
        w_read, w_write = w[:, :R], w[:, R:]
        erase = torch.sigmoid(erase.reshape(B, W, Md))
        add = torch.tanh(add.reshape(B, W, Md))
        M_erase = torch.prod(1.0 - torch.matmul(w_write.unsqueeze(3), erase.unsqueeze(2)), dim=1)
        M_write = torch.sum(torch.matmul(w_write.unsqueeze(3), add.unsqueeze(2)), dim=1)
        M = M_prev * M_erase + M_write
        read = torch.matmul(w_read, M if s.write_first else M_prev)
        logit = torch.matmul(hc, p[O.CELL + "/weights"]) + p[O.CELL + "/biases"]
        out = torch.softmax(logit, dim=-1)
        return out, logit, M, w, read, ctrl

    def run(self, inputs, state=None):
        """LoopNTMTracker: [B,T,D] -> (outputs, logits, final_state).  Builds an autograd graph when
        the parameters require gradients (used as the gradient oracle of the training path)."""
        if not any(v.requires_grad for v in self.p.values()):
            with torch.no_grad():
                return self._run(inputs, state)
        return self._run(inputs, state)

    def _run(self, inputs, state=None):
        inputs = inputs.to(self.dtype)
        B, T, _ = inputs.shape
        state = state or self.zero_state(B)
        xs = inputs.permute(1, 0, 2).contiguous()        # unstack_into_tensorarray (utility.py:61-91)
        M, w, read, ctrl = state["M"], state["w"], state["read"], state["controller_state"]
        outs, logits, Ms, ws, reads = [], [], [], [], []
        for t in range(T):
            o, lg, M, w, read, ctrl = self.step(xs[t], M, w, read, ctrl)
            outs.append(o); logits.append(lg)
            Ms.append(M); ws.append(w); reads.append(read)     # TensorArray writes (:57-61)
        outputs = torch.stack(outs, 0).permute(1, 0, 2).contiguous()
        out_logits = torch.stack(logits, 0).permute(1, 0, 2).contiguous()
        return outputs, out_logits, {"M": M, "w": w, "read": read, "controller_state": ctrl}
