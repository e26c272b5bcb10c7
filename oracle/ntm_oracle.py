"""CPU oracle for the NTM-cell hot path (TEST INFRASTRUCTURE -- not a product path).

NumPy restatement of the reference's arithmetic for one NTM cell step and the
T-step driver.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline legs may import this module; the product package
(``ntm_tracker_b200``) never does, and has no CPU fallback.

What it follows (all paths relative to /root/reference):
  * cell step ............. ntm_cell.py:53-253
  * initial state ......... ntm_cell.py:284-315
  * _linear ............... ntm_cell.py:317-370     (x @ W + b)
  * similarity ............ ops.py:135-158          (column-normalised, quirk 1)
  * circular convolution .. ops.py:180-242          (Py2 floor-div taps, quirk 2)
  * T-step driver ......... ntm_tracker_new.py:13-64
  * LSTM controller ....... TensorFlow 1.0/1.1 tf.contrib.rnn.BasicLSTMCell /
    MultiRNNCell (third-party, NOT under /root/reference; version unpinned by
    the reference -- API usage implies TF 1.0-1.1).  Published algorithm:
    z = [inp, h] @ W + b ; i, j, f, o = split4(z) ;
    c' = c*sigmoid(f + forget_bias) + sigmoid(i)*tanh(j) ; h' = tanh(c')*sigmoid(o) ;
    non-tuple state = concat([c, h]) per layer, layers concatenated.

Pinning status: the reference's only test on this path (ops_test.py) is stale
(it pins the superseded row-wise "smooth" cosine).  The oracle is therefore
pinned against *the reference's own source executed here* under a NumPy shim of
the TF1 API (oracle/tf1_shim.py + oracle/make_golden.py -> tests/golden/*.npz);
the TF kernels themselves are restated, not run ("parity unpinned at the TF
boundary" -- see DESIGN.md).

Parameter dictionary keys are the reference's TF variable names.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------- #
# Shapes
# --------------------------------------------------------------------------- #


@dataclasses.dataclass(frozen=True)
class NTMShape:
    """Constructor arguments of NTMCell (ntm_cell.py:18-20) + the input width."""

    output_dim: int
    input_dim: int
    mem_size: int = 128
    mem_dim: int = 20
    shift_range: int = 1
    controller_hidden_size: int = 100
    controller_num_layers: int = 10
    write_head_size: int = 3
    read_head_size: int = 3
    write_first: bool = False

    @property
    def num_heads(self) -> int:
        return self.read_head_size + self.write_head_size

    @property
    def shift_space(self) -> int:
        return 2 * self.shift_range + 1

    @property
    def param_size(self) -> int:
        """Width of the 'unpack_mem_params' projection (ntm_cell.py:113-126)."""
        H, M, W, S = self.num_heads, self.mem_dim, self.write_head_size, self.shift_space
        return H * M + H + H + S * H + H + 2 * M * W


SCOPE = "ntm-tracker"
CELL = SCOPE + "/ntm-cell"


def lstm_names(layer: int) -> Tuple[str, str]:
    base = "%s/lstm-controller/cell_%d/basic_lstm_cell" % (CELL, layer)
    return base + "/weights", base + "/biases"


def param_shapes(s: NTMShape) -> Dict[str, Tuple[int, ...]]:
    C = s.controller_hidden_size
    shapes: Dict[str, Tuple[int, ...]] = {
        SCOPE + "/init_state/M": (s.mem_size, s.mem_dim),
        SCOPE + "/init_state/w": (s.num_heads, s.mem_size),
        SCOPE + "/init_state/read": (s.read_head_size, s.mem_dim),
    }
    for l in range(s.controller_num_layers):
        in_l = (s.input_dim + s.read_head_size * s.mem_dim) if l == 0 else C
        wn, bn = lstm_names(l)
        shapes[wn] = (in_l + C, 4 * C)
        shapes[bn] = (4 * C,)
    shapes[CELL + "/addressing/weights"] = (C, s.param_size)
    shapes[CELL + "/addressing/biases"] = (s.param_size,)
    shapes[CELL + "/weights"] = (C, s.output_dim)
    shapes[CELL + "/biases"] = (s.output_dim,)
    return shapes


def init_params(s: NTMShape, seed: int, scale: float = 0.05,
                random_biases: bool = False) -> Dict[str, np.ndarray]:
    """Uniform(-scale, scale) matrices, zero biases (ntm_cell.py:369,
    direct_offset_output.py:42,528).  ``random_biases`` exercises the bias path
    (a trained checkpoint has non-zero biases)."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shp in param_shapes(s).items():
        if name.endswith("biases") and not random_biases:
            out[name] = np.zeros(shp, np.float32)
        else:
            out[name] = rng.uniform(-scale, scale, size=shp).astype(np.float32)
    return out


# --------------------------------------------------------------------------- #
# Elementwise helpers (TF semantics)
# --------------------------------------------------------------------------- #


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softplus(x):
    # log(1 + exp(x)), overflow-safe (tf.nn.softplus)
    return np.logaddexp(0.0, x)


def _softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def _l2_normalize(x, axis, eps=1e-12):
    # tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), eps))
    ss = np.sum(np.square(x), axis=axis, keepdims=True)
    return x / np.sqrt(np.maximum(ss, eps))


def shift_offsets(shift_space: int):
    """ops.py:204-209 under Python-2 floor division: start = -S/2 -> (-S)//2."""
    start = (-shift_space) // 2
    return list(range(start, shift_space + start))


# --------------------------------------------------------------------------- #
# ops.py restatements
# --------------------------------------------------------------------------- #


def batched_smooth_cosine_similarity(memory, keys):
    """ops.py:135-158.  memory [B,N,M], keys [B,H,M] -> [B,H,N].

    NOTE quirk 1: memory is transposed to [B,M,N] and l2-normalised along the
    LAST axis, i.e. each column d is normalised over the N rows."""
    mem_t = np.transpose(memory, (0, 2, 1))
    mem_t = _l2_normalize(mem_t, 2)
    keys = _l2_normalize(keys, 2)
    return np.matmul(keys, mem_t)


def rowwise_smooth_cosine_similarity(memory, keys):
    """The SUPERSEDED semantics pinned by ops_test.py:20-34 (and by the legacy
    ops.smooth_cosine_similarity, ops.py:161-178): dot / (|m|*|k| + 1e-3).
    Kept only so the reference's stale known-answer test can be replayed."""
    dots = np.einsum("bhd,bnd->bhn", keys, memory)
    mn = np.sqrt(np.sum(memory ** 2, axis=2))[:, None, :]
    kn = np.sqrt(np.sum(keys ** 2, axis=2))[:, :, None]
    return dots / (mn * kn + 1e-3)


def batched_circular_convolution(w, kernel):
    """ops.py:180-242.  w [B,H,N], kernel [B,H,S] -> [B,H,N].
    circular_shift(x, j)[n] = x[(n + j) mod N]; taps j from shift_offsets()."""
    S = kernel.shape[-1]
    out = np.zeros_like(w)
    for t, j in enumerate(shift_offsets(S)):
        out = out + np.roll(w, -j, axis=-1) * kernel[..., t:t + 1]
    return out


# --------------------------------------------------------------------------- #
# The cell
# --------------------------------------------------------------------------- #


def zero_state(params, s: NTMShape, batch: int, dtype=np.float64):
    """ntm_cell.py:284-315."""
    M = np.tanh(params[SCOPE + "/init_state/M"].astype(dtype))
    w = _sigmoid(params[SCOPE + "/init_state/w"].astype(dtype))
    r = np.tanh(params[SCOPE + "/init_state/read"].astype(dtype))
    C, L = s.controller_hidden_size, s.controller_num_layers
    return {
        "M": np.repeat(M[None], batch, 0),
        "w": np.repeat(w[None], batch, 0),
        "read": np.repeat(r[None], batch, 0),
        "controller_state": np.zeros((batch, 2 * C * L), dtype),
    }


def controller(params, s: NTMShape, inp, ctrl_state, dtype):
    """MultiRNNCell([BasicLSTMCell(C, forget_bias=0.0, state_is_tuple=False)]*L)."""
    C = s.controller_hidden_size
    new_states = []
    cur = inp
    for l in range(s.controller_num_layers):
        st = ctrl_state[:, 2 * C * l: 2 * C * (l + 1)]
        c, h = st[:, :C], st[:, C:]
        wn, bn = lstm_names(l)
        z = np.concatenate([cur, h], 1) @ params[wn].astype(dtype) + params[bn].astype(dtype)
        i, j, f, o = np.split(z, 4, axis=1)
        new_c = c * _sigmoid(f + 0.0) + _sigmoid(i) * np.tanh(j)
        new_h = np.tanh(new_c) * _sigmoid(o)
        new_states += [new_c, new_h]
        cur = new_h
    return cur, np.concatenate(new_states, 1)


def cell_step(params, s: NTMShape, x, state, dtype=np.float64, debug=False):
    """One NTMCell.__call__ (ntm_cell.py:53-253).

    Returns (output, logit, new_state, debug_dict_or_None)."""
    x = np.asarray(x, dtype)
    M_prev = np.asarray(state["M"], dtype)
    w_prev = np.asarray(state["w"], dtype)
    read_prev = np.asarray(state["read"], dtype)
    ctrl = np.asarray(state["controller_state"], dtype)
    B = x.shape[0]
    H, R, W = s.num_heads, s.read_head_size, s.write_head_size
    Md, S = s.mem_dim, s.shift_space

    # controller (ntm_cell.py:101-105)
    u = np.concatenate([x, read_prev.reshape(B, R * Md)], 1)
    hc, ctrl_new = controller(params, s, u, ctrl, dtype)

    # head parameters (ntm_cell.py:113-130)
    mc = hc @ params[CELL + "/addressing/weights"].astype(dtype) \
        + params[CELL + "/addressing/biases"].astype(dtype)
    sizes = [Md * H, H, H, S * H, H, Md * W, Md * W]
    offs = np.cumsum([0] + sizes)
    k, beta, g, sw, gamma, erase, add = [mc[:, offs[i]:offs[i + 1]] for i in range(7)]

    k = np.tanh(k.reshape(B, H, Md))                                  # :133
    similarity = batched_smooth_cosine_similarity(M_prev, k)          # :136
    beta = _softplus(beta)[..., None]                                 # :140
    w_c = _softmax(similarity * beta, axis=-1)                        # :142
    g = _sigmoid(g)[..., None]                                        # :151
    w_g = w_c * g + w_prev * (1.0 - g)                                # :153-156
    sw = _softmax(sw.reshape(B, H, S), axis=-1)                       # :161
    w_conv = batched_circular_convolution(w_g, sw)                    # :165
    gamma = (_softplus(gamma) + 1.0)[..., None]                       # :169-170
    powed = np.power(w_conv, gamma)                                   # :173
    w = powed / (np.sum(powed, axis=2, keepdims=True) + 1e-3)         # :175-176
    w_read, w_write = w[:, :R], w[:, R:]                              # :181-184

    erase = _sigmoid(erase.reshape(B, W, Md))                         # :193-194
    add = np.tanh(add.reshape(B, W, Md))                              # :195-196
    M_erase = np.prod(1.0 - w_write[:, :, :, None] * erase[:, :, None, :], axis=1)  # :202-204
    M_write = np.sum(w_write[:, :, :, None] * add[:, :, None, :], axis=1)           # :206-208
    M = M_prev * M_erase + M_write                                    # :210
    read = np.matmul(w_read, M if s.write_first else M_prev)          # :212-215

    logit = hc @ params[CELL + "/weights"].astype(dtype) + params[CELL + "/biases"].astype(dtype)  # :220
    out = _softmax(logit, axis=-1)                                    # :221

    new_state = {"M": M, "w": w, "read": read, "controller_state": ctrl_new}
    dbg = None
    if debug:  # ntm_cell.py:230-250 (key 'bega' [sic] is beta)
        dbg = {"k": k, "gamma": gamma, "add": add, "erase": erase, "bega": beta, "g": g,
               "sw": sw, "similarity": similarity, "w_content_focused": w_c,
               "w_gated": w_g, "w_conv": w_conv, "w_conv_powed": powed, "w": w,
               "w_read": w_read, "w_write": w_write, "M": M, "M_prev": M_prev,
               "M_write": M_write, "M_erase": M_erase}
    return out, logit, new_state, dbg


def run_sequence(params, s: NTMShape, inputs, state=None, dtype=np.float64,
                 history=False):
    """LoopNTMTracker.__call__ (ntm_tracker_new.py:13-64): inputs [B,T,D]
    batch-major -> (outputs [B,T,O], logits [B,T,O], final_state[, history])."""
    inputs = np.asarray(inputs, dtype)
    B, T, _ = inputs.shape
    if state is None:
        state = zero_state(params, s, B, dtype)
    outs, logits = [], []
    hist = {"M": [], "w": [], "read": []}
    for t in range(T):
        o, lg, state, _ = cell_step(params, s, inputs[:, t], state, dtype)
        outs.append(o)
        logits.append(lg)
        if history:
            for kk in hist:
                hist[kk].append(state[kk])
    res = (np.stack(outs, 1), np.stack(logits, 1), state)
    if history:
        res = res + ({kk: np.stack(v, 1) for kk, v in hist.items()},)
    return res


# --------------------------------------------------------------------------- #
# Synthetic workloads
# --------------------------------------------------------------------------- #

CONFIGS = {
    # name: (shape kwargs, B, T)
    "c1_copy": (dict(output_dim=4, input_dim=4, mem_size=128, mem_dim=20, shift_range=1,
                     controller_hidden_size=100, controller_num_layers=1,
                     write_head_size=1, read_head_size=1), 16, 20),
    "c2_tracker": (dict(output_dim=2, input_dim=514, mem_size=128, mem_dim=512, shift_range=1,
                        controller_hidden_size=200, controller_num_layers=1,
                        write_head_size=1, read_head_size=4), 64, 32),
    "c3_sweep": (dict(output_dim=2, input_dim=514, mem_size=128, mem_dim=512, shift_range=1,
                      controller_hidden_size=200, controller_num_layers=1,
                      write_head_size=1, read_head_size=4), 4096, 64),
    "c5_train": (dict(output_dim=2, input_dim=514, mem_size=128, mem_dim=512, shift_range=1,
                      controller_hidden_size=200, controller_num_layers=1,
                      write_head_size=1, read_head_size=4), 256, 32),
    "c4_large": (dict(output_dim=2, input_dim=514, mem_size=1024, mem_dim=256, shift_range=1,
                      controller_hidden_size=200, controller_num_layers=1,
                      write_head_size=1, read_head_size=4), 512, 128),
}


def copy_task_inputs(B: int, T: int, width: int, seed: int) -> np.ndarray:
    """main.py:1546-1559: bits [B,width,length] + indicator channel; inputs =
    concat(bits, delimiter, zeros) -> transposed to [B, 2*length+1, width+1].
    T is forced to the caller's value by truncation/zero-padding (BASELINE C1
    quotes T=20)."""
    rng = np.random.RandomState(seed)
    length = max(1, (T - 1) // 2)
    bits = (rng.rand(B, length, width) < 0.5).astype(np.float32)
    x = np.zeros((B, T, width + 1), np.float32)
    x[:, :length, :width] = bits
    if length < T:
        x[:, length, width] = 1.0
    return x


def tracker_inputs(B: int, T: int, seed: int, scale: float = 1.0,
                   feat: int = 512, frame: int = 65) -> np.ndarray:
    """direct_offset_output.py:463-500: channels [0,feat) = post-ReLU conv4_3
    features max(0, N(0,1))*scale; channel feat = frame delimiter (1 on the last
    row of every `frame`-row block, features zero on that row); channel feat+1 =
    target indicator, non-zero only within the first frame's feature rows."""
    rng = np.random.RandomState(seed)
    x = np.zeros((B, T, feat + 2), np.float32)
    x[:, :, :feat] = np.maximum(rng.standard_normal((B, T, feat)), 0.0).astype(np.float32) * scale
    t = np.arange(T)
    delim = (t % frame) == (frame - 1)
    x[:, delim, :feat] = 0.0
    x[:, delim, feat] = 1.0
    first = t < min(frame - 1, T)
    x[:, first, feat + 1] = (rng.rand(B, int(first.sum())) < 0.1).astype(np.float32)
    return x


# --------------------------------------------------------------------------- #
# Data formats either side of the path
# --------------------------------------------------------------------------- #


def serialize_tracker_inputs(features, target, delimiter_first=False):
    """direct_offset_output.py:439-500 (training layout, delimiter row last in every frame) /
    test_tracker.py:385-404 (serve layout, delimiter row first).
    features [B,L,F,C], target [B,F] -> [B, L*(F+1), C+2].  The target channel sits on the F feature rows
    of the FIRST frame (training: steps 0..F-1, :490-494; serve: rows [feat_f, 0, gt_f] built before the
    delimiter [0..0,1,0] is prepended, so steps 1..F) and is zero everywhere else."""
    features = np.asarray(features)
    B, L, F, Cc = features.shape
    tgt = np.zeros((B, L, F, 1), features.dtype)
    tgt[:, 0, :, 0] = np.asarray(target, features.dtype)                                  # gts[:, 0, :], :456
    padded = np.concatenate([features, np.zeros((B, L, F, 1), features.dtype), tgt], 3)   # :463-464, :496-498
    delim = np.zeros((B, L, 1, Cc + 2), features.dtype)
    delim[..., Cc] = 1.0                                                                  # :467-475
    frames = np.concatenate([delim, padded] if delimiter_first else [padded, delim], 2)  # :478-479 / test_tracker.py:404
    return frames.reshape(B, L * (F + 1), Cc + 2)                                         # :481-486


def gather_offsets(output_logits, num_features):
    """direct_offset_output.py:581-593: drop the first frame, take each frame's delimiter row, tanh."""
    lg = np.asarray(output_logits)
    B, T, Od = lg.shape
    L = T // (num_features + 1)
    g = lg[:, num_features + 1:, :].reshape(B, L - 1, num_features + 1, Od)[:, :, num_features, :]
    return np.tanh(g)
