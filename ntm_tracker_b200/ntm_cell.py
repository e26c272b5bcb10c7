"""NTMCell -- drop-in for the reference's ``ntm_cell.NTMCell`` (ntm_cell.py:17-315).

Same constructor arguments, same state dict {'M','w','read','controller_state'},
same ``__call__(inputs, prev_state, M_prev, w_prev, read_prev, controller_state,
scope)`` returning the 8-tuple of ntm_cell.py:252-253, same ``zero_state`` /
``state_placeholder``.  The arithmetic runs in libntm_b200.so (hand-written
sm_100a CUDA behind the C ABI of include/ntm_b200.h); PyTorch is used only for
device memory and streams.  There is no CPU fallback: without the library or
without a B200 every compute call raises.

Deviations from the reference, all declared:
  * tensors are torch CUDA tensors, not TF graph nodes; variables live in
    ``cell.variables`` keyed by the reference's TF variable names;
  * the 19-entry ``debug`` dict (ntm_cell.py:230-250) is filled only when
    ``cell.debug = True`` (otherwise ``{}``);
  * ``scope`` is accepted and ignored (there is no graph to name).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _cabi

SCOPE = "ntm-tracker"


def _as_initializer(initializer):
    """None -> TF's default glorot-uniform; (lo, hi) -> uniform; callable(shape) -> values."""
    if initializer is None:
        def glorot(shape):
            fan = sum(shape) if len(shape) >= 2 else 2 * shape[0]
            lim = math.sqrt(6.0 / fan)
            return torch.empty(shape).uniform_(-lim, lim)
        return glorot
    if isinstance(initializer, (tuple, list)):
        lo, hi = initializer
        return lambda shape: torch.empty(shape).uniform_(lo, hi)
    return initializer


def random_uniform_initializer(minval=-0.1, maxval=0.1):
    """Stand-in for tf.random_uniform_initializer (ntm_tracker_new.py:6)."""
    return (minval, maxval)


class NTMCell(object):
    def __init__(self, output_dim, mem_size=128, mem_dim=20, shift_range=1,
                 controller_hidden_size=100, controller_num_layers=10,
                 write_head_size=3, read_head_size=3, write_first=False,
                 device=None, scope=SCOPE):
        self.mem_size = mem_size
        self.mem_dim = mem_dim
        self.controller_hidden_size = controller_hidden_size
        self.controller_num_layers = controller_num_layers
        self.write_head_size = write_head_size
        self.read_head_size = read_head_size
        self.shift_range = shift_range
        self.output_dim = output_dim
        self.write_first = write_first

        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        self.scope = scope
        self.variables = {}
        self.input_dim = None
        self.debug = False
        self._dirty = True
        self._packed = None
        self._ws = {}
        self._zero_ctrl = {}
        # validate the constructor arguments now, like _linear / circular_shift would
        # at graph-construction time (ValueError / AssertionError in the reference)
        self._check_shape(input_dim=1)

    # ------------------------------------------------------------------ names --
    @property
    def num_heads(self):
        return self.read_head_size + self.write_head_size

    @property
    def param_size(self):
        H, M, W, S = self.num_heads, self.mem_dim, self.write_head_size, 2 * self.shift_range + 1
        return H * M + 3 * H + S * H + 2 * M * W

    def _cell(self, suffix):
        return "%s/ntm-cell/%s" % (self.scope, suffix)

    def _lstm(self, l, what):
        return self._cell("lstm-controller/cell_%d/basic_lstm_cell/%s" % (l, what))

    def variable_shapes(self, input_dim):
        """TF variable name -> shape, in creation order of the reference graph."""
        C_, R, M = self.controller_hidden_size, self.read_head_size, self.mem_dim
        shapes = {
            self.scope + "/init_state/M": (self.mem_size, M),
            self.scope + "/init_state/w": (self.num_heads, self.mem_size),
            self.scope + "/init_state/read": (R, M),
        }
        for l in range(self.controller_num_layers):
            in_l = input_dim + R * M if l == 0 else C_
            shapes[self._lstm(l, "weights")] = (in_l + C_, 4 * C_)
            shapes[self._lstm(l, "biases")] = (4 * C_,)
        shapes[self._cell("addressing/weights")] = (C_, self.param_size)
        shapes[self._cell("addressing/biases")] = (self.param_size,)
        shapes[self._cell("weights")] = (C_, self.output_dim)
        shapes[self._cell("biases")] = (self.output_dim,)
        return shapes

    # -------------------------------------------------------------- variables --
    def _create(self, names, input_dim, initializer):
        init = _as_initializer(initializer)
        shapes = self.variable_shapes(input_dim)
        for name in names:
            if name in self.variables:
                continue
            shp = shapes[name]
            if name.endswith("biases"):
                v = torch.zeros(shp)                       # ntm_cell.py:369, bias_start = 0
            else:
                v = torch.as_tensor(np.asarray(init(shp)), dtype=torch.float32).reshape(shp)
            self.variables[name] = v.to(self.device, torch.float32).contiguous()
            self._dirty = True

    def build(self, input_dim, initializer=None):
        """Create every variable the reference graph would create for this input width."""
        if self.input_dim is not None and self.input_dim != input_dim:
            raise ValueError("cell was built for input_dim=%d, got %d" % (self.input_dim, input_dim))
        self._check_shape(input_dim)
        self.input_dim = int(input_dim)
        self._create(list(self.variable_shapes(input_dim)), input_dim, initializer)
        return self

    def load_reference_weights(self, mapping):
        """Load variables keyed by the reference's TF variable names (a
        tf.train.Saver checkpoint read into a dict; ':0' suffixes accepted)."""
        clean = {k[:-2] if k.endswith(":0") else k: v for k, v in mapping.items()}
        w0 = clean[self._lstm(0, "weights")]
        input_dim = int(w0.shape[0]) - self.read_head_size * self.mem_dim - self.controller_hidden_size
        if input_dim < 1:
            raise ValueError("layer-0 LSTM weights have %d rows, too few for this cell" % w0.shape[0])
        self._check_shape(input_dim)
        shapes = self.variable_shapes(input_dim)
        missing = sorted(set(shapes) - set(clean))
        if missing:
            raise ValueError("missing variables: %s" % missing)
        for name, shp in shapes.items():
            v = torch.as_tensor(np.asarray(clean[name]), dtype=torch.float32)
            if tuple(v.shape) != tuple(shp):
                raise ValueError("variable %s: expected shape %s, got %s" % (name, shp, tuple(v.shape)))
            self.variables[name] = v.to(self.device).contiguous()
        self.input_dim = input_dim
        self._dirty = True
        return self

    def mark_weights_dirty(self):
        """Call after modifying ``cell.variables`` in place (forces a re-pack)."""
        self._dirty = True

    # ------------------------------------------------------------------ state --
    def zero_state(self, batch_size, initializer=None):
        """ntm_cell.py:284-315.  M = tanh(var), w = sigmoid(var) (NOT normalised),
        read = tanh(var), each shared by every sequence of the batch (stride-0
        views -- what tf.stack([M]*batch_size) expresses); controller_state zeros."""
        names = self.ensure_init_state(initializer)
        vM, vw, vr = (self.variables[n] for n in names)
        if self.device.type == "cuda":      # one launch of the library's own kernel (ntm_b200_zero_state)
            M, w, read = torch.empty_like(vM), torch.empty_like(vw), torch.empty_like(vr)
            with torch.cuda.device(self.device):
                stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                _cabi.check(_cabi.load().ntm_b200_zero_state(
                    vM.data_ptr(), vM.numel(), vw.data_ptr(), vw.numel(), vr.data_ptr(), vr.numel(),
                    M.data_ptr(), w.data_ptr(), read.data_ptr(), stream), "zero_state")
        else:                               # variables kept on the host (host-side logic tests; nothing can run there)
            M, w, read = torch.tanh(vM), torch.sigmoid(vw), torch.tanh(vr)
        B = int(batch_size)
        CL2 = 2 * self.controller_hidden_size * self.controller_num_layers
        zc = self._zero_ctrl.get(B)
        if zc is None:                      # read-only input of every call: one zero tensor per batch size
            zc = self._zero_ctrl[B] = torch.zeros(B, CL2, device=self.device)
            if len(self._zero_ctrl) > 8:
                self._zero_ctrl = {B: zc}
        return {
            "M": M.unsqueeze(0).expand(B, -1, -1),
            "w": w.unsqueeze(0).expand(B, -1, -1),
            "read": read.unsqueeze(0).expand(B, -1, -1),
            "controller_state": zc,
        }

    def ensure_init_state(self, initializer=None):
        """Create the three init_state variables if they do not exist yet; returns their names (M, w, read)."""
        names = [self.scope + "/init_state/" + n for n in ("M", "w", "read")]
        self._create(names, self.input_dim or 1, initializer)
        return names

    def state_placeholder(self, batch_size):
        """ntm_cell.py:255-282: feedable state buffers (here: preallocated tensors)."""
        B = int(batch_size)
        z = lambda *s: torch.zeros(*s, device=self.device)
        return {
            "M": z(B, self.mem_size, self.mem_dim),
            "w": z(B, self.num_heads, self.mem_size),
            "read": z(B, self.read_head_size, self.mem_dim),
            "controller_state": z(B, 2 * self.controller_hidden_size * self.controller_num_layers),
        }

    # -------------------------------------------------------------- C structs --
    def _shape_struct(self, input_dim):
        return _cabi.Shape(int(input_dim), int(self.output_dim), int(self.mem_size), int(self.mem_dim),
                           int(self.shift_range), int(self.controller_hidden_size),
                           int(self.controller_num_layers), int(self.write_head_size),
                           int(self.read_head_size), int(bool(self.write_first)))

    def _check_shape(self, input_dim):
        lib = _cabi.load()
        plan = _cabi.Plan()
        shp = self._shape_struct(input_dim)
        _cabi.check(lib.ntm_b200_query(C.byref(shp), 1, 1, C.byref(plan)), "NTMCell")

    def plan(self, batch_size, steps=1):
        """Launch geometry the library will use (cluster size, smem, workspace)."""
        lib = _cabi.load()
        plan = _cabi.Plan()
        shp = self._shape_struct(self.input_dim or 1)
        _cabi.check(lib.ntm_b200_query(C.byref(shp), int(batch_size), int(steps), C.byref(plan)), "query")
        return {n: getattr(plan, n) for n, _ in _cabi.Plan._fields_}

    def _weights_struct(self):
        w = _cabi.Weights()
        for l in range(self.controller_num_layers):
            w.lstm_w[l] = self.variables[self._lstm(l, "weights")].data_ptr()
            w.lstm_b[l] = self.variables[self._lstm(l, "biases")].data_ptr()
        w.addr_w = self.variables[self._cell("addressing/weights")].data_ptr()
        w.addr_b = self.variables[self._cell("addressing/biases")].data_ptr()
        w.out_w = self.variables[self._cell("weights")].data_ptr()
        w.out_b = self.variables[self._cell("biases")].data_ptr()
        return w

    @staticmethod
    def _state_struct(st, inner):
        """ctypes State + the tensors kept alive for the call."""
        keep, ptrs, strides = [], [], []
        for key in ("M", "w", "read", "controller_state"):
            t = st[key]
            if t.dtype != torch.float32:
                t = t.float()
            B = t.shape[0]
            ok = t[0].is_contiguous() if B > 0 else True
            if not ok or (B > 1 and t.stride(0) not in (0, inner[key])):
                t = t.contiguous()
            keep.append(t)
            ptrs.append(t.data_ptr())
            strides.append(int(t.stride(0)) if B > 1 else inner[key])
        return _cabi.State(*ptrs, *strides), keep

    def _state_geometry(self, B):
        H, R, N, M = self.num_heads, self.read_head_size, self.mem_size, self.mem_dim
        CL2 = 2 * self.controller_hidden_size * self.controller_num_layers
        inner = {"M": N * M, "w": H * N, "read": R * M, "controller_state": CL2}
        want = {"M": (B, N, M), "w": (B, H, N), "read": (B, R, M), "controller_state": (B, CL2)}
        return inner, want

    def _device_state(self, state, want, continuation=False):
        """The reference takes NumPy state through feed_dict (test_tracker.py:284-299): accept host / NumPy state
        here too -- the kernels only ever see device pointers."""
        dev, conv = self.device, {}
        for k, s in want.items():
            v = state[k]
            if not torch.is_tensor(v):
                v = torch.as_tensor(np.asarray(v))
            if tuple(v.shape) != s:
                raise ValueError("state['%s'] has shape %s, expected %s" % (k, tuple(v.shape), s))
            if v.device != dev or v.dtype != torch.float32:
                if continuation:
                    raise ValueError("continuation needs the previous call's device state, got state['%s'] on %s"
                                     % (k, v.device))
                v = v.to(dev, torch.float32)
            conv[k] = v
        return conv

    def _output_state(self, state, want, out_state):
        dev = self.device
        if out_state is None:
            return {k: torch.empty(s, dtype=torch.float32, device=dev) for k, s in want.items()}
        for k, s in want.items():
            v = out_state[k]
            if tuple(v.shape) != s or v.device != dev or v.dtype != torch.float32 or not v.is_contiguous():
                raise ValueError("out_state['%s'] must be a dense float32 %s tensor on %s" % (k, s, dev))
            if v.data_ptr() == state[k].data_ptr():
                raise ValueError("out_state['%s'] aliases the input state" % k)
        return {k: out_state[k] for k in want}

    def _packed_weights(self, shp, wts, plan, stream):
        if self._dirty or self._packed is None:
            lib = _cabi.load()
            self._packed = torch.empty(int(plan.packed_bytes), dtype=torch.uint8, device=self.device)
            _cabi.check(lib.ntm_b200_pack_weights(C.byref(shp), C.byref(wts), self._packed.data_ptr(),
                                                  self._packed.numel(), stream), "pack_weights")
            self._dirty = False
        return self._packed

    # ------------------------------------------------------------------- run --
    def _run(self, inputs, state, steps, history=None, workspace=None, continuation=False, out_state=None):
        """inputs [B, steps, D] float32 CUDA contiguous -> (logits, outputs, new_state, taps).
        `workspace`: caller-provided scratch tensor (else cached per geometry).  `continuation`: this call
        advances the sequences of the previous call on the same workspace; `state` (that call's new_state) is
        updated in place and returned (ntm_b200_forward_seq_continue).  `out_state`: dense state buffers (as
        ``state_placeholder`` makes them) to receive the new state instead of freshly allocated ones."""
        if not torch.cuda.is_available():
            raise RuntimeError("ntm_tracker_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        # the library launches on the CURRENT device with the stream it is handed: pin both to the cell's device
        with torch.cuda.device(self.device):
            return self._run_on_device(inputs, state, steps, history, workspace, continuation, out_state)

    def _run_on_device(self, inputs, state, steps, history, workspace, continuation, out_state=None):
        lib = _cabi.load()
        B, T, D = inputs.shape
        if T != steps:
            raise ValueError("inputs have %d steps, expected %d" % (T, steps))
        if self.input_dim is None:
            self.build(D)
        if D != self.input_dim:
            raise ValueError("inputs have width %d, the cell was built for %d" % (D, self.input_dim))
        dev = self.device
        if inputs.device != dev or inputs.dtype != torch.float32 or not inputs.is_contiguous():
            inputs = inputs.to(dev, torch.float32).contiguous()
        shp = self._shape_struct(D)
        plan = _cabi.Plan()
        _cabi.check(lib.ntm_b200_query(C.byref(shp), B, T, C.byref(plan)), "query")
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        wts = self._weights_struct()
        self._packed_weights(shp, wts, plan, stream)
        ws = workspace if workspace is not None else self._ws.get((B, T))
        if ws is None or ws.numel() < plan.workspace_bytes:
            if workspace is not None:
                raise ValueError("workspace too small: %d < %d bytes" % (ws.numel(), plan.workspace_bytes))
            self.finish()                                # the old workspace holds the error flag of earlier calls
            ws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device=dev)
            self._ws = {(B, T): ws}                      # keep only the latest geometry
        inner, want = self._state_geometry(B)
        state = self._device_state(state, want, continuation)
        if continuation:
            if self.debug or history is not None or any(not state[k].is_contiguous() for k in want):
                raise ValueError("continuation needs a dense state and neither debug taps nor history")
            new_state = state
        else:
            new_state = self._output_state(state, want, out_state)
        sin, keep_in = self._state_struct(state, inner)
        sout, keep_out = self._state_struct(new_state, inner)
        logits = torch.empty(B, T, self.output_dim, dtype=torch.float32, device=dev)
        outputs = torch.empty_like(logits)
        taps = None
        if self.debug:
            taps = torch.zeros(B, int(plan.debug_floats_per_sequence), dtype=torch.float32, device=dev)
        hstruct = None
        if history is not None:      # training mode: record what the backward pass needs
            hstruct = _cabi.History(*[history[k].data_ptr() if history.get(k) is not None else None
                                      for k, _ in _cabi.History._fields_])
        if continuation:
            _cabi.check(lib.ntm_b200_forward_seq_continue(
                C.byref(shp), C.byref(wts), self._packed.data_ptr(), B, T, inputs.data_ptr(),
                C.byref(sout), logits.data_ptr(), outputs.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                "forward_seq_continue")
        else:
            _cabi.check(lib.ntm_b200_forward_seq_train(
                C.byref(shp), C.byref(wts), self._packed.data_ptr(), B, T, inputs.data_ptr(),
                C.byref(sin), C.byref(sout), logits.data_ptr(), outputs.data_ptr(),
                taps.data_ptr() if taps is not None else None,
                C.byref(hstruct) if hstruct is not None else None, ws.data_ptr(), ws.numel(), stream),
                "forward_seq")
        self._last_ws = ws
        return logits, outputs, new_state, taps

    def _run_features(self, features, target, state, delimiter_first=False, out_state=None):
        """Frames in feature layout -- features [B, L, F, Cch] (CUDA), target [B, F] -- through
        ntm_b200_forward_seq_features: T = L*(F+1) steps with the delimiter / target channels synthesised by the
        library (direct_offset_output.py:439-500; serve layout test_tracker.py:385-404 when `delimiter_first`).
        Returns (logits, outputs, new_state) like ``_run``."""
        if not torch.cuda.is_available():
            raise RuntimeError("ntm_tracker_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        if features.dim() != 4 or target.dim() != 2 or tuple(target.shape) != (features.shape[0], features.shape[2]):
            raise ValueError("expected features [B,L,F,C] and target [B,F], got %s and %s"
                             % (tuple(features.shape), tuple(target.shape)))
        dev = self.device
        with torch.cuda.device(dev):
            lib = _cabi.load()
            features = features.to(dev, torch.float32).contiguous()
            target = target.to(dev, torch.float32).contiguous()
            B, L, F, Cch = features.shape
            D, T = Cch + 2, L * (F + 1)
            if self.input_dim is None:
                self.build(D)
            if D != self.input_dim:
                raise ValueError("features have %d channels (+2), the cell was built for input width %d" % (Cch, self.input_dim))
            shp = self._shape_struct(D)
            plan = _cabi.Plan()
            _cabi.check(lib.ntm_b200_query(C.byref(shp), B, T, C.byref(plan)), "query")
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            wts = self._weights_struct()
            self._packed_weights(shp, wts, plan, stream)
            need = int(lib.ntm_b200_features_workspace_bytes(C.byref(shp), B, L, F))
            if need < 0:
                raise ValueError("bad shape for the feature-layout call")
            ws = self._ws.get(("features", B, L, F))
            if ws is None or ws.numel() < need:
                self.finish()
                ws = torch.empty(need, dtype=torch.uint8, device=dev)
                self._ws = {("features", B, L, F): ws}
            inner, want = self._state_geometry(B)
            conv = self._device_state(state, want)
            new_state = self._output_state(conv, want, out_state)
            sin, keep_in = self._state_struct(conv, inner)
            sout, keep_out = self._state_struct(new_state, inner)
            logits = torch.empty(B, T, self.output_dim, dtype=torch.float32, device=dev)
            outputs = torch.empty_like(logits)
            _cabi.check(lib.ntm_b200_forward_seq_features(
                C.byref(shp), C.byref(wts), self._packed.data_ptr(), B, L, F, features.data_ptr(), target.data_ptr(),
                int(bool(delimiter_first)), C.byref(sin), C.byref(sout), logits.data_ptr(), outputs.data_ptr(),
                ws.data_ptr(), ws.numel(), stream), "forward_seq_features")
            self._last_ws = ws
            return logits, outputs, new_state

    def finish(self):
        """Synchronise and surface device-side failures of earlier calls."""
        lib = _cabi.load()
        ws = getattr(self, "_last_ws", None)
        if ws is not None:
            with torch.cuda.device(self.device):
                stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                _cabi.check(lib.ntm_b200_finish(ws.data_ptr(), stream), "finish")

    def _prepare_inputs(self, inputs, ndim):
        if isinstance(inputs, np.ndarray):
            inputs = torch.from_numpy(inputs)
        if inputs.dim() != ndim:
            # _linear: "linear is expecting 2D arguments" (ntm_cell.py:337-338)
            raise ValueError("expected a %dD inputs tensor, got shape %s" % (ndim, tuple(inputs.shape)))
        return inputs.to(self.device, torch.float32, non_blocking=True).contiguous()

    def __call__(self, inputs, prev_state, M_prev=None, w_prev=None,
                 read_prev=None, controller_state=None, scope=None):
        """One cell step (ntm_cell.py:53-253).  ``prev_state`` wins over the
        explicit tensors when it is not None (ntm_cell.py:84-95)."""
        if prev_state is not None:
            state = {k: prev_state[k] for k in ("M", "w", "read", "controller_state")}
        else:
            state = {"M": M_prev, "w": w_prev, "read": read_prev, "controller_state": controller_state}
            if any(v is None for v in state.values()):
                raise ValueError("either prev_state or all of M_prev, w_prev, read_prev, "
                                 "controller_state must be given")
        x = self._prepare_inputs(inputs, 2)
        state = {k: (torch.as_tensor(v) if not torch.is_tensor(v) else v).to(self.device) for k, v in state.items()}
        logits, outputs, new_state, taps = self._run(x.unsqueeze(1), state, 1)
        ntm_output_logit = logits[:, 0]
        ntm_output = outputs[:, 0]
        debug = self._debug_dict(taps, state, new_state) if taps is not None else {}
        return (ntm_output, ntm_output_logit, new_state, debug, new_state["M"], new_state["w"],
                new_state["read"], new_state["controller_state"])

    def _debug_dict(self, taps, prev, new):
        """The reference's `debug` dict (ntm_cell.py:230-250) from the kernel's taps."""
        B = taps.shape[0]
        H, R, W, N, M = self.num_heads, self.read_head_size, self.write_head_size, self.mem_size, self.mem_dim
        S = 2 * self.shift_range + 1
        sizes = [H * M, H, H, S * H, H, W * M, W * M] + [H * N] * 5
        parts = torch.split(taps, sizes, dim=1)
        k, beta, g, sw, gamma, erase, add, sim, wc, wgated, wconv, powed = parts
        w = new["w"]
        w_write = w[:, R:]
        erase = erase.reshape(B, W, M)
        add = add.reshape(B, W, M)
        M_erase = torch.prod(1.0 - w_write.unsqueeze(3) * erase.unsqueeze(2), dim=1)
        M_write = torch.sum(w_write.unsqueeze(3) * add.unsqueeze(2), dim=1)
        return {
            "k": k.reshape(B, H, M), "gamma": gamma.reshape(B, H, 1), "add": add, "erase": erase,
            "bega": beta.reshape(B, H, 1), "g": g.reshape(B, H, 1), "sw": sw.reshape(B, H, S),
            "similarity": sim.reshape(B, H, N), "w_content_focused": wc.reshape(B, H, N),
            "w_gated": wgated.reshape(B, H, N), "w_conv": wconv.reshape(B, H, N),
            "w_conv_powed": powed.reshape(B, H, N), "w": w, "w_read": w[:, :R], "w_write": w_write,
            "M": new["M"], "M_prev": prev["M"], "M_write": M_write, "M_erase": M_erase,
        }
