#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

TEST INFRASTRUCTURE.  Run in the authoring container only (needs
/root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

How: /root/reference/{utils,ops,utility,ntm_cell,ntm_tracker_new}.py are read
from where they lie (never copied), compiled with Python-2 semantics restored
(`xrange`, list-returning `range`, and -- for files without
`from __future__ import division` -- integer `/` as floor division, which is
what gives ops.py:204 its {-2,-1,0} shift taps), and executed against the
NumPy TF1 shim in oracle/tf1_shim/.  Parameters come from
oracle.ntm_oracle.init_params keyed by TF variable names; the shim raises if
the reference asks for a variable name/shape the table does not have, so the
name mapping of SURVEY.md s5 is pinned too.

The outputs are what tests/test_oracle_golden.py holds the NumPy restatement
(oracle/ntm_oracle.py) to, and what the GPU parity tests hold the CUDA path to.
"""
import ast
import builtins
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "tf1_shim"))
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)

from oracle import ntm_oracle as O  # noqa: E402


def _py2div(a, b):
    ints = (int, np.integer)
    if isinstance(a, ints) and isinstance(b, ints):
        return a // b
    return a / b


class _Py2Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()),
                         args=[node.left, node.right], keywords=[]), node)
        return node


def load_reference_module(name):
    path = os.path.join(REF, name + ".py")
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    future_div = any(isinstance(n, ast.ImportFrom) and n.module == "__future__"
                     and any(a.name == "division" for a in n.names) for n in tree.body)
    if not future_div:
        tree = ast.fix_missing_locations(_Py2Div().visit(tree))
    mod = types.ModuleType(name)
    mod.__file__ = path
    mod.__dict__.update(xrange=builtins.range, _py2div=_py2div,
                        range=lambda *a: list(builtins.range(*a)))
    sys.modules[name] = mod
    exec(compile(tree, path, "exec"), mod.__dict__)
    return mod


def load_reference():
    for n in ("utils", "ops", "utility", "ntm_cell", "ntm_tracker_new"):
        load_reference_module(n)
    return sys.modules["ops"], sys.modules["ntm_cell"], sys.modules["ntm_tracker_new"]


def cell_kwargs(s):
    return dict(mem_size=s.mem_size, mem_dim=s.mem_dim, shift_range=s.shift_range,
                controller_hidden_size=s.controller_hidden_size,
                controller_num_layers=s.controller_num_layers,
                write_head_size=s.write_head_size, read_head_size=s.read_head_size,
                write_first=s.write_first)


def run_reference_loop(ntm_tracker_new, s, params, x):
    """LoopNTMTracker(T, O, init, **kw)(inputs) exactly as direct_offset_output.py:528-543."""
    tf.reset_store(params)
    tracker = ntm_tracker_new.LoopNTMTracker(
        x.shape[1], s.output_dim, tf.random_uniform_initializer(-.05, .05), **cell_kwargs(s))
    res = tracker(tf.Tensor(x))
    created = tf.created_variables()
    assert set(created) == set(params), (sorted(set(created) ^ set(params)))
    return res[0].a, res[1].a


def run_reference_stepwise(ntm_cell, s, params, x, debug_step=None):
    """The serve-path usage (test_tracker.py:284-299,331-342): one cell step at a
    time with the state dict carried by the caller."""
    tf.reset_store(params)
    B, T, _ = x.shape
    outs, logits, dbg = [], [], None
    with tf.variable_scope("ntm-tracker"):
        cell = ntm_cell.NTMCell(s.output_dim, **cell_kwargs(s))
        state = cell.zero_state(B)
        for t in range(T):
            o, lg, state, debug, M, w, read, cs = cell(tf.Tensor(x[:, t]), state)
            outs.append(o.a)
            logits.append(lg.a)
            if debug_step == t:
                dbg = {k: np.array(v.a) for k, v in debug.items()}
    st = {k: np.array(v.a) for k, v in state.items()}
    return np.stack(outs, 1), np.stack(logits, 1), st, dbg


def shape_to_npz(s):
    return np.array([s.output_dim, s.input_dim, s.mem_size, s.mem_dim, s.shift_range,
                     s.controller_hidden_size, s.controller_num_layers,
                     s.write_head_size, s.read_head_size, int(s.write_first)], np.int64)


CASES = {
    # name: (NTMShape, B, T, seed, random_biases, input kind, subsample stride for M)
    "small_r2w1_l2": (O.NTMShape(output_dim=3, input_dim=5, mem_size=16, mem_dim=8,
                                 controller_hidden_size=12, controller_num_layers=2,
                                 write_head_size=1, read_head_size=2), 3, 6, 11, True, "normal", 1),
    "small_writefirst_s2": (O.NTMShape(output_dim=2, input_dim=7, mem_size=24, mem_dim=12,
                                       shift_range=2, controller_hidden_size=16,
                                       controller_num_layers=1, write_head_size=2,
                                       read_head_size=3, write_first=True), 2, 5, 12, True, "normal", 1),
    "c1_copy": (O.NTMShape(**O.CONFIGS["c1_copy"][0]), 16, 20, 1235, False, "copy", 1),
    "c2_tracker_b2t4": (O.NTMShape(**O.CONFIGS["c2_tracker"][0]), 2, 4, 1236, False, "tracker", 4),
    "defaults_r3w3_l3": (O.NTMShape(output_dim=4, input_dim=6, mem_size=32, mem_dim=20,
                                    controller_hidden_size=20, controller_num_layers=3,
                                    write_head_size=3, read_head_size=3), 2, 4, 13, True, "normal", 1),
}


def make_inputs(kind, s, B, T, seed):
    if kind == "copy":
        return O.copy_task_inputs(B, T, s.input_dim - 1, seed)
    if kind == "tracker":
        return O.tracker_inputs(B, T, seed, feat=s.input_dim - 2, frame=3)
    return np.random.RandomState(seed).standard_normal((B, T, s.input_dim)).astype(np.float32)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    tf.set_precision(np.float64)
    ops, ntm_cell, ntm_tracker_new = load_reference()

    # -- 1. the reference's own known-answer test, ops_test.py:20-34 ------------
    memory = np.array([[[1, 2, 3], [2, 2, 2], [3, 2, 1], [0, 2, 4]]], np.float32)
    keys = np.array([[[2, 2, 2], [1, 2, 3]]], np.float32)
    shipped = ops.batched_smooth_cosine_similarity(tf.Tensor(memory), tf.Tensor(keys)).a
    stale = np.array([[[0.92574867671153, 0.99991667361053, 0.92574867671153, 0.77454667246876],
                       [0.999928, 0.925749, 0.714235, 0.956126]]])  # ops_test.py:27-34
    np.savez(os.path.join(out_dir, "kat_similarity.npz"), memory=memory, keys=keys,
             shipped_code=shipped, ops_test_expected=stale)
    print("kat_similarity: shipped code ->", np.round(shipped, 6).tolist())

    # -- 2. circular convolution taps (ops.py:180-242) ----------------------------
    rng = np.random.RandomState(5)
    for S in (3, 5):
        w = rng.rand(2, 3, 10)
        kern = rng.rand(2, 3, S)
        res = ops.batched_circular_convolution(tf.Tensor(w), tf.Tensor(kern)).a
        np.savez(os.path.join(out_dir, "kat_circular_conv_s%d.npz" % S), w=w, kernel=kern, out=res)

    # -- 3. cell / loop cases ---------------------------------------------------------
    for name, (s, B, T, seed, rb, kind, sub) in CASES.items():
        params = O.init_params(s, seed, 0.05, random_biases=rb)
        x = make_inputs(kind, s, B, T, seed + 100)
        lo, ll = run_reference_loop(ntm_tracker_new, s, params, x)
        so, sl, st, dbg = run_reference_stepwise(ntm_cell, s, params, x, debug_step=min(1, T - 1))
        assert np.array_equal(lo, so) and np.array_equal(ll, sl), "loop and stepwise paths disagree"
        blob = dict(shape=shape_to_npz(s), B=B, T=T, seed=seed, random_biases=int(rb),
                    kind=kind, m_stride=sub, inputs=x, outputs=lo, logits=ll,
                    final_w=st["w"], final_read=st["read"],
                    final_controller_state=st["controller_state"],
                    final_M=st["M"][:, ::sub, ::sub])
        for k, v in dbg.items():
            v = v[:2]            # debug taps for the first two sequences only (fixture size)
            if v.ndim >= 3 and v.shape[-1] == s.mem_dim and v.shape[-2] == s.mem_size:
                v = v[..., ::sub, ::sub]
            blob["dbg_" + k] = v
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **blob)
        print("%-22s B=%d T=%d  sum(w[0,0])=%.4f  |logit|max=%.4f" % (
            name, B, T, st["w"][0, 0].sum(), np.abs(ll).max()))


if __name__ == "__main__":
    main()
