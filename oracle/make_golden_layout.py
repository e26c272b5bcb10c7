#!/usr/bin/env python
"""Generate tests/golden/layout_*.npz by EXECUTING THE REFERENCE'S OWN STATEMENTS for the two
input layouts either side of the NTM path.

TEST INFRASTRUCTURE.  Run in the authoring container only (needs /root/reference):

    python oracle/make_golden_layout.py

  * serve layout    -- the body of ``_preprocess_image`` after its ``sess.run`` (test_tracker.py:380-404:
                       pad column, ground-truth column on the first frame only, delimiter row [0..0,1,0]
                       PREPENDED) is lifted out of the reference file by line number and executed as is on
                       NumPy arrays, with ``preprocess.generate_gt`` stubbed to return the given map;
  * training layout -- the statements of ``ntm_offsets`` that build the tracker inputs
                       (direct_offset_output.py:439-500) and that gather the outputs (:581-593) are lifted
                       out the same way and executed against the NumPy TF1 shim (oracle/tf1_shim/).

Nothing is copied: the statements are read from where they lie, compiled and run; only their OUTPUTS
are stored.  tests/test_oracle_golden.py holds oracle.serialize_tracker_inputs / gather_offsets to them
bit-exactly and tests/test_gpu_io.py holds the CUDA serialiser to them.
"""
import ast
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


def lifted(path, func, lo, hi):
    """Statements of function `func` in `path` whose lines lie in [lo, hi], compiled as a module."""
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == func:
            fn = node
    assert fn is not None, func
    body = [st for st in fn.body if st.lineno >= lo and (st.end_lineno or st.lineno) <= hi]
    assert body, (func, lo, hi)
    mod = ast.Module(body=body, type_ignores=[])
    return compile(ast.fix_missing_locations(mod), path, "exec")


def serve_rows(features, gt, is_first_frame):
    """test_tracker.py:380-404 on one frame's [F, C] features (everything between the sess.run that
    produces `features` and the `return features`)."""
    code = lifted(os.path.join(REF, "test_tracker.py"), "_preprocess_image", 380, 404)
    pre = types.SimpleNamespace(generate_gt=lambda *a: gt, apply_transformation=lambda *a: None)
    env = dict(np=np, features=features, is_first_frame=is_first_frame, preprocess=pre,
               self=types.SimpleNamespace(normalized_bbox=None, transformation=None),
               FLAGS=types.SimpleNamespace(cropbox_grid=None, bbox_grid=None))
    exec(code, env)
    return np.asarray(env["features"])


def training_inputs(features, batch_gt, B, L):
    """direct_offset_output.py:439-500 on features [B*L, F, C] and batch_gt [B*L, F]."""
    if not hasattr(tf, "tile"):
        tf.tile = lambda x, multiples, name=None: tf.Tensor(np.tile(tf._np(x), multiples))
    tf.float32 = getattr(tf, "float32", np.float32)
    code = lifted(os.path.join(REF, "direct_offset_output.py"), "ntm_offsets", 439, 500)
    F, C = features.shape[1], features.shape[2]
    env = dict(tf=tf, FLAGS=types.SimpleNamespace(batch_size=B, sequence_length=L),
               features=tf.Tensor(features), batch_gt=tf.Tensor(batch_gt), num_features=F, num_channels=C,
               print=lambda *a, **k: None)
    exec(code, env)
    return np.asarray(env["inputs"].a), F


def training_gather(output_logits, B, L, F):
    """direct_offset_output.py:581-593."""
    code = lifted(os.path.join(REF, "direct_offset_output.py"), "ntm_offsets", 581, 593)
    env = dict(tf=tf, FLAGS=types.SimpleNamespace(batch_size=B, sequence_length=L),
               output_logits=tf.Tensor(output_logits), num_features=F, print=lambda *a, **k: None)
    exec(code, env)
    return np.asarray(env["output_sigmoids"].a)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    tf.set_precision(np.float64)
    rng = np.random.RandomState(2024)
    # ---- serve layout: three frames of one sequence, first frame carries the ground-truth map ----
    F, C = 6, 5
    feats = np.maximum(rng.standard_normal((3, F, C)), 0.0)
    gt = rng.rand(F, 1)
    rows = [serve_rows(feats[i], gt, i == 0) for i in range(3)]
    serve = np.stack(rows, 0)                      # [3 frames, F+1, C+2]
    assert serve.shape == (3, F + 1, C + 2)
    np.savez(os.path.join(out_dir, "layout_serve.npz"), features=feats.astype(np.float32),
             gt=gt[:, 0].astype(np.float32), rows=serve.astype(np.float32))
    print("serve: delimiter row", serve[0, 0].tolist(), " first feature row tail", serve[0, 1, C:].tolist())
    # ---- training layout ----
    B, L = 2, 3
    feats = np.maximum(rng.standard_normal((B * L, F, C)), 0.0)
    gts = rng.rand(B * L, F)
    x, _ = training_inputs(feats, gts, B, L)
    assert x.shape == (B, L * (F + 1), C + 2)
    logits = rng.standard_normal((B, L * (F + 1), 2))
    off = training_gather(logits, B, L, F)
    np.savez(os.path.join(out_dir, "layout_train.npz"), features=feats.reshape(B, L, F, C).astype(np.float32),
             target=gts.reshape(B, L, F)[:, 0].astype(np.float32), inputs=x.astype(np.float32),
             logits=logits.astype(np.float32), offsets=off.astype(np.float32))
    print("train: inputs", x.shape, " offsets", off.shape)


if __name__ == "__main__":
    main()
