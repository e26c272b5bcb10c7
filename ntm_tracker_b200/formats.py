"""Per-frame on-disk formats of the reference's dataset -- host-side readers.

``preprocess.py:322-334`` writes, per object and frame, ``<frame>.txt`` (one CSV line of 11 fields:
cropbox y1,x1,y2,x2, transformed bbox (4 values), image path, y_offset, x_offset) and ``<frame>.bin``
(the gt_width x gt_width ground-truth map as raw float64, 8x8 by default);
``direct_offset_output.py:159-224`` reads them back with ``tf.decode_csv`` (record defaults
``[.0]*8 + [''] + [.0]*2``) and ``tf.decode_raw(value, tf.float64)`` cast to float32.  The image
decoding / VGG part of that pipeline is out of scope; these readers recover the
numeric side: cropbox, offsets (the regression targets) and the target map that feeds
``serialize.tracker_inputs``.
"""
import os

import numpy as np


def read_frame_txt(path):
    """-> dict(cropbox [y1,x1,y2,x2] float32, bbox [4] float32, image_path str, y_offset, x_offset)."""
    with open(path) as f:
        line = f.readline().strip()
    parts = line.split(",")
    if len(parts) != 11:
        raise ValueError("%s: expected 11 comma-separated fields, got %d" % (path, len(parts)))
    vals = [float(p) for p in parts[:8]]
    return {"cropbox": np.asarray(vals[:4], np.float32), "bbox": np.asarray(vals[4:8], np.float32),
            "image_path": parts[8], "y_offset": float(parts[9]), "x_offset": float(parts[10])}


def read_frame_gt(path, gt_width=8):
    """``<frame>.bin``: gt_width*gt_width float64 -> float32 [gt_width, gt_width]."""
    raw = np.fromfile(path, dtype=np.float64)
    if raw.size != gt_width * gt_width:
        raise ValueError("%s: expected %d float64 values, got %d" % (path, gt_width * gt_width, raw.size))
    return raw.astype(np.float32).reshape(gt_width, gt_width)


def load_sequence(directory, frames, gt_width=8, reverse_image=False):
    """Frames (file names without suffix) of one object -> (cropboxes [L,4], offsets [L,2] as
    (y, x), gts [L, gt_width*gt_width], image paths).  ``reverse_image`` negates x offsets
    (direct_offset_output.py:186-187)."""
    crops, offs, gts, paths = [], [], [], []
    for name in frames:
        rec = read_frame_txt(os.path.join(directory, name + ".txt"))
        crops.append(rec["cropbox"])
        offs.append([rec["y_offset"], -rec["x_offset"] if reverse_image else rec["x_offset"]])
        gts.append(read_frame_gt(os.path.join(directory, name + ".bin"), gt_width).reshape(-1))
        paths.append(rec["image_path"])
    return np.stack(crops), np.asarray(offs, np.float32), np.stack(gts), paths
