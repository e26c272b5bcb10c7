"""CPU: readers of the reference's per-frame files (written here exactly as preprocess.py:322-334 does)."""
import os

import numpy as np
import pytest

from ntm_tracker_b200.formats import load_sequence, read_frame_gt, read_frame_txt


def write_frame(d, name, crop, bbox, img, yo, xo, gt):
    gt.astype(np.float64).tofile(os.path.join(d, name + ".bin"))
    with open(os.path.join(d, name + ".txt"), "w") as f:
        f.write("{crop[0]},{crop[1]},{crop[2]},{crop[3]},{bbox[0]},{bbox[1]},{bbox[2]},{bbox[3]},{image_path},{y_offset},{x_offset}".format(
            crop=crop, bbox=bbox, image_path=img, y_offset=yo, x_offset=xo))


def test_round_trip(tmp_path):
    d = str(tmp_path)
    rng = np.random.RandomState(0)
    gts = [rng.rand(8, 8) for _ in range(3)]
    for i, g in enumerate(gts):
        write_frame(d, "%06d" % i, [0.1, 0.2, 0.8, 0.9], [10, 20, 30, 40], "/data/img%d.JPEG" % i, 0.25 * i, -0.5 * i, g)
    rec = read_frame_txt(os.path.join(d, "000001.txt"))
    assert rec["image_path"] == "/data/img1.JPEG" and rec["y_offset"] == 0.25 and rec["x_offset"] == -0.5
    np.testing.assert_allclose(rec["cropbox"], [0.1, 0.2, 0.8, 0.9], rtol=1e-6)
    np.testing.assert_allclose(read_frame_gt(os.path.join(d, "000002.bin")), gts[2].astype(np.float32))
    crops, offs, g, paths = load_sequence(d, ["%06d" % i for i in range(3)], reverse_image=True)
    assert crops.shape == (3, 4) and offs.shape == (3, 2) and g.shape == (3, 64) and len(paths) == 3
    np.testing.assert_allclose(offs[:, 1], [0.0, 0.5, 1.0])          # x offsets negated


def test_malformed_files_raise(tmp_path):
    p = str(tmp_path / "x.txt")
    open(p, "w").write("1,2,3")
    with pytest.raises(ValueError):
        read_frame_txt(p)
    b = str(tmp_path / "x.bin")
    np.zeros(10).tofile(b)
    with pytest.raises(ValueError):
        read_frame_gt(b)
