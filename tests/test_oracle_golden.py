"""CPU: the NumPy oracle (oracle/ntm_oracle.py) against the golden vectors that
oracle/make_golden.py recorded by executing the reference's own source files
(/root/reference/{ops,ntm_cell,ntm_tracker_new}.py) under the TF1 shim.

Mirrors the reference's test style (ops_test.py: assertAllClose on a KAT;
dnc/addressing_test.py: random inputs vs explicit NumPy loops, degenerate-input
NaN checks)."""
import glob
import os

import numpy as np
import pytest

from oracle import ntm_oracle as O

CASES = ["small_r2w1_l2", "small_writefirst_s2", "c1_copy", "c2_tracker_b2t4", "defaults_r3w3_l3"]


def load_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    v = [int(t) for t in z["shape"]]
    s = O.NTMShape(output_dim=v[0], input_dim=v[1], mem_size=v[2], mem_dim=v[3], shift_range=v[4],
                   controller_hidden_size=v[5], controller_num_layers=v[6],
                   write_head_size=v[7], read_head_size=v[8], write_first=bool(v[9]))
    params = O.init_params(s, int(z["seed"]), 0.05, random_biases=bool(z["random_biases"]))
    return z, s, params


def test_kat_similarity_shipped_code(golden_dir):
    """ops_test.py:20-34 inputs; expected = what ops.py:147-156 computes today."""
    z = np.load(os.path.join(golden_dir, "kat_similarity.npz"))
    got = O.batched_smooth_cosine_similarity(z["memory"].astype(np.float64), z["keys"].astype(np.float64))
    np.testing.assert_allclose(got, z["shipped_code"], rtol=0, atol=1e-12)
    # the values SURVEY.md s4 recomputed independently
    np.testing.assert_allclose(
        got, [[[0.75920624, 0.80810034, 0.8569944, 0.7103121],
               [0.7778448, 0.70288837, 0.62793195, 0.85280126]]], atol=1e-6)


def test_kat_similarity_stale_reference_vector(golden_dir):
    """The reference's own golden vector pins the SUPERSEDED row-wise smooth
    cosine: it must match the legacy formula and must NOT match shipped code."""
    z = np.load(os.path.join(golden_dir, "kat_similarity.npz"))
    legacy = O.rowwise_smooth_cosine_similarity(z["memory"].astype(np.float64), z["keys"].astype(np.float64))
    np.testing.assert_allclose(legacy, z["ops_test_expected"], atol=1e-6)
    assert np.abs(z["shipped_code"] - z["ops_test_expected"]).max() > 0.05


@pytest.mark.parametrize("S", [3, 5])
def test_kat_circular_convolution(golden_dir, S):
    z = np.load(os.path.join(golden_dir, "kat_circular_conv_s%d.npz" % S))
    np.testing.assert_allclose(O.batched_circular_convolution(z["w"], z["kernel"]), z["out"], atol=1e-14)
    # explicit loop form of the spec line in SURVEY.md s8a (taps {-2,-1,0} for S=3)
    offs = O.shift_offsets(S)
    assert offs == ({3: [-2, -1, 0], 5: [-3, -2, -1, 0, 1]}[S])
    w, k = z["w"], z["kernel"]
    N = w.shape[-1]
    exp = np.zeros_like(w)
    for n in range(N):
        for t, j in enumerate(offs):
            exp[..., n] += k[..., t] * w[..., (n + j) % N]
    np.testing.assert_allclose(z["out"], exp, atol=1e-14)


@pytest.mark.parametrize("name", CASES)
def test_sequence_matches_reference_source(golden_dir, name):
    z, s, params = load_case(golden_dir, name)
    sub = int(z["m_stride"])
    out, logit, st = O.run_sequence(params, s, z["inputs"], dtype=np.float64)
    tol = dict(rtol=0, atol=1e-12)
    np.testing.assert_allclose(out, z["outputs"], **tol)
    np.testing.assert_allclose(logit, z["logits"], **tol)
    np.testing.assert_allclose(st["w"], z["final_w"], **tol)
    np.testing.assert_allclose(st["read"], z["final_read"], **tol)
    np.testing.assert_allclose(st["controller_state"], z["final_controller_state"], **tol)
    np.testing.assert_allclose(st["M"][:, ::sub, ::sub], z["final_M"], **tol)


@pytest.mark.parametrize("name", CASES)
def test_debug_taps_match_reference_source(golden_dir, name):
    """The 19 intermediates of ntm_cell.py:230-250 at step 1."""
    z, s, params = load_case(golden_dir, name)
    sub = int(z["m_stride"])
    x = z["inputs"].astype(np.float64)
    state = O.zero_state(params, s, x.shape[0])
    dbg = None
    for t in range(2 if x.shape[1] > 1 else 1):
        _, _, state, dbg = O.cell_step(params, s, x[:, t], state, debug=True)
    keys = [k[4:] for k in z.files if k.startswith("dbg_")]
    assert len(keys) == 19
    for k in keys:
        got = dbg[k][:2]
        if got.ndim >= 3 and got.shape[-1] == s.mem_dim and got.shape[-2] == s.mem_size:
            got = got[..., ::sub, ::sub]
        np.testing.assert_allclose(np.squeeze(got), np.squeeze(z["dbg_" + k]), rtol=0, atol=1e-12, err_msg=k)


def test_fp32_oracle_within_budget_of_fp64(golden_dir):
    """The fp32 restatement stays inside the 1e-4 parity budget of the fp64 one
    (the budget the CUDA path is held to on read / w / M / logits)."""
    z, s, params = load_case(golden_dir, "c1_copy")
    o64, l64, s64 = O.run_sequence(params, s, z["inputs"], dtype=np.float64)
    o32, l32, s32 = O.run_sequence(params, s, z["inputs"], dtype=np.float32)
    for k in ("M", "w", "read"):
        assert np.abs(s32[k] - s64[k]).max() < 1e-5, k
    assert np.abs(l32 - l64).max() < 1e-5


def test_quirks_hold(golden_dir):
    """Parity-critical quirks 3 and 4 of SURVEY.md s0: weightings sum to ~0.97
    (the +1e-3 in the sharpening denominator), initial weighting sums to ~N/2."""
    z, s, params = load_case(golden_dir, "c1_copy")
    st0 = O.zero_state(params, s, 1)
    assert abs(st0["w"].sum(-1).mean() - s.mem_size / 2) < 2.0
    assert 0.95 < z["final_w"].sum(-1).mean() < 0.999


def test_degenerate_inputs_no_nan():
    """dnc/addressing_test.py-style divide-by-zero check: all-zero memory and
    all-zero keys must not produce NaN/Inf (l2_normalize's 1e-12 floor)."""
    s = O.NTMShape(output_dim=2, input_dim=3, mem_size=8, mem_dim=4, controller_hidden_size=6,
                   controller_num_layers=1, write_head_size=1, read_head_size=1)
    params = {k: np.zeros_like(v) for k, v in O.init_params(s, 0).items()}
    st = O.zero_state(params, s, 2)
    st["M"][:] = 0.0
    st["w"][:] = 0.0
    out, logit, st2, _ = O.cell_step(params, s, np.zeros((2, 3)), st)
    for v in (out, logit, st2["M"], st2["w"], st2["read"], st2["controller_state"]):
        assert np.isfinite(v).all()


def test_all_golden_files_present(golden_dir):
    have = {os.path.basename(p) for p in glob.glob(os.path.join(golden_dir, "*.npz"))}
    need = {c + ".npz" for c in CASES} | {"kat_similarity.npz", "kat_circular_conv_s3.npz",
                                          "kat_circular_conv_s5.npz", "layout_serve.npz", "layout_train.npz"}
    assert need <= have


def test_serve_layout_matches_reference_rows(golden_dir):
    """test_tracker.py:380-404 executed by oracle/make_golden_layout.py: delimiter row [0..0,1,0] FIRST,
    then rows [feat_f, 0, gt_f] on the first frame and [feat_f, 0, 0] afterwards -- feature f of the
    first frame is step f + 1 and the delimiter's target bit is 0."""
    z = np.load(os.path.join(golden_dir, "layout_serve.npz"))
    feats, gt, rows = z["features"], z["gt"], z["rows"]
    nfr, F, Cc = feats.shape
    # frame by frame, the way ResidentTracker.track feeds them (target zeros after the first frame)
    for i in range(nfr):
        tgt = gt[None] if i == 0 else np.zeros((1, F), np.float32)
        got = O.serialize_tracker_inputs(feats[i][None, None], tgt, delimiter_first=True)
        assert np.array_equal(got[0], rows[i]), i
    # all frames of the sequence in one call
    got = O.serialize_tracker_inputs(feats[None], gt[None], delimiter_first=True)
    assert np.array_equal(got[0], rows.reshape(nfr * (F + 1), Cc + 2))
    assert rows[0, 0, Cc] == 1.0 and rows[0, 0, Cc + 1] == 0.0
    assert np.array_equal(rows[0, 1:, Cc + 1], gt) and np.abs(rows[1:, :, Cc + 1]).sum() == 0


def test_training_layout_and_gather_match_reference_statements(golden_dir):
    """direct_offset_output.py:439-500 and :581-593 executed under the TF1 shim."""
    z = np.load(os.path.join(golden_dir, "layout_train.npz"))
    got = O.serialize_tracker_inputs(z["features"], z["target"], delimiter_first=False)
    assert np.array_equal(got, z["inputs"])
    F = z["features"].shape[2]
    np.testing.assert_allclose(O.gather_offsets(z["logits"], F), z["offsets"], rtol=0, atol=1e-6)
