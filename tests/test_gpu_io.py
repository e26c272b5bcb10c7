"""GPU: the input serialiser and the output gather (csrc/ntm_b200_io.cu) against the NumPy
restatement of direct_offset_output.py:439-500 / :581-593 -- bit-exact for the copy, 1e-6 for tanh."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import ntm_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,F,C,first", [(3, 4, 64, 512, False), (2, 3, 64, 512, True), (5, 2, 7, 9, False),
                                           (1, 6, 3, 1, True)])
def test_serialize_matches_reference_layout(B, L, F, C, first):
    from ntm_tracker_b200.serialize import tracker_inputs
    rng = np.random.RandomState(B * 100 + L)
    feat = np.maximum(rng.standard_normal((B, L, F, C)), 0).astype(np.float32)
    tgt = rng.rand(B, F).astype(np.float32)
    got = tracker_inputs(torch.from_numpy(feat).cuda(), torch.from_numpy(tgt).cuda(), delimiter_first=first)
    exp = O.serialize_tracker_inputs(feat, tgt, delimiter_first=first)
    assert got.shape == exp.shape == (B, L * (F + 1), C + 2)
    assert np.array_equal(got.cpu().numpy(), exp)
    # structure: one delimiter bit per frame, target only within the first F steps
    x = got.cpu().numpy()
    assert x[:, :, C].sum() == B * L
    first_frame = x[:, :F + 1, C + 1]
    assert np.abs(x[:, F + 1:, C + 1]).sum() == 0                       # target only inside the first frame ...
    assert np.array_equal(first_frame[:, 1:] if first else first_frame[:, :F], tgt)   # ... on its feature rows


def test_serialize_matches_reference_golden_layouts(golden_dir):
    """Against the rows the reference's OWN statements produce (oracle/make_golden_layout.py): serve
    layout test_tracker.py:380-404 (per frame and all frames at once), training layout
    direct_offset_output.py:439-500, gather :581-593."""
    import os
    from ntm_tracker_b200.serialize import gather_offsets, tracker_inputs
    z = np.load(os.path.join(golden_dir, "layout_serve.npz"))
    feats, gt, rows = z["features"], z["gt"], z["rows"]
    nfr, F, Cc = feats.shape
    for i in range(nfr):
        tgt = gt[None] if i == 0 else np.zeros((1, F), np.float32)
        got = tracker_inputs(torch.from_numpy(feats[i][None, None]).cuda(), torch.from_numpy(tgt).cuda(),
                             delimiter_first=True)
        assert np.array_equal(got.cpu().numpy()[0], rows[i]), i
    got = tracker_inputs(torch.from_numpy(feats[None]).cuda(), torch.from_numpy(gt[None]).cuda(), delimiter_first=True)
    assert np.array_equal(got.cpu().numpy()[0], rows.reshape(nfr * (F + 1), Cc + 2))
    z = np.load(os.path.join(golden_dir, "layout_train.npz"))
    got = tracker_inputs(torch.from_numpy(z["features"]).cuda(), torch.from_numpy(z["target"]).cuda())
    assert np.array_equal(got.cpu().numpy(), z["inputs"])
    off = gather_offsets(torch.from_numpy(z["logits"]).cuda(), z["features"].shape[2]).cpu().numpy()
    assert np.abs(off - z["offsets"]).max() <= 1e-6


@pytest.mark.parametrize("B,L,F,Od", [(4, 5, 64, 2), (2, 2, 3, 5)])
def test_gather_offsets(B, L, F, Od):
    from ntm_tracker_b200.serialize import gather_offsets
    lg = np.random.RandomState(L).standard_normal((B, L * (F + 1), Od)).astype(np.float32)
    got = gather_offsets(torch.from_numpy(lg).cuda(), F).cpu().numpy()
    exp = O.gather_offsets(lg, F)
    assert got.shape == (B, L - 1, Od)
    assert np.abs(got - exp).max() <= 1e-6


def test_serialized_inputs_drive_the_tracker():
    """features -> serialiser -> LoopNTMTracker -> gather, end to end on the device, vs the oracle."""
    from ntm_tracker_b200 import LoopNTMTracker
    from ntm_tracker_b200.serialize import gather_offsets, tracker_inputs
    B, L, F, C = 3, 3, 4, 6
    s = O.NTMShape(output_dim=2, input_dim=C + 2, mem_size=16, mem_dim=8, controller_hidden_size=10,
                   controller_num_layers=1, write_head_size=1, read_head_size=2)
    params = O.init_params(s, 3, 0.05)
    rng = np.random.RandomState(1)
    feat = np.maximum(rng.standard_normal((B, L, F, C)), 0).astype(np.float32)
    tgt = rng.rand(B, F).astype(np.float32)
    x = tracker_inputs(torch.from_numpy(feat).cuda(), torch.from_numpy(tgt).cuda())
    trk = LoopNTMTracker(L * (F + 1), 2, mem_size=16, mem_dim=8, controller_hidden_size=10,
                         controller_num_layers=1, write_head_size=1, read_head_size=2)
    trk.cell.load_reference_weights(params)
    _, logits = trk(x)
    off = gather_offsets(logits, F).cpu().numpy()
    _, rl, _ = O.run_sequence(params, s, O.serialize_tracker_inputs(feat, tgt))
    assert np.abs(off - O.gather_offsets(rl, F)).max() <= 1e-4


def test_bad_shapes_raise():
    from ntm_tracker_b200.serialize import gather_offsets, tracker_inputs
    with pytest.raises(ValueError):
        tracker_inputs(torch.zeros(2, 3, 4, 5).cuda(), torch.zeros(2, 3).cuda())
    with pytest.raises(ValueError):
        gather_offsets(torch.zeros(2, 10, 2).cuda(), 3)


def test_resident_tracker_session_matches_stepwise_oracle():
    """Serve path (test_tracker.py:284-299): three frames of F+1 cell steps each, state carried on
    the device between frames, against the oracle stepping the same rows one at a time."""
    from ntm_tracker_b200 import NTMCell, ResidentTracker
    F, C = 5, 6
    s = O.NTMShape(output_dim=2, input_dim=C + 2, mem_size=16, mem_dim=8, controller_hidden_size=10,
                   controller_num_layers=1, write_head_size=1, read_head_size=2)
    params = O.init_params(s, 4, 0.05)
    cell = NTMCell(2, mem_size=16, mem_dim=8, controller_hidden_size=10, controller_num_layers=1,
                   write_head_size=1, read_head_size=2)
    cell.load_reference_weights(params)
    sess = ResidentTracker(cell, num_features=F, batch_size=1).reset()
    rng = np.random.RandomState(5)
    state = O.zero_state(params, s, 1)
    for frame in range(3):
        feat = np.maximum(rng.standard_normal((1, F, C)), 0).astype(np.float32)
        tgt = rng.rand(1, F).astype(np.float32) if frame == 0 else np.zeros((1, F), np.float32)
        got = sess.track(torch.from_numpy(feat).cuda(), torch.from_numpy(tgt).cuda()).cpu().numpy()
        rows = O.serialize_tracker_inputs(feat[:, None], tgt, delimiter_first=True)[0]
        for r in rows:
            _, lg, state, _ = O.cell_step(params, s, r[None], state)
        assert np.abs(got - np.tanh(lg)).max() <= 1e-4
    for k in ("M", "w", "read"):
        assert np.abs(sess.state[k].cpu().numpy() - state[k]).max() <= 1e-4


@pytest.mark.parametrize("mode", ["stream", "resident"])
@pytest.mark.parametrize("B,L,F,first", [(130, 2, 7, False), (3, 1, 64, True), (129, 3, 3, True)])
def test_feature_layout_call_is_bit_identical_to_serialised_call(B, L, F, first, mode, monkeypatch):
    """ntm_b200_forward_seq_features (delimiter / target channels synthesised inside the library; in streaming mode
    by the input projection's pack pass, without ever materialising the [B, T, 514] rows) against the serialiser
    followed by the ordinary call: logits, outputs and the final state must agree bit for bit, in both layouts
    (direct_offset_output.py:439-500, test_tracker.py:385-404) and both execution modes."""
    from ntm_tracker_b200 import LoopNTMTracker
    from ntm_tracker_b200.serialize import tracker_inputs
    monkeypatch.setenv("NTM_B200_MODE", mode)
    Cch, T = 512, L * (F + 1)
    rng = np.random.RandomState(B + 7 * L + F)
    feat = torch.from_numpy(np.maximum(rng.standard_normal((B, L, F, Cch)), 0).astype(np.float32)).cuda()
    tgt = torch.from_numpy((rng.rand(B, F) < 0.3).astype(np.float32)).cuda()
    torch.manual_seed(5)
    trk = LoopNTMTracker(T, 2, mem_size=128, mem_dim=512, controller_hidden_size=200, controller_num_layers=1,
                         write_head_size=1, read_head_size=4)
    trk.cell.build(Cch + 2)
    out_a, log_a = trk(tracker_inputs(feat, tgt, delimiter_first=first))
    st_a = {k: v.clone() for k, v in trk.final_state.items()}
    out_b, log_b = trk.call_features(feat, tgt, delimiter_first=first)
    trk.cell.finish()
    assert torch.equal(log_a, log_b) and torch.equal(out_a, out_b)
    for k in st_a:
        assert torch.equal(st_a[k], trk.final_state[k]), k
    assert torch.isfinite(log_b).all()
