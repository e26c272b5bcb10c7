"""GPU: the training path (forward with history -> hand-written memory-backward kernel + library
GEMMs -> gradients) against PyTorch autograd through the fp64 op-for-op restatement of the
reference graph (oracle/ntm_ref_torch.py), for the reference's loss
(direct_offset_output.py:581-606: l2_loss of tanh(logits at delimiter steps) vs offsets).
Mirrors dnc/access_test.py's gradient checks (there: tf.test.compute_gradient_error)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ntm_oracle as O  # noqa: E402
from oracle.ntm_ref_torch import TorchRefNTM  # noqa: E402

pytestmark = pytest.mark.gpu
GRAD_TOL = 5e-4      # of the largest entry of each variable's gradient (fp32 kernels vs fp64 autograd)


@pytest.fixture(autouse=True, params=["resident", "stream"])
def forward_mode(request):
    """Every training test runs with both forward implementations recording the history: the persistent
    resident kernel and the streaming mode (where pass 2 of step t writes straight into slot t+1 of the
    recorded memories); shapes the streaming kernels do not cover fall back to the resident kernel."""
    import os
    os.environ["NTM_B200_MODE"] = request.param
    yield request.param
    os.environ.pop("NTM_B200_MODE", None)


def reference_grads(s, params, x, gather, targets):
    ref = TorchRefNTM(s, params, dtype=torch.float64, requires_grad=True)
    _, logits, _ = ref.run(torch.from_numpy(x))
    y = torch.tanh(logits[:, gather])
    loss = 0.5 * torch.sum((y - torch.from_numpy(targets).double()) ** 2)
    loss.backward()
    return float(loss), {k: v.grad.numpy() for k, v in ref.p.items()}


def kwargs_of(s):
    return dict(mem_size=s.mem_size, mem_dim=s.mem_dim, shift_range=s.shift_range,
                controller_hidden_size=s.controller_hidden_size,
                controller_num_layers=s.controller_num_layers,
                write_head_size=s.write_head_size, read_head_size=s.read_head_size,
                write_first=s.write_first)


CASES = [
    # N, M, R, W, C, L, D, O, shift_range, write_first, B, T, frame
    (16, 8, 2, 1, 12, 1, 5, 2, 1, False, 3, 6, 2),
    (24, 12, 3, 2, 10, 2, 7, 3, 2, True, 2, 6, 3),
    (32, 20, 1, 1, 16, 1, 4, 2, 1, False, 4, 8, 2),
    (33, 21, 4, 3, 9, 2, 3, 2, 1, True, 2, 4, 2),
    (128, 64, 4, 1, 24, 1, 10, 2, 1, False, 3, 6, 2),
]


@pytest.mark.parametrize("N,M,R,W,C,L,D,Odim,sr,wf,B,T,frame", CASES)
def test_gradients_match_autograd(N, M, R, W, C, L, D, Odim, sr, wf, B, T, frame):
    from ntm_tracker_b200 import LoopNTMTracker, NTMTrainer
    from ntm_tracker_b200.training import delimiter_steps
    s = O.NTMShape(output_dim=Odim, input_dim=D, mem_size=N, mem_dim=M, shift_range=sr,
                   controller_hidden_size=C, controller_num_layers=L, write_head_size=W,
                   read_head_size=R, write_first=wf)
    params = O.init_params(s, 100 + N, 0.3, random_biases=True)      # larger scale: gradients well away from 0
    rng = np.random.RandomState(7 + M)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    gather = delimiter_steps(T, frame)
    assert len(gather) >= 1
    targets = rng.uniform(-0.5, 0.5, (B, len(gather), Odim)).astype(np.float32)
    ref_loss, ref = reference_grads(s, params, x, gather, targets)

    trk = LoopNTMTracker(T, Odim, **kwargs_of(s))
    trk.cell.load_reference_weights(params)
    trainer = NTMTrainer(trk, frame=frame)
    loss, grads = trainer.loss_and_grads(torch.from_numpy(x).cuda(), torch.from_numpy(targets).cuda())
    trk.cell.finish()
    assert abs(float(loss) - ref_loss) <= 1e-4 * max(1.0, abs(ref_loss))
    assert set(grads) == set(ref)
    for name, g in ref.items():
        got = grads[name].detach().cpu().numpy()
        scale = max(1e-3, np.abs(g).max())
        err = np.abs(got - g).max() / scale
        assert err <= GRAD_TOL, "%s: rel-to-max error %.3e (max |g| %.3e)" % (name, err, np.abs(g).max())


def test_gradients_match_autograd_at_c5_shape():
    """BASELINE configs[4] at its own shape: N128 M512 4R+1W LSTM-200 D514, T = 32 (B = 4 sequences: they are
    independent, the batch only adds terms to the weight-gradient sums), reference initialisation scale
    (direct_offset_output.py:42), tracker-style inputs with a frame of 8 rows -> loss on steps 15, 23, 31.
    fp64 autograd through oracle/ntm_ref_torch.py; tolerance 5e-4 of each variable's largest entry."""
    from ntm_tracker_b200 import LoopNTMTracker, NTMTrainer
    from ntm_tracker_b200.training import delimiter_steps
    kw, _, T = O.CONFIGS["c5_train"]
    s = O.NTMShape(**kw)
    B, frame = 4, 8
    params = O.init_params(s, 1234, 0.05, random_biases=True)
    x = O.tracker_inputs(B, T, 77, feat=s.input_dim - 2, frame=frame)
    gather = delimiter_steps(T, frame)
    assert gather == [15, 23, 31]
    targets = np.random.RandomState(78).uniform(-0.5, 0.5, (B, len(gather), s.output_dim)).astype(np.float32)
    ref_loss, ref = reference_grads(s, params, x, gather, targets)
    trk = LoopNTMTracker(T, s.output_dim, **kwargs_of(s))
    trk.cell.load_reference_weights(params)
    trainer = NTMTrainer(trk, frame=frame)
    loss, grads = trainer.loss_and_grads(torch.from_numpy(x).cuda(), torch.from_numpy(targets).cuda())
    trk.cell.finish()
    assert abs(float(loss) - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    assert set(grads) == set(ref)
    worst = {}
    for name, g in ref.items():
        got = grads[name].detach().cpu().numpy()
        worst[name.split("/", 1)[1]] = float(np.abs(got - g).max() / max(1e-7, np.abs(g).max()))
    print("C5-shape gradient parity, rel-to-max error per variable:", worst)
    assert max(worst.values()) <= GRAD_TOL, worst


def test_train_step_reduces_loss_and_matches_reference_optimizer():
    """Three clip + RMSProp steps (tf.clip_by_global_norm(5), RMSProp(1e-4, 0.95, 0.9), slots
    rms=1 / momentum=0) against the same update done on the autograd gradients."""
    from ntm_tracker_b200 import LoopNTMTracker, NTMTrainer
    from ntm_tracker_b200.training import delimiter_steps
    s = O.NTMShape(output_dim=2, input_dim=6, mem_size=16, mem_dim=8, controller_hidden_size=10,
                   controller_num_layers=1, write_head_size=1, read_head_size=2)
    params = O.init_params(s, 5, 0.3, random_biases=True)
    rng = np.random.RandomState(3)
    B, T, frame = 4, 6, 2
    x = rng.standard_normal((B, T, 6)).astype(np.float32)
    gather = delimiter_steps(T, frame)
    targets = rng.uniform(-0.5, 0.5, (B, len(gather), 2)).astype(np.float32)
    trk = LoopNTMTracker(T, 2, **kwargs_of(s))
    trk.cell.load_reference_weights(params)
    trainer = NTMTrainer(trk, learning_rate=1e-2, frame=frame)
    cur = {k: v.astype(np.float64) for k, v in params.items()}
    rms = {k: np.ones_like(v) for k, v in cur.items()}
    mom = {k: np.zeros_like(v) for k, v in cur.items()}
    losses = []
    for it in range(3):
        ref_loss, g = reference_grads(s, {k: v.astype(np.float32) for k, v in cur.items()}, x, gather, targets)
        gn = np.sqrt(sum((v ** 2).sum() for v in g.values()))
        scale = 5.0 / max(gn, 5.0)
        for k in cur:
            gk = g[k] * scale
            rms[k] = 0.95 * rms[k] + 0.05 * gk * gk
            mom[k] = 0.9 * mom[k] + 1e-2 * gk / np.sqrt(rms[k] + 1e-10)
            cur[k] = cur[k] - mom[k]
        loss, gnorm = trainer.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(targets).cuda())
        losses.append(float(loss))
        assert abs(float(loss) - ref_loss) <= 1e-3 * max(1.0, abs(ref_loss))
        assert abs(gnorm - gn) <= 2e-3 * max(1.0, gn)
    trk.cell.finish()
    for k, v in cur.items():
        got = trk.cell.variables[k].detach().cpu().numpy()
        assert np.abs(got - v).max() <= 2e-4, k
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("nrows,ncols,K,ks", [(256, 200, 3616, 15), (256, 2248, 800, 1), (70, 130, 100, 2),
                                              (300, 800, 2048, 4), (5, 7, 9, 1)])
def test_tile_gemm_matches_fp64_matmul(nrows, ncols, K, ks):
    """The tcgen05 tile-record GEMM of the backward pass on its own (out = A @ B^T): pack kernels, K slices,
    ragged edges, against a float64 matmul.  Budget: the 3-term bf16 split keeps ~2^-16 of each product."""
    import ctypes as C
    from ntm_tracker_b200 import _cabi
    lib = _cabi.load()
    g = torch.Generator().manual_seed(nrows * 7 + K)
    a = torch.randn(nrows, K + 3, generator=g).cuda()            # padded leading dimensions
    b = torch.randn(ncols, K + 5, generator=g).cuda()
    out = torch.full((nrows, ncols), float("nan")).cuda()
    need = lib.ntm_b200_gemm_nt_workspace_bytes(nrows, ncols, K, ks)
    ws = torch.empty(need, dtype=torch.uint8).cuda()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(lib.ntm_b200_gemm_nt(a.data_ptr(), K + 3, b.data_ptr(), K + 5, out.data_ptr(), ncols, nrows, ncols, K,
                                     ks, ws.data_ptr(), need, stream), "gemm_nt")
    torch.cuda.synchronize()
    ref = a[:, :K].double() @ b[:, :K].double().t()
    err = float((out.double() - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= 2e-5 * scale, (err, scale)
