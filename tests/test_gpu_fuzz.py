"""GPU: random-shape parity (a fixed-seed subset of tools/fuzz_parity.py): shapes the enumerated tests do not list,
streaming and resident mode against the fp64 NumPy oracle, max-abs 1e-4 on logits / w / read / M."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import ntm_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _shapes(n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for ci in range(n):
        N = int(rng.choice([64, 128, 256, 512, 1024]))
        M = int(rng.choice([64, 128, 256, 512]))
        if N == 1024:
            M = min(M, 256)
        out.append(dict(N=N, M=M, R=int(rng.randint(1, 5)), W=int(rng.randint(1, 4)), sr=int(rng.randint(0, 4)),
                        C=int(rng.choice([24, 40, 104])), wf=bool(rng.randint(0, 2)), D=int(rng.choice([10, 66, 130, 514])),
                        B=int(rng.choice([3, 130])), seed=1000 + ci))
    return out


@pytest.mark.parametrize("mode", ["stream", "resident"])
@pytest.mark.parametrize("cfg", _shapes(8, 7), ids=lambda c: "N%d_M%d_R%dW%d_S%d_C%d_D%d_B%d%s" % (
    c["N"], c["M"], c["R"], c["W"], 2 * c["sr"] + 1, c["C"], c["D"], c["B"], "_wf" if c["wf"] else ""))
def test_random_shape_matches_oracle(cfg, mode, monkeypatch):
    from ntm_tracker_b200 import LoopNTMTracker
    monkeypatch.setenv("NTM_B200_MODE", mode)
    s = O.NTMShape(output_dim=2, input_dim=cfg["D"], mem_size=cfg["N"], mem_dim=cfg["M"], shift_range=cfg["sr"],
                   controller_hidden_size=cfg["C"], controller_num_layers=1, write_head_size=cfg["W"],
                   read_head_size=cfg["R"], write_first=cfg["wf"])
    params = O.init_params(s, cfg["seed"], 0.2, random_biases=True)
    T = 3
    x = np.random.RandomState(cfg["seed"]).standard_normal((cfg["B"], T, cfg["D"])).astype(np.float32)
    trk = LoopNTMTracker(T, 2, mem_size=cfg["N"], mem_dim=cfg["M"], shift_range=cfg["sr"], controller_hidden_size=cfg["C"],
                         controller_num_layers=1, write_head_size=cfg["W"], read_head_size=cfg["R"], write_first=cfg["wf"])
    trk.cell.load_reference_weights(params)
    out, lg = trk(torch.from_numpy(x).cuda())
    trk.cell.finish()
    _, rl, rst = O.run_sequence(params, s, x)
    st = trk.final_state
    err = max(float(np.abs(lg.cpu().numpy() - rl).max()), float(np.abs(st["w"].cpu().numpy() - rst["w"]).max()),
              float(np.abs(st["read"].cpu().numpy() - rst["read"]).max()), float(np.abs(st["M"].cpu().numpy() - rst["M"]).max()))
    assert np.isfinite(err) and err <= TOL, err
