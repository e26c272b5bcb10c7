"""CPU: the op-for-op torch restatement that bench.py times as the CPU baseline
(oracle/ntm_ref_torch.py) agrees with the NumPy oracle, so the baseline number is
for the same computation."""
import numpy as np
import torch

from oracle import ntm_oracle as O
from oracle.ntm_ref_torch import TorchRefNTM


def test_torch_reference_matches_oracle():
    s = O.NTMShape(output_dim=3, input_dim=9, mem_size=24, mem_dim=12, shift_range=2,
                   controller_hidden_size=14, controller_num_layers=2, write_head_size=2,
                   read_head_size=3)
    params = O.init_params(s, 3, 0.05, random_biases=True)
    x = np.random.RandomState(4).standard_normal((4, 5, 9)).astype(np.float32)
    ro, rl, rs = O.run_sequence(params, s, x, dtype=np.float64)
    out, logits, st = TorchRefNTM(s, params).run(torch.from_numpy(x))
    assert np.abs(logits.numpy() - rl).max() < 1e-5
    assert np.abs(out.numpy() - ro).max() < 1e-5
    for k in ("M", "w", "read", "controller_state"):
        assert np.abs(st[k].numpy() - rs[k]).max() < 1e-5, k


def test_torch_reference_write_first():
    s = O.NTMShape(output_dim=2, input_dim=4, mem_size=16, mem_dim=8, controller_hidden_size=8,
                   controller_num_layers=1, write_head_size=1, read_head_size=1, write_first=True)
    params = O.init_params(s, 9, 0.05)
    x = np.random.RandomState(1).standard_normal((2, 3, 4)).astype(np.float32)
    _, rl, rs = O.run_sequence(params, s, x)
    _, logits, st = TorchRefNTM(s, params).run(torch.from_numpy(x))
    assert np.abs(logits.numpy() - rl).max() < 1e-5
    assert np.abs(st["read"].numpy() - rs["read"]).max() < 1e-5
