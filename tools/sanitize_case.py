"""Small end-to-end invocation for compute-sanitizer (forward both GEMM paths, training backward)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ntm_oracle as O
from ntm_tracker_b200 import LoopNTMTracker, NTMTrainer
from ntm_tracker_b200.training import delimiter_steps
kw = dict(output_dim=2, input_dim=18, mem_size=64, mem_dim=160, shift_range=1, controller_hidden_size=24,
          controller_num_layers=1, write_head_size=1, read_head_size=2)
s = O.NTMShape(**kw)
params = O.init_params(s, 1, 0.05)
ckw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
B, T = 5, 4
x = np.random.RandomState(0).standard_normal((B, T, 18)).astype(np.float32)
for mode in ("tensor", "simt"):
    if mode == "simt": os.environ["NTM_B200_DISABLE_TC"] = "1"
    trk = LoopNTMTracker(T, 2, **ckw); trk.cell.load_reference_weights(params)
    out, lg = trk(torch.from_numpy(x).cuda()); trk.cell.finish()
    _, rl, _ = O.run_sequence(params, s, x)
    print(mode, "max err", float(np.abs(lg.cpu().numpy() - rl).max()))
os.environ.pop("NTM_B200_DISABLE_TC", None)
tr = NTMTrainer(trk, frame=2)
tg = torch.zeros(B, len(delimiter_steps(T, 2)), 2).cuda()
loss, g = tr.loss_and_grads(torch.from_numpy(x).cuda(), tg); trk.cell.finish()
print("train loss", float(loss))
