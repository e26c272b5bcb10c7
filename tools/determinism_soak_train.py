"""Soak for the training path: loss and the flat gradient of the same batch, recomputed, compared bit for bit (forward
chain and reverse-time loop both run with programmatic dependent launch).  usage: determinism_soak_train.py [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs_torch, TRAIN_FRAME
from ntm_tracker_b200 import LoopNTMTracker, NTMTrainer
from ntm_tracker_b200.training import delimiter_steps
from oracle import ntm_oracle as O
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kw, B, T = O.CONFIGS["c5_train"]
ckw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
dev = torch.device("cuda", 0)
torch.manual_seed(3)
trk = LoopNTMTracker(T, 2, (-0.05, 0.05), device=dev, **ckw); trk.cell.build(514, (-0.05, 0.05))
tr = NTMTrainer(trk, frame=TRAIN_FRAME)
x = make_inputs_torch("tracker", B, T, 514, 1).to(dev)
tg = (torch.rand(B, len(delimiter_steps(T, TRAIN_FRAME)), 2) - 0.5).to(dev)
ref, bad = None, 0
for i in range(reps):
    loss, _ = tr.loss_and_grads(x, tg)
    cur = (loss.clone(), tr._grad.clone())
    trk.cell.finish()
    if ref is None:
        ref = cur
        assert torch.isfinite(cur[1]).all()
    else:
        bad += 0 if (torch.equal(ref[0], cur[0]) and torch.equal(ref[1], cur[1])) else 1
print("c5_train B=%d T=%d: %d repeats of loss + gradient, %d differ from the first" % (B, T, reps, bad))
sys.exit(1 if bad else 0)
