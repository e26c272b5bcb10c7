"""Soak: the same full-size call repeated, results compared bit for bit (streaming mode with programmatic dependent
launch along the per-timestep chain: a consumer that started reading before its producer finished would show up as
a run-to-run difference).  usage: determinism_soak.py [workload] [batch] [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs_torch
from ntm_tracker_b200 import LoopNTMTracker
from oracle import ntm_oracle as O
wl = sys.argv[1] if len(sys.argv) > 1 else "c3_sweep"
kw, B, T = O.CONFIGS[wl]
if len(sys.argv) > 2 and int(sys.argv[2]) > 0:
    B = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
ckw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
dev = torch.device("cuda", 0)
torch.manual_seed(3)
trk = LoopNTMTracker(T, kw["output_dim"], (-0.05, 0.05), device=dev, **ckw); trk.cell.build(kw["input_dim"], (-0.05, 0.05))
state = trk.cell.zero_state(B, (-0.05, 0.05))
x = make_inputs_torch("tracker", B, T, kw["input_dim"], 1).to(dev)
ref = None
bad = 0
for i in range(reps):
    out, lg = trk(x, state)
    st = trk.final_state
    cur = (lg.clone(), st["M"].clone(), st["w"].clone(), st["read"].clone(), st["controller_state"].clone())
    trk.cell.finish()
    if ref is None:
        ref = cur
        assert torch.isfinite(lg).all()
    else:
        same = all(torch.equal(a, b) for a, b in zip(ref, cur))
        bad += 0 if same else 1
print("%s B=%d T=%d: %d repeats, %d differ from the first" % (wl, B, T, reps, bad))
sys.exit(1 if bad else 0)
