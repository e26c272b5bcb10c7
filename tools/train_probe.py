import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import make_inputs_torch
from ntm_tracker_b200 import LoopNTMTracker, NTMTrainer
from oracle import ntm_oracle as O
kw, B, T = O.CONFIGS["c5_train"]
ckw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
dev = torch.device("cuda", 0)
trk = LoopNTMTracker(T, 2, (-0.05, 0.05), device=dev, **ckw); trk.cell.build(514, (-0.05, 0.05))
tr = NTMTrainer(trk, frame=8)
x = make_inputs_torch("tracker", B, T, 514, 1).to(dev)
tg = (torch.rand(B, 3, 2) - 0.5).to(dev)
for _ in range(3): tr.train_step(x, tg)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): tr.train_step(x, tg)
torch.cuda.synchronize(); print("train_step ms %.2f" % ((time.perf_counter() - t0) / 3 * 1e3))
st = trk.cell.zero_state(B, (-0.05, 0.05))
t0 = time.perf_counter()
for _ in range(3): trk(x, st)
torch.cuda.synchronize(); print("forward only ms %.2f" % ((time.perf_counter() - t0) / 3 * 1e3))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_step(x, tg); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
