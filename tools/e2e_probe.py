import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs_torch
from ntm_tracker_b200 import LoopNTMTracker
from oracle import ntm_oracle as O
kw, B, T = O.CONFIGS["c3_sweep"]
ckw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
dev = torch.device("cuda", 0)
trk = LoopNTMTracker(T, 2, (-0.05, 0.05), device=dev, **ckw); trk.cell.build(514, (-0.05, 0.05))
state = trk.cell.zero_state(B, (-0.05, 0.05))
xh = make_inputs_torch("tracker", B, T, 514, 1).pin_memory()
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for chunks in (1, 2, 4, 8):
    trk.host_chunks = chunks
    print("host_chunks", chunks, "ms/call %.1f" % timeit(lambda: trk(xh, state)), flush=True)
xd = xh.to(dev)
print("device-resident ms/call %.1f" % timeit(lambda: trk(xd, state)))
print("H2D only ms %.1f" % timeit(lambda: xh.to(dev, non_blocking=True)))
# --- components of the chunked path
n = 4
bounds = [(i * B // n, (i + 1) * B // n) for i in range(n)]
print("slice pinned?", xh[0:1024].is_pinned())
print("4 slice copies ms %.1f" % timeit(lambda: [xh[lo:hi].to(dev, non_blocking=True) for lo, hi in bounds]))
xs = [xd[lo:hi].contiguous() for lo, hi in bounds]
sts = [{k: v[lo:hi] for k, v in state.items()} for lo, hi in bounds]
print("4 chunk runs (device) ms %.1f" % timeit(lambda: [trk.cell._run(x_, s_, T) for x_, s_ in zip(xs, sts)]))
print("1 chunk run (device, B=1024) ms %.1f" % timeit(lambda: trk.cell._run(xs[0], sts[0], T)))
x2 = xd[:1036].contiguous(); s2 = {k: v[:1036] for k, v in state.items()}
print("1 run B=1036 (14 full waves) ms %.1f" % timeit(lambda: trk.cell._run(x2, s2, T)))
