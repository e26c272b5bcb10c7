"""A/B timing of one workload under several NTM_B200_EXP settings in ONE process (the library reads the
environment switches once per C-ABI call).  usage: ab_probe.py <workload> <host|dev> <exp> [<exp> ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_inputs_torch
from ntm_tracker_b200 import LoopNTMTracker
from oracle import ntm_oracle as O
wl, where = sys.argv[1], sys.argv[2]
exps = sys.argv[3:] or ["0"]
kw, B, T = O.CONFIGS[wl]
ckw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
dev = torch.device("cuda", 0)
trk = LoopNTMTracker(T, kw["output_dim"], (-0.05, 0.05), device=dev, **ckw); trk.cell.build(kw["input_dim"], (-0.05, 0.05))
state = trk.cell.zero_state(B, (-0.05, 0.05))
xh = make_inputs_torch("tracker", B, T, kw["input_dim"], 1).pin_memory()
x = xh if where == "host" else xh.to(dev)
def timeit(n=4):
    trk(x, state); trk(x, state); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): trk(x, state)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
smi = None
if os.environ.get("AB_SMI"):     # the bench's clock sampler, to see what it costs
    import subprocess
    smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-lms", "100"],
                           stdout=subprocess.DEVNULL)
if os.environ.get("AB_CUTS"):      # first block boundaries of the host pipeline, e.g. AB_CUTS=1,4,12
    LoopNTMTracker.first_cuts = tuple(int(v) for v in os.environ["AB_CUTS"].split(","))
for rep in range(2):
    for e in exps:
        os.environ["NTM_B200_EXP"] = e
        ms = timeit()
        print("%s %s EXP=%s: %.2f ms/call = %.3f M seq-steps/s" % (wl, where, e, ms, B * T / ms / 1e3), flush=True)
if smi is not None:
    smi.terminate()
