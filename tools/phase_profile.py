#!/usr/bin/env python
"""Per-phase time of the persistent kernel (in-kernel SM-clock counters, see
ntm_b200_phase_cycles).  Usage: python tools/phase_profile.py [workload] [B] [T]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as entry

entry.build()
from bench import make_inputs_torch  # noqa: E402
from ntm_tracker_b200 import LoopNTMTracker, _cabi  # noqa: E402
from oracle import ntm_oracle as O  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_tracker"
kw, B, T = O.CONFIGS[wl]
if len(sys.argv) > 2:
    B = int(sys.argv[2])
if len(sys.argv) > 3:
    T = int(sys.argv[3])
dev = torch.device("cuda", 0)
cell_kw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
torch.manual_seed(0)
trk = LoopNTMTracker(T, kw["output_dim"], (-0.05, 0.05), device=dev, **cell_kw)
trk.cell.build(kw["input_dim"], (-0.05, 0.05))
state = trk.cell.zero_state(B, (-0.05, 0.05))
x = make_inputs_torch("c1_copy" if wl == "c1_copy" else "tracker", B, T, kw["input_dim"], 1).to(dev)
lib = _cabi.load()
lib.ntm_b200_set_profiling(1)
for _ in range(3):
    trk(x, state)
trk.cell.finish()
plan = trk.cell.plan(B, T)
info = _cabi.last_launch_info()
ncta = info["ctas"]
buf = (C.c_int64 * (16 * ncta))()
lib.ntm_b200_phase_cycles(trk.cell._last_ws.data_ptr(), buf, ncta)
a, b = C.c_float(), C.c_float()
lib.ntm_b200_last_kernel_ms(C.byref(a), C.byref(b))
cyc = np.array(buf, dtype=np.int64).reshape(ncta, 16)
waves = -(-B // info["sequences_resident"])
steps = waves * T
names = ["A gemm", "A barrier", "B lstm", "B barrier", "C gemm", "C barrier", "D.1a pass1 compute",
         "D barrier", "prologue", "epilogue", "D.0b activations", "D.1b csync+gather",
         "D.2 addressing", "D.3 pass2+csync", "D.4 finalize", "D.0a param loads"]
mhz = 1965.0
res = {"workload": wl, "B": B, "T": T, "ncta": ncta, "waves": waves, "launch": info, "seq_kernel_ms": b.value,
       "xproj_ms": a.value, "us_per_step": b.value * 1e3 / steps, "phases_us_per_step": {}}
for i, n in enumerate(names):
    per = cyc[:, i] / mhz / (waves if i in (8, 9) else steps)
    res["phases_us_per_step"][n] = {"mean": round(float(per.mean()), 3), "max": round(float(per.max()), 3),
                                    "min": round(float(per.min()), 3)}
print(json.dumps(res, indent=1))
