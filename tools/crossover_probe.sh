#!/bin/bash
# resident vs streaming mode at a few batch sizes (C2 shapes, T=16): where does the automatic choice belong?
for B in "$@"; do
  for mode in resident stream; do
    NTM_B200_MODE=$mode timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --batch $B --seq-len 16 > gpurun_out/x_${mode}_$B.json 2> gpurun_out/x_${mode}_$B.err
    python - <<PY
import json
d = json.load(open("gpurun_out/x_${mode}_$B.json"))
print("B=$B $mode value=%.3fM ms=%.3f" % (d["value"]/1e6, d["ms_per_step"]))
PY
  done
done
