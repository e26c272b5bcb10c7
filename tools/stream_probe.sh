#!/bin/bash
# usage: tools/stream_probe.sh <tag> [env assignments...] -- short device-resident bench of the streaming mode
tag=$1; shift
env NTM_B200_MODE=stream "$@" timeout 300 python bench.py $BENCH_ARGS --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-serve > gpurun_out/probe_$tag.json 2> gpurun_out/probe_$tag.err || tail -5 gpurun_out/probe_$tag.err
python - <<PY
import json
d = json.load(open("gpurun_out/probe_$tag.json"))
r = d["roofline"]
print("$tag", "value=%.3fM" % (d["value"]/1e6), "ms/step=%.2f" % d["ms_per_step"], "mem_us=%.1f" % (r.get("kernel_ms",0)*1e3), "ctrl_ms=%.2f head_ms=%.2f" % (r.get("controller_gemm_lstm_ms_per_step",0), r.get("head_param_gemm_ms_per_step",0)), "frac=%.3f" % r["frac"], "occ=%s" % r.get("ctas_per_sm"))
PY
python - <<PY
import json
d = json.load(open("gpurun_out/probe_$tag.json"))
ph = d["roofline"].get("mem_kernel_phase_ns")
if ph: print("   phases(ns):", {k: round(v) for k, v in ph.items()})
PY
