#!/usr/bin/env python
"""bench.py -- sequence-timesteps/s of the NTM-cell hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A "step" is one pass of the hot path (LoopNTMTracker call: T cell steps for the
rank's share of the B sequences) over one batch of synthetic input.  Default
workload: BASELINE.json configs[2], the one the metric ("sequence-timesteps/sec
at 1/2/4/8 B200") is quoted on -- tracker NTM, N=128 M=512 4R+1W LSTM-200, D=514,
B=4096 sequences x T=64, sharded over the ranks with no per-step collective
(strong scaling: total work fixed).  --workload c2_tracker / c1_copy / c4_large
select the other BASELINE configs (B fixed per GPU -> weak scaling).

Prints ONE JSON line on stdout (rank 0).  `value` = whole-job throughput with
inputs resident in HBM; `e2e` = the same metric through the public API with
pinned HOST inputs and host results (H2D and D2H inside the timed region).
`--impl reference` times the CPU restatement of the reference TF graph
(oracle/ntm_ref_torch.py; the reference itself is TF1/Python-2 source and cannot
run here) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (oracle config name, scaling)
    "c3_sweep": ("c3_sweep", "strong"),
    "c2_tracker": ("c2_tracker", "weak"),
    "c1_copy": ("c1_copy", "weak"),
    "c4_large": ("c4_large", "weak"),
    # BASELINE configs[4]: forward + backward + NCCL gradient all-reduce + clip + RMSProp, 256 sequences
    # per GPU, T = 32.  Synthetic frames of 8 rows (7 feature rows + delimiter) -> loss on 3 delimiter steps.
    "c5_train": ("c5_train", "weak"),
}
TRAIN_FRAME = 8
INIT_SCALE = 0.05       # direct_offset_output.py:42
FEATURE_SCALE = 1.0     # synthetic conv4_3 features = max(0, N(0,1)) * FEATURE_SCALE


def algorithmic_bytes_per_seqstep(kw):
    """3*N*M*4 (two reads + one write of the memory) + 6*H*N*4
    (weighting passes) + (D + O)*4 (frame in, logits out)."""
    N, M = kw["mem_size"], kw["mem_dim"]
    H = kw["read_head_size"] + kw["write_head_size"]
    return 3 * N * M * 4 + 6 * H * N * 4 + (kw["input_dim"] + kw["output_dim"]) * 4


def backward_bytes_per_seqstep(kw):
    """Training, memory/addressing backward kernel: read M_prev, read + write dM (3*N*M*4),
    w_prev / recorded similarities in, dw in / out (4*H*N*4), raw head parameters in, d_raw out."""
    N, M = kw["mem_size"], kw["mem_dim"]
    R, W = kw["read_head_size"], kw["write_head_size"]
    H, S = R + W, 2 * kw["shift_range"] + 1
    P = H * M + 3 * H + S * H + 2 * M * W
    return 3 * N * M * 4 + 4 * H * N * 4 + 2 * (P + kw["output_dim"]) * 4


def history_bytes_per_seqstep(kw):
    """Training history spill (write in the forward, read in the backward): M and w entering the step."""
    N, M = kw["mem_size"], kw["mem_dim"]
    H = kw["read_head_size"] + kw["write_head_size"]
    return 2 * (N * M + H * N) * 4


def make_inputs_torch(kind, B, T, D, seed):
    """Same layout as oracle.ntm_oracle.{tracker,copy_task}_inputs, generated with torch (fast)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    if kind == "c1_copy":
        x = torch.zeros(B, T, D)
        length = max(1, (T - 1) // 2)
        x[:, :length, :D - 1] = (torch.rand(B, length, D - 1, generator=g) < 0.5).float()
        if length < T:
            x[:, length, D - 1] = 1.0
        return x
    feat, frame = D - 2, 65
    x = torch.zeros(B, T, D)
    x[:, :, :feat] = torch.randn(B, T, feat, generator=g).clamp_min_(0.0) * FEATURE_SCALE
    t = torch.arange(T)
    delim = (t % frame) == (frame - 1)
    x[:, delim, :feat] = 0.0
    x[:, delim, feat] = 1.0
    first = t < min(frame - 1, T)
    x[:, first, feat + 1] = (torch.rand(B, int(first.sum()), generator=g) < 0.1).float()
    return x


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [v.strip() for v in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def cpu_reference_run(cfg_name, steps, warmup, threads=None, full=False):
    """Time the op-for-op torch-CPU restatement of the reference TF graph on a
    bounded sample of the workload (`full`: ONE pass over every sequence and every step of the workload, in
    batch chunks of 256 so that the per-step M history the reference's TensorArrays keep stays at a few GB).
    Returns (seq-steps/s, cores, sample text, ms/step)."""
    import torch
    from oracle import ntm_oracle as O
    from oracle.ntm_ref_torch import TorchRefNTM
    kw, B, T = O.CONFIGS[cfg_name]
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs, Ts = min(B, 64), min(T, 16)
    s = O.NTMShape(**kw)
    params = O.init_params(s, 1234, INIT_SCALE)
    kind = "c1_copy" if cfg_name == "c1_copy" else "tracker"
    x = make_inputs_torch(kind, Bs, Ts, s.input_dim, 99)
    ref = TorchRefNTM(s, params)
    if full:
        chunk = min(B, 256)
        xs = make_inputs_torch(kind, chunk, T, s.input_dim, 99)
        ref.run(xs[:8, :2])
        t0 = time.perf_counter()
        for _ in range(0, B, chunk):
            ref.run(xs)
        el = time.perf_counter() - t0
        n = (B + chunk - 1) // chunk * chunk
        sample = "%s FULL workload: %d sequences x %d steps, one pass, batch chunks of %d" % (cfg_name, n, T, chunk)
        return n * T / el, cores, sample, el * 1e3
    for _ in range(warmup):
        ref.run(x)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref.run(x)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    sample = "%s shapes, first %d of %d sequences x first %d of %d steps, %d runs, median" % (
        cfg_name, Bs, B, Ts, T, steps)
    return Bs * Ts / med, cores, sample, med * 1e3


_REAL_STDOUT = None


def protect_stdout():
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr, so that
    stdout carries exactly one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def run_reference_impl(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    protect_stdout()
    cfg_name, scaling = WORKLOADS.get(args.workload, WORKLOADS["c2_tracker"])   # "serve" runs the c2_tracker cell
    from oracle import ntm_oracle as O
    kw, B, T = O.CONFIGS[cfg_name]
    val, cores, sample, ms = cpu_reference_run(cfg_name, max(args.steps, 1), max(args.warmup, 1),
                                               full=args.full_reference)
    line = {
        "impl": "reference", "metric": "sequence_timesteps_per_s", "value": val, "unit": "seq-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict(workload=args.workload, batch=B, T=T, **kw),
        "cpu_baseline": {"value": val, "unit": "seq-steps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "seq-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is TF1/Python-2 graph code (not installable here); this is the fp32 "
                "op-for-op torch-CPU restatement oracle/ntm_ref_torch.py on the host cores",
    }
    emit(line)
    return 0


def _peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    return peaks, hbm_peak, peak_src


def _traffic(workload, units_per_launch=None):
    """ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (profiles/traffic.json,
    which names the capture each figure comes from), or None.  The capture is of a launch over
    `<workload>_units_per_launch` sequences; a launch over a different number (a rank's shard) is charged in proportion."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        t = d.get(workload)
        cap = d.get(workload + "_units_per_launch")
        if t is not None and cap and units_per_launch and units_per_launch != cap:
            t = t * float(units_per_launch) / float(cap)
        return t
    except Exception:
        return None


def measure(workload, ctx, steps, warmup, want_e2e=True, want_clocks=True, batch=0, seq_len=0):
    """One workload on this rank's GPU: W warm-up steps, K timed steps (CUDA events, max over ranks), a second
    profiled region for the per-kernel durations, the end-to-end leg.  Returns the JSON line (rank 0) or None."""
    import ctypes as C
    import torch
    from ntm_tracker_b200 import LoopNTMTracker, _cabi
    from ntm_tracker_b200.sharding import max_over_ranks, shard_range
    from oracle import ntm_oracle as O    # shapes / config table only (no oracle compute here)
    world, rank, dev, barrier = ctx["world"], ctx["rank"], ctx["dev"], ctx["barrier"]

    cfg_name, scaling = WORKLOADS[workload]
    kw, B, T = O.CONFIGS[cfg_name]
    if batch:
        B = batch
    if seq_len:
        T = seq_len
    if scaling == "strong":
        B_total = B
        lo, hi = shard_range(B, world, rank)
        B_local = hi - lo
    else:
        B_local = B
        B_total = B * world
    D, Odim = kw["input_dim"], kw["output_dim"]
    cell_kw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}

    torch.manual_seed(1234)      # identical weights on every rank (replicated)
    trk = LoopNTMTracker(T, Odim, (-INIT_SCALE, INIT_SCALE), device=dev, **cell_kw)
    trk.cell.build(D, (-INIT_SCALE, INIT_SCALE))
    state = trk.cell.zero_state(B_local, (-INIT_SCALE, INIT_SCALE))
    training = workload == "c5_train"
    trainer = targets = None
    if training:
        from ntm_tracker_b200 import NTMTrainer
        from ntm_tracker_b200.training import delimiter_steps
        trainer = NTMTrainer(trk, frame=TRAIN_FRAME)
        n_t = len(delimiter_steps(T, TRAIN_FRAME))
        targets = (torch.rand(B_local, n_t, Odim, generator=torch.Generator().manual_seed(7 + rank)) - 0.5).to(dev)
    kind = "c1_copy" if cfg_name == "c1_copy" else "tracker"
    x_host = make_inputs_torch(kind, B_local, T, D, 1000 + rank).pin_memory()
    x_dev = x_host.to(dev)
    lib = _cabi.load()
    lib.ntm_b200_set_profiling(0)
    plan = trk.cell.plan(B_local, T)

    input_bytes = x_dev.numel() * 4
    flush = None
    l2_note = "inputs (%.0f MB/rank) exceed the 126 MB L2" % (input_bytes / 1e6)
    if input_bytes < 256e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        l2_note = "L2 flushed between timed iterations (256 MiB write)"

    last_loss = [None]

    def one_step(x=None):
        if training:      # forward (with history) + backward + gradient all-reduce + clip + RMSProp
            last_loss[0], _ = trainer.train_step(x_dev if x is None else x, targets, sync=False)
        else:
            trk(x_dev if x is None else x, state)

    for _ in range(max(warmup, 3)):
        one_step()
    trk.cell.finish()
    # Python's cyclic collector: a full collection over the interpreter's ~10^6 objects (torch, numpy) costs 20-40 ms
    # -- a whole C3 pass -- and used to land inside a timed step now and then (one 40.7 ms call among 18.4 ms ones on
    # 2 GPUs).  Collect now, park what exists in the permanent generation; nothing is disabled.
    gc.collect()
    gc.freeze()

    # ---------------- device-resident timing: K steps, CUDA events, max over ranks ---------
    # Two back-to-back timed regions of K steps each.  Region 1 is the one `value` comes from: nothing but
    # the hot path between the step events.  Region 2 repeats the same K steps with the library's per-kernel
    # CUDA events switched on (recorded on the launching stream, between the kernels of every step); the
    # roofline's kernel duration comes from there.  They are separate because those ~4 extra event records
    # per timestep cost the streaming mode up to 10 % at small per-GPU batches (512 sequences), and the
    # headline must not pay for its own instrumentation; both step times are reported.
    sampler = ClockSampler(ctx["local_rank"])
    if rank == 0 and want_clocks:
        sampler.start()

    def timed_region(profiled):
        lib.ntm_b200_set_profiling(1 if profiled else 0)
        one_step()                              # settle (event creation, workspace) outside the region
        trk.cell.finish()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        k_seq, k_xp, k_stream, k_bwd = [], [], [], []
        barrier()
        l0 = lib.ntm_b200_launch_count()
        t0 = time.perf_counter()
        for i in range(steps):
            if flush is not None:
                flush.fill_(i & 0xff)
            evs[i][0].record()
            one_step()
            evs[i][1].record()
            evs[i][1].synchronize()
            if profiled:
                a, b = C.c_float(), C.c_float()
                lib.ntm_b200_last_kernel_ms(C.byref(a), C.byref(b))
                k_xp.append(a.value); k_seq.append(b.value)
                k_stream.append(_cabi.last_stream_ms())
                if training:
                    k_bwd.append(_cabi.last_backward_ms())
        barrier()
        wall_s = time.perf_counter() - t0
        ms = [e0.elapsed_time(e1) for e0, e1 in evs]
        return ms, wall_s, k_seq, k_xp, k_stream, lib.ntm_b200_launch_count() - l0, k_bwd

    step_ms, wall, _, _, _, launches, _ = timed_region(False)
    prof_step_ms, _, seq_ms, xp_ms, stream_ms, _, bwd_ms = timed_region(True)
    info = _cabi.last_launch_info()
    phase_ns = None
    if info.get("streaming") and not training:      # in-kernel phase stamps of the last profiled memory-kernel launch
        try:
            phase_ns = _cabi.stream_phase_ns()
        except Exception:
            phase_ns = None
    lib.ntm_b200_set_profiling(0)
    trk.cell.finish()
    total_ms = max_over_ranks(sum(step_ms), dev)
    value = B_total * T * steps / (total_ms / 1e3)

    # ---------------- end-to-end: pinned host inputs in, host results out ------------------
    e2e = None
    if want_e2e:
        d2h = 4
        # warm-up in exactly the shape of the timed loop (results kept across the next call, the loss read back): the
        # second page-locked result buffer and the scalar read-back's staging are first-use allocations
        # (cudaHostAlloc: one 114 ms call among 34 ms ones when they landed inside the timed region)
        for _ in range(3):
            if training:
                one_step(x_host)
                float(last_loss[0])
            else:
                out_h, log_h = trk(x_host, state)
        barrier()
        t0 = time.perf_counter()
        per_call = []
        for i in range(steps):
            tc = time.perf_counter()
            if training:
                one_step(x_host)                  # H2D copy of the frames, train step, D2H of the loss
                float(last_loss[0])
            else:
                out_h, log_h = trk(x_host, state)     # H2D copy, kernels, D2H of outputs + logits
                d2h = int(out_h.numel() + log_h.numel()) * 4
            per_call.append((time.perf_counter() - tc) * 1e3)
        barrier()
        if rank == 0:
            sys.stderr.write("[bench] %s e2e ms per call: %s\n" % (workload, " ".join("%.2f" % v for v in per_call)))
        e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
        e2e = {"value": B_total * T * steps / e2e_s, "unit": "seq-steps/s",
               "h2d_bytes_per_step": int(input_bytes) * world, "d2h_bytes_per_step": d2h * world}
    clocks = sampler.stop() if (rank == 0 and want_clocks) else None
    del trainer, trk, x_dev, state, flush
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---------------- roofline of the dominant kernel ------------------------------------------
    peaks, hbm_peak, peak_src = _peaks()
    NM4 = kw["mem_size"] * kw["mem_dim"] * 4
    abytes = algorithmic_bytes_per_seqstep(kw)
    step_avg = sum(prof_step_ms) / len(prof_step_ms)
    ms_per_step = total_ms / steps
    traffic = _traffic(workload, B_local)
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    smem_peak = 128.0 * 148 * sm_mhz * 1e6 / 1e9          # 128 B/clk/SM at the sampled SM clock
    streaming = bool(info.get("streaming"))
    # streaming design's own ceiling: one HBM read + one HBM write of the memory per sequence-step
    stream_ceiling = hbm_peak * 1e9 / (2 * NM4) * world
    common = {"bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
              "profiled_ms_per_step": step_avg, "xproj_ms": sum(xp_ms) / len(xp_ms)}

    if training:
        # dominant kernel of the training step: the fused memory/addressing backward (one launch per timestep)
        n = max(len(bwd_ms), 1)
        bb = backward_bytes_per_seqstep(kw)
        mem_bwd_ms = sum(m["memory_backward"] for m in bwd_ms) / n if bwd_ms else 0.0
        launch_ms = max(mem_bwd_ms / T, 1e-9)
        achieved = bb * B_local / (launch_ms / 1e3) / 1e9
        step_bytes = abytes + bb + history_bytes_per_seqstep(kw)
        fwd_ms = sum(seq_ms) / len(seq_ms)
        roofline = dict(common, **{
            "kernel": "mem_backward_kernel", "achieved": achieved, "frac": achieved / hbm_peak, "traffic": traffic,
            "frac_dram": (traffic / (launch_ms / 1e3) / 1e9 / hbm_peak) if traffic else None,
            "algorithmic_bytes_per_seq_step": bb, "units_per_launch": B_local, "kernel_ms": launch_ms,
            "launches_per_step": T, "kernel_share_of_step": mem_bwd_ms / step_avg,
            "step_algorithmic_bytes_per_seq_step": step_bytes,
            "frac_step": step_bytes * B_local * T / (ms_per_step / 1e3) / 1e9 / hbm_peak,
            "forward_ms_per_step": fwd_ms, "forward_execution": "streaming" if streaming else "resident",
            "backward_ms_per_step": sum(m["total"] for m in bwd_ms) / n if bwd_ms else None,
            "memory_backward_ms_per_step": mem_bwd_ms,
            "backward_loop_rest_ms_per_step": sum(m["loop_rest"] for m in bwd_ms) / n if bwd_ms else None,
            "weight_grad_ms_per_step": sum(m["weight_grads"] for m in bwd_ms) / n if bwd_ms else None,
            "note": "frac = algorithmic bytes of the memory/addressing backward kernel / its launch time; frac_step = "
                    "(forward + backward + history spill) algorithmic bytes / whole train-step time",
        })
    elif streaming and stream_ms and stream_ms[-1]["steps"] > 0:
        # streaming mode: the dominant kernel is the fused addressing/memory kernel, launched once per
        # timestep for the rank's B_local sequences; it IS bound by HBM (M is read and written per step)
        n = len(stream_ms)
        mem_launch_ms = max(sum(m["memory"] for m in stream_ms) / n / T, 1e-9)
        achieved = abytes * B_local / (mem_launch_ms / 1e3) / 1e9
        roofline = dict(common, **{
            "kernel": "mem_step_tma_kernel" if info.get("ctas_per_sm", 0) else "mem_step_kernel",
            "achieved": achieved, "frac": achieved / hbm_peak, "traffic": traffic,
            # what the DRAM controllers moved (ncu capture named in profiles/traffic.json) over the live launch time
            "frac_dram": (traffic / (mem_launch_ms / 1e3) / 1e9 / hbm_peak) if traffic else None,
            # whole step by algorithmic bytes (GEMMs, LSTM, x-projection included in the time)
            "frac_step": abytes * B_local * T / (ms_per_step / 1e3) / 1e9 / hbm_peak,
            # against what this design could reach at best: 2*N*M*4 bytes of HBM per sequence-step
            "frac_of_streaming_ceiling": value / stream_ceiling,
            "streaming_ceiling_seq_steps_per_s": stream_ceiling,
            "algorithmic_bytes_per_seq_step": abytes, "units_per_launch": B_local,
            "kernel_ms": mem_launch_ms, "launches_per_step": T, "ctas_per_sm": info.get("ctas_per_sm"),
            "kernel_share_of_step": mem_launch_ms * T / step_avg,
            "controller_gemm_lstm_ms_per_step": sum(m["controller"] for m in stream_ms) / n,
            "head_param_gemm_ms_per_step": sum(m["head_params"] for m in stream_ms) / n,
            "memory_kernel_ms_per_step": mem_launch_ms * T,
            "init_ms_per_step": sum(m["init"] for m in stream_ms) / n,
            "mem_kernel_phase_ns": phase_ns,
            "note": "streaming mode: memory streamed from HBM once per sequence-step (second pass from L2); frac counts "
                    "the 3 algorithmic passes of SURVEY s8(d), so it can exceed 1 -- frac_dram is the DRAM-level figure",
        })
    else:
        seq_avg_ms = max(sum(seq_ms) / len(seq_ms), 1e-9)
        achieved = abytes * B_local * T / (seq_avg_ms / 1e3) / 1e9
        roofline = dict(common, **{
            "kernel": "ntm_seq_kernel", "achieved": achieved, "frac": achieved / hbm_peak, "traffic": traffic,
            "frac_dram": (traffic / (seq_avg_ms / 1e3) / 1e9 / hbm_peak) if traffic else None,
            "frac_step": abytes * B_local * T / (ms_per_step / 1e3) / 1e9 / hbm_peak,
            "algorithmic_bytes_per_seq_step": abytes, "kernel_ms": seq_avg_ms,
            "kernel_share_of_step": seq_avg_ms / step_avg,
            "note": "state is shared-memory resident, so the level that actually bounds the fused step is "
                    "SMEM/FP32, not HBM: see smem_*",
            "smem_peak_gbs": smem_peak, "smem_frac": achieved / smem_peak,
        })

    return {
        "metric": "sequence_timesteps_per_s", "value": value, "unit": "seq-steps/s",
        "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload=workload, batch=B_total, batch_per_gpu=B_local, T=T,
                       parallelism=("dp%d (sequences sharded, gradient all-reduce over NCCL)" if training else
                                    "dp%d (sequences sharded, no per-step collective)") % world,
                       l2=l2_note, cluster_size=plan["cluster_size"],
                       sequences_resident=plan["sequences_resident"],
                       smem_bytes_per_cta=plan["smem_bytes_per_cta"],
                       mode=("train: fwd+bwd+allreduce+clip+RMSProp, frame=%d" % TRAIN_FRAME) if training else "forward",
                       execution=("streaming (lockstep over the shard, memory in HBM)" if streaming
                                  else "resident (persistent kernel, memory in shared memory)"),
                       **kw),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": None,
        "wall_s_timed_region": wall,
    }


def measure_serve(dev, with_cpu, frames=30):
    """Serve path (test_tracker.py:284-299,331-342): batch 1, one frame = delimiter row + 64 feature rows = 65 cell
    steps, state carried between frames.  ResidentTracker keeps the state on the device and runs a frame as one
    launch; the reference makes 65 sess.run calls per frame with the whole state fed and fetched as NumPy.  Reports
    microseconds per frame: features already on the device, features from pinned host memory with the offsets read
    back (e2e), and the CPU port stepping the same rows one step per call."""
    import torch
    from ntm_tracker_b200 import NTMCell, ResidentTracker
    from oracle import ntm_oracle as O
    kw, _, _ = O.CONFIGS["c2_tracker"]
    D, Odim, F = kw["input_dim"], kw["output_dim"], 64
    cell_kw = {k: v for k, v in kw.items() if k not in ("input_dim", "output_dim")}
    torch.manual_seed(1234)
    cell = NTMCell(Odim, device=dev, **cell_kw)
    cell.build(D, (-INIT_SCALE, INIT_SCALE))
    cell.zero_state(1, (-INIT_SCALE, INIT_SCALE))
    g = torch.Generator().manual_seed(5)
    feats_h = (torch.randn(frames, 1, F, D - 2, generator=g).clamp_min_(0.0) * FEATURE_SCALE).pin_memory()
    target_h = (torch.rand(1, F, generator=g) < 0.1).float().pin_memory()
    feats_d, target_d = feats_h.to(dev), target_h.to(dev)
    rt = ResidentTracker(cell, F, 1).reset()
    for i in range(5):
        rt.track(feats_d[i], target_d if i == 0 else None)
    cell.finish()
    rt.reset()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
    for i in range(frames):
        evs[i][0].record()
        rt.track(feats_d[i], target_d if i == 0 else None)
        evs[i][1].record()
    torch.cuda.synchronize(dev)
    dev_us = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    rt.reset()
    e2e_us = []
    for i in range(frames):
        t0 = time.perf_counter()
        off = rt.track(feats_h[i].to(dev, non_blocking=True), target_d if i == 0 else None).cpu()
        e2e_us.append((time.perf_counter() - t0) * 1e6)
    cell.finish()
    e2e_us.sort()
    out = {"workload": "serve: batch 1, %d steps per frame (delimiter + %d feature rows), state resident on the device, "
                       "c2_tracker cell" % (F + 1, F),
           "frames": frames, "steps_per_frame": F + 1,
           "us_per_frame_device": dev_us[len(dev_us) // 2], "us_per_frame_device_min": dev_us[0],
           "us_per_frame_e2e": e2e_us[len(e2e_us) // 2],
           "frames_per_s_e2e": 1e6 / e2e_us[len(e2e_us) // 2],
           "h2d_bytes_per_frame": int(feats_h[0].numel()) * 4, "d2h_bytes_per_frame": int(off.numel()) * 4}
    if with_cpu:
        from oracle.ntm_ref_torch import TorchRefNTM
        s = O.NTMShape(**kw)
        params = {k: v.detach().cpu().numpy() for k, v in cell.variables.items()}
        ref = TorchRefNTM(s, params)
        cores = os.cpu_count() or 1
        best = None
        for thr in sorted(set([1, min(cores, 8)])):       # batch-1 steps: a few threads at most help
            torch.set_num_threads(thr)
            st, times = None, []
            for i in range(3):
                rows = torch.zeros(1, F + 1, D)
                rows[:, 0, D - 2] = 1.0                              # delimiter row first (test_tracker.py:400-404)
                rows[:, 1:, :D - 2] = feats_h[i, 0]
                if i == 0:
                    rows[:, 1:, D - 1] = target_h[0]
                t0 = time.perf_counter()
                for t in range(F + 1):                               # one "sess.run" per row, state fed back
                    _, _, st = ref.run(rows[:, t:t + 1], st)
                times.append((time.perf_counter() - t0) * 1e6)
            med = sorted(times)[1]
            if best is None or med < best[0]:
                best = (med, thr)
        torch.set_num_threads(cores)
        out["cpu_port_us_per_frame"] = best[0]
        out["cpu_port_threads"] = best[1]
        out["speedup_e2e_vs_cpu_port"] = best[0] / out["us_per_frame_e2e"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3_sweep", choices=sorted(WORKLOADS) + ["serve"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the c5_train companion measurement of the default line")
    ap.add_argument("--no-serve", action="store_true", help="skip the serve-latency companion measurement")
    ap.add_argument("--full-reference", action="store_true",
                    help="--impl reference: one pass over the FULL workload (all sequences x all steps) instead of the bounded sample")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch (debug)")
    ap.add_argument("--seq-len", type=int, default=0, help="override the workload's T (debug)")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_impl(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    protect_stdout()

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    ctx = {"world": world, "rank": rank, "local_rank": local_rank, "dev": dev, "barrier": barrier}
    want_cpu = not args.no_cpu_baseline

    if args.workload == "serve":
        line = None
        if rank == 0:
            sv = measure_serve(dev, want_cpu)
            steps_s = sv["steps_per_frame"] * 1e6 / sv["us_per_frame_device"]
            line = {"metric": "sequence_timesteps_per_s", "value": steps_s, "unit": "seq-steps/s", "n_gpus": 1,
                    "steps": sv["frames"], "warmup": 5, "ms_per_step": sv["us_per_frame_device"] / 1e3,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic", "config": {"workload": "serve"},
                    "e2e": {"value": sv["steps_per_frame"] * 1e6 / sv["us_per_frame_e2e"], "unit": "seq-steps/s",
                            "h2d_bytes_per_step": sv["h2d_bytes_per_frame"], "d2h_bytes_per_step": sv["d2h_bytes_per_frame"]},
                    "serve": sv}
    else:
        line = measure(args.workload, ctx, args.steps, args.warmup, want_e2e=not args.no_e2e,
                       batch=args.batch, seq_len=args.seq_len)
        # companion numbers on the default line: the training step (BASELINE configs[4]; the path's only collective,
        # the NCCL gradient all-reduce, runs when N > 1) and the serve path's per-frame latency (rank 0)
        if args.workload == "c3_sweep" and not args.no_train and not args.batch and not args.seq_len:
            tl = measure("c5_train", ctx, max(3, min(args.steps, 5)), 3, want_e2e=True, want_clocks=False)
            if line is not None and tl is not None:
                r = tl["roofline"]
                line["train"] = {
                    "workload": "c5_train", "value": tl["value"], "unit": tl["unit"], "ms_per_step": tl["ms_per_step"],
                    "n_gpus": world, "scaling": "weak", "batch": tl["config"]["batch"], "T": tl["config"]["T"],
                    "mode": tl["config"]["mode"], "parallelism": tl["config"]["parallelism"],
                    "e2e": tl["e2e"], "gpu_launches": tl["gpu_launches"],
                    "roofline": {k: r.get(k) for k in (
                        "kernel", "bound", "achieved", "peak", "unit", "frac", "frac_step", "kernel_ms",
                        "kernel_share_of_step", "forward_ms_per_step", "backward_ms_per_step",
                        "memory_backward_ms_per_step", "backward_loop_rest_ms_per_step", "weight_grad_ms_per_step")},
                }
        if args.workload == "c3_sweep" and not args.no_serve and rank == 0 and line is not None:
            line["serve"] = measure_serve(dev, want_cpu)

    # the CPU baseline runs LAST: its 16 busy host threads would otherwise sit next to the GPU legs that follow
    # (a training e2e measured right after it lost 2.7x to the leftover thread pool)
    if rank == 0 and line is not None and want_cpu and world == 1 and args.workload in WORKLOADS:
        v, cores, sample, _ = cpu_reference_run(WORKLOADS[args.workload][0], 5, 2)
        line["cpu_baseline"] = {"value": v, "unit": "seq-steps/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0 and line is not None:
        emit(line)
    if world > 1:
        barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
