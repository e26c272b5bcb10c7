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


def cpu_reference_run(cfg_name, steps, warmup, threads=None):
    """Time the op-for-op torch-CPU restatement of the reference TF graph on a
    bounded sample of the workload.  Returns (seq-steps/s, cores, sample text, ms/step)."""
    import numpy as np
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
    cfg_name, scaling = WORKLOADS[args.workload]
    from oracle import ntm_oracle as O
    kw, B, T = O.CONFIGS[cfg_name]
    val, cores, sample, ms = cpu_reference_run(cfg_name, max(args.steps, 1), max(args.warmup, 1))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3_sweep", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    from ntm_tracker_b200 import LoopNTMTracker, _cabi
    from ntm_tracker_b200.sharding import max_over_ranks, shard_range
    from oracle import ntm_oracle as O    # shapes / config table only (no oracle compute here)
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg_name, scaling = WORKLOADS[args.workload]
    kw, B, T = O.CONFIGS[cfg_name]
    if args.batch:
        B = args.batch
    if args.seq_len:
        T = args.seq_len
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
    training = args.workload == "c5_train"
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

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    last_loss = [None]

    def one_step(x=None):
        if training:      # forward (with history) + backward + gradient all-reduce + clip + RMSProp
            last_loss[0], _ = trainer.train_step(x_dev if x is None else x, targets, sync=False)
        else:
            trk(x_dev if x is None else x, state)

    for _ in range(max(args.warmup, 3)):
        one_step()
    trk.cell.finish()

    # ---------------- device-resident timing: K steps, CUDA events, max over ranks ---------
    # Two back-to-back timed regions of K steps each.  Region 1 is the one `value` comes from: nothing but
    # the hot path between the step events.  Region 2 repeats the same K steps with the library's per-kernel
    # CUDA events switched on (recorded on the launching stream, between the kernels of every step); the
    # roofline's kernel duration comes from there.  They are separate because those ~4 extra event records
    # per timestep cost the streaming mode up to 10 % at small per-GPU batches (512 sequences), and the
    # headline must not pay for its own instrumentation; both step times are reported.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def timed_region(profiled):
        lib.ntm_b200_set_profiling(1 if profiled else 0)
        one_step()                              # settle (event creation, workspace) outside the region
        trk.cell.finish()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        k_seq, k_xp, k_stream = [], [], []
        barrier()
        l0 = lib.ntm_b200_launch_count()
        t0 = time.perf_counter()
        for i in range(args.steps):
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
        barrier()
        wall_s = time.perf_counter() - t0
        ms = [e0.elapsed_time(e1) for e0, e1 in evs]
        return ms, wall_s, k_seq, k_xp, k_stream, lib.ntm_b200_launch_count() - l0

    step_ms, wall, _, _, _, launches = timed_region(False)
    prof_step_ms, _, seq_ms, xp_ms, stream_ms, _ = timed_region(True)
    trk.cell.finish()
    total_ms = max_over_ranks(sum(step_ms), dev)
    value = B_total * T * args.steps / (total_ms / 1e3)

    # ---------------- end-to-end: pinned host inputs in, host results out ------------------
    e2e = None
    if not args.no_e2e:
        d2h = 4
        for _ in range(2):
            one_step(x_host) if training else trk(x_host, state)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            if training:
                one_step(x_host)                  # H2D copy of the frames, train step, D2H of the loss
                float(last_loss[0])
            else:
                out_h, log_h = trk(x_host, state)     # H2D copy, kernels, D2H of outputs + logits
                d2h = int(out_h.numel() + log_h.numel()) * 4
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
        e2e = {"value": B_total * T * args.steps / e2e_s, "unit": "seq-steps/s",
               "h2d_bytes_per_step": int(input_bytes) * world, "d2h_bytes_per_step": d2h * world}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel (the persistent sequence kernel) ------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    abytes = algorithmic_bytes_per_seqstep(kw)
    if training:    # + HBM spill of the history (write forward, read backward), SURVEY.md s8d
        abytes += 2 * (kw["mem_size"] * kw["mem_dim"] + (kw["read_head_size"] + kw["write_head_size"]) * kw["mem_size"]) * 4
    seq_avg_ms = max(sum(seq_ms) / len(seq_ms), 1e-9)
    achieved = abytes * B_local * T / (seq_avg_ms / 1e3) / 1e9
    info = _cabi.last_launch_info()
    streaming = bool(info.get("streaming")) and stream_ms and stream_ms[-1]["steps"] > 0 and not training
    sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    smem_peak = 128.0 * 148 * sm_mhz * 1e6 / 1e9          # 128 B/clk/SM at the sampled SM clock
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass
    roofline = {
        "kernel": ("stream_forward (mem_step_tma_kernel + gemm_ws_kernel + lstm_stream_kernel, whole forward)"
                   if info.get("streaming") else "ntm_seq_kernel"), "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
        "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_seq_step": abytes, "kernel_ms": seq_avg_ms,
        "kernel_share_of_step": seq_avg_ms / (sum(prof_step_ms) / len(prof_step_ms)),
        "profiled_ms_per_step": sum(prof_step_ms) / len(prof_step_ms),
        "xproj_ms": sum(xp_ms) / len(xp_ms),
        "note": "state is shared-memory resident, so the level that actually bounds the fused step is "
                "SMEM/FP32, not HBM: see smem_*",
        "smem_peak_gbs": smem_peak, "smem_frac": achieved / smem_peak,
    }

    if streaming:
        # streaming mode: the dominant kernel is the fused addressing/memory kernel, launched once per
        # timestep for the rank's B_local sequences; it IS bound by HBM (M is read and written per step)
        n = len(stream_ms)
        mem_launch_ms = max(sum(m["memory"] for m in stream_ms) / n / T, 1e-9)
        achieved = abytes * B_local / (mem_launch_ms / 1e3) / 1e9
        step_avg = sum(prof_step_ms) / len(prof_step_ms)
        roofline = {
            "kernel": "mem_step_tma_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
            "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_seq_step": abytes, "units_per_launch": B_local,
            "kernel_ms": mem_launch_ms, "launches_per_step": T, "ctas_per_sm": info.get("ctas_per_sm"),
            "kernel_share_of_step": mem_launch_ms * T / step_avg, "profiled_ms_per_step": step_avg,
            "controller_gemm_lstm_ms_per_step": sum(m["controller"] for m in stream_ms) / n,
            "head_param_gemm_ms_per_step": sum(m["head_params"] for m in stream_ms) / n,
            "memory_kernel_ms_per_step": mem_launch_ms * T,
            "init_ms_per_step": sum(m["init"] for m in stream_ms) / n,
            "xproj_ms": sum(xp_ms) / len(xp_ms),
            "mem_kernel_phase_ns": _cabi.stream_phase_ns(),
            "note": "streaming mode: memory streamed from HBM once per sequence-step (second pass from L2); "
                    "algorithmic bytes count 3 passes, so frac can exceed what the DRAM traffic alone implies",
        }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, _ = cpu_reference_run(cfg_name, 5, 2)
        cpu = {"value": v, "unit": "seq-steps/s", "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": "sequence_timesteps_per_s", "value": value, "unit": "seq-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload=args.workload, batch=B_total, batch_per_gpu=B_local, T=T,
                       parallelism="dp%d (sequences sharded, no per-step collective)" % world,
                       l2=l2_note, cluster_size=plan["cluster_size"],
                       sequences_resident=plan["sequences_resident"],
                       smem_bytes_per_cta=plan["smem_bytes_per_cta"],
                       mode=("train: fwd+bwd+allreduce+clip+RMSProp, frame=%d" % TRAIN_FRAME) if training else "forward",
                       execution=("streaming (lockstep over the shard, memory in HBM)" if info.get("streaming")
                                  else "resident (persistent kernel, memory in shared memory)"),
                       **kw),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "wall_s_timed_region": wall,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
