#!/usr/bin/env python
"""bench.py — contact windows/s at batch=4096 per B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16x3|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (contact_cnn forward + argmax + contact
bits) over one batch of 4096 synthetic z-scored
150x54 fp32 windows per GPU (BASELINE.json configs[1]).  Weak scaling: every
rank classifies its own 4096 windows; the only collective is the one-time NCCL
broadcast of the packed weights, outside the timed region.

One JSON line on stdout (rank 0):
  value        whole-job windows/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through ContactEngine.classify_host_async() with two batches in flight: pinned HOST
               windows in, host class/bits out, every batch's H2D + D2H inside the timed region
               (e2e.one_batch_at_a_time: the blocking classify_host call)
  roofline     dominant kernel: algorithmic FLOPs / its CUDA-event duration vs measured bf16 peak
               (burst or sustained, chosen from the timed region's length and clocks; both fractions printed)
  cpu_baseline the reference's own contact_cnn (oracle/_ref, byte-compiled from /root/reference; else the oracle
               port) on this box's host cores
  stream       BASELINE configs[2]: dce_stream over a long contiguous synthetic log (kernel-only and, through
               ContactEngine.stream_host, end to end from a pinned HOST log: 216 B per window over PCIe)
  batch_32768_per_gpu   BASELINE configs[3]: 32 768 windows per rank in ONE dce_forward call (262 144 over 8 GPUs)
  h2d          the box's pinned host->device copy rate with all ranks copying at once (the e2e ceiling), and
               classify_host(zero_copy=True) beside the staged default
  latency_b1   BASELINE.json's second headline: p50 per-window microseconds at batch 1
  torch_eager_gpu  the same module's stock PyTorch layers in eager mode on the same GPU (informational)
--impl reference times that CPU path alone, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_WINDOW = 39_368_192          # SURVEY.md §8d / BASELINE.md §3 (algorithmic; split passes count once)
BYTES_PER_WINDOW = 150 * 54 * 4 + 64  # batch mode: window read + logits written
LAYER_FLOP = {                        # per window, BASELINE.md §3
    "conv": 3_110_400 + 3_686_400 + 3_686_400 + 7_372_800,
    "conv1": 3_110_400, "conv2": 3_686_400, "conv3": 3_686_400, "conv4": 7_372_800,
    "fc1": 19_398_656, "fc2": 2_097_152, "fc3": 16_384,
}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def workload_name(batch: int) -> str:
    return (f"batch={batch} synthetic z-scored 150x54 fp32 windows per GPU -> 16 logits + class + 4 contact bits "
            f"(BASELINE configs[1])")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.1] or [r for _, r in self.rows]
        sm, smax, reasons, power = [], [], set(), []
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def reference_forward(params):
    """-> (callable x -> logits, kind, description): the reference's own ``contact_cnn`` from oracle/_ref (the module
    byte-compiled from /root/reference/src/contact_cnn.py by oracle/make_ref.py) when it travelled with the snapshot,
    else the oracle port (the same ATen ops restated)."""
    from oracle import make_ref
    if make_ref.available():
        ref_cnn, _ = make_ref.load()
        model = ref_cnn()
        model.load_state_dict(params)
        model = model.eval()
        return model, "reference", "oracle/_ref contact_cnn (the reference's own module, byte-compiled from /root/reference/src/contact_cnn.py)"
    from oracle import contact_oracle as oracle
    return (lambda x: oracle.forward_torch(params, x)), "port", "oracle.forward_torch (the reference's contact_cnn ops restated)"


def cpu_forward_baseline(batch: int, budget_s: float = 20.0, max_runs: int = 5):
    """The reference's CPU forward on all host cores, on a bounded sample of the workload."""
    import torch
    from deep_contact_estimator_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = synth.make_params(0)
    fwd, kind, what = reference_forward(params)
    x = synth.make_windows(batch, seed=1)
    with torch.no_grad():
        fwd(x[:256])                                          # warm-up
        times, t_begin = [], time.perf_counter()
        while len(times) < max_runs and (time.perf_counter() - t_begin < budget_s or not times):
            t0 = time.perf_counter()
            y = fwd(x)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": batch / med, "unit": "windows/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{len(times)} x forward of {batch} windows through {what}, fp32, torch {torch.__version__} CPU, "
                      f"median {med * 1e3:.1f} ms",
            "ms_per_batch": med * 1e3}, y


def gpu_latency_b1(eng, dev, calls: int = 500):
    """BASELINE.json's second headline: p50 per-window microseconds at batch 1 (latency mode, configs[4]): host wall
    clock from "new data in host memory" to "class + contact bits visible on the host", three forms of the same fused
    kernel (dce_latency.cuh) — the row server (headline), the window server, one launch per step — plus a plain-C
    caller; `*_at_1khz` = one step per millisecond (the GPU idles in between, as in the control loop)."""
    import numpy as np
    import torch
    from deep_contact_estimator_b200 import synth
    run = eng.latency_runner(1)
    run.x_host.copy_(synth.make_windows(1, seed=6))
    for _ in range(20):
        run.step()
    host, gpu = [], []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(calls):
        t0 = time.perf_counter()
        run.step()
        host.append((time.perf_counter() - t0) * 1e6)
    for _ in range(calls):
        a.record(run.stream); run.enqueue(); b.record(run.stream)
        b.synchronize()
        gpu.append(a.elapsed_time(b) * 1e3)
    a.record(run.stream)
    for _ in range(calls):
        run.enqueue()
    b.record(run.stream)
    b.synchronize()
    b2b = a.elapsed_time(b) * 1e3 / calls
    # the same call paced at the 1 kHz of the control loop (BASELINE configs[4]): the GPU idles ~960 us between steps
    paced = []
    t_next = time.perf_counter()
    for _ in range(300):
        t_next += 1e-3
        while time.perf_counter() < t_next:
            pass
        t0 = time.perf_counter()
        run.step()
        paced.append((time.perf_counter() - t0) * 1e6)
    launch_form = {"host_us_p50": float(np.percentile(host, 50)), "host_us_p99": float(np.percentile(host, 99)),
                   "host_us_p50_at_1khz": float(np.percentile(paced, 50)), "host_us_p99_at_1khz": float(np.percentile(paced, 99)),
                   "gpu_us_p50": float(np.percentile(gpu, 50)), "gpu_us_p99": float(np.percentile(gpu, 99)), "gpu_us_back_to_back": b2b,
                   "launches_per_call": run.launches,
                   "what": "LatencyRunner.step(): one launch of the fused latency kernel per step + a stream synchronise; the window "
                           "(32.4 KB) is read from pinned host memory, class + bits written to pinned host memory"}
    bits_launch = run.bits_host.tolist()

    def percentiles(fn, n, pace=0.0):
        t, t_next = [], time.perf_counter()
        for i in range(n):
            if pace:
                t_next += pace
                while time.perf_counter() < t_next:
                    pass
            t0 = time.perf_counter()
            fn(i)
            t.append((time.perf_counter() - t0) * 1e6)
        return float(np.percentile(t, 50)), float(np.percentile(t, 99))

    # the resident servers: no launch, no stream synchronise, no CUDA call per step (dce_latency_server_start / _row_server_start)
    windows = synth.make_windows(64, seed=6)
    wn = [windows[i].numpy() for i in range(64)]
    srv = eng.latency_runner(1, persistent=True, idle_timeout_s=1.0)
    xv = srv.x_host.numpy()
    for i in range(20):
        srv.step()
    s50, s99 = percentiles(lambda i: srv.step(), calls)

    def fresh(i):
        np.copyto(xv[0], wn[i % 64])
        srv.step()
    f50, f99 = percentiles(fresh, calls)
    p50, p99 = percentiles(fresh, 300, pace=1e-3)
    server = {"host_us_p50": f50, "host_us_p99": f99, "host_us_p50_at_1khz": p50, "host_us_p99_at_1khz": p99,
              "host_us_p50_same_window": s50, "host_us_p99_same_window": s99, "device_us": float(srv._c[srv._DEVICE_NS]) * 1e-3,
              "bits_equal_launch_form": bool((np.copyto(xv[0], synth.make_windows(1, seed=6)[0].numpy()), srv.step())[1][1].tolist() == bits_launch),
              "what": "LatencyRunner(persistent=True).step() with a NEW 32.4 KB window written to pinned memory every step: resident "
                      "cooperative kernel, doorbell word in pinned memory, class + bits + step number back in one 16-byte store"}
    srv.close()
    log = synth.make_sensor_log(150 + calls, seed=9).numpy()
    rr = eng.row_runner(idle_timeout_s=1.0)
    for t in range(150):
        rr.push(log[t])
    r50, r99 = percentiles(lambda i: rr.push(log[150 + i]), calls)
    dev_us = rr.device_us
    q50, q99 = percentiles(lambda i: rr.push(log[150 + i % calls]), 300, pace=1e-3)
    rows = {"host_us_p50": r50, "host_us_p99": r99, "host_us_p50_at_1khz": q50, "host_us_p99_at_1khz": q99, "device_us": dev_us,
            "h2d_bytes_per_step": 304, "d2h_bytes_per_step": 16,
            "what": "RowRunner.push(row): ONE NEW 54-float sensor row per step (what a 1 kHz estimator receives); the resident kernel keeps "
                    "the 150-row ring on the device, z-scores the window (utils/data_handler.py:55-56) and classifies it"}
    rr.close()
    return {"host_us_p50": rows["host_us_p50"], "host_us_p99": rows["host_us_p99"],
            "host_us_p50_at_1khz": rows["host_us_p50_at_1khz"], "host_us_p99_at_1khz": rows["host_us_p99_at_1khz"],
            "headline": "row_server", "row_server": rows, "window_server": server, "launch_per_step": launch_form,
            "c_caller": c_caller_latency(), "calls": calls, "bits": bits_launch,
            "floor_us": 43.4e6 / 6.5e12 * 1e6,
            "what": "batch=1 per-step host wall clock, new data every step -> class + 4 contact bits visible on the host.  Headline = the row "
                    "server (one new sensor row per tick); window_server = same with a whole new window per step; launch_per_step = "
                    "round 1's form; c_caller = examples/realtime_step.c (no Python).  floor_us = 43.4 MB of weights / 6.5 TB/s"}


def c_caller_latency():
    """examples/realtime_step.c compiled with gcc and run on the GPU: the same two forms from plain C (random weights),
    so what Python / ctypes add to a step can be read off.  Informational; never fails the bench."""
    try:
        import shutil
        import tempfile
        gcc = shutil.which("gcc")
        pkg = os.path.join(ROOT, "deep_contact_estimator_b200")
        exe = os.path.join(tempfile.mkdtemp(), "realtime_step")
        cmd = [gcc, "-std=c99", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
               os.path.join(ROOT, "examples", "realtime_step.c"), "-L", pkg, "-ldce_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
               "-Wl,-rpath," + pkg, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe]
        subprocess.run(cmd, check=True, capture_output=True, timeout=120)
        out = subprocess.run([exe, "--gpu", "1000"], capture_output=True, text=True, timeout=120).stdout
        import re
        res = {}
        for key, label in (("launch_per_step", "launch per step"), ("window_server", "resident server"), ("row_server", "row server")):
            m = re.search(label + r"\s*: host p50 ([0-9.]+) us\s+p99 ([0-9.]+) us", out)
            if m:
                res[key] = {"host_us_p50": float(m.group(1)), "host_us_p99": float(m.group(2))}
        res["classes_equal"] = "classes equal to the launch-per-step form: yes" in out
        return res
    except Exception as e:                                    # pragma: no cover - informational only
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def cpu_latency_b1(params, calls: int = 100):
    """The reference's forward at batch 1 on the host cores, p50 microseconds per window."""
    import numpy as np
    import torch
    from deep_contact_estimator_b200 import synth
    from oracle import contact_oracle as oracle
    fwd, kind, what = reference_forward(params)
    x = synth.make_windows(1, seed=5)
    t = []
    with torch.no_grad():
        for i in range(calls + 5):
            t0 = time.perf_counter()
            logits = fwd(x)
            oracle.decimal2binary(oracle.argmax_class(logits))
            if i >= 5:
                t.append((time.perf_counter() - t0) * 1e6)
    return {"host_us_p50": float(np.percentile(t, 50)), "host_us_p99": float(np.percentile(t, 99)), "calls": calls,
            "what": f"batch=1 forward + argmax + bits through {what} on the host cores", "kind": kind}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores — oracle/_ref
    (its contact_cnn byte-compiled from /root/reference, which travels with the snapshot) or, without it, the
    oracle port.  Under torchrun rank 0 alone runs and prints; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from deep_contact_estimator_b200 import synth
    from oracle import contact_oracle as oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = synth.make_params(0)
    fwd, kind, what = reference_forward(params)
    batch = args.batch
    x = synth.make_windows(batch, seed=1)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            fwd(x[: min(batch, 512)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            logits = fwd(x)
            cls = oracle.argmax_class(logits)                 # torch.max(output, 1) + decimal2binary: src/inference_one_seq.py:26-27
            oracle.decimal2binary(cls)
        dt = time.perf_counter() - t0
    v = batch * args.steps / dt
    line = {
        "impl": "reference", "metric": "contact windows/sec at batch=4096", "value": v, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(batch), "precision": "fp32 (ATen CPU kernels)",
                   "weights": "seeded random init (synth.make_params(0))",
                   "host": f"{cores} CPU threads, torch {torch.__version__}"},
        "cpu_baseline": {"value": v, "unit": "windows/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{args.steps} steps x {batch} windows through {what}"},
        "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "latency_b1": cpu_latency_b1(params),
    }
    print(json.dumps(line), flush=True)


def dominant_kernel(profile_runs):
    """profile_runs: list of [(name, ms)] per profiled step -> (avg ms per kernel name, launches per step)."""
    agg, count = {}, {}
    for run in profile_runs:
        for name, ms in run:
            agg[name] = agg.get(name, 0.0) + ms
            count[name] = count.get(name, 0) + 1
    n = max(len(profile_runs), 1)
    per_step = {k: v / n for k, v in agg.items()}
    launches = {k: count[k] / n for k in count}
    return per_step, launches


def torch_eager_gpu(dev, x, ours_logits, steps: int = 20):
    """the stock PyTorch layers of the same model (what the reference runs on a GPU: cuDNN / cuBLAS eager
    kernels, PyTorch's default TF32 policy) on the same B200, same batch, CUDA-event timed.  Informational — it is not
    the reference arm (the reference ships no GPU kernel of its own) — and never allowed to fail the bench."""
    try:
        import torch
        import deep_contact_estimator_b200 as dce
        from deep_contact_estimator_b200 import synth
        model = dce.contact_cnn()
        model.load_state_dict(synth.make_params(0))
        model = model.eval().to(dev)
        with torch.no_grad():
            for _ in range(3):
                y = model._forward_torch(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                y = model._forward_torch(x)
                pred = torch.max(y, 1)[1]
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        num = (y - ours_logits).abs().amax(dim=1)
        den = ours_logits.abs().amax(dim=1)
        return {"value": x.shape[0] / (ms * 1e-3), "unit": "windows/s", "ms_per_step": ms, "steps": steps,
                "what": "stock nn.Conv1d / nn.Linear layers of the same module in PyTorch eager mode on this GPU, "
                        "forward + argmax, inputs resident",
                "flags": {"cudnn.allow_tf32": bool(torch.backends.cudnn.allow_tf32),
                          "cuda.matmul.allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
                          "torch": torch.__version__},
                "normwise_diff_vs_this_path": float((num / den).max()),
                "classes_equal_to_this_path": bool((pred == ours_logits.argmax(1)).all())}
    except Exception as e:                                    # pragma: no cover - informational only
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def stream_leg(eng, dev, rank, sync_all, max_over_ranks, steps: int):
    """BASELINE configs[2]: streaming inference over a long contiguous synthetic sensor log (one log per rank, weak
    scaling).  `value`: ONE dce_stream call over the device-resident log (window extraction + z-score + forward +
    argmax + bits, overlapping windows read once).  `e2e`: ContactEngine.stream_host from a PINNED HOST log — the
    216 B/step upload runs on a side stream under the kernels, classes + bits are read back — so, unlike the batch
    e2e (32 400 B per window over PCIe), it is not bound by the wire."""
    import torch
    from deep_contact_estimator_b200 import synth
    log = synth.make_sensor_log(steps, seed=2 + rank)
    n = steps - 149
    pinned = log.pin_memory()
    logd = log.to(dev)
    eng.stream(logd, 0, min(n, 65536))
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, launches = 2, 0
    e0.record()
    for _ in range(reps):
        eng.stream(logd)
        launches += eng.last_launches
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    del logd
    out_b = torch.empty((n, 4), dtype=torch.uint8).pin_memory()
    out_c = torch.empty((n,), dtype=torch.int32).pin_memory()
    eng.stream_host(pinned[: 300000], out_bits_host=out_b[: 300000 - 149], out_cls_host=out_c[: 300000 - 149])      # warm-up: copy stream, staging
    sync_all()
    t0 = time.perf_counter()
    eng.stream_host(pinned, out_bits_host=out_b, out_cls_host=out_c)                 # returns after the D2H read completed
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    return {"steps_per_gpu": steps, "windows_per_gpu": n, "ms": ms, "e2e_ms": e2e_ms, "launches": launches // reps,
            "h2d_bytes": steps * 216, "d2h_bytes": n * 8}


def big_batch_leg(eng, dev, rank, sync_all, max_over_ranks, per_gpu: int = 32768, reps: int = 3):
    """BASELINE configs[3]: 262 144 windows sharded over 8 GPUs = 32 768 windows per rank in ONE dce_forward call (the
    library walks them in internal chunks of 4 096).  Inputs resident in HBM (1.06 GB per rank, far beyond L2), generated
    on the device with seed 1000 + rank."""
    import torch
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.randn(per_gpu, 150, 54, generator=g, device=dev, dtype=torch.float32)
    x = ((x - x.mean(dim=1, keepdim=True)) / x.std(dim=1, keepdim=True)).contiguous()
    eng.classify(x[:8192], want_logits=True)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(reps):
        eng.classify(x, want_logits=True)
        launches += eng.last_launches
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    return {"windows_per_gpu": per_gpu, "ms_per_call": ms, "launches_per_call": launches // reps}


def h2d_leg(dev, pinned, sync_all, max_over_ranks, reps: int = 6):
    """The box's pinned host -> device copy rate with EVERY rank copying at once (plain cudaMemcpyAsync of one 132.7 MB
    batch, CUDA-event timed, max over ranks): the ceiling of the batch-mode e2e figure at this N."""
    import torch
    dst = torch.empty_like(pinned, device=dev)
    dst.copy_(pinned, non_blocking=True)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(pinned, non_blocking=True)
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    return pinned.numel() * 4 / (ms * 1e-3) / 1e9


def kernel_flops(name: str) -> int:
    """Algorithmic FLOPs per window a kernel covers, from its profiler name (dce_forward_profile): a kernel's name
    lists the layers it fuses (`tc_block1` = conv1 + conv2, `tc_block2` = conv3 + conv4, `tc_fc2_fc3` = fc.3 + fc.6)."""
    if "block1" in name:
        return LAYER_FLOP["conv1"] + LAYER_FLOP["conv2"]
    if "block2" in name:
        return LAYER_FLOP["conv3"] + LAYER_FLOP["conv4"]
    total = sum(LAYER_FLOP[key] for key in ("conv1", "conv2", "conv3", "conv4", "fc1", "fc2", "fc3") if key in name)
    if total:
        return total
    if "conv" in name:
        return LAYER_FLOP["conv"]
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="windows per GPU per step")
    ap.add_argument("--precision", default=None, choices=[None, "bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true",
                    help="only the device-resident `value` loop and the per-kernel times (the ncu launch list of THIS command is "
                         "the step's four kernels and nothing else); prints the same line with the other legs null")
    ap.add_argument("--no-latency", action="store_true", help="skip the latency_b1 leg (profiler runs: its resident servers wait for host doorbells)")
    ap.add_argument("--stream-steps", type=int, default=2_000_000, help="rows of the synthetic log of the `stream` leg per GPU (0: skip)")
    ap.add_argument("--big-batch", type=int, default=32768, help="windows per GPU of the one-call big-batch leg, BASELINE configs[3] (0: skip)")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE",
                    help="dce_weights_set_option(KEY, VALUE) before timing (A/B of ablation switches; recorded in config.options)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import deep_contact_estimator_b200 as dce
    from deep_contact_estimator_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision or dce.default_precision()
    B = args.batch

    # weights: rank 0 packs (K0), one NCCL broadcast of the packed buffer, outside the timed region
    eng = dce.ContactEngine(synth.make_params(0) if rank == 0 else None, dev, precision)
    options = {}
    for kv in args.set:
        key, _, val = kv.partition("=")
        if eng.set_option(key, int(val)) != 0:
            raise SystemExit(f"bench.py: dce_weights_set_option({key!r}, {val}) was rejected")
        options[key] = int(val)
    bcast_ms = None
    if world > 1:
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        eng.broadcast_weights(src=0)
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3

    # inputs: NBUF distinct batches resident in HBM, rotated so every step reads data no longer in L2
    NBUF = 4
    host = [synth.make_windows(B, seed=1000 * (rank + 1) + i) for i in range(NBUF)]
    xs = [h.to(dev) for h in host]
    pinned = [h.pin_memory() for h in host]
    out_bits_host = torch.empty((B, 4), dtype=torch.uint8).pin_memory()
    out_cls_host = torch.empty((B,), dtype=torch.int32).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (`value`) ---------------------------------
    for i in range(args.warmup):
        eng.classify(xs[i % NBUF], want_logits=True)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    sync_all()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        eng.classify(xs[i % NBUF], want_logits=True)
        launches += eng.last_launches
    e1.record()
    sync_all()
    t_wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- per-kernel durations for the roofline (CUDA events around every launch) ----
    prof_runs = [eng.profile_forward(xs[i % NBUF]) for i in range(min(args.steps, 50))]
    per_kernel_ms, per_kernel_launches = dominant_kernel(prof_runs)

    if args.kernel_only:
        if rank == 0:
            step_kernel_ms = sum(per_kernel_ms.values())
            print(json.dumps({"metric": "contact windows/sec at batch=4096", "value": value, "unit": "windows/s", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3_f32acc" if precision == "bf16x3" else "f32",
                              "data": "synthetic", "config": {"workload": workload_name(B), "precision": precision, "kernel_only": True},
                              "e2e": None, "gpu_launches": launches, "clocks": clocks,
                              "kernels_ms_per_step": {k: round(v, 4) for k, v in per_kernel_ms.items()},
                              "kernel_share_of_step": {k: round(v / step_kernel_ms, 4) for k, v in per_kernel_ms.items()}}), flush=True)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- end to end from pinned host memory (`e2e`) --------------------------------
    for i in range(2):
        eng.classify_host(pinned[i % NBUF], out_bits_host, out_cls_host)
    sync_all()
    t0 = time.perf_counter()
    e2e_launches = 0
    e2e_steps = min(args.steps, 100)
    for i in range(e2e_steps):
        eng.classify_host(pinned[i % NBUF], out_bits_host, out_cls_host)     # returns after the D2H read completed
        e2e_launches += eng.last_launches
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    # the same call with TWO batches in flight (classify_host_async): the upload of batch i+1 runs under the kernel tail of
    # batch i; every batch still crosses PCIe in (132.7 MB) and out (32 KB) inside the timed region, each into its own buffers
    outs = [(torch.empty((B, 4), dtype=torch.uint8).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory()) for _ in range(2)]
    eng.classify_host_async(pinned[0], *outs[0]).wait()
    sync_all()
    t0 = time.perf_counter()
    pending = None
    for i in range(e2e_steps):
        h = eng.classify_host_async(pinned[i % NBUF], *outs[i & 1])
        if pending is not None:
            pending.wait()
        pending = h
    pending.wait()
    torch.cuda.synchronize()
    e2e2_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e2_value = world * B * e2e_steps / (e2e2_ms * 1e-3)
    e2e2_ok = bool(torch.equal(outs[(e2e_steps - 1) & 1][1], eng.classify(xs[(e2e_steps - 1) % NBUF], want_logits=False)[1].cpu()))

    # ---- the box's concurrent pinned H2D rate (the ceiling of `e2e`), and the zero-copy variant of the same call ----
    h2d_gbs = h2d_leg(dev, pinned[0], sync_all, max_over_ranks)
    zc = None
    try:
        eng.classify_host(pinned[0], out_bits_host, out_cls_host, zero_copy=True)
        sync_all()
        t0 = time.perf_counter()
        zc_steps = min(args.steps, 20)
        for i in range(zc_steps):
            eng.classify_host(pinned[i % NBUF], out_bits_host, out_cls_host, zero_copy=True)
        zc_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / zc_steps
        zc = {"value": world * B / (zc_ms * 1e-3), "unit": "windows/s", "ms_per_step": zc_ms, "steps": zc_steps,
              "what": "classify_host(zero_copy=True): no staging copy, block1's bulk-TMA loader reads the pinned windows in place over PCIe"}
    except Exception as e:                                    # informational leg
        zc = {"error": f"{type(e).__name__}: {e}"[:200]}
        sync_all()

    # ---- BASELINE configs[2] and configs[3] ----------------------------------------
    stream = stream_leg(eng, dev, rank, sync_all, max_over_ranks, args.stream_steps) if args.stream_steps >= 300000 else None
    big = big_batch_leg(eng, dev, rank, sync_all, max_over_ranks, args.big_batch) if args.big_batch > 0 else None

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    latency = None if args.no_latency else gpu_latency_b1(eng, dev)
    eager = torch_eager_gpu(dev, xs[0], eng.classify(xs[0], want_logits=True)[0])
    peaks, peak_src = load_peaks()
    dom = max(per_kernel_ms, key=per_kernel_ms.get) if per_kernel_ms else None
    roofline = None
    if dom:
        step_kernel_ms = sum(per_kernel_ms.values())
        n_launch = per_kernel_launches[dom]
        avg_launch_ms = per_kernel_ms[dom] / n_launch
        flops_per_launch = kernel_flops(dom) * B / n_launch
        achieved_tflops = flops_per_launch / (avg_launch_ms * 1e-3) / 1e12
        # which measured peak applies: the burst figure for a short timed region at full clocks (a kernel "timed
        # alone"), the sustained one for a long, power-capped region — both fractions are printed either way
        p_burst = peaks.get("bf16_tflops") or peaks.get("bf16_tflops_sustained")
        p_sust = peaks.get("bf16_tflops_sustained") or p_burst
        capped = bool(clocks and ("sw_power_cap" in (clocks.get("reasons") or []) or
                                  (clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] < 0.97 * clocks["sm_max_mhz"])))
        use_burst = (ms_total < 100.0) and not capped
        # the dominant kernel's duration comes from individually synchronised profiled steps (a kernel "timed alone"): burst
        # peak, whatever the length of the main timed region; the WHOLE-STEP figure is judged against the regime of that region
        peak = p_burst
        step_peak = p_burst if use_burst else p_sust
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and B == 4096:
            with open(tpath) as f:
                traffic = json.load(f).get(dom)
        roofline = {
            "bound": "tensor", "kernel": dom, "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved_tflops / peak, "traffic": traffic,
            "frac_of_burst_peak": achieved_tflops / p_burst, "frac_of_sustained_peak": achieved_tflops / p_sust,
            "peak_source": f"{peak_src} cuBLAS bf16 burst ({p_burst} TFLOP/s): the kernel's duration is the mean over individually "
                           f"synchronised profiled steps (dce_forward_profile), i.e. a kernel timed alone",
            "note": "algorithmic FLOPs (split-precision passes count once; ceiling of frac is 1/3 with three bf16 passes)",
            "kernel_share_of_step": per_kernel_ms[dom] / step_kernel_ms,
            "kernels_ms_per_step": {k: round(v, 4) for k, v in per_kernel_ms.items()},
            "whole_step": {"achieved_tflops": value / world * FLOP_PER_WINDOW / 1e12,
                           "peak": step_peak, "frac": value / world * FLOP_PER_WINDOW / 1e12 / step_peak,
                           "peak_source": (f"burst: timed region of {ms_total:.1f} ms at full clocks" if use_burst else
                                           f"sustained: timed region of {ms_total:.1f} ms" + (", power-capped / clocks below max" if capped else "")),
                           "hbm_read_gbs": value / world * BYTES_PER_WINDOW / 1e9,
                           "hbm_frac": value / world * BYTES_PER_WINDOW / 1e9 / peaks["hbm_gbs"]},
        }

    cpu = None
    if not args.no_cpu_baseline:     # rank 0, after every rank's GPU work is done (the others idle at the final barrier); a shorter sample at N > 1
        cpu, _ = cpu_forward_baseline(B, budget_s=20.0 if world == 1 else 6.0, max_runs=5 if world == 1 else 2)

    line = {
        "metric": "contact windows/sec at batch=4096", "value": value, "unit": "windows/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3_f32acc"}.get(precision, "f32"), "data": "synthetic",
        "config": {"workload": workload_name(B),
                   "precision": precision, **({"options": options} if options else {}),
                   "weights": "seeded random init (synth.make_params(0))",
                   "l2": f"inputs rotate over {NBUF} resident batches of {B * 32400 / 1e6:.1f} MB each (> 126 MB L2 in total)",
                   "parallelism": f"window-range shards x{world}, one-time NCCL weight broadcast"
                                  + (f" ({bcast_ms:.1f} ms, untimed)" if bcast_ms else "")},
        "e2e": {"value": e2e2_value, "unit": "windows/s", "h2d_bytes_per_step": B * 32400, "d2h_bytes_per_step": B * 8,
                "ms_per_step": e2e2_ms / e2e_steps, "steps": e2e_steps, "results_match_device_path": e2e2_ok,
                "api": "ContactEngine.classify_host_async (pinned host windows -> host cls+bits), two batches in flight: every batch is "
                       "uploaded (132.7 MB) and its result read back (32 KB) inside the timed region, into its own pinned buffers; the "
                       "upload of batch i+1 runs under the kernel tail of batch i",
                "one_batch_at_a_time": {"value": e2e_value, "unit": "windows/s", "ms_per_step": e2e_ms / e2e_steps,
                                        "api": "ContactEngine.classify_host: returns after its own result is on the host, so the ~0.25 ms "
                                               "kernel tail after the last uploaded chunk is exposed every step"}},
        "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "latency_b1": latency, "torch_eager_gpu": eager,
        "h2d": {"pinned_h2d_gbs_per_gpu_all_ranks_copying": h2d_gbs, "e2e_h2d_gbs_per_gpu": B * 32400 / (e2e2_ms / e2e_steps * 1e-3) / 1e9,
                "e2e_frac_of_h2d": (B * 32400 / (e2e2_ms / e2e_steps * 1e-3) / 1e9) / h2d_gbs, "zero_copy": zc},
    }
    if stream:
        n_w = stream["windows_per_gpu"]
        line["stream"] = {
            "workload": f"dce_stream over one contiguous synthetic {stream['steps_per_gpu']}-step x 54 fp32 sensor log per GPU "
                        f"({n_w} windows; BASELINE configs[2], synth.make_sensor_log seed 2 + rank)",
            "value": world * n_w / (stream["ms"] * 1e-3), "unit": "windows/s", "ms_per_log": stream["ms"],
            "gpu_launches_per_log": stream["launches"],
            "e2e": {"value": world * n_w / (stream["e2e_ms"] * 1e-3), "unit": "windows/s", "ms_per_log": stream["e2e_ms"],
                    "h2d_bytes_per_step": stream["h2d_bytes"], "d2h_bytes_per_step": stream["d2h_bytes"],
                    "api": "ContactEngine.stream_host (pinned host log -> chunked upload under the kernels -> host cls+bits)"}}
    if big:
        tot = world * big["windows_per_gpu"]
        line["batch_%d" % tot if world == 8 else "batch_32768_per_gpu"] = {
            "workload": f"{tot} windows over {world} GPU(s): ONE dce_forward call of {big['windows_per_gpu']} resident windows per rank "
                        f"(BASELINE configs[3] is this at 8 GPUs = 262144)",
            "value": tot / (big["ms_per_call"] * 1e-3), "unit": "windows/s", "ms_per_call": big["ms_per_call"],
            "gpu_launches_per_call": big["launches_per_call"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
