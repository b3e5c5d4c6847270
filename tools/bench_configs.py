#!/usr/bin/env python
"""Secondary BASELINE.json configs (bench.py covers configs[1] and the scaling run):
   --stream   configs[2]: inference over a 10M-step contiguous synthetic sensor log, 1 GPU
   --latency  configs[4]: batch=1 latency, CUDA-graph replay, p50/p99 per-window microseconds
Prints one JSON line per config.  Parity for both is asserted on a prefix against the oracle."""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
from oracle import contact_oracle as oracle


def stream(args, dev, eng):
    T = args.steps_log
    log = synth.make_sensor_log(T, seed=2)
    logd = log.to(dev)
    n = T - 149
    eng.stream(logd, 0, min(n, 8192))                       # warm-up (workspace, attributes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, cls, bits = eng.stream(logd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # parity on the first windows vs the oracle (reference inference() semantics)
    k = 512
    _, wc, wb = oracle.inference_stream(synth.make_params(0), log[: k + 149], batch_size=128)
    ok = bool(np.array_equal(bits[:k].cpu().numpy(), wb.numpy()))
    # end to end from a pinned host log: chunked H2D overlapped with the kernels + D2H of classes and bits
    pinned = log.pin_memory()
    out = torch.empty((n, 4), dtype=torch.uint8).pin_memory()
    outc = torch.empty((n,), dtype=torch.int32).pin_memory()
    eng.stream_host(pinned[: 300_000], out_bits_host=out[: 300_000 - 149], out_cls_host=outc[: 300_000 - 149])   # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.stream_host(pinned, out_bits_host=out, out_cls_host=outc)
    e2e = time.perf_counter() - t0
    ok = ok and bool(torch.equal(out, bits.cpu()))
    print(json.dumps({"config": "stream", "steps": T, "windows": n, "ms": ms, "windows_per_s": n / ms * 1e3,
                      "launches": eng.last_launches, "bits_match_oracle_prefix": ok,
                      "e2e_s": e2e, "e2e_windows_per_s": n / e2e, "e2e_api": "ContactEngine.stream_host (pinned host log -> host cls+bits, chunked H2D overlapped)", "h2d_bytes": T * 216, "d2h_bytes": n * 8,
                      "hbm_read_bytes_per_window_algorithmic": 216}))


def latency(args, dev, eng):
    x1 = synth.make_windows(1, seed=5)
    want = oracle.forward_torch(synth.make_params(0), x1).numpy()
    run = eng.latency_runner(1, want_logits=True)
    run.x_host.copy_(x1)
    for _ in range(20):
        run.step()
    err = oracle.normwise_rel_err(run.logits_host.numpy(), want)
    host, gpu = [], []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.calls):
        t0 = time.perf_counter()
        run.step()                                 # one graph launch (kernel reads/writes pinned host memory), then stream sync
        host.append((time.perf_counter() - t0) * 1e6)
    for _ in range(args.calls):
        a.record(run.stream); run.enqueue(); b.record(run.stream); b.synchronize()
        gpu.append(a.elapsed_time(b) * 1e3)
    a.record(run.stream)
    for _ in range(args.calls):
        run.enqueue()
    b.record(run.stream); b.synchronize()
    print(json.dumps({"config": "latency", "batch": 1, "calls": args.calls, "precision": eng.precision,
                      "gpu_us_p50": float(np.percentile(gpu, 50)), "gpu_us_p99": float(np.percentile(gpu, 99)),
                      "gpu_us_back_to_back": a.elapsed_time(b) * 1e3 / args.calls,
                      "host_us_p50": float(np.percentile(host, 50)), "host_us_p99": float(np.percentile(host, 99)),
                      "launches_per_call": run.launches, "normwise_err": err,
                      "bits": run.bits_host.numpy().tolist()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stream", action="store_true"); ap.add_argument("--latency", action="store_true")
    ap.add_argument("--steps-log", type=int, default=10_000_000); ap.add_argument("--calls", type=int, default=1000)
    ap.add_argument("--precision", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = dce.ContactEngine(synth.make_params(0), dev, args.precision)
    if args.stream or not args.latency:
        stream(args, dev, eng)
    if args.latency or not args.stream:
        latency(args, dev, eng)


if __name__ == "__main__":
    main()
