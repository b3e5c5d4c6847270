#!/usr/bin/env python
"""Where a batch-1 step's time goes outside the kernel: host submission rate, cooperative vs plain launch,
device-resident vs zero-copy input.   python tools/latency_diag.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth

dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
N = 2000
for coop, tma in ((1, 1), (1, 0), (0, 1)):
    eng.set_option(b"latency_coop", coop)
    eng.set_option(b"latency_tma_in", tma)
    run = eng.latency_runner(1)
    run.x_host.copy_(synth.make_windows(1, seed=6))
    for _ in range(50):
        run.step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(run.stream)
    t0 = time.perf_counter()
    for _ in range(N):
        run.enqueue()
    t1 = time.perf_counter()
    b.record(run.stream); b.synchronize()
    print(f"zero-copy coop={coop} tma_in={tma}: host enqueue {1e6 * (t1 - t0) / N:.1f} us/call, device back-to-back {a.elapsed_time(b) * 1e3 / N:.1f} us/call", flush=True)
    # device-resident input/outputs, same graph mechanics
    x = synth.make_windows(1, seed=6).to(dev)
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        for _ in range(3):
            eng.classify(x)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = eng.classify(x)
        torch.cuda.synchronize()
        a.record(s)
        t0 = time.perf_counter()
        for _ in range(N):
            g.replay()
        t1 = time.perf_counter()
        b.record(s); b.synchronize()
    print(f"device-resident coop={coop} tma_in={tma}: host enqueue {1e6 * (t1 - t0) / N:.1f} us/call, device back-to-back {a.elapsed_time(b) * 1e3 / N:.1f} us/call", flush=True)
eng.set_option(b"latency_coop", 1); eng.set_option(b"latency_tma_in", 1)

# direct C-ABI launches (no CUDA graph) on the runner's stream: host cost per call and device period
import ctypes
from deep_contact_estimator_b200 import _lib
run = eng.latency_runner(1)
run.x_host.copy_(synth.make_windows(1, seed=6))
P = dce.ContactEngine._p
args = (eng._handle, P(run.x_host), 1, None, P(run.cls_host), P(run.bits_host), P(run._ws), run._ws.numel(),
        _lib.PRECISIONS[eng.precision], ctypes.c_void_p(run.stream.cuda_stream))
for _ in range(20):
    eng.lib.dce_forward(*args)
run.stream.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(run.stream)
t0 = time.perf_counter()
for _ in range(N):
    eng.lib.dce_forward(*args)
t1 = time.perf_counter()
b.record(run.stream); b.synchronize()
print(f"zero-copy direct dce_forward: host enqueue {1e6 * (t1 - t0) / N:.1f} us/call, device back-to-back {a.elapsed_time(b) * 1e3 / N:.1f} us/call", flush=True)
host = []
for _ in range(500):
    t0 = time.perf_counter()
    eng.lib.dce_forward(*args)
    run.stream.synchronize()
    host.append((time.perf_counter() - t0) * 1e6)
print(f"zero-copy direct dce_forward + stream sync: host wall p50 {np.percentile(host, 50):.1f} us  p99 {np.percentile(host, 99):.1f} us", flush=True)
host = []
for _ in range(500):
    t0 = time.perf_counter()
    run.step()
    host.append((time.perf_counter() - t0) * 1e6)
print(f"runner.step() (graph): host wall p50 {np.percentile(host, 50):.1f} us  p99 {np.percentile(host, 99):.1f} us", flush=True)
