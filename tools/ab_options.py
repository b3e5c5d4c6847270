#!/usr/bin/env python
"""A/B of a dce_set_option switch on the batch-4096 step: python tools/ab_options.py fuse_fc3 [fuse_block2 ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
xs = [synth.make_windows(4096, seed=5 + i).to(dev) for i in range(4)]
def step_us(n=60):
    for i in range(5): eng.classify(xs[i % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): eng.classify(xs[i % 4])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n
for key in [k.encode() for k in sys.argv[1:]] or [b"fuse_fc3"]:
    res = {}
    for v in (1, 0, 1, 0):
        eng.set_option(key, v)
        res.setdefault(v, []).append(round(step_us(), 1))
    eng.set_option(key, 1)
    prof = {}
    for i in range(10):
        for n, ms in eng.profile_forward(xs[i % 4]): prof[n] = prof.get(n, 0) + ms * 100
    print(key.decode(), "step us: on", res[1], "off", res[0], "| per-kernel (on):", {k: round(v, 1) for k, v in prof.items()}, flush=True)
