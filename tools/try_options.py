#!/usr/bin/env python
"""A/B of handle options on the GPU: bit-equality of the results against the default kernels at several batch sizes,
then the batch-4096 step and per-kernel CUDA-event times, two passes (the second runs power-capped).  Run under a timeout:
    timeout 180 python tools/try_options.py block2_dbg=2 [key=value ...]  [-- second set ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce          # noqa: E402
from deep_contact_estimator_b200 import synth      # noqa: E402

dev = torch.device("cuda", 0)
sets, cur = [], {}
for a in sys.argv[1:]:
    if a == "--":
        sets.append(cur); cur = {}
    else:
        k, v = a.split("="); cur[k] = int(v)
sets.append(cur)
sets = [{}] + [s for s in sets if s]
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")


DEFAULTS = {"fuse_argmax": 1, "fuse_block1": 1, "fuse_block2": 1, "fuse_fc3": 1, "latency_kernel": 1, "latency_coop": 1, "latency_tma_in": 1}


def apply(opts):
    for s in sets:
        for k in s:
            assert eng.set_option(k, DEFAULTS.get(k, 0)) == 0
    for k, v in opts.items():
        assert eng.set_option(k, v) == 0, k


for B in (7, 300, 1000, 4096, 4097):
    x = synth.make_windows(B, seed=50 + B % 7).to(dev)
    apply({})
    want = eng.classify(x)
    torch.cuda.synchronize()
    for opts in sets[1:]:
        apply(opts)
        got = eng.classify(x)
        torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(got, want))
        err = (got[0] - want[0]).abs().max().item()
        print(f"B={B} {opts}: bit-identical {same}, max |dlogit| {err:.3e}, classes equal {bool(torch.equal(got[1], want[1]))}", flush=True)

xs = [synth.make_windows(4096, seed=5 + i).to(dev) for i in range(4)]


def step_us(n=60):
    for i in range(5):
        eng.classify(xs[i % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        eng.classify(xs[i % 4])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


# interleaved A/B: every round profiles one step of every set, so all sets see the same clocks / power state;
# the median over rounds is reported (the first rounds run hotter clocks than a long run sustains)
import statistics
for i in range(300):                       # ~0.1 s of load: leave the idle state
    eng.classify(xs[i % 4])
torch.cuda.synchronize()
rounds = 15
acc = [dict() for _ in sets]
steps = [[] for _ in sets]
for r in range(rounds):
    for si, opts in enumerate(sets):
        apply(opts)
        for i in range(2):
            for name, ms in eng.profile_forward(xs[(r + i) % 4]):
                acc[si].setdefault(name, []).append(ms * 1e3)
        steps[si].append(step_us(20))
for si, opts in enumerate(sets):
    med = {k: round(statistics.median(v), 1) for k, v in acc[si].items()}
    print(f"{opts}: step median {statistics.median(steps[si]):.1f} us (min {min(steps[si]):.1f}) | per-kernel median us: {med} sum {sum(med.values()):.1f}", flush=True)
apply({})
