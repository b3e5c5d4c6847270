#!/usr/bin/env python
"""Throughput of batch-4096 steps issued alternately on two streams (two engines = two workspaces) against one stream:
does the fill / drain of each kernel overlap with the other lane's kernels?"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce          # noqa: E402
from deep_contact_estimator_b200 import synth      # noqa: E402

dev = torch.device("cuda", 0)
params = synth.make_params(0)
engs = [dce.ContactEngine(params, dev, "bf16x3") for _ in range(2)]
streams = [torch.cuda.Stream(dev) for _ in range(2)]
xs = [synth.make_windows(4096, seed=5 + i).to(dev) for i in range(4)]
torch.cuda.synchronize()


def run(lanes, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    torch.cuda.synchronize()
    a.record(main)
    for s in streams[:lanes]:
        s.wait_stream(main)
    outs = []
    for i in range(n):
        with torch.cuda.stream(streams[i % lanes]):
            outs.append(engs[i % lanes].classify(xs[i % 4]))
    for s in streams[:lanes]:
        main.wait_stream(s)
    b.record(main)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n, outs


for lanes in (1, 2):
    run(lanes, 50)
for rep in range(3):
    for lanes, n in ((1, 20), (2, 20), (1, 400), (2, 400)):
        us, outs = run(lanes, n)
        print(f"lanes {lanes} steps {n}: {us:.1f} us per step = {4096 / us:.2f} M windows/s", flush=True)
ref = engs[0].classify(xs[0]); torch.cuda.synchronize()
_, outs = run(2, 8)
print("lane results equal:", all(torch.equal(a, b) for a, b in zip(outs[4], ref)), all(torch.equal(a, b) for a, b in zip(outs[0], ref)))
