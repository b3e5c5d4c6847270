#!/usr/bin/env python
"""Per-role clock64 timeline of CTA 0 of the fused block1 kernel (cycles relative to the first event).
Needs a trace build: DCE_TRACE=1 python -m deep_contact_estimator_b200.build --force"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
x = synth.make_windows(4096, seed=1).to(dev)
for _ in range(3): eng.classify(x)
eng.set_option(b"trace", 1)
eng.classify(x); torch.cuda.synchronize()
buf = eng.read_trace(60 * 16)
t = np.array(buf, dtype=np.int64).reshape(60, 16)
t0 = t[t > 0].min()
names = ["cv:start", "cv:x0empty", "cv:landed", "cv:done", "mma:c1", "mma:c2", "e1:start", "e1:d1full", "e1:ld_done", "e1:x1empty", "e1:done", "e2:start", "e2:d2full", "e2:done", "-", "cv:read"]
print("tile " + " ".join(f"{n:>10s}" for n in names))
for k in range(2, 12):
    print(f"{k:4d} " + " ".join(f"{(t[k, e] - t0) if t[k, e] else 0:10d}" for e in range(16)))
d = np.diff(t[5:30, 5])
print("period between successive conv2 issues (cycles):", d.mean(), d.min(), d.max())
