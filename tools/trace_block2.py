#!/usr/bin/env python
"""clock64 timeline of CTA 0 of the fused block2 kernel (build with DCE_TRACE=1)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
x = synth.make_windows(4096, seed=1).to(dev)
for _ in range(3): eng.classify(x)
eng.set_option(b"trace", 1); eng.set_option(b"trace_layer", 6)
eng.classify(x); torch.cuda.synchronize()
buf = eng.read_trace(60 * 16)
t = np.array(buf, dtype=np.int64).reshape(60, 16)
t0 = t[t > 0].min()
names = ["c3:pre", "c3:go", "c4:pre", "c4:go", "c4:issued", "e1:start", "e1:d3full", "e1:ld", "e1:x3empty", "e1:done", "e2:start", "e2:d4full", "e2:done", "e2:ld", "e2:pooled", "e2:cvt0"]
print("tile " + " ".join(f"{n:>10s}" for n in names))
for k in range(3, 11):
    print(f"{k:4d} " + " ".join(f"{(t[k, e] - t0) if t[k, e] else 0:10d}" for e in range(len(names))))
