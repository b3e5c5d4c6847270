#!/usr/bin/env python
"""One batch-4096 step per option set, for ncu:   python tools/profile_step.py [key=value ...] [-- key=value ...]
(3 warm-up steps, then 2 steps with the options of every set; run it under `ncu -k regex:... -s <skip>`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
sets, cur = [], {}
for a in sys.argv[1:]:
    if a == "--":
        sets.append(cur); cur = {}
    else:
        k, v = a.split("="); cur[k] = int(v)
sets.append(cur)
xs = [synth.make_windows(4096, seed=5 + i).to(dev) for i in range(2)]
for s in sets:
    for k, v in s.items():
        assert eng.set_option(k, v) == 0, k
    for i in range(3 if s is sets[0] else 2):
        eng.classify(xs[i % 2])
    torch.cuda.synchronize()
    for k in s:
        eng.set_option(k, 0)
