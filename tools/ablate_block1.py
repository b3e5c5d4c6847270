#!/usr/bin/env python
"""Timing ablations of the fused block1 kernel (results are invalid when a flag is set)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth

dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
xs = [synth.make_windows(4096, seed=10 + i).to(dev) for i in range(3)]
key = b"tapgemm_dbg" if "--tapgemm" in sys.argv else b"block1_dbg"
for dbg in [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [0, 1, 2, 4, 8, 9, 11, 15]:
    eng.set_option(key, dbg)
    for i in range(3):
        eng.classify(xs[i % 3])
    tot = {}
    for i in range(10):
        for name, ms in eng.profile_forward(xs[i % 3]):
            tot[name] = tot.get(name, 0) + ms / 10
    print(f"dbg={dbg:2d}  " + "  ".join(f"{k} {v*1e3:6.1f}" for k, v in tot.items()))
