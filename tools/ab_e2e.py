#!/usr/bin/env python
"""Host-to-host step time of ContactEngine.classify_host at batch 4096 for several upload chunk sizes and for the
zero-copy form (kernels read the pinned windows in place over PCIe).   python tools/ab_e2e.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth

dev = torch.device("cuda", 0)
eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
B = 4096
xs = [synth.make_windows(B, seed=5 + i).pin_memory() for i in range(4)]
ob, oc = torch.empty((B, 4), dtype=torch.uint8).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory()
_, ref_cls, _ = eng.classify(xs[0].to(dev), want_logits=False)


def run(n=40, **kw):
    for i in range(3):
        eng.classify_host(xs[i % 4], ob, oc, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        eng.classify_host(xs[i % 4], ob, oc, **kw)
    ms = (time.perf_counter() - t0) * 1e3 / n
    eng.classify_host(xs[0], ob, oc, **kw)
    return ms, bool(torch.equal(oc, ref_cls.cpu()))


for kw in ({"chunk": 512}, {"chunk": 1024}, {"chunk": 2048}, {"chunk": 4096}, {"zero_copy": True}):
    ms, ok = run(**kw)
    print(f"{kw}: {ms:.3f} ms per step = {B / ms * 1e3 / 1e6:.3f} M windows/s, {B * 32400 / ms / 1e6:.1f} GB/s over PCIe, classes match: {ok}", flush=True)
