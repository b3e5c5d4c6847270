#!/usr/bin/env python
"""First GPU run of the experimental options (dce_set_option: "fc_f16f8", "conv_f16f8", "block2_cluster", "fc_cluster";
DESIGN.md §3, §8): parity against the oracle on 512 windows, then an A/B of the batch-4096 step with per-kernel times.
Run it under a timeout — none of them has been on a GPU yet:
    timeout 180 python tools/try_f16f8.py [--fc-only | --cluster-only]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce          # noqa: E402
from deep_contact_estimator_b200 import synth      # noqa: E402
from oracle import contact_oracle as oracle        # noqa: E402

dev = torch.device("cuda", 0)
# (fc_f16f8, conv_f16f8, block2_cluster, fc_cluster): bf16x3 everywhere | FC layers in fp16 + e4m3 | + X2 and block2 |
# + block1's convolutions (every MMA of the path) | bf16x3 with block2 in clusters of 2 / 4 sharing the weight stream |
# bf16x3 with fc.0 / fc.3 as CTA pairs sharing the activation slabs | everything on
MODES = [(0, 0, 0, 0), (1, 0, 0, 0), (1, 1, 0, 0), (1, 2, 0, 0), (0, 0, 2, 0), (0, 0, 4, 0), (0, 0, 0, 2), (1, 2, 4, 2)]
if "--fc-only" in sys.argv:
    MODES = [(0, 0, 0, 0), (1, 0, 0, 0)]
if "--cluster-only" in sys.argv:
    MODES = [(0, 0, 0, 0), (0, 0, 2, 0), (0, 0, 4, 0), (0, 0, 0, 2)]


def set_mode(eng, fc, conv, cl, fcl):
    assert eng.lib.dce_set_option(b"fc_f16f8", fc) == 0 and eng.lib.dce_set_option(b"conv_f16f8", conv) == 0
    assert eng.lib.dce_set_option(b"block2_cluster", cl) == 0 and eng.lib.dce_set_option(b"fc_cluster", fcl) == 0


for scale in (1.0, 50.0):
    params = synth.make_params(0, logit_scale=scale)
    eng = dce.ContactEngine(params, dev, "bf16x3")
    x = synth.make_windows(512, seed=1)
    with torch.no_grad():
        want = oracle.forward_torch(params, x)
    for fc, conv, cl, fcl in MODES:
        set_mode(eng, fc, conv, cl, fcl)
        logits, cls, bits = eng.classify(x.to(dev))
        torch.cuda.synchronize()
        err = oracle.normwise_rel_err(logits.cpu().numpy(), want.numpy())
        same = bool((cls.cpu().long() == oracle.argmax_class(want)).all())
        print(f"logit scale {scale:4.0f}  fc_f16f8={fc} conv_f16f8={conv} block2_cluster={cl} fc_cluster={fcl}: normwise err {err:.2e}, "
              f"classes exact {same}, {eng.last_launches} launches, range status {eng.f16f8_status():#x}", flush=True)
    set_mode(eng, 0, 0, 0, 0)

# the per-call form: precision "f16f8" (DCE_PREC_F16F8) = fc_f16f8 1 + conv_f16f8 2 without touching the global options
if "--fc-only" not in sys.argv and "--cluster-only" not in sys.argv:
    e8 = dce.ContactEngine(synth.make_params(0), dev, "f16f8")
    x = synth.make_windows(512, seed=1)
    with torch.no_grad():
        want = oracle.forward_torch(synth.make_params(0), x)
    logits, cls, _ = e8.classify(x.to(dev))
    torch.cuda.synchronize()
    print(f'precision "f16f8": normwise err {oracle.normwise_rel_err(logits.cpu().numpy(), want.numpy()):.2e}, classes exact '
          f"{bool((cls.cpu().long() == oracle.argmax_class(want)).all())}", flush=True)

eng = dce.ContactEngine(synth.make_params(0), dev, "bf16x3")
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


for fc, conv, cl, fcl in MODES + MODES:
    set_mode(eng, fc, conv, cl, fcl)
    us = round(step_us(), 1)
    prof = {}
    for i in range(10):
        for name, ms in eng.profile_forward(xs[i % 4]):
            prof[name] = prof.get(name, 0) + ms * 100
    print(f"fc_f16f8={fc} conv_f16f8={conv} block2_cluster={cl} fc_cluster={fcl}: step {us} us | per-kernel us:", {k: round(v, 1) for k, v in prof.items()}, flush=True)
set_mode(eng, 0, 0, 0, 0)
