#!/usr/bin/env python
"""Small-shape run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize.py [--experimental]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
from oracle import contact_oracle as oracle
dev = torch.device("cuda", 0)
params = synth.make_params(0)
for precision in ("bf16x3", "fp32"):
    eng = dce.ContactEngine(params, dev, precision)
    for B in (1, 3, 37, 260):
        x = synth.make_windows(B, seed=B)
        lo, cl, bi = eng.classify(x.to(dev)); torch.cuda.synchronize()
        want = oracle.forward_torch(params, x).detach().numpy()
        assert np.array_equal(cl.cpu().numpy(), want.argmax(1)), (precision, B)
    log = synth.make_sensor_log(150 + 333, seed=2)          # odd row count: exercises the unaligned stream tail
    _, cl, bi = eng.stream(log.to(dev)); torch.cuda.synchronize()
    _, wc, wb = oracle.inference_stream(params, log)
    assert np.array_equal(bi.cpu().numpy(), wb.numpy()), precision
    for key in (b"fuse_block1", b"fuse_block2", b"fuse_fc3", b"latency_kernel"):
        eng.lib.dce_set_option(key, 0)
        x = synth.make_windows(37 if key != b"latency_kernel" else 3, seed=37)
        lo, cl, bi = eng.classify(x.to(dev)); torch.cuda.synchronize()
        assert np.array_equal(cl.cpu().numpy(), oracle.forward_torch(params, x).detach().numpy().argmax(1)), key
        eng.lib.dce_set_option(key, 1)
if "--experimental" in sys.argv:
    # the option-gated kernels: per-call f16f8 precision, then the cluster variants
    eng = dce.ContactEngine(params, dev, "f16f8")
    for B in (5, 37, 260):
        x = synth.make_windows(B, seed=B)
        lo, cl, bi = eng.classify(x.to(dev)); torch.cuda.synchronize()
        assert np.array_equal(cl.cpu().numpy(), oracle.forward_torch(params, x).detach().numpy().argmax(1)), ("f16f8", B)
    _, cl, bi = eng.stream(log.to(dev)); torch.cuda.synchronize()
    assert np.array_equal(bi.cpu().numpy(), wb.numpy()), "f16f8 stream"
    for precision in ("bf16x3", "f16f8"):
        eng = dce.ContactEngine(params, dev, precision)
        for key, val in ((b"block2_cluster", 2), (b"block2_cluster", 4), (b"fc_cluster", 2)):
            eng.lib.dce_set_option(key, val)
            for B in (7, 260):
                x = synth.make_windows(B, seed=B)
                lo, cl, bi = eng.classify(x.to(dev)); torch.cuda.synchronize()
                assert np.array_equal(cl.cpu().numpy(), oracle.forward_torch(params, x).detach().numpy().argmax(1)), (precision, key, B)
            eng.lib.dce_set_option(key, 0)
print("sanitize run ok")
