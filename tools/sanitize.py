#!/usr/bin/env python
"""Small-shape run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize.py"""
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
    for key in (b"fuse_block1", b"fuse_block2", b"fuse_fc3", b"fuse_argmax", b"latency_kernel"):
        eng.set_option(key, 0)
        x = synth.make_windows(37 if key != b"latency_kernel" else 3, seed=37)
        lo, cl, bi = eng.classify(x.to(dev)); torch.cuda.synchronize()
        assert np.array_equal(cl.cpu().numpy(), oracle.forward_torch(params, x).detach().numpy().argmax(1)), key
        eng.set_option(key, 1)
# the resident servers (a few steps each; generous idle timeout: the sanitizer slows the kernels down a lot)
eng = dce.ContactEngine(params, dev, "bf16x3")
xs = synth.make_windows(6, seed=3)
want = oracle.forward_torch(params, xs).detach().numpy().argmax(1)
run = eng.latency_runner(1, persistent=True, idle_timeout_s=30.0)
for i in range(6):
    cls, bits = run.step(xs[i])
    assert int(cls[0]) == int(want[i]), ("window server", i)
run.close()
log = synth.make_sensor_log(150 + 4, seed=2)
_, wc, wb = oracle.inference_stream(params, log)
rr = eng.row_runner(idle_timeout_s=30.0)
got = [rr.push(log[t]) for t in range(log.shape[0])][149:]
rr.close()
assert [g[0] for g in got] == wc.tolist(), "row server"
print("sanitize run ok")
