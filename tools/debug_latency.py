#!/usr/bin/env python
"""Fused latency kernel (csrc/dce_latency.cuh) against the oracle, phase by phase: logits / classes / bits for
B = 1..4 in batch and stream mode, and — on a mismatch — the intermediate buffers P1, A4, H1, H2 it left in the
workspace.  Also times graph replays with and without the cooperative launch attribute.
   timeout -s KILL 120 python tools/debug_latency.py"""
import os, sys, time
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_contact_estimator_b200 as dce
from deep_contact_estimator_b200 import synth
from oracle import contact_oracle as oracle

dev = torch.device("cuda", 0)
params = synth.make_params(0)
P = params


def stages(x):
    """oracle intermediates in the kernel's layouts: p1 [B][75][64], a4 [B][t*128+c], h1, h2, logits"""
    with torch.no_grad():
        y = x.permute(0, 2, 1)
        y = F.relu(F.conv1d(y, P["block1.0.weight"], P["block1.0.bias"], padding=1))
        y = F.max_pool1d(F.relu(F.conv1d(y, P["block1.2.weight"], P["block1.2.bias"], padding=1)), 2, 2)
        p1 = y.permute(0, 2, 1).contiguous()
        y = F.relu(F.conv1d(y, P["block2.0.weight"], P["block2.0.bias"], padding=1))
        y = F.max_pool1d(F.relu(F.conv1d(y, P["block2.2.weight"], P["block2.2.bias"], padding=1)), 2, 2)
        a4 = y.permute(0, 2, 1).reshape(x.shape[0], -1)
        h1 = F.relu(F.linear(y.reshape(x.shape[0], -1), P["fc.0.weight"], P["fc.0.bias"]))
        h2 = F.relu(F.linear(h1, P["fc.3.weight"], P["fc.3.bias"]))
        lg = F.linear(h2, P["fc.6.weight"], P["fc.6.bias"])
    return [t.numpy() for t in (p1, a4, h1, h2, lg)]


def ws_views(eng, n):
    ws = eng._workspace
    al = lambda v: (v + 255) // 256 * 256
    o = 256; out = []
    for cnt in (n * 75 * 64, n * 4736, n * 2048):
        out.append(ws[o:o + cnt * 4].view(torch.float32).cpu().numpy()); o = al(o + cnt * 4)
    return out


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


ok = True
for precision in ("bf16x3", "fp32"):
    eng = dce.ContactEngine(params, dev, precision)
    for B in (1, 2, 3, 4):
        x = synth.make_windows(B, seed=10 + B)
        want = stages(x)
        for rep in range(3):                           # repeated calls: the barrier counters must re-arm
            lg, cl, bi = eng.classify(x.to(dev)); torch.cuda.synchronize()
        err = oracle.normwise_rel_err(lg.cpu().numpy(), want[4])
        good = err < 1e-5 and np.array_equal(cl.cpu().numpy(), want[4].argmax(1)) and \
            np.array_equal(bi.cpu().numpy(), oracle.decimal2binary_numpy(want[4].argmax(1)))
        print(f"[{precision}] batch B={B}: launches {eng.last_launches} err {err:.2e} {'ok' if good else 'MISMATCH'}", flush=True)
        if not good:
            ok = False
            got = ws_views(eng, B)
            for name, g, w in zip(("p1", "a4", "h1"), got, want[:3]):
                print(f"    {name}: rel err {rel(g, w.reshape(-1)):.3e}  nan {int(np.isnan(g).sum())}")
            print("    sync counters", eng._workspace[:8].view(torch.int32).cpu().numpy())
    log = synth.make_sensor_log(150 + 7, seed=2)
    for first, n in ((0, 1), (1, 3), (4, 4)):
        _, wc, wb = oracle.inference_stream(params, log)
        lg, cl, bi = eng.stream(log.to(dev), first, n, want_logits=True); torch.cuda.synchronize()
        ds = torch.stack([(log[i:i + 150] - log[i:i + 150].mean(0)) / log[i:i + 150].std(0) for i in range(first, first + n)])
        err = oracle.normwise_rel_err(lg.cpu().numpy(), stages(ds)[4])
        good = err < 1e-5 and np.array_equal(bi.cpu().numpy(), wb.numpy()[first:first + n])
        print(f"[{precision}] stream first={first} n={n}: err {err:.2e} {'ok' if good else 'MISMATCH'}", flush=True)
        ok &= good

# timing: CUDA-graph replay, cooperative attribute on / off, vs the per-layer kernels
eng = dce.ContactEngine(params, dev, "bf16x3")
x = synth.make_windows(1, seed=5).to(dev)
for label, opts in (("fused coop", {b"latency_kernel": 1, b"latency_coop": 1}), ("fused non-coop", {b"latency_kernel": 1, b"latency_coop": 0}),
                    ("per-layer", {b"latency_kernel": 0})):
    for k, v in opts.items():
        eng.set_option(k, v)
    try:
        s = torch.cuda.Stream(dev)
        with torch.cuda.stream(s):
            for _ in range(3):
                eng.classify(x)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                lg, cl, bi = eng.classify(x)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(300)]
        for a, b in ev:
            a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        us = np.array([a.elapsed_time(b) * 1e3 for a, b in ev])
        print(f"{label}: graph replay p50 {np.percentile(us, 50):.1f} us  p99 {np.percentile(us, 99):.1f} us  launches {eng.last_launches}", flush=True)
    except Exception as e:                                # e.g. cooperative launch not capturable
        print(f"{label}: FAILED {type(e).__name__}: {e}", flush=True)
        ok = False
eng.set_option(b"latency_kernel", 1); eng.set_option(b"latency_coop", 1)
# phase timeline (clock64 of the first and last CTA, written behind the barrier counters)
NAMES = ["start", "A done", "bar1", "B done", "bar2", "C done", "bar3", "D done", "E done (last CTA only)", "exit"]
for mode in ("batch", "stream"):
    logd = synth.make_sensor_log(151, seed=2).to(dev)
    for _ in range(5):
        eng.classify(x) if mode == "batch" else eng.stream(logd, 0, 1)
    torch.cuda.synchronize()
    tr = eng._workspace[64:64 + 24 * 8].view(torch.int64).cpu().numpy().reshape(2, 12)
    for which, row in zip(("cta 0", "last cta"), tr):
        d = (row[1:10] - row[0]) / 1.965e3
        print(f"trace[{mode}] {which} (us since start): " + "  ".join(f"{n} {v:.2f}" for n, v in zip(NAMES[1:], d)), flush=True)
print("debug_latency", "OK" if ok else "FAILED")
