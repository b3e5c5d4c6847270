#!/usr/bin/env python
"""GPU debugging aid: run the bf16x3 path on a few windows, decode every
activation tape out of the workspace and compare layer by layer with the oracle
(torch CPU fp32).  Not part of the product path.
    python tools/debug_tc_layers.py [B] [--layerwise] [--block2]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deep_contact_estimator_b200 as dce          # noqa: E402
from deep_contact_estimator_b200 import synth      # noqa: E402
from oracle import contact_oracle as oracle        # noqa: E402

GUARD = 8


def align(v, a=256):
    return (v + a - 1) // a * a


def tape(rows, kch):
    m_tiles = (rows + 127) // 128
    m_tiles += m_tiles & 1
    cap = GUARD + m_tiles * 128 + 136
    return dict(rows=rows, m_tiles=m_tiles, cap=cap, kch=kch, bytes=align(cap * 16 * kch * 2))


def rowmajor(rows, kch):            # the fc.0 operand: [part][row][kch * 8] bf16, no guard rows (make_rowmajor, dce_tc.cuh)
    m_tiles = (rows + 127) // 128
    m_tiles += m_tiles & 1
    cap = m_tiles * 128
    return dict(rows=rows, m_tiles=m_tiles, cap=cap, kch=kch, bytes=align(cap * 16 * kch * 2), rowmajor=True)


def workspace(n):
    o, W = 512, {}                  # [0, 256): the latency kernel's barrier counters, [256, 512): fc.3 tickets (make_workspace, dce_tc.cuh)
    for name, t in (("x0", tape(n * 152, 8)), ("x1", tape(n * 152, 8)), ("x2", tape(n * 76, 8)),
                    ("x3", tape(n * 76, 16)), ("x4", rowmajor(n, 592)), ("h1", tape(n, 256))):
        t["off"] = o; W[name] = t; o = align(o + t["bytes"])
    W["h2"] = dict(off=o); o = align(o + n * 512 * 4)
    return W


def decode(ws, t):
    if t.get("rowmajor"):
        raw = ws[t["off"]: t["off"] + t["cap"] * 16 * t["kch"] * 2].view(torch.bfloat16).view(2, t["cap"], t["kch"] * 8).float()
        return (raw[0] + raw[1])[:t["rows"]].cpu()
    raw = ws[t["off"]: t["off"] + t["cap"] * 16 * t["kch"] * 2].view(torch.bfloat16).view(2, t["kch"], t["cap"], 8).float()
    v = raw[0] + raw[1]                                         # [kch][cap][8]
    return v.permute(1, 0, 2).reshape(t["cap"], t["kch"] * 8)[GUARD:GUARD + t["rows"]].cpu()


def report(name, got, want):
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    bad = ((got - want).abs() > 1e-3 * max(ref, 1e-6)).sum().item()
    print(f"{name:8s} max|d| {err:.3e}  max|ref| {ref:.3e}  rel {err / max(ref, 1e-30):.2e}  bad elems {bad}/{want.numel()}")
    return err / max(ref, 1e-30)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    layerwise = "--layerwise" in sys.argv
    dev = torch.device("cuda", 0)
    params = synth.make_params(0)
    x = synth.make_windows(B, seed=1)
    eng = dce.ContactEngine(params, dev, "bf16x3")
    eng.set_option(b"fuse_block1", 0 if layerwise else 1)
    fused2 = "--block2" in sys.argv
    eng.set_option(b"fuse_block2", 1 if fused2 else 0)
    logits, cls, bits = eng.classify(x.to(dev))
    torch.cuda.synchronize()
    ws = eng._workspace
    W = workspace(B)
    p = params
    with torch.no_grad():
        h0 = x.permute(0, 2, 1)
        a1 = F.relu(F.conv1d(h0, p["block1.0.weight"], p["block1.0.bias"], padding=1))
        a2 = F.max_pool1d(F.relu(F.conv1d(a1, p["block1.2.weight"], p["block1.2.bias"], padding=1)), 2, 2)
        a3 = F.relu(F.conv1d(a2, p["block2.0.weight"], p["block2.0.bias"], padding=1))
        a4 = F.max_pool1d(F.relu(F.conv1d(a3, p["block2.2.weight"], p["block2.2.bias"], padding=1)), 2, 2)
        f1 = F.relu(F.linear(a4.reshape(B, -1), p["fc.0.weight"], p["fc.0.bias"]))
        f2 = F.relu(F.linear(f1, p["fc.3.weight"], p["fc.3.bias"]))
        want = F.linear(f2, p["fc.6.weight"], p["fc.6.bias"])

    def conv_tape(name, act, rw, tv, ch):
        got = decode(ws, W[name]).reshape(B, rw, -1)
        guard = got[:, tv:, :].abs().max().item()
        print(f"{name}: guard rows max |v| = {guard:.3e}")
        return report(name, got[:, :tv, :ch], act.permute(0, 2, 1))

    if layerwise:
        x0 = decode(ws, W["x0"]).reshape(B, 152, 64)
        report("x0", x0[:, :150, :54], x)
        print("x0 pad channels / guard rows:", x0[:, :, 54:].abs().max().item(), x0[:, 150:, :].abs().max().item())
        conv_tape("x1", a1, 152, 150, 64)
    conv_tape("x2", a2, 76, 75, 64)
    if not fused2:
        conv_tape("x3", a3, 76, 75, 128)
    x4 = decode(ws, W["x4"])      # [B][592*8], k' = t*128 + c
    report("x4", x4.reshape(B, 37, 128), a4.permute(0, 2, 1))
    report("h1", decode(ws, W["h1"]), f1)
    h2 = ws[W["h2"]["off"]: W["h2"]["off"] + B * 512 * 4].view(torch.float32).view(B, 512).cpu()
    report("h2", h2, f2)
    report("logits", logits.cpu(), want)
    print("normwise", oracle.normwise_rel_err(logits.cpu().numpy(), want.numpy()),
          "argmax equal", bool((cls.cpu().long() == want.argmax(1)).all()))


if __name__ == "__main__":
    main()
