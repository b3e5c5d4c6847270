"""CPU study: which operand splittings keep contact_cnn's logits within the 1e-4 parity bar?

The tensor-core path stores every fp32 operand as bf16 hi + bf16 lo and issues three MMAs per K-step
(DESIGN.md §3), which caps the step at 1/3 of the bf16 peak.  This script emulates cheaper splittings on the
CPU — operands rounded exactly as the hardware formats would round them, products and sums in float64 so that
only the OPERAND rounding is measured (fp32 accumulation adds ~1e-7) — and reports the norm-wise logit error
against the float64 forward of the same weights, plus the number of argmax flips:

    bf16x3        xh*wh + xh*wl + xl*wh, all bf16                              3 bf16 passes   (today)
    f16x1         fp16(x) * fp16(w)                                            1 pass
    f16x3         as bf16x3 with fp16 parts                                    3 passes
    f16+e4m3      fp16 main + two correction terms xl8*w8 + x8*wl8 whose       1 + 2 * 1/2 = 2 passes
                  operands are e4m3 with STATIC power-of-two scales (2^12,
                  2^3 sw, 2^1, 2^14 sw; sw = the layer's weight scale): both
                  products carry 2^15, undone by scale-input-d = 15
    f16+e5m2      same with e5m2                                               2 passes
    f16+mxe4m3    same with e4m3 and one power-of-two scale per 32 K-elements  2 passes
                  (kind::mxf8f6f4 block scaling)
    f16+e2m1mx    corrections in block-scaled fp4 (e2m1)                       1 + 2 * 1/4 = 1.5 passes

fp8 MMAs run at twice the bf16 rate on sm_100a, so "fp16 main + fp8 corrections" costs two bf16-pass equivalents
instead of three.  Usage:  python tools/emulate_split_precision.py [--windows 64] [--logit-scale 1]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deep_contact_estimator_b200 import synth  # noqa: E402

D = torch.float64


def rnd(t: torch.Tensor, dtype) -> torch.Tensor:
    return t.to(torch.float32).to(dtype).to(D)


def rnd_scaled(t, dtype, scale):
    sat = {torch.float8_e4m3fn: 448.0, torch.float8_e5m2: 57344.0}.get(dtype, 65504.0)
    return rnd((t * scale).clamp(-sat, sat), dtype) / scale


def pow2_floor(v: torch.Tensor) -> torch.Tensor:
    return torch.exp2(torch.floor(torch.log2(v.clamp_min(1e-300))))


def e2m1(t: torch.Tensor) -> torch.Tensor:
    """Round to the fp4 e2m1 grid {0, .5, 1, 1.5, 2, 3, 4, 6} (nearest, saturating)."""
    grid = torch.tensor([0., .5, 1., 1.5, 2., 3., 4., 6.], dtype=D)
    a = t.abs().clamp_max(6.0)
    idx = (a.unsqueeze(-1) - grid).abs().argmin(-1)
    return torch.sign(t) * grid[idx]


def block_q(t: torch.Tensor, kdim: int, fmt: str) -> torch.Tensor:
    """One power-of-two scale per 32 consecutive elements along `kdim` (mx block scaling), element format `fmt`."""
    t = t.movedim(kdim, -1)
    k = t.shape[-1]
    pad = (-k) % 32
    tp = F.pad(t, (0, pad)).reshape(*t.shape[:-1], -1, 32)
    amax = tp.abs().amax(-1, keepdim=True)
    top = {"e4m3": 256.0, "e5m2": 32768.0, "e2m1": 4.0}[fmt]           # largest binade of the element format
    scale = pow2_floor(amax) / top
    scale = torch.where(amax > 0, scale, torch.ones_like(scale))
    if fmt == "e2m1":
        q = e2m1(tp / scale) * scale
    else:
        sat = {"e4m3": 448.0, "e5m2": 57344.0}[fmt]                     # saturating conversion, as cvt.satfinite
        q = rnd((tp / scale).clamp(-sat, sat), {"e4m3": torch.float8_e4m3fn, "e5m2": torch.float8_e5m2}[fmt]) * scale
    q = q.reshape(*t.shape[:-1], -1)[..., :k]
    return q.movedim(-1, kdim)


class Scheme:
    """product(x, w, op, kx, kw): op(x', w') summed over the scheme's operand pairs; kx / kw = the K dim of x / w."""

    def __init__(self, name):
        self.name = name

    def __call__(self, x, w, op, kx, kw):
        n = self.name
        if n == "f64":
            return op(x, w)
        if n in ("bf16x3", "f16x3"):
            dt = torch.bfloat16 if n == "bf16x3" else torch.float16
            xh, wh = rnd(x, dt), rnd(w, dt)
            xl, wl = rnd(x - xh, dt), rnd(w - wh, dt)
            return op(xh, wl) + op(xl, wh) + op(xh, wh)
        if n == "bf16x1":
            return op(rnd(x, torch.bfloat16), rnd(w, torch.bfloat16))
        # fp16 main term; the weight is pre-scaled into fp16's normal range by a per-layer power of two
        sw = 1.0 / pow2_floor(w.abs().max())
        x = x.clamp(-65504.0, 65504.0)                                   # split16_f16f8 saturates activations at the fp16 maximum
        xh, wh = rnd(x, torch.float16), rnd(w * sw, torch.float16) / sw
        main = op(xh, wh)
        if n == "f16x1":
            return main
        xl, wl = x - xh, w - wh
        if n in ("f16+e4m3", "f16+e5m2"):
            # STATIC power-of-two operand scales, the same for every layer and input (tools/microbench/umma_f16f8.cu):
            # both correction products carry 2^15 and are scaled back by the first fp16 MMA's scale-input-d
            dt = torch.float8_e4m3fn if n == "f16+e4m3" else torch.float8_e5m2
            xl8 = rnd_scaled(xl, dt, 2.0 ** 12)
            w8 = rnd_scaled(w, dt, sw * 2.0 ** 3)
            x8 = rnd_scaled(x, dt, 2.0 ** 1)
            wl8 = rnd_scaled(wl, dt, sw * 2.0 ** 14)
            return main + op(xl8, w8) + op(x8, wl8)
        if n in ("f16+mxe4m3", "f16+mxe5m2", "f16+e2m1mx"):
            fmt = {"f16+mxe4m3": "e4m3", "f16+mxe5m2": "e5m2", "f16+e2m1mx": "e2m1"}[n]
            return main + op(block_q(xl, kx, fmt), block_q(w, kw, fmt)) + op(block_q(x, kx, fmt), block_q(wl, kw, fmt))
        raise ValueError(n)


def forward(params, x, scheme: Scheme):
    """contact_cnn.forward (/root/reference/src/contact_cnn.py:60-66), float64 accumulation, operands per `scheme`."""
    p = {k: v.to(D) for k, v in params.items()}
    h = x.to(D).permute(0, 2, 1)

    def conv(h, name):
        y = scheme(h, p[name + ".weight"], lambda a, b: F.conv1d(a, b, None, padding=1), 1, 1)
        return F.relu(y + p[name + ".bias"][None, :, None])

    def lin(h, name, relu=True, sch=scheme):
        y = sch(h, p[name + ".weight"], lambda a, b: a @ b.T, 1, 1) + p[name + ".bias"]
        return F.relu(y) if relu else y

    h = conv(h, "block1.0")
    h = conv(h, "block1.2")
    h = F.max_pool1d(h, 2, 2)
    h = conv(h, "block2.0")
    h = conv(h, "block2.2")
    h = F.max_pool1d(h, 2, 2)
    # the kernels' fc.0 K order is k' = t*128 + c, so a 32-element scale block is 32 channels at one t
    B = h.shape[0]
    hk = h.permute(0, 2, 1).reshape(B, -1)
    w0 = p["fc.0.weight"].reshape(2048, 128, 37).permute(0, 2, 1).reshape(2048, -1)
    y = scheme(hk, w0, lambda a, b: a @ b.T, 1, 1) + p["fc.0.bias"]
    h = F.relu(y)
    h = lin(h, "fc.3")
    return lin(h, "fc.6", relu=False, sch=Scheme("f64"))              # fc.6 runs in fp32 FMAs (epilogue of fc.3)


def normwise(got, want):
    num = (got - want).abs().amax(-1)
    den = want.abs().amax(-1)
    return (num / den)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=64)
    ap.add_argument("--logit-scale", type=float, default=1.0)
    ap.add_argument("--weight-spread", type=float, default=0.0,
                    help="multiply every weight by 10**U(-s, s) per output row: trained-like dynamic range")
    ap.add_argument("--outliers", type=float, default=0.0,
                    help="multiply a random 0.1 %% of every weight tensor by this factor: heavy-tailed weights stress the "
                         "static per-layer scales of the e4m3 corrections (small weights fall below their range)")
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    params = synth.make_params(0, logit_scale=args.logit_scale)
    if args.weight_spread:
        g = torch.Generator().manual_seed(7)
        for k in list(params):
            if k.endswith("weight"):
                s = 10.0 ** ((torch.rand(params[k].shape[0], generator=g) * 2 - 1) * args.weight_spread)
                params[k] = params[k] * s.reshape(-1, *([1] * (params[k].dim() - 1)))
    if args.outliers:
        g = torch.Generator().manual_seed(11)
        for k in list(params):
            if k.endswith("weight"):
                m = torch.rand(params[k].shape, generator=g) < 1e-3
                params[k] = torch.where(m, params[k] * args.outliers, params[k])
    x = synth.make_windows(args.windows, seed=1)
    want = forward(params, x, Scheme("f64"))
    print(f"{args.windows} windows, logit scale {args.logit_scale}, weight spread {args.weight_spread}, outliers x{args.outliers}; "
          f"error = max|d| / max|logit| per window")
    print(f"{'scheme':<12} {'passes':>6} {'max':>10} {'median':>10} {'argmax flips':>13}")
    for name, cost in (("bf16x1", 1), ("bf16x3", 3), ("f16x1", 1), ("f16x3", 3), ("f16+e4m3", 2), ("f16+e5m2", 2),
                       ("f16+mxe4m3", 2), ("f16+mxe5m2", 2), ("f16+e2m1mx", 1.5)):
        got = forward(params, x, Scheme(name))
        e = normwise(got, want)
        flips = int((got.argmax(-1) != want.argmax(-1)).sum())
        print(f"{name:<12} {cost:>6} {float(e.max()):>10.2e} {float(e.median()):>10.2e} {flips:>13d}")


if __name__ == "__main__":
    main()
