#!/usr/bin/env python
"""CPU emulation of the index arithmetic of csrc/dce_latency.cuh (pack layouts + phases A-E), checked against
the oracle.  No GPU: this pins the layout formulas, not the synchronisation.   python tools/emulate_latency.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_contact_estimator_b200 import synth
from oracle import contact_oracle as oracle

params = synth.make_params(0)
P = {k: v.detach().numpy().astype(np.float32) for k, v in params.items()}
conv = lambda w: np.ascontiguousarray(w.transpose(2, 1, 0)).reshape(-1)          # wp[tap][cin][cout]
w1, w2, w3, w4 = (conv(P[k]) for k in ("block1.0.weight", "block1.2.weight", "block2.0.weight", "block2.2.weight"))
kp = np.arange(4736); t, c = kp // 128, kp % 128
f1p = np.ascontiguousarray(P["fc.0.weight"][:, c * 37 + t].T)                     # [k'][n]
f2p = np.ascontiguousarray(P["fc.3.weight"].T)                                    # [k][n]
f3t = np.ascontiguousarray(P["fc.6.weight"].T)                                    # [512][16]
# pack_w4q_kernel
i = np.arange(4 * 384 * 32); o, k, q = i & 31, (i >> 5) % 384, i // (384 * 32)
w4q = w4[k * 128 + q * 32 + o]
# pack_f1s_kernel (float4 granularity)
i = np.arange(128 * 4736 * 4); sl = i // (4736 * 4); r = i % (4736 * 4); s = r // 2048; r = r - s * 2048
rows = np.where(s < 9, 512, 128); j, row = r // rows, r % rows; k = s * 512 + row
f1s = f1p.reshape(-1)[((k * 2048 + sl * 16 + j * 4)[:, None] + np.arange(4)[None, :])].reshape(-1)
i = np.arange(128 * 2048); sl, row = i // 2048, i % 2048
f2s = f2p.reshape(-1)[((row * 512 + sl * 4)[:, None] + np.arange(4)[None, :])].reshape(-1)

relu = lambda v: np.maximum(v, 0)
def conv_rows(xflat, w, CIN, CT, R):            # out[r][o] = sum_k xflat[r*CIN + k] * w[k*CT + o]
    W = w.reshape(3 * CIN, CT)
    return np.stack([xflat[r * CIN: r * CIN + 3 * CIN] @ W for r in range(R)])

def run(x, stream=False, first=0, B=1):
    p1 = np.zeros((B, 75, 64), np.float32); a4 = np.zeros((B, 4736), np.float32)
    for it in range(B * 75):
        b, tp = divmod(it, 75)
        xw = x[first + b: first + b + 150] if stream else x[b]
        if stream:
            mean = xw.mean(0); std = np.sqrt(((xw - mean) ** 2).sum(0) / 149)
        xin = np.zeros(6 * 54, np.float32)
        for i in range(324):
            j, c = divmod(i, 54); row = 2 * tp - 2 + j
            if 0 <= row < 150:
                v = xw[row, c]
                xin[i] = (v - mean[c]) / std[c] if stream else v
        c1 = relu(conv_rows(xin, w1, 54, 64, 4) + P["block1.0.bias"])
        for r in range(4):
            if not 0 <= 2 * tp - 1 + r < 150: c1[r] = 0
        c2 = relu(conv_rows(c1.reshape(-1), w2, 64, 64, 2) + P["block1.2.bias"])
        p1[b, tp] = c2.max(0)
    for it in range(B * 37):
        b, t = divmod(it, 37)
        pin = np.zeros((6, 64), np.float32)
        for j in range(6):
            row = 2 * t - 2 + j
            if 0 <= row < 75: pin[j] = p1[b, row]
        c3 = relu(conv_rows(pin.reshape(-1), w3, 64, 128, 4) + P["block2.0.bias"])
        for r in range(4):
            if not 0 <= 2 * t - 1 + r < 75: c3[r] = 0
        for q in range(4):
            c4 = relu(conv_rows(c3.reshape(-1), w4q[q * 12288:(q + 1) * 12288], 128, 32, 2) + P["block2.2.bias"][q * 32:(q + 1) * 32])
            a4[b, t * 128 + q * 32: t * 128 + q * 32 + 32] = c4.max(0)
    h1 = np.zeros((B, 2048), np.float32); h2 = np.zeros((B, 512), np.float32); lg = np.zeros((B, 16), np.float32)
    for b in range(B):
        for cta in range(128):
            acc = np.zeros(16, np.float64)
            for s in range(10):
                rows = 512 if s < 9 else 128
                st = f1s[cta * 4736 * 16 + s * 512 * 16: cta * 4736 * 16 + s * 512 * 16 + rows * 16].reshape(4, rows, 4)   # [j][row][e]
                xk = a4[b, s * 512: s * 512 + rows]
                acc += np.einsum("r,jre->je", xk.astype(np.float64), st.astype(np.float64)).reshape(16)
            h1[b, cta * 16: cta * 16 + 16] = relu(acc + P["fc.0.bias"][cta * 16: cta * 16 + 16])
        for cta in range(128):
            st = f2s[cta * 8192:(cta + 1) * 8192].reshape(2048, 4)
            h2[b, cta * 4: cta * 4 + 4] = relu(h1[b].astype(np.float64) @ st.astype(np.float64) + P["fc.3.bias"][cta * 4: cta * 4 + 4])
        lg[b] = h2[b].astype(np.float64) @ f3t.astype(np.float64) + P["fc.6.bias"]
    return lg

def check(batch: int = 2):
    """-> (batch-mode error, stream-mode error), norm-wise relative to the oracle"""
    x = synth.make_windows(batch, seed=1)
    want = oracle.forward_torch(params, x).detach().numpy()
    e_batch = oracle.normwise_rel_err(run(x.numpy(), B=batch), want)
    log = synth.make_sensor_log(150 + batch, seed=2)
    got = run(log.numpy(), stream=True, first=1, B=batch)
    ds = torch.stack([(log[i:i + 150] - log[i:i + 150].mean(0)) / log[i:i + 150].std(0) for i in range(1, 1 + batch)])
    e_stream = oracle.normwise_rel_err(got, oracle.forward_torch(params, ds).detach().numpy())
    return e_batch, e_stream


if __name__ == "__main__":
    eb, es = check()
    print("batch  err", eb)
    print("stream err", es)
