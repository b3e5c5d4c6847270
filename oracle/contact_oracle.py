"""CPU oracle for the contact-classification hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, what the reference computes on the path
SURVEY.md §8(a) scopes: window extraction + z-score, ``contact_cnn.forward``,
argmax, ``decimal2binary``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the
product package (``deep_contact_estimator_b200``) never does and fails loudly
when its CUDA library is missing.

Where the arithmetic lives: the reference's own arithmetic is PyTorch/ATen
(third-party, not under /root/reference; pinned only through the Docker base
images pytorch 1.6.0 / 1.8.0, docker/cuda10_1/Dockerfile:1,
docker/cuda11_1/Dockerfile:1).  The primary oracle therefore calls the same
ATen CPU ops through ``torch.nn.functional`` (torch 2.11 in this image), and a
second, independent restatement in numpy float64 acts as the arbiter when two
fp32 results disagree in the last bits.

Parity pinning: the reference has no tests, golden vectors or fixtures for this
path.  The oracle is pinned instead against outputs of the
reference itself run in the build container: ``oracle/make_golden.py`` imports
``contact_cnn`` / ``contact_dataset`` / ``inference`` from /root/reference,
checks this file reproduces them bit-for-bit there, and commits small logits /
bits fixtures under ``tests/golden/``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

WINDOW = 150
CHANNELS = 54


# --------------------------------------------------------------------------
# torch fp32 restatement (same ATen kernels the reference module dispatches to)
# --------------------------------------------------------------------------

def forward_torch(params, x: torch.Tensor) -> torch.Tensor:
    """``contact_cnn.forward`` (src/contact_cnn.py:60-66) in eval mode.

    ``params``: mapping with the reference's 14 state_dict keys.
    ``x``: ``(B, 150, 54)`` -> logits ``(B, 16)``.  Dropout layers
    (src/contact_cnn.py:23,41,51,55) are identity under ``model.eval()``
    (src/inference_one_seq.py:156).
    """
    p = params
    h = x.permute(0, 2, 1)                                                        # :61
    h = F.relu(F.conv1d(h, p["block1.0.weight"], p["block1.0.bias"], padding=1))  # :11-16
    h = F.relu(F.conv1d(h, p["block1.2.weight"], p["block1.2.bias"], padding=1))  # :17-22
    h = F.max_pool1d(h, kernel_size=2, stride=2)                                  # :24-25  150 -> 75
    h = F.relu(F.conv1d(h, p["block2.0.weight"], p["block2.0.bias"], padding=1))  # :29-34
    h = F.relu(F.conv1d(h, p["block2.2.weight"], p["block2.2.bias"], padding=1))  # :35-40
    h = F.max_pool1d(h, kernel_size=2, stride=2)                                  # :42-43  75 -> 37 (floor)
    h = h.reshape(h.shape[0], -1)                                                 # :64  index = c*37 + t
    h = F.relu(F.linear(h, p["fc.0.weight"], p["fc.0.bias"]))                     # :48-50
    h = F.relu(F.linear(h, p["fc.3.weight"], p["fc.3.bias"]))                     # :52-54
    return F.linear(h, p["fc.6.weight"], p["fc.6.bias"])                          # :56-57


def normalize_window(w: torch.Tensor) -> torch.Tensor:
    """One window's z-score: minus column mean, divided by the UNBIASED column
    std, no epsilon (utils/data_handler.py:55-56)."""
    return (w - torch.mean(w, dim=0)) / torch.std(w, dim=0)


def extract_windows(data: torch.Tensor, first: int, count: int, window: int = WINDOW) -> torch.Tensor:
    """Windows ``first .. first+count`` of a ``(T, 54)`` log, stride 1, each
    normalised independently (utils/data_handler.py:24,55) and stacked as
    DataLoader's default collate does (src/inference_one_seq.py:151)."""
    return torch.stack([normalize_window(data[i:i + window, :]) for i in range(first, first + count)])


def window_label(labels: torch.Tensor, idx: int, window: int = WINDOW) -> torch.Tensor:
    """Label of window ``idx`` is the label of its LAST row (utils/data_handler.py:57)."""
    return labels[idx + window - 1]


def num_windows(steps: int, window: int = WINDOW) -> int:
    """utils/data_handler.py:24."""
    return steps - window + 1


def argmax_class(logits: torch.Tensor) -> torch.Tensor:
    """``_, prediction = torch.max(output, 1)`` (src/inference_one_seq.py:26)."""
    return torch.max(logits, 1)[1]


def decimal2binary(x: torch.Tensor) -> torch.Tensor:
    """Class 0..15 -> 4 contact bits, MSB first: bit3 -> leg 0 (RF) ... bit0 ->
    leg 3 (LH) (src/inference_one_seq.py:59-62, utils/mat2numpy.py:33-46)."""
    mask = 2 ** torch.arange(3, -1, -1).to(x.device, x.dtype)
    return x.unsqueeze(-1).bitwise_and(mask).ne(0).byte()


def inference_stream(params, data: torch.Tensor, first: int = 0, count: int | None = None,
                     batch_size: int = 256, window: int = WINDOW):
    """``inference()`` (src/inference_one_seq.py:19-30) over a resident log:
    returns ``(logits (N,16) f32, cls (N,) i64, bits (N,4) u8)``."""
    n = num_windows(data.shape[0], window) - first if count is None else count
    outs = []
    with torch.no_grad():
        for s in range(first, first + n, batch_size):
            c = min(batch_size, first + n - s)
            outs.append(forward_torch(params, extract_windows(data, s, c, window)))
    logits = torch.cat(outs, 0) if outs else torch.empty(0, 16)
    cls = argmax_class(logits) if n else torch.empty(0, dtype=torch.int64)
    return logits, cls, decimal2binary(cls)


# --------------------------------------------------------------------------
# numpy float64 restatement (independent of ATen; the arbiter)
# --------------------------------------------------------------------------

def _np(p):
    return {k: np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float64)
            for k, v in p.items()}


def _conv1d_k3_relu_np(h: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Conv1d(kernel 3, stride 1, zero padding 1) + ReLU; ``h`` is (B, Cin, T),
    ``w`` is (Cout, Cin, 3): out[b,o,t] = b[o] + sum_{c,k} w[o,c,k] * h[b,c,t+k-1]."""
    B, _, T = h.shape
    hp = np.zeros((B, h.shape[1], T + 2), dtype=np.float64)
    hp[:, :, 1:T + 1] = h
    out = np.broadcast_to(b[None, :, None], (B, w.shape[0], T)).copy()
    for k in range(3):
        out += np.einsum("oc,bct->bot", w[:, :, k], hp[:, :, k:k + T])
    return np.maximum(out, 0.0)


def _maxpool2_np(h: np.ndarray) -> np.ndarray:
    """MaxPool1d(2, 2): floor(T/2) outputs, a trailing odd sample is dropped."""
    T2 = h.shape[2] // 2
    return np.maximum(h[:, :, 0:2 * T2:2], h[:, :, 1:2 * T2:2])


def forward_numpy64(params, x) -> np.ndarray:
    """src/contact_cnn.py:60-66 in float64 with explicit loops over taps."""
    p = _np(params)
    h = np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)
    h = h.transpose(0, 2, 1)
    h = _conv1d_k3_relu_np(h, p["block1.0.weight"], p["block1.0.bias"])
    h = _conv1d_k3_relu_np(h, p["block1.2.weight"], p["block1.2.bias"])
    h = _maxpool2_np(h)
    h = _conv1d_k3_relu_np(h, p["block2.0.weight"], p["block2.0.bias"])
    h = _conv1d_k3_relu_np(h, p["block2.2.weight"], p["block2.2.bias"])
    h = _maxpool2_np(h)
    h = h.reshape(h.shape[0], -1)
    h = np.maximum(h @ p["fc.0.weight"].T + p["fc.0.bias"], 0.0)
    h = np.maximum(h @ p["fc.3.weight"].T + p["fc.3.bias"], 0.0)
    return h @ p["fc.6.weight"].T + p["fc.6.bias"]


def normalize_window_numpy64(w) -> np.ndarray:
    w = np.asarray(w, dtype=np.float64)
    return (w - w.mean(axis=0)) / w.std(axis=0, ddof=1)


def decimal2binary_numpy(x) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64)
    return ((x[..., None] & np.array([8, 4, 2, 1])) != 0).astype(np.uint8)


# --------------------------------------------------------------------------
# the parity metric (norm-wise, per window)
# --------------------------------------------------------------------------

def normwise_rel_err(got, want) -> float:
    """max over windows of max|got - want| / max|want|: element-wise relative
    error is ill-posed because logits cross zero."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if want.size == 0:
        return 0.0
    num = np.abs(got - want).max(axis=-1)
    den = np.abs(want).max(axis=-1)
    return float((num / den).max())
