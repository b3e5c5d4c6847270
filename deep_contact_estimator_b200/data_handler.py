"""``contact_dataset`` — same surface as /root/reference/utils/data_handler.py:13-61.

The whole log lives on ``device`` as float32 ``(T, 54)`` (``self.data``) with
int64 labels (``self.label``); ``__getitem__`` keeps the reference semantics
(used by the training loop and by per-batch DataLoader iteration), while the
fast path in ``inference.py`` hands ``self.data`` to the streaming kernel in a
single call instead of iterating.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset


def ingest_log(data, device, chunk_rows: int = 1 << 18) -> torch.Tensor:
    """``torch.FloatTensor(np.load(data_path)).to(device)`` (utils/data_handler.py:21-26) for a CUDA device
    without the host-side float64 -> float32 pass: the log is uploaded in chunks through two pinned staging
    buffers and, when it is float64 (the format utils/mat2numpy.py:73,80 writes), converted on the device by
    ``dce_ingest_f64`` while the next chunk is in flight.  Same values as the host cast (round to nearest)."""
    device = torch.device(device)
    t = torch.as_tensor(data)
    if device.type != "cuda" or t.dim() != 2 or t.dtype not in (torch.float64, torch.float32) or t.numel() == 0:
        return t.to(torch.float32).to(device)
    import ctypes
    from . import _lib
    lib = _lib.load()
    t = t.contiguous()
    rows, ch = t.shape
    out = torch.empty((rows, ch), dtype=torch.float32, device=device)
    is64 = t.dtype == torch.float64
    chunk_rows = max(1, min(chunk_rows, rows))
    stage_h = [torch.empty((chunk_rows, ch), dtype=t.dtype).pin_memory() for _ in range(2)]
    stage_d = [torch.empty((chunk_rows, ch), dtype=t.dtype, device=device) for _ in range(2)] if is64 else None
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device)
        done = [torch.cuda.Event() for _ in range(2)]
        for i, r0 in enumerate(range(0, rows, chunk_rows)):
            r1 = min(rows, r0 + chunk_rows)
            k = i & 1
            done[k].synchronize()                          # the copy that last used this staging pair has finished
            stage_h[k][: r1 - r0].copy_(t[r0:r1])          # pageable -> pinned on the host, overlaps the previous chunk's DMA
            if is64:
                stage_d[k][: r1 - r0].copy_(stage_h[k][: r1 - r0], non_blocking=True)
                rc = lib.dce_ingest_f64(ctypes.c_void_p(stage_d[k].data_ptr()), ctypes.c_void_p(out[r0:].data_ptr()),
                                        (r1 - r0) * ch, ctypes.c_void_p(stream.cuda_stream))
                _lib.check(rc, "dce_ingest_f64")
            else:
                out[r0:r1].copy_(stage_h[k][: r1 - r0], non_blocking=True)
            done[k].record(stream)
        stream.synchronize()
    return out


class contact_dataset(Dataset):
    def __init__(self, data_path=None, label_path=None, window_size=150, device="cuda", data=None, label=None):
        # utils/data_handler.py:21-27; `data=` / `label=` accept in-memory arrays (tests, synthetic logs)
        if data is None:
            data = np.load(data_path)
        if label is None:
            label = np.load(label_path)
        label = torch.as_tensor(label)
        self.num_data = data.shape[0] - window_size + 1
        self.window_size = window_size
        self.data = ingest_log(data, device)                # float64 .npy -> float32 on `device`
        self.label = label.to(torch.int64).to(device)

    def __len__(self):
        return self.num_data

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        w = self.data[idx:idx + self.window_size, :]
        # unbiased std, no epsilon: utils/data_handler.py:55-56
        this_data = (w - torch.mean(w, dim=0)) / torch.std(w, dim=0)
        this_label = self.label[idx + self.window_size - 1]          # :57
        return {"data": this_data, "label": this_label}
