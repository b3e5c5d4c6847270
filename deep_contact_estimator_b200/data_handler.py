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


class contact_dataset(Dataset):
    def __init__(self, data_path=None, label_path=None, window_size=150, device="cuda", data=None, label=None):
        # utils/data_handler.py:21-27; `data=` / `label=` accept in-memory arrays (tests, synthetic logs)
        if data is None:
            data = np.load(data_path)
        if label is None:
            label = np.load(label_path)
        data = torch.as_tensor(data)
        label = torch.as_tensor(label)
        self.num_data = data.shape[0] - window_size + 1
        self.window_size = window_size
        self.data = data.to(torch.float32).to(device)
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
