"""The reference's inference / evaluation loops, same signatures, one launch.

    inference(dataloader, model, device)                 src/inference_one_seq.py:19-30
    inference_and_compute_acc(dataloader, model, device) src/inference_one_seq.py:33-57
    compute_accuracy(dataloader, model)                  src/test.py:72-107
    decimal2binary(x)                                    src/inference_one_seq.py:59-62

When the dataloader wraps a ``contact_dataset`` whose log is resident on a CUDA
device and iterates it in order, the whole ``for sample in dataloader`` loop
(window extraction + z-score + forward + argmax + bits + ``torch.cat``) becomes
one ``dce_stream`` call over the log.  Otherwise (custom datasets, shuffling
samplers, CPU) the loop runs per batch: on CUDA each batch is one
``dce_forward`` call, on CPU the stock module runs, as in the reference.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import SequentialSampler

from .contact_cnn import contact_cnn
from .synth import WINDOW


def decimal2binary(x: torch.Tensor) -> torch.Tensor:
    """Class 0..15 -> 4 contact bits, MSB first (bit3 = leg 0 RF ... bit0 = leg 3 LH)."""
    mask = 2 ** torch.arange(4 - 1, -1, -1).to(x.device, x.dtype)
    return x.unsqueeze(-1).bitwise_and(mask).ne(0).byte()


def _stream_source(dataloader, model):
    """The dataset's resident log if the one-launch path applies, else None."""
    ds = getattr(dataloader, "dataset", None)
    data = getattr(ds, "data", None)
    if not isinstance(model, contact_cnn) or model.training:
        return None
    if data is None or not torch.is_tensor(data) or not data.is_cuda or data.dim() != 2:
        return None
    if getattr(ds, "window_size", None) != WINDOW or data.dtype != torch.float32:
        return None
    if not isinstance(getattr(dataloader, "sampler", None), SequentialSampler):
        return None
    if getattr(dataloader, "drop_last", False):
        return None
    if next(model.parameters()).device != data.device:
        return None
    return ds


def _classify_loader(dataloader, model, want_labels: bool):
    """-> (cls int64 (N,), bits u8 (N,4), labels int64 (N,) or None), all on device."""
    ds = _stream_source(dataloader, model)
    if ds is not None:
        eng = model.engine(ds.data.device)
        _, cls, bits = eng.stream(ds.data, 0, len(ds), want_logits=False)
        labels = ds.label.reshape(-1)[WINDOW - 1:WINDOW - 1 + len(ds)] if want_labels else None   # utils/data_handler.py:57
        return cls.long(), bits, labels
    cls_all, bits_all, lab_all = [], [], []
    with torch.no_grad():
        for sample in dataloader:
            x = sample["data"]
            if x.is_cuda and isinstance(model, contact_cnn) and not model.training:
                _, cls, bits = model.engine(x.device).classify(x.float(), want_logits=False)
                cls = cls.long()
            else:
                out = model(x)
                _, cls = torch.max(out, 1)
                bits = decimal2binary(cls)
            cls_all.append(cls); bits_all.append(bits)
            if want_labels:
                lab_all.append(sample["label"].reshape(-1))
    dev = cls_all[0].device if cls_all else torch.device("cpu")
    cls = torch.cat(cls_all) if cls_all else torch.empty(0, dtype=torch.int64, device=dev)
    bits = torch.cat(bits_all) if bits_all else torch.empty(0, 4, dtype=torch.uint8, device=dev)
    labels = (torch.cat(lab_all) if lab_all else torch.empty(0, dtype=torch.int64, device=dev)) if want_labels else None
    return cls, bits, labels


def inference(dataloader, model, device):
    """-> ``(N,4)`` uint8 contact bits on ``device`` (src/inference_one_seq.py:19-30)."""
    _, bits, _ = _classify_loader(dataloader, model, want_labels=False)
    return bits.to(device)


def _counts(model, cls, labels):
    if cls.is_cuda and isinstance(model, contact_cnn):
        c = model.engine(cls.device).accuracy_counts(cls, labels).cpu().numpy()
        return int(c[0]), c[1:5].astype(np.float64)
    bin_pred, bin_gt = decimal2binary(cls), decimal2binary(labels)
    return int((cls == labels).sum().item()), (bin_pred == bin_gt).sum(dim=0).cpu().numpy().astype(np.float64)


def inference_and_compute_acc(dataloader, model, device):
    """-> ``(bits (N,4) u8, class accuracy, per-leg accuracy[4])`` (src/inference_one_seq.py:33-57).

    Labels are flattened before comparison: with the ``(T,1)`` label files
    ``utils/mat2numpy.py:76`` writes, the reference's ``prediction==gt_label``
    broadcasts ``(B,)`` against ``(B,1)`` and reports class accuracies above 1
    (SURVEY.md §8a note E); per-leg numbers are identical.
    """
    cls, bits, labels = _classify_loader(dataloader, model, want_labels=True)
    n = cls.numel()
    correct, per_leg = _counts(model, cls, labels)
    return bits.to(device), correct / n, per_leg / n


def compute_accuracy(dataloader, model):
    """-> ``(acc, acc_per_leg[4], bin_pred_arr, bin_gt_arr, pred_arr, gt_arr)`` with the numpy
    array types the reference's vstack/hstack accumulation yields (src/test.py:72-107)."""
    cls, bits, labels = _classify_loader(dataloader, model, want_labels=True)
    n = cls.numel()
    correct, per_leg = _counts(model, cls, labels)
    bin_gt = decimal2binary(labels)
    return (correct / n, per_leg / n,
            bits.cpu().numpy().astype(np.float64), bin_gt.cpu().numpy().astype(np.float64),
            cls.cpu().numpy().astype(np.float64), labels.cpu().numpy().astype(np.float64))
