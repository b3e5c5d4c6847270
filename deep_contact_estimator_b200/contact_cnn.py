"""``contact_cnn`` — the reference's nn.Module surface over the B200 kernels.

Same constructor (no arguments), same submodule names ``block1`` / ``block2`` /
``fc`` and therefore the same 14 ``state_dict`` keys as
/root/reference/src/contact_cnn.py:7-58, so reference checkpoints load
unchanged (``model.load_state_dict(ckpt['model_state_dict'])``,
src/inference_one_seq.py:153-156).

``forward`` (src/contact_cnn.py:60-66):
  * CUDA input, eval mode, no grad  -> libdce_b200.so (one C-ABI call);
    raises if the library is missing — there is no silent fallback on a GPU;
  * training / grad enabled / CPU   -> the stock PyTorch layers (the reference's
    own path; src/train.py:100 needs autograd).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from .engine import ContactEngine, default_precision
from .synth import PARAM_NAMES, WINDOW, CHANNELS


class contact_cnn(nn.Module):
    def __init__(self):
        super().__init__()
        # layer order and hyper-parameters: src/contact_cnn.py:10-58
        self.block1 = nn.Sequential(
            nn.Conv1d(54, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(),
            nn.Conv1d(64, 64, kernel_size=3, stride=1, padding=1), nn.ReLU(),
            nn.Dropout(p=0.5), nn.MaxPool1d(kernel_size=2, stride=2))
        self.block2 = nn.Sequential(
            nn.Conv1d(64, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(),
            nn.Conv1d(128, 128, kernel_size=3, stride=1, padding=1), nn.ReLU(),
            nn.Dropout(p=0.5), nn.MaxPool1d(kernel_size=2, stride=2))
        self.fc = nn.Sequential(
            nn.Linear(4736, 2048), nn.ReLU(), nn.Dropout(p=0.5),
            nn.Linear(2048, 512), nn.ReLU(), nn.Dropout(p=0.5),
            nn.Linear(512, 16))
        self.precision: Optional[str] = None        # None -> $DCE_PRECISION or "bf16x3"
        self._engine: Optional[ContactEngine] = None
        self._engine_key = None

    def __getstate__(self):
        # copy.deepcopy / pickling of the whole module: the engine wraps a device handle of THIS process and is
        # rebuilt lazily from the parameters, so it is never part of the module's state
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_key"] = None
        return state

    # -- stock path (training, CPU) ------------------------------------------
    def _forward_torch(self, x: torch.Tensor) -> torch.Tensor:
        x = x.permute(0, 2, 1)
        x = self.block2(self.block1(x))
        return self.fc(x.view(x.shape[0], -1))

    # -- native path ---------------------------------------------------------------
    def _native_ok(self, x: torch.Tensor) -> bool:
        if self.training or torch.is_grad_enabled():
            return False
        if os.environ.get("DCE_BACKEND", "b200") == "torch":     # explicit opt-out, never implicit
            return False
        return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3
                and tuple(x.shape[1:]) == (WINDOW, CHANNELS))

    def _params_key(self, device):
        sd = [self.get_parameter(k) for k in PARAM_NAMES]
        return (str(device), self.precision or default_precision(),
                tuple((p.data_ptr(), p._version) for p in sd))

    def engine(self, device=None) -> ContactEngine:
        """The packed-weight engine for this module's current parameters; rebuilt
        when a parameter is replaced or modified in place (load_state_dict,
        .to(), optimizer steps bump ``_version``)."""
        device = torch.device(device) if device is not None else self.get_parameter(PARAM_NAMES[0]).device
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = self._params_key(device)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = ContactEngine({k: self.get_parameter(k) for k in PARAM_NAMES}, device,
                                         self.precision or default_precision())
            self._engine_key = key
        return self._engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._native_ok(x):
            return self.engine(x.device).forward(x)
        return self._forward_torch(x)
