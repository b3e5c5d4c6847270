"""Seeded synthetic weights and sensor data for parity tests and bench.py.

The reference ships no checkpoint and no sample data, so every
test and benchmark regenerates its inputs from seeds.  Shapes and names follow
the reference: the 14 ``state_dict`` entries of ``contact_cnn``
(/root/reference/src/contact_cnn.py:10-58) and the ``(T, 54)`` sensor log that
``contact_dataset`` slides a 150-row window over
(/root/reference/utils/data_handler.py:15-27).

Everything is generated on the CPU with an explicit ``torch.Generator`` so the
same seed gives the same bits in this container and on the GPU box.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

WINDOW = 150      # config/*.yaml: window_size
CHANNELS = 54     # [q(12), qd(12), acc(3), omega(3), p(12), v(12)]  utils/mat2numpy.py:73
CLASSES = 16      # 4 legs -> 2**4 contact states

# name -> shape, in state_dict order (src/contact_cnn.py:10-58).
PARAM_SHAPES = OrderedDict([
    ("block1.0.weight", (64, 54, 3)),   ("block1.0.bias", (64,)),
    ("block1.2.weight", (64, 64, 3)),   ("block1.2.bias", (64,)),
    ("block2.0.weight", (128, 64, 3)),  ("block2.0.bias", (128,)),
    ("block2.2.weight", (128, 128, 3)), ("block2.2.bias", (128,)),
    ("fc.0.weight", (2048, 4736)),      ("fc.0.bias", (2048,)),
    ("fc.3.weight", (512, 2048)),       ("fc.3.bias", (512,)),
    ("fc.6.weight", (16, 512)),         ("fc.6.bias", (16,)),
])
PARAM_NAMES = list(PARAM_SHAPES.keys())


def make_params(seed: int = 0, logit_scale: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """A seeded ``state_dict`` with PyTorch's default init distribution.

    ``nn.Conv1d`` / ``nn.Linear`` draw weight and bias from
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)); we draw the same distribution from our
    own generator so the values do not depend on module construction order.
    ``logit_scale`` multiplies the last layer (SURVEY.md §8c asks for a second
    weight set whose logits are O(1)).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    fan_in = 1
    for name, shape in PARAM_SHAPES.items():
        if name.endswith("weight"):
            fan_in = math.prod(shape[1:])
        bound = 1.0 / math.sqrt(fan_in)
        t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
        if name.startswith("fc.6"):
            t = t * logit_scale
        out[name] = t
    return out


def zscore_windows(x: torch.Tensor) -> torch.Tensor:
    """Per-window, per-channel z-score over time with the unbiased std, exactly
    as utils/data_handler.py:55-56 does for one window."""
    return (x - x.mean(dim=1, keepdim=True)) / x.std(dim=1, keepdim=True)


def make_windows(batch: int, seed: int = 1) -> torch.Tensor:
    """``(batch, 150, 54)`` float32 z-scored synthetic windows (SURVEY.md §8d
    config 1/2): the shape ``DataLoader`` hands ``contact_cnn.forward``."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(batch, WINDOW, CHANNELS, generator=g, dtype=torch.float32)
    return zscore_windows(x).contiguous()


def make_sensor_log(steps: int, seed: int = 2) -> torch.Tensor:
    """``(steps, 54)`` float32 synthetic proprioceptive log (SURVEY.md §8d
    config 3): per-channel random walk + white noise with per-channel scale and
    offset, so no channel is constant inside a window and the z-score is
    exercised with non-zero means."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    scale = 10.0 ** (torch.rand(CHANNELS, generator=g) * 2.0 - 1.0)      # [0.1, 10]
    offset = torch.rand(CHANNELS, generator=g) * 6.0 - 3.0               # [-3, 3]
    walk = torch.cumsum(torch.randn(steps, CHANNELS, generator=g, dtype=torch.float64) * 0.01, dim=0)
    noise = torch.randn(steps, CHANNELS, generator=g, dtype=torch.float64)
    log = walk + noise * scale.double() + offset.double()
    return log.float().contiguous()


def make_labels(steps: int, seed: int = 3) -> torch.Tensor:
    """``(steps,)`` int64 contact-state labels in [0, 16) (utils/mat2numpy.py:199-200)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, CLASSES, (steps,), generator=g, dtype=torch.int64)
