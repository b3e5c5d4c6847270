"""Real-time front end of the latency path: the in-tree analogue of the upstream
TensorRT runner (/root/reference/README.md:67-83), without its transport.

One call per control tick: the newest `leg_control_data_lcmt` and `microstrain_lcmt` messages go in, a
`contact_t` message comes out.  Inside, with a :class:`ContactEngine`: the 54-channel row `[q, qd, acc, omega, p, v]`
(utils/mat2numpy.py:73) goes to a :class:`RowRunner` — 216 bytes into pinned memory; the resident kernel keeps
the 150-row ring on the device, z-scores the window (utils/data_handler.py:55-56) and returns class and contact
bits, with no CUDA call and no host arithmetic per tick.  (With an injected `runner` — tests use a CPU stand-in —
the ring and the z-score are kept on the host with the reference's own expression.)  liblcm is not
installed in this image, so the subscribe / publish calls stay with the caller; message bytes use the
reference's wire format (`lcm_wire`, pinned to the generated encoders by tests/golden/lcm_bytes.npz).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import lcm_wire
from .synth import WINDOW, CHANNELS


class RealtimeContactEstimator:
    def __init__(self, engine=None, runner=None, window: int = WINDOW):
        """``engine``: a :class:`ContactEngine` (a ``LatencyRunner`` is created on it); or pass any ``runner`` with
        ``x_host`` ((1,150,54) float32) and ``step() -> (cls, bits)`` (tests inject a CPU stand-in)."""
        if window != WINDOW:
            raise ValueError(f"the kernels are built for window_size {WINDOW} (config/*.yaml)")
        self.rows = None
        if runner is None:
            if engine is None:
                raise ValueError("pass a ContactEngine or a runner")
            self.rows = engine.row_runner()
        self.runner = runner
        # double-length ring: row i is written at i and i + 150, so the newest 150 rows are always one contiguous slice
        self._ring = torch.zeros((2 * WINDOW, CHANNELS), dtype=torch.float32)
        self._pos = 0
        self.rows_seen = 0

    @property
    def ready(self) -> bool:
        return self.rows_seen >= WINDOW

    def push_row(self, row: Sequence[float]) -> Optional[Tuple[int, Tuple[int, int, int, int]]]:
        """Append one 54-vector; once 150 rows are in: ``(class 0..15, (RF, LF, RH, LH) contact bits)``."""
        r = torch.as_tensor(row, dtype=torch.float32).reshape(CHANNELS)
        if self.rows is not None:                                    # ring + z-score + classification on the device
            out = self.rows.push(r)
            self.rows_seen += 1
            return out if self.ready else None
        self._ring[self._pos] = r
        self._ring[self._pos + WINDOW] = r
        self._pos = (self._pos + 1) % WINDOW
        self.rows_seen += 1
        if not self.ready:
            return None
        w = self._ring[self._pos:self._pos + WINDOW]                 # oldest .. newest
        x = self.runner.x_host[0]
        torch.sub(w, torch.mean(w, dim=0), out=x)                    # utils/data_handler.py:55-56
        x.div_(torch.std(w, dim=0))
        cls, bits = self.runner.step()
        return int(cls[0]), tuple(int(b) for b in bits[0])

    def push_messages(self, leg_control_data: bytes, microstrain: bytes, timestamp: float = 0.0) -> Optional[bytes]:
        """The tick as the LCM thread sees it: encoded `leg_control_data_lcmt` + `microstrain_lcmt` in, encoded
        `contact_t` out (None until the first 150 ticks have arrived)."""
        q, qd, p, v, _tau = lcm_wire.decode_leg_control_data(leg_control_data)
        _quat, _rpy, omega, acc, _good, _bad = lcm_wire.decode_microstrain(microstrain)
        out = self.push_row(list(q) + list(qd) + list(acc) + list(omega) + list(p) + list(v))
        if out is None:
            return None
        return lcm_wire.encode_contact(4, timestamp, out[1])

    def close(self):
        if self.rows is not None:
            self.rows.close()
