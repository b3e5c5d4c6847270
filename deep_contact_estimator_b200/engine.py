"""ContactEngine: the host-side owner of a ``dce_weights`` handle on one B200.

PyTorch is plumbing here: it owns device memory (inputs, outputs, workspace)
and the stream; the arithmetic is in libdce_b200.so (include/dce.h).

Replaces the per-batch body of the reference loops:
    output = model(input_data); _, prediction = torch.max(output, 1);
    bin_pred = decimal2binary(prediction)
(/root/reference/src/inference_one_seq.py:25-27, src/test.py:87-94) and, for a
device-resident log, the whole ``for sample in dataloader`` loop
(src/inference_one_seq.py:23-28) including ``contact_dataset.__getitem__``
(utils/data_handler.py:55-57).
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import Mapping, Optional, Tuple

import torch

from . import _lib
from .synth import PARAM_NAMES, PARAM_SHAPES, WINDOW, CHANNELS, CLASSES


def default_precision() -> str:
    return os.environ.get("DCE_PRECISION", "bf16x3")


class ContactEngine:
    """Packed weights + workspace on one CUDA device.

    ``params``: mapping with the reference's 14 ``state_dict`` keys
    (src/contact_cnn.py:10-58); tensors may live anywhere, they are copied to
    ``device`` as fp32 and repacked once (K0).

    Streams: calls enqueue on the CURRENT torch stream and share one workspace (and, for the host-buffer entry
    points, one staging area and copy stream) per engine, as include/dce.h's "one workspace serves one call at a
    time" requires — issue the calls of one engine from one stream (or order the streams yourself); use one engine
    per concurrent stream otherwise.  ``LatencyRunner`` owns a workspace and stream of its own.
    """

    def __init__(self, params: Optional[Mapping[str, torch.Tensor]], device, precision: Optional[str] = None):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ContactEngine needs a CUDA (B200) device; CPU tensors use the stock PyTorch module")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision or default_precision()
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self._handle = ctypes.c_void_p()
        _lib.check(self.lib.dce_weights_create(ctypes.byref(self._handle), self.device.index), "dce_weights_create")
        self._workspace: Optional[torch.Tensor] = None
        self.last_launches = 0
        if params is not None:
            self.pack(params)

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.dce_weights_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- K0 ---------------------------------------------------------------
    def pack(self, params: Mapping[str, torch.Tensor]):
        missing = [k for k in PARAM_NAMES if k not in params]
        if missing:
            raise KeyError(f"state_dict is missing {missing}")
        dev = []
        for k in PARAM_NAMES:
            t = params[k].detach()
            if tuple(t.shape) != PARAM_SHAPES[k]:
                raise ValueError(f"{k}: expected shape {PARAM_SHAPES[k]}, got {tuple(t.shape)}")
            dev.append(t.to(device=self.device, dtype=torch.float32).contiguous())
        arr = (ctypes.c_void_p * len(dev))(*[t.data_ptr() for t in dev])
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            _lib.check(self.lib.dce_weights_pack(self._handle, arr, ctypes.c_void_p(stream.cuda_stream)), "dce_weights_pack")
            stream.synchronize()            # the fp32 staging copies in `dev` may be freed after this
        return self

    def packed_view(self) -> torch.Tensor:
        """The packed-weight buffer as a uint8 tensor sharing memory with the
        handle (for ``torch.distributed.broadcast`` over NCCL)."""
        n = self.lib.dce_weights_packed_bytes(self._handle)
        ptr = self.lib.dce_weights_packed_ptr(self._handle)
        iface = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        holder = type("_DcePacked", (), {"__cuda_array_interface__": iface})()
        with torch.cuda.device(self.device):
            return torch.as_tensor(holder, device=self.device)

    def broadcast_weights(self, src: int = 0, group=None):
        """One-time NCCL broadcast of the packed weights from rank ``src``
        (SURVEY.md §8e).  No steady-state collective exists on this path."""
        import torch.distributed as dist
        view = self.packed_view()
        dist.broadcast(view, src=src, group=group)
        _lib.check(self.lib.dce_weights_adopt(self._handle), "dce_weights_adopt")
        return self

    # -- workspace ----------------------------------------------------------
    def _ws(self, n: int) -> torch.Tensor:
        need = self.lib.dce_workspace_bytes(n, _lib.PRECISIONS[self.precision])
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.zeros(need, dtype=torch.uint8, device=self.device)   # tape padding rows start finite
        return self._workspace

    def _outs(self, n: int, want_logits: bool, want_cls: bool, want_bits: bool):
        logits = torch.empty((n, CLASSES), dtype=torch.float32, device=self.device) if want_logits else None
        cls = torch.empty((n,), dtype=torch.int32, device=self.device) if want_cls else None
        bits = torch.empty((n, 4), dtype=torch.uint8, device=self.device) if want_bits else None
        return logits, cls, bits

    @staticmethod
    def _p(t: Optional[torch.Tensor]):
        return ctypes.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)

    # -- K1 -------------------------------------------------------------------
    def classify(self, x: torch.Tensor, want_logits: bool = True, want_cls: bool = True, want_bits: bool = True
                 ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
        """``(B,150,54)`` fp32 CUDA windows -> ``(logits (B,16) f32, cls (B,) i32, bits (B,4) u8)``."""
        if x.dim() != 3 or tuple(x.shape[1:]) != (WINDOW, CHANNELS):
            raise ValueError(f"expected (B,{WINDOW},{CHANNELS}), got {tuple(x.shape)}")
        if x.device != self.device or x.dtype != torch.float32:
            raise ValueError(f"expected float32 on {self.device}, got {x.dtype} on {x.device}")
        x = x.contiguous()
        if x.data_ptr() % 16:             # a view into a larger buffer at an odd offset
            x = x.clone()
        n = x.shape[0]
        ops = _lib.torch_ops()
        if ops is not None and n:
            # contact_cnn.forward -> torch.ops.dce.forward -> dce_forward (SURVEY.md §8b "who calls it")
            logits, cls, bits = ops.forward(self._handle.value, x, self._ws(n), _lib.PRECISIONS[self.precision],
                                            want_logits, want_cls, want_bits)
            self.last_launches = self.lib.dce_last_launch_count()
            return (logits if want_logits else None), (cls if want_cls else None), (bits if want_bits else None)
        logits, cls, bits = self._outs(n, want_logits, want_cls, want_bits)
        if n == 0:
            return logits, cls, bits
        ws = self._ws(n)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.dce_forward(self._handle, self._p(x), n, self._p(logits), self._p(cls), self._p(bits),
                                      self._p(ws), ws.numel(), _lib.PRECISIONS[self.precision],
                                      ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dce_forward")
        self.last_launches = self.lib.dce_last_launch_count()
        return logits, cls, bits

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.classify(x, True, False, False)[0]

    def profile_forward(self, x: torch.Tensor):
        """Per-kernel CUDA-event durations of one ``dce_forward`` (synchronises):
        ``[(kernel name, ms), ...]`` in launch order."""
        x = x.contiguous()
        n = x.shape[0]
        logits, cls, bits = self._outs(n, True, True, True)
        ws = self._ws(n)
        cap = 64
        ms = (ctypes.c_float * cap)()
        names = (ctypes.c_char_p * cap)()
        cnt = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.dce_forward_profile(self._handle, self._p(x), n, self._p(logits), self._p(cls), self._p(bits),
                                              self._p(ws), ws.numel(), _lib.PRECISIONS[self.precision],
                                              ctypes.c_void_p(stream.cuda_stream), cap, ms, names, ctypes.byref(cnt))
        _lib.check(rc, "dce_forward_profile")
        return [(names[i].decode(), float(ms[i])) for i in range(cnt.value)]

    def profile_stream(self, data: torch.Tensor, first_window: int, n_windows: int):
        """Per-kernel CUDA-event durations of one ``dce_stream`` call (synchronises)."""
        data = data.contiguous()
        if data.data_ptr() % 16:
            data = data.clone()
        logits, cls, bits = self._outs(n_windows, False, True, True)
        ws = self._ws(n_windows)
        cap = 64
        ms = (ctypes.c_float * cap)(); names = (ctypes.c_char_p * cap)(); cnt = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.dce_stream_profile(self._handle, self._p(data), data.shape[0], first_window, n_windows, None,
                                             self._p(cls), self._p(bits), self._p(ws), ws.numel(),
                                             _lib.PRECISIONS[self.precision], ctypes.c_void_p(stream.cuda_stream),
                                             cap, ms, names, ctypes.byref(cnt))
        _lib.check(rc, "dce_stream_profile")
        return [(names[i].decode(), float(ms[i])) for i in range(cnt.value)]

    # -- host-buffer entry point (what a user with numpy / pinned data calls) ----
    def classify_host_async(self, x_host: torch.Tensor, out_bits_host: torch.Tensor, out_cls_host: torch.Tensor,
                            chunk: int = 1024) -> "HostBatch":
        """``classify_host`` without the final wait: enqueues the chunked upload, the kernels and the device->host read
        of the results and returns a :class:`HostBatch` whose ``wait()`` blocks until ``out_cls_host`` / ``out_bits_host``
        (pinned, caller-owned, one pair per batch in flight) hold the answer.  Keeping two batches in flight lets the
        upload of batch i+1 run under the kernel tail of batch i (the ~0.25 ms after the last chunk has landed), which a
        call that waits for its own result cannot hide.  Inputs must stay untouched until ``wait()`` returns."""
        return self.classify_host(x_host, out_bits_host, out_cls_host, chunk=chunk, _async=True)

    def classify_host(self, x_host: torch.Tensor, out_bits_host: Optional[torch.Tensor] = None,
                      out_cls_host: Optional[torch.Tensor] = None, chunk: int = 1024, zero_copy: bool = False, _async: bool = False):
        """``x_host``: ``(B,150,54)`` fp32 on the HOST (pinned for full speed).
        Copies chunk by chunk on a side stream so the host->device transfer of
        chunk i+1 overlaps the kernels of chunk i; returns host ``(cls, bits)``
        after the device->host read of the results has completed.

        ``zero_copy=True`` (pinned ``x_host`` and pinned outputs only; not yet measured at this size): no
        staging copy at all — one ``dce_forward`` over the whole batch whose first kernel's bulk-TMA loader reads
        the windows in place over PCIe, and whose last kernel writes classes and bits straight into the pinned
        outputs, as ``LatencyRunner`` does for one window."""
        if (x_host.dim() != 3 or tuple(x_host.shape[1:]) != (WINDOW, CHANNELS) or x_host.dtype != torch.float32
                or x_host.is_cuda):
            raise ValueError(f"expected a host (B,{WINDOW},{CHANNELS}) float32 tensor, got {tuple(x_host.shape)} "
                             f"{x_host.dtype} on {x_host.device}")
        if chunk < 1:
            raise ValueError("chunk must be positive")
        n = x_host.shape[0]
        if out_bits_host is None:
            out_bits_host = torch.empty((n, 4), dtype=torch.uint8)
            out_bits_host = out_bits_host.pin_memory() if n else out_bits_host
        if out_cls_host is None:
            out_cls_host = torch.empty((n,), dtype=torch.int32)
            out_cls_host = out_cls_host.pin_memory() if n else out_cls_host
        if tuple(out_bits_host.shape) != (n, 4) or out_bits_host.dtype != torch.uint8 or \
                tuple(out_cls_host.shape) != (n,) or out_cls_host.dtype != torch.int32:
            raise ValueError("out_bits_host must be (B,4) uint8 and out_cls_host (B,) int32")
        if n == 0:
            return out_cls_host, out_bits_host
        if zero_copy:
            if not (x_host.is_pinned() and out_bits_host.is_pinned() and out_cls_host.is_pinned()):
                raise ValueError("zero_copy needs pinned host tensors (the kernels dereference them directly)")
            x_host = x_host.contiguous()
            ws = self._ws(n)
            with torch.cuda.device(self.device):
                compute = torch.cuda.current_stream(self.device)
                rc = self.lib.dce_forward(self._handle, self._p(x_host), n, None, self._p(out_cls_host), self._p(out_bits_host),
                                          self._p(ws), ws.numel(), _lib.PRECISIONS[self.precision],
                                          ctypes.c_void_p(compute.cuda_stream))
                _lib.check(rc, "dce_forward")
                self.last_launches = self.lib.dce_last_launch_count()
                compute.synchronize()
            return out_cls_host, out_bits_host
        with torch.cuda.device(self.device):
            compute = torch.cuda.current_stream(self.device)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            fresh = False
            if getattr(self, "_stage", None) is None:
                self._stage = [torch.empty((chunk, WINDOW, CHANNELS), dtype=torch.float32, device=self.device) for _ in range(2)]
                self._stage_free = [torch.cuda.Event() for _ in range(2)]
                fresh = True
            if self._stage[0].shape[0] != chunk:
                self._stage = [torch.empty((chunk, WINDOW, CHANNELS), dtype=torch.float32, device=self.device) for _ in range(2)]
                fresh = True
            copy = self._copy_stream
            if fresh:
                copy.wait_stream(compute)      # new staging memory may have been handed back by work still queued on `compute`
            # (otherwise the copy stream does NOT wait for earlier compute work: the per-buffer `_stage_free` events below are the
            # only dependency, so the upload of this batch runs under the kernel tail of the previous one)
            bits_dev = torch.empty((n, 4), dtype=torch.uint8, device=self.device)
            cls_dev = torch.empty((n,), dtype=torch.int32, device=self.device)
            ws = self._ws(min(n, chunk))
            launches = 0
            for i, s0 in enumerate(range(0, n, chunk)):
                m = min(chunk, n - s0)
                buf = self._stage[i & 1]
                with torch.cuda.stream(copy):
                    copy.wait_event(self._stage_free[i & 1])          # kernels of chunk i-2 are done with this buffer
                    buf[:m].copy_(x_host[s0:s0 + m], non_blocking=True)
                    ready = torch.cuda.Event(); ready.record(copy)
                compute.wait_event(ready)
                rc = self.lib.dce_forward(self._handle, self._p(buf), m, None, ctypes.c_void_p(cls_dev[s0:].data_ptr()),
                                          ctypes.c_void_p(bits_dev[s0:].data_ptr()), self._p(ws), ws.numel(),
                                          _lib.PRECISIONS[self.precision], ctypes.c_void_p(compute.cuda_stream))
                _lib.check(rc, "dce_forward")
                launches += self.lib.dce_last_launch_count()
                self._stage_free[i & 1].record(compute)
            out_bits_host.copy_(bits_dev, non_blocking=True)
            out_cls_host.copy_(cls_dev, non_blocking=True)
            self.last_launches = launches
            if _async:
                if not (out_bits_host.is_pinned() and out_cls_host.is_pinned()):
                    raise ValueError("classify_host_async needs pinned output tensors (the device->host read must not block)")
                done = torch.cuda.Event()
                done.record(compute)
                return HostBatch(done, out_cls_host, out_bits_host, (bits_dev, cls_dev))
            compute.synchronize()
        return out_cls_host, out_bits_host

    # -- K2 -------------------------------------------------------------------
    def stream(self, data: torch.Tensor, first_window: int = 0, n_windows: Optional[int] = None,
               want_logits: bool = False, want_cls: bool = True, want_bits: bool = True):
        """Whole-log inference: ``data`` is the device-resident ``(T,54)`` fp32
        sensor log (utils/data_handler.py:26); returns results for windows
        ``first_window .. first_window + n_windows``."""
        if data.dim() != 2 or data.shape[1] != CHANNELS:
            raise ValueError(f"expected (T,{CHANNELS}), got {tuple(data.shape)}")
        if data.device != self.device or data.dtype != torch.float32:
            raise ValueError(f"expected float32 on {self.device}, got {data.dtype} on {data.device}")
        data = data.contiguous()
        if data.data_ptr() % 16:          # e.g. log[k:] with odd k: rows are 216 B, the ABI wants a 16-byte aligned base
            data = data.clone()
        T = data.shape[0]
        total = max(T - WINDOW + 1, 0)
        if n_windows is None:
            n_windows = total - first_window
        if first_window < 0 or n_windows < 0 or first_window + n_windows > total:
            raise ValueError(f"window range [{first_window}, {first_window + n_windows}) outside [0, {total})")
        ops = _lib.torch_ops()
        if ops is not None and n_windows:
            logits, cls, bits = ops.stream(self._handle.value, data, first_window, n_windows, self._ws(n_windows),
                                           _lib.PRECISIONS[self.precision], want_logits, want_cls, want_bits)
            self.last_launches = self.lib.dce_last_launch_count()
            return (logits if want_logits else None), (cls if want_cls else None), (bits if want_bits else None)
        logits, cls, bits = self._outs(n_windows, want_logits, want_cls, want_bits)
        if n_windows == 0:
            return logits, cls, bits
        ws = self._ws(n_windows)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.dce_stream(self._handle, self._p(data), T, first_window, n_windows,
                                     self._p(logits), self._p(cls), self._p(bits), self._p(ws), ws.numel(),
                                     _lib.PRECISIONS[self.precision], ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dce_stream")
        self.last_launches = self.lib.dce_last_launch_count()
        return logits, cls, bits

    def stream_host(self, log_host: torch.Tensor, chunk_rows: int = 1 << 18, out_bits_host: Optional[torch.Tensor] = None,
                    out_cls_host: Optional[torch.Tensor] = None):
        """Whole-log inference from a HOST log (``(T,54)`` float32, pinned for full speed): the log is uploaded in
        chunks on a side stream while ``dce_stream`` classifies the windows whose rows have already arrived, so
        the 216 B/step upload hides behind the kernels.  Returns host ``(cls (N,) int32, bits (N,4) uint8)`` after
        the device->host read has completed; results are bit-identical to ``stream()`` on the resident log
        (calls end on the 32-window tile boundaries of the statistics kernel)."""
        if log_host.dim() != 2 or log_host.shape[1] != CHANNELS or log_host.dtype != torch.float32 or log_host.is_cuda:
            raise ValueError(f"expected a host (T,{CHANNELS}) float32 log")
        T = log_host.shape[0]
        n = max(T - WINDOW + 1, 0)
        if out_bits_host is None:
            out_bits_host = torch.empty((n, 4), dtype=torch.uint8)
            out_bits_host = out_bits_host.pin_memory() if n else out_bits_host
        if out_cls_host is None:
            out_cls_host = torch.empty((n,), dtype=torch.int32)
            out_cls_host = out_cls_host.pin_memory() if n else out_cls_host
        if n == 0:
            return out_cls_host, out_bits_host
        with torch.cuda.device(self.device):
            compute = torch.cuda.current_stream(self.device)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            copy = self._copy_stream
            copy.wait_stream(compute)
            log_dev = torch.empty((T, CHANNELS), dtype=torch.float32, device=self.device)
            log_dev.record_stream(copy)
            cls = torch.empty((n,), dtype=torch.int32, device=self.device)
            bits = torch.empty((n, 4), dtype=torch.uint8, device=self.device)
            ws = self._ws(n)
            from .sharding import upload_schedule
            launches, up = 0, 0
            for rows_up, first, end in upload_schedule(T, chunk_rows):
                with torch.cuda.stream(copy):
                    for r0 in range(up, rows_up, chunk_rows):
                        log_dev[r0:min(rows_up, r0 + chunk_rows)].copy_(log_host[r0:min(rows_up, r0 + chunk_rows)], non_blocking=True)
                    ready = torch.cuda.Event(); ready.record(copy)
                up = rows_up
                compute.wait_event(ready)             # windows [first, end): their rows (and statistics tiles) have arrived
                rc = self.lib.dce_stream(self._handle, self._p(log_dev), T, first, end - first, None,
                                         ctypes.c_void_p(cls[first:].data_ptr()), ctypes.c_void_p(bits[first:].data_ptr()),
                                         self._p(ws), ws.numel(), _lib.PRECISIONS[self.precision],
                                         ctypes.c_void_p(compute.cuda_stream))
                _lib.check(rc, "dce_stream")
                launches += self.lib.dce_last_launch_count()
            out_bits_host.copy_(bits, non_blocking=True)
            out_cls_host.copy_(cls, non_blocking=True)
            compute.synchronize()
        self.last_launches = launches
        return out_cls_host, out_bits_host

    # -- K3: control-loop runner -------------------------------------------------
    def latency_runner(self, n: int = 1, want_logits: bool = False, use_graph: bool = False, persistent: bool = False,
                       idle_timeout_s: float = 2.0) -> "LatencyRunner":
        """Batch-``n`` (<= 4) control-loop path: see :class:`LatencyRunner`."""
        return LatencyRunner(self, n, want_logits, use_graph, persistent, idle_timeout_s)

    def row_runner(self, idle_timeout_s: float = 2.0) -> "RowRunner":
        """The control loop fed one new 54-channel row per tick: see :class:`RowRunner`."""
        return RowRunner(self, idle_timeout_s)

    def set_option(self, key, value: int) -> int:
        """``dce_weights_set_option``: an ablation / debugging switch of THIS engine's handle (include/dce.h);
        returns the ABI's code (0, or -1 for an unknown key)."""
        if isinstance(key, str):
            key = key.encode()
        return int(self.lib.dce_weights_set_option(self._handle, key, int(value)))

    def read_trace(self, n: int = 60 * 16):
        """clock64 samples of CTA 0 recorded by the kernel ``trace_layer`` selects (DCE_TRACE=1 builds, option "trace")."""
        buf = (ctypes.c_longlong * n)()
        _lib.check(self.lib.dce_debug_read_trace(self._handle, buf, n), "dce_debug_read_trace")
        return list(buf)

    # -- small helpers on the same ABI ------------------------------------------
    def decimal2binary(self, x: torch.Tensor) -> torch.Tensor:
        flat = x.reshape(-1).to(torch.int64).contiguous()
        out = torch.empty((flat.numel(), 4), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.dce_decimal2binary(self._p(flat), flat.numel(), self._p(out), ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dce_decimal2binary")
        return out.reshape(*x.shape, 4)

    def accuracy_counts(self, cls: torch.Tensor, labels: torch.Tensor, counts: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``dce_accuracy_counts``: counts[0] += #(cls == label); counts[1+l] += #(bit l agrees); counts[5:21] the four
        per-leg 2x2 confusion matrices; counts[21:277] the 16x16 class confusion matrix (int64[277], on device)."""
        if counts is None:
            counts = torch.zeros(_lib.DCE_NUM_COUNTS, dtype=torch.int64, device=self.device)
        cls = cls.to(torch.int32).contiguous()
        labels = labels.reshape(-1).to(torch.int64).contiguous()
        if cls.numel() != labels.numel():
            raise ValueError("cls and labels differ in length")
        ops = _lib.torch_ops()
        if ops is not None and cls.numel():
            return ops.accuracy_counts(cls, labels, counts)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.dce_accuracy_counts(self._p(cls), self._p(labels), cls.numel(), self._p(counts),
                                              ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "dce_accuracy_counts")
        return counts


_resident = {}          # device index -> the runner whose resident server was started last (one server holds every SM)


def _claim_device(eng, runner):
    """A resident server occupies the whole GPU: refuse to start a second one while another runner's is alive (it could
    only wait for the first to idle out)."""
    idx = eng.device.index
    other = _resident.get(idx)
    if other is not None and other is not runner and other._c is not None and other._c[other._ALIVE]:
        raise RuntimeError(f"cuda:{idx} already runs a resident latency server ({type(other).__name__}); close() it first")
    _resident[idx] = runner


class HostBatch:
    """A batch in flight (``ContactEngine.classify_host_async``): ``wait()`` -> host ``(cls, bits)``."""

    def __init__(self, done, cls_host, bits_host, keep):
        self._done, self.cls_host, self.bits_host, self._keep = done, cls_host, bits_host, keep

    def wait(self):
        self._done.synchronize()
        self._keep = None
        return self.cls_host, self.bits_host


class LatencyRunner:
    """The 1 kHz control-loop call (BASELINE configs[4]).  The newest window(s) sit in PINNED host memory and the
    fused latency kernel reads them in place over PCIe (zero-copy: 32.4 KB per window, once) and writes class,
    contact bits (and logits) straight back into pinned host memory — no copy-engine hops on either side, which
    cost more than the kernel itself at this size.  A step is one ``dce_forward`` call (a single kernel launch) and
    one stream synchronise; ``use_graph=True`` replays a captured CUDA graph instead (measured 4 us slower per step
    from Python, `tools/latency_diag.py`, but it is the form a larger captured control loop would embed).

        run = engine.latency_runner()
        run.x_host[0] = newest_window          # (150, 54) float32, z-scored (utils/data_handler.py:55-56)
        cls, bits = run.step()                 # pinned host tensors, valid until the next step()

    ``persistent=True`` is the resident form (``dce_latency_server_start``, include/dce.h): ONE cooperative kernel
    stays on the GPU and serves a step per doorbell — ``step()`` writes the window, increments a word in pinned
    memory and spins on the answer word; no launch, no stream synchronisation, no CUDA call at all in the loop.
    The server holds every SM while it lives; it retires on ``close()`` or after ``idle_timeout_s`` without a step
    (the next ``step()`` starts it again), so other work on the device is delayed by at most that long.

    Replaces one iteration of the reference loop at its default batch_size 1
    (/root/reference/src/inference_one_seq.py:23-28, config/inference_one_seq_params.yaml:10).
    """

    _SEQ_IN, _QUIT, _SEQ_OUT, _CLS0, _BITS0, _DEVICE_NS, _ALIVE = 0, 1, 16, 17, 18, 19, 20     # 32-bit words of dce_latency_ctrl

    def __init__(self, eng: ContactEngine, n: int = 1, want_logits: bool = False, use_graph: bool = False,
                 persistent: bool = False, idle_timeout_s: float = 2.0):
        if not 1 <= n <= 4:
            raise ValueError("the latency path takes 1..4 windows per step")
        if persistent and use_graph:
            raise ValueError("persistent and use_graph exclude each other")
        self.eng, self.n, self.use_graph, self.persistent, self.idle_timeout_s = eng, n, use_graph, persistent, float(idle_timeout_s)
        dev = eng.device
        self.x_host = torch.zeros((n, WINDOW, CHANNELS), dtype=torch.float32).pin_memory()
        self.cls_host = torch.zeros((n,), dtype=torch.int32).pin_memory()
        self.bits_host = torch.zeros((n, 4), dtype=torch.uint8).pin_memory()
        self.logits_host = torch.zeros((n, CLASSES), dtype=torch.float32).pin_memory() if want_logits else None
        self.stream = torch.cuda.Stream(dev)
        # its own zero-filled workspace: the engine's may be in use by (or regrown for) other calls on other streams
        ws = torch.zeros(eng.lib.dce_workspace_bytes(n, _lib.PRECISIONS[eng.precision]), dtype=torch.uint8, device=dev)
        P = ContactEngine._p

        # pinned host pointers are device-accessible under unified addressing: the kernel dereferences them directly.
        # The buffers live as long as the runner, so the ctypes arguments are built once (a few microseconds per step
        # of a ~40 us call); only the handle is looked up per call, so a closed engine fails cleanly.
        forward = eng.lib.dce_forward
        rest = (P(self.x_host), n, P(self.logits_host), P(self.cls_host), P(self.bits_host), P(ws), ws.numel(),
                _lib.PRECISIONS[eng.precision], ctypes.c_void_p(self.stream.cuda_stream))

        index = dev.index

        def launch():
            # the C ABI launches on the calling thread's current device (include/dce.h): make it the engine's
            if torch.cuda.current_device() != index:
                with torch.cuda.device(index):
                    rc = forward(eng._handle, *rest)
            else:
                rc = forward(eng._handle, *rest)
            if rc:
                _lib.check(rc, "dce_forward")

        self._launch = launch
        with torch.cuda.device(dev), torch.cuda.stream(self.stream):
            for _ in range(2):                                   # kernel attributes set before the capture
                launch()
            self.stream.synchronize()
            self.launches = eng.lib.dce_last_launch_count()
            self.graph = None
            if use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.stream):
                    launch()
        self._ws = ws
        self._ctrl = self._c = None
        self.server_starts = 0
        if persistent:
            self._ctrl = torch.zeros(32, dtype=torch.int32).pin_memory()        # dce_latency_ctrl: 128 bytes, page aligned
            self._c = self._ctrl.numpy().view("uint32")
            if n == 1 and not want_logits:
                # results inside the control block: class, bits and the step number arrive as one 16-byte store
                self.cls_host = self._ctrl[self._CLS0:self._CLS0 + 1]
                self.bits_host = self._ctrl[self._BITS0:self._BITS0 + 1].view(torch.uint8).reshape(1, 4)
            start = eng.lib.dce_latency_server_start
            sargs = (P(self.x_host), n, P(self.logits_host), P(self.cls_host), P(self.bits_host), P(self._ctrl), P(ws), ws.numel(),
                     ctypes.c_double(self.idle_timeout_s), ctypes.c_void_p(self.stream.cuda_stream))

            def start_server():
                _claim_device(eng, self)
                self.stream.synchronize()                            # a retired server has left the stream
                with torch.cuda.device(index):
                    _lib.check(start(eng._handle, *sargs), "dce_latency_server_start")
                self._seq = 0
                self.server_starts += 1
                c, t0 = self._c, time.perf_counter()
                while not c[self._ALIVE]:
                    if time.perf_counter() - t0 > 10.0:
                        raise RuntimeError("the latency server did not come up within 10 s (is the GPU busy with other kernels?)")
            self._start_server = start_server
            start_server()

    def _step_persistent(self):
        c = self._c
        if not c[self._ALIVE]:                                       # retired after idle_timeout_s without a step
            self._start_server()
        self._seq += 1
        seq = self._seq
        c[self._SEQ_IN] = seq                                        # the doorbell (the window was written before it)
        spins = 0
        while c[self._SEQ_OUT] != seq:
            spins += 1
            if (spins & 1023) == 0 and not c[self._ALIVE]:           # it retired just as the doorbell rang: start over
                if c[self._SEQ_OUT] == seq:
                    break
                self._start_server()
                self._seq = seq = 1
                c[self._SEQ_IN] = seq
        return self.cls_host, self.bits_host

    def close(self):
        """Retire the resident server (persistent mode); idempotent."""
        if self._c is not None and self._c[self._ALIVE]:
            self._c[self._QUIT] = 1
            self.stream.synchronize()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enqueue(self):
        """Launch one step without waiting (``self.stream``)."""
        if self.graph is None:
            self._launch()                                       # enqueues on self.stream (passed to the C ABI)
        else:
            with torch.cuda.device(self.eng.device), torch.cuda.stream(self.stream):
                self.graph.replay()

    def step(self, window: Optional[torch.Tensor] = None):
        """Classify ``self.x_host`` (or ``window``, copied into it first); returns host ``(cls, bits)``."""
        if window is not None:
            self.x_host.copy_(window.reshape(self.x_host.shape))
        if self.persistent:
            return self._step_persistent()
        self.enqueue()
        self.stream.synchronize()
        return self.cls_host, self.bits_host


class RowRunner:
    """The 1 kHz control loop fed ONE NEW SENSOR ROW per tick (``dce_latency_row_server_start``, include/dce.h): the
    resident kernel keeps the last 150 rows in a device ring, z-scores the newest window itself
    (utils/data_handler.py:55-56) and classifies it.  ``push(row)`` writes 216 bytes + tags into pinned memory and
    spins on the answer word: no CUDA call, no host arithmetic.

        rr = engine.row_runner()
        for row in sensor_rows:                # (54,) float32: [q, qd, acc, omega, p, v] (utils/mat2numpy.py:73)
            cls, bits = rr.push(row)           # valid once 150 rows have been pushed (rr.ready)

    The server holds every SM while it lives and retires after ``idle_timeout_s`` without a row (the next push
    restarts it; the ring is kept), or on ``close()``.
    """

    _SEQ_OUT, _CLS0, _BITS0, _DEVICE_NS, _ALIVE = 80, 81, 82, 83, 84

    def __init__(self, eng: ContactEngine, idle_timeout_s: float = 2.0):
        import numpy as np
        self.eng, self.idle_timeout_s = eng, float(idle_timeout_s)
        dev = eng.device
        self.stream = torch.cuda.Stream(dev)
        self._ctrl = torch.zeros(128, dtype=torch.int32).pin_memory()          # dce_latency_row_ctrl: 512 bytes
        self._c = self._ctrl.numpy().view("uint32")
        self._chunks = self._c[:76].reshape(19, 4)
        self._floats = self._chunks[:18, :3].view("float32")                   # (18, 3): the 54 floats of the row
        self._tags = self._chunks[:, 3]
        self.ring = torch.zeros((2 * WINDOW, CHANNELS), dtype=torch.float32, device=dev)
        self._ws = torch.zeros(eng.lib.dce_workspace_bytes(1, _lib.PRECISIONS[eng.precision]), dtype=torch.uint8, device=dev)
        self.rows_seen = 0
        self.server_starts = 0
        self._seq = 0
        self._np = np
        P = ContactEngine._p
        start = eng.lib.dce_latency_row_server_start
        args = (P(self._ctrl), P(self.ring), P(self._ws), self._ws.numel(), ctypes.c_double(self.idle_timeout_s),
                ctypes.c_void_p(self.stream.cuda_stream))
        index = dev.index

        def start_server():
            _claim_device(eng, self)
            torch.cuda.synchronize(dev)                          # ring / workspace initialised; a retired server has left the stream
            with torch.cuda.device(index):
                _lib.check(start(eng._handle, *args), "dce_latency_row_server_start")
            self._seq = 0
            self.server_starts += 1
            t0 = time.perf_counter()
            while not self._c[self._ALIVE]:
                if time.perf_counter() - t0 > 10.0:
                    raise RuntimeError("the row server did not come up within 10 s (is the GPU busy with other kernels?)")
        self._start_server = start_server
        start_server()

    @property
    def ready(self) -> bool:
        return self.rows_seen >= WINDOW

    @property
    def device_us(self) -> float:
        """Device time of the last step: row seen -> results written (microseconds)."""
        return float(self._c[self._DEVICE_NS]) * 1e-3

    def push(self, row):
        """Append one 54-vector (numpy array / sequence / CPU tensor) and classify the newest window:
        ``(class 0..15, (RF, LF, RH, LH) contact bits)``; meaningful once ``ready``."""
        c = self._c
        if not c[self._ALIVE]:
            self._start_server()
        r = row.numpy() if torch.is_tensor(row) else self._np.asarray(row, dtype="float32")
        slot = self.rows_seen % WINDOW
        self.rows_seen += 1
        self._seq += 1
        seq = self._seq
        r = r.reshape(18, 3)
        self._floats[...] = r
        self._chunks[18, 1] = slot
        self._tags[...] = seq                                    # the doorbell: every chunk's tag, after its data
        spins = 0
        while c[self._SEQ_OUT] != seq:
            spins += 1
            if (spins & 1023) == 0 and not c[self._ALIVE]:       # it retired just as the row arrived: start over, same row, same slot
                if c[self._SEQ_OUT] == seq:
                    break
                self._start_server()                             # (clears the control block; the ring in device memory is kept)
                self._seq = seq = 1
                self._floats[...] = r
                self._chunks[18, 1] = slot
                self._tags[...] = seq
        b = int(c[self._BITS0])
        return int(c[self._CLS0]), (b & 1, (b >> 8) & 1, (b >> 16) & 1, (b >> 24) & 1)

    def close(self):
        if self._c is not None and self._c[self._ALIVE]:
            self._chunks[18, 0] = 1                               # quit
            self.stream.synchronize()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
