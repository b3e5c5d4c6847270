"""Window-range sharding across ranks.

Every window is classified independently (utils/data_handler.py:55-57 recomputes
its statistics per window; the model is stateless in eval), so the path shards
by window range with NO data-path collective: rank r owns windows
``[r*N/R, (r+1)*N/R)`` and, in stream mode, reads rows ``[start, end+149)`` of
the log — a 149-row read-only halo.  The only collectives are the one-time
weight broadcast (``ContactEngine.broadcast_weights``) and an optional final
all-gather of the ``(N,4)`` uint8 contact bits for callers that need the full
result on one rank (``save2mat`` / ``save2lcm``, src/inference_one_seq.py:172-176).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .synth import WINDOW


def window_range(n_windows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[start, end)`` for ``rank``; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world) or n_windows < 0:
        raise ValueError("bad rank/world/n_windows")
    base, rem = divmod(n_windows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rows_for_windows(start: int, end: int, window: int = WINDOW) -> Tuple[int, int]:
    """Rows of the sensor log a rank must hold to classify windows ``[start, end)``."""
    return (start, start) if end <= start else (start, end + window - 1)


def upload_schedule(total_rows: int, chunk_rows: int, window: int = WINDOW, tile: int = 32) -> List[Tuple[int, int, int]]:
    """Schedule of ``ContactEngine.stream_host``: the log is uploaded in chunks of ``chunk_rows`` rows and, after
    each chunk, the windows whose rows have all arrived are classified.  Returns ``(rows_uploaded, first, end)``
    per call: windows ``[first, end)`` run once rows ``[0, rows_uploaded)`` are on the device.  Every call but the
    last ends on a multiple of ``tile`` whose whole statistics tile (``tile + window - 1`` rows from its first
    window) has been uploaded, so chunked and one-shot results are bit-identical (csrc/dce_tc.cuh,
    ``window_stats_kernel``)."""
    if total_rows < 0 or chunk_rows <= 0:
        raise ValueError("bad total_rows / chunk_rows")
    n = max(total_rows - window + 1, 0)
    span = tile + window - 1
    out, done = [], 0
    for r0 in range(0, total_rows, chunk_rows):
        r1 = min(total_rows, r0 + chunk_rows)
        hi = n if r1 == total_rows else min(n, max(done, ((r1 - span) // tile + 1) * tile if r1 >= span else 0))
        if hi > done:
            out.append((r1, done, hi))
            done = hi
    return out


def all_gather_bits(bits: torch.Tensor, n_windows: int, group=None) -> torch.Tensor:
    """Assemble the full ``(N,4)`` uint8 result from per-rank shards (ragged
    shards are padded to the largest one for the collective)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes: List[int] = [window_range(n_windows, r, world)[1] - window_range(n_windows, r, world)[0] for r in range(world)]
    m = max(sizes) if sizes else 0
    pad = torch.zeros((m, 4), dtype=torch.uint8, device=bits.device)
    pad[: bits.shape[0]] = bits
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)
