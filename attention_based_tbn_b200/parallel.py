"""One process per GPU: clip sharding and the final score gather.

Replaces the reference's single-process ``nn.DataParallel`` (core/models/dataparallel.py,
model_builder.py:73-75: per-forward parameter broadcast + input scatter + output gather on GIL-bound
threads).  Clips are independent in eval (BatchNorm uses running statistics), so the batch is
partitioned across ranks -- contiguous chunks, remainder to the low ranks, all segments of a clip
on one rank so the consensus is local -- with NO collective on the data path; the only
exchange is one ``all_gather`` of the (B_local, sum n_cls) fp32 scores (+ the (B_local*N, S)
attention weights), a few hundred KB, latency bound.  Works on NCCL (GPU) and
gloo (CPU tests).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .sampling import shard_clips


def world() -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_input(inputs: Dict[str, torch.Tensor], rank: Optional[int] = None, world_size: Optional[int] = None):
    """Slice every (B, ...) tensor of the input dict to this rank's clips; other entries pass through."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    some = next(v for v in inputs.values() if torch.is_tensor(v))
    rng = shard_clips(some.shape[0], world_size, rank)
    out = {}
    for k, v in inputs.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == some.shape[0]:
            out[k] = v[rng.start:rng.stop]
        elif isinstance(v, (list, tuple)) and len(v) == some.shape[0]:
            out[k] = v[rng.start:rng.stop]
        else:
            out[k] = v
    return out


def _all_gather_rows(t: torch.Tensor, counts: List[int]) -> torch.Tensor:
    """all_gather of row blocks with per-rank row counts (padded to the max, then trimmed)."""
    mx = max(counts)
    pad = t.new_zeros((mx,) + tuple(t.shape[1:]))
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def gather_outputs(local: "OrderedDict[str, torch.Tensor]", num_clips: int, num_segments: int) -> "OrderedDict[str, torch.Tensor]":
    """Score gather: every rank receives the full-batch output dict in clip order."""
    rank, ws = world()
    if ws == 1:
        return local
    counts = [len(shard_clips(num_clips, ws, r)) for r in range(ws)]
    out = OrderedDict()
    for k, v in local.items():
        rows = [c * num_segments for c in counts] if k == "weights" else counts
        out[k] = _all_gather_rows(v.contiguous(), rows)
    return out


def data_parallel_forward(model, inputs: Dict[str, torch.Tensor], gather: bool = True):
    """Shard -> local forward -> gather.  ``inputs`` is the FULL batch on every rank (or use
    ``shard_input`` upstream and call ``model`` directly)."""
    some = next(v for v in inputs.values() if torch.is_tensor(v) and v.dim() == 5)
    b, n = some.shape[:2]
    local = model(shard_input(inputs))
    return gather_outputs(local, b, n) if gather else local
