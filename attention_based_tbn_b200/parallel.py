"""One process per GPU: clip sharding and the final score gather.

Replaces the reference's single-process ``nn.DataParallel`` (core/models/dataparallel.py,
model_builder.py:73-75: per-forward parameter broadcast + input scatter + output gather on GIL-bound
threads).  Clips are independent in eval (BatchNorm uses running statistics), so the batch is
partitioned across ranks -- contiguous chunks, remainder to the low ranks, all segments of a clip
on one rank so the consensus is local -- with NO collective on the data path; the only
exchange is ONE ``all_gather_into_tensor`` per step: every output of a clip (its class scores and
the attention weights of its N segments) is packed into one fp32 row, rows are padded to the largest
shard, gathered, and split back into the reference's output dict -- a few hundred KB, latency bound
(SURVEY.md section 8e).  Works on NCCL (GPU) and gloo (CPU tests).

Run the forward + gather loop under ``torch.cuda.stream(<a stream of your own>)``: issued on the legacy default stream
the per-step gather was measured to cost, intermittently, a whole extra forward (bench.py).
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


def pack_rows(local: "OrderedDict[str, torch.Tensor]", clips: int, rows: int) -> tuple:
    """Pack the first ``clips`` clips of an output dict into a (rows, width) fp32 matrix, one row per clip.
    Returns (matrix, layout) with layout = [(key, width, trailing shape per clip), ...].  A tensor with k*clips_total
    rows (the (B*N, 1, S) attention weights) contributes k rows' worth of columns per clip."""
    some = next(iter(local.values()))
    have = next(v.shape[0] for k, v in local.items() if k != "weights")  # clips the local forward produced
    layout, width = [], 0
    for k, v in local.items():
        per = v.shape[0] // max(have, 1)  # rows of this tensor per clip (N for the weights, 1 for scores)
        tail = (per,) + tuple(v.shape[1:])
        w = 1
        for s in tail:
            w *= s
        layout.append((k, w, tail, v.dtype))
        width += w
    mat = some.new_zeros((rows, width), dtype=torch.float32)
    col = 0
    for (k, w, tail, _), v in zip(layout, local.values()):
        per = tail[0]
        if clips > 0:
            mat[:clips, col:col + w] = v[:clips * per].reshape(clips, w)
        col += w
    return mat, layout


def unpack_rows(mat: torch.Tensor, layout: list) -> "OrderedDict[str, torch.Tensor]":
    """Inverse of ``pack_rows`` on a (clips, width) matrix."""
    out = OrderedDict()
    col = 0
    b = mat.shape[0]
    for k, w, tail, dtype in layout:
        t = mat[:, col:col + w]
        per = tail[0]
        out[k] = (t if (per == 1 and len(tail) == 2) else t.reshape((b * per,) + tuple(tail[1:]))).to(dtype)
        col += w
    return out


def gather_outputs(local: "OrderedDict[str, torch.Tensor]", num_clips: int, num_segments: int,
                   local_clips: Optional[int] = None) -> "OrderedDict[str, torch.Tensor]":
    """Score gather: every rank receives the full-batch output dict in clip order (one collective).
    ``local_clips``: how many of the local rows are real (default: this rank's shard size; a rank whose shard is
    empty still calls this with the outputs of a stand-in clip and contributes zero rows)."""
    rank, ws = world()
    if ws == 1:
        return local
    counts = [len(shard_clips(num_clips, ws, r)) for r in range(ws)]
    mine = counts[rank] if local_clips is None else local_clips
    mx = max(max(counts), 1)
    mat, layout = pack_rows(local, mine, mx)
    full = mat.new_empty((ws * mx, mat.shape[1]))
    dist.all_gather_into_tensor(full, mat)
    if all(c == mx for c in counts):
        rows = full
    else:
        rows = torch.cat([full[r * mx:r * mx + c] for r, c in enumerate(counts)], 0)
    return unpack_rows(rows, layout)


def data_parallel_forward(model, inputs: Dict[str, torch.Tensor], gather: bool = True):
    """Shard -> local forward -> gather.  ``inputs`` is the FULL batch on every rank (or use
    ``shard_input`` upstream and call ``model`` directly).  With fewer clips than ranks (the last partial batch of
    an evaluation) the ranks without work run the forward on clip 0 as a stand-in, so that every rank reaches the
    collective with the same layout, and contribute no rows."""
    some = next(v for v in inputs.values() if torch.is_tensor(v) and v.dim() == 5)
    b, n = some.shape[:2]
    rank, ws = world()
    mine = len(shard_clips(b, ws, rank))
    if mine == 0:
        local = model(shard_input(inputs, rank=0, world_size=b))  # clip 0 only
    else:
        local = model(shard_input(inputs))
    return gather_outputs(local, b, n, local_clips=mine) if gather else local


def all_reduce_gradients(params, bucket_bytes: int = 32 << 20, average: bool = True) -> int:
    """Data-parallel gradient exchange of a training step: NCCL ``all_reduce`` (sum, then / world) of the ``.grad`` of
    ``params`` in flat buckets of ~``bucket_bytes`` -- what replaces ``nn.DataParallel``'s reduce-add of replica
    gradients (reference model_builder.py:73-75; train.py:81-94 then clips and steps).  On one NVSwitch box a bucket is
    latency / launch bound, not link bound, so few large buckets are used (27-38 M fp32 parameters = 4-5 buckets).
    Buckets are issued asynchronously in reverse parameter order (the order backward produces gradients) and awaited
    together.  Returns the number of collectives.  Parameters without a gradient are skipped on EVERY rank identically
    (requires_grad decides, not ``grad is None``: a rank with a missing gradient contributes zeros)."""
    rank, ws = world()
    plist = [p for p in params if p.requires_grad]
    if ws == 1 or not plist:
        return 0
    buckets, cur, size = [], [], 0
    for p in reversed(plist):
        cur.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    work = []
    for b in buckets:
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in b])
        work.append((b, flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)))
    for b, flat, w in work:
        w.wait()
        if average:
            flat /= ws
        off = 0
        for p in b:
            n = p.numel()
            g = flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
    return len(buckets)
