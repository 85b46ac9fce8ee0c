"""Segment sampling / indexing for the TBN hot path (host side, integer, bit-exact).

These indices define the inputs of the forward (which frames / flow stacks / audio windows make
up the (B, N, ...) tensors), so they are part of the hot-path contract (SURVEY.md section 8a, row
a16) and must match the reference bit for bit, including the order of draws from numpy's global
RNG in train mode.  They stay on the host (a few integers per clip).

Reference: core/dataset/dataset.py:156-172 (sync / async derivation, flow window expansion),
194-239 (``_get_offsets``), 439-451 (audio window start), 534-548 (prior attention weights);
core/dataset/epic_record.py:24-46 (start/end/num_frames); core/utils/misc.py:7-26 (modality order).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np


@dataclass(frozen=True)
class SegmentRecord:
    """One trimmed action segment (a row of the EPIC annotation file)."""

    start_frame_1b: int  # annotation "start_frame" (1-based)
    stop_frame_1b: int  # annotation "stop_frame"

    @property
    def start_frame(self) -> Dict[str, int]:  # epic_record.py:24-30
        s = self.start_frame_1b - 1
        return {"RGB": s, "Flow": s // 2, "Audio": s}

    @property
    def end_frame(self) -> Dict[str, int]:  # epic_record.py:32-38
        e = self.stop_frame_1b - 2
        return {"RGB": e, "Flow": e // 2, "Audio": e}

    @property
    def num_frames(self) -> Dict[str, int]:  # epic_record.py:40-46
        s, e = self.start_frame, self.end_frame
        return {m: e[m] - s[m] for m in ("RGB", "Flow", "Audio")}


def segment_offsets(rec: SegmentRecord, modality: str, num_segments: int, mode: str, frame_len: int = 1,
                    rng=None) -> np.ndarray:
    """``Video_Dataset._get_offsets`` (dataset.py:194-239) -> int64 (num_segments,).

    ``rng=None`` draws from numpy's global RNG exactly like the reference (seeded in main.py:20)."""
    nf = rec.num_frames[modality]
    if mode == "train":
        seg_len = (nf - frame_len + 1) // num_segments
    else:
        seg_len = nf // num_segments
    start = rec.start_frame[modality]
    if seg_len > 0:
        if mode == "train":
            randint = np.random.randint if rng is None else rng.randint
            offsets = randint(seg_len, size=num_segments)
        else:
            offsets = seg_len // 2
            if modality == "Flow":
                offsets = max(offsets - (frame_len // 2), 0)
        return (start + np.arange(0, num_segments) * seg_len + offsets).astype(np.int64)
    return start + np.zeros((num_segments), dtype=np.int64)


def sample_indices(rec: SegmentRecord, modality: List[str], num_segments: int, mode: str, sampling: str = "sync",
                   flow_win: int = 5, rng=None) -> "OrderedDict[str, np.ndarray]":
    """Per-modality segment indices of ``__getitem__`` (dataset.py:156-165)."""
    idx: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for m_no, m in enumerate(modality):
        if m_no > 0 and sampling == "sync":
            idx[m] = idx[modality[0]]
            if m == "Flow":
                idx[m] = (idx[m] / 2).astype(np.int64)
        else:
            idx[m] = segment_offsets(rec, m, num_segments, mode, flow_win if m == "Flow" else 1, rng)
    return idx


def flow_frame_indices(idx: np.ndarray, flow_win: int, num_segments: int) -> np.ndarray:
    """Each flow index expands to ``flow_win`` consecutive frames (dataset.py:168-172)."""
    return (idx.repeat(flow_win) + np.tile(np.arange(flow_win), num_segments)).astype(np.int64)


def audio_window_start(frame_idx: int, n_samples: int, vid_fps: int = 60, audio_length: float = 2.1,
                       sampling_rate: int = 24000) -> int:
    """First audio sample of the window centred on ``frame_idx`` (dataset.py:439-451)."""
    min_len = int(audio_length * sampling_rate)
    # Quirk kept from the reference: a clip shorter than the window is zero-padded, but max_len
    # is NOT refreshed afterwards, so the clamp below can return a negative start for it.
    max_len = n_samples
    start_sec = float(frame_idx / vid_fps) - (audio_length / 2)
    start = int(max(0, start_sec * sampling_rate))
    if start + min_len > max_len:
        start = max_len - min_len
    return start


def attention_prior(prior_type: str, audio_length: float) -> np.ndarray:
    """Fixed / prior attention weights (S,1) float32 (dataset.py:534-548; 'loud' needs the spectrogram
    and belongs to the data loader, out of scope)."""
    win = round(audio_length * (25 / 4))
    if prior_type == "uniform":
        return np.ones((win, 1), dtype=np.float32) / win
    if prior_type == "gaussian":  # cv2.getGaussianKernel(win, sigma=1), float64 then .float()
        xs = np.arange(win, dtype=np.float64) - (win - 1) / 2.0
        k = np.exp(-(xs * xs) / 2.0)
        return (k / k.sum()).reshape(win, 1).astype(np.float32)
    raise ValueError(f"prior_type '{prior_type}' is not computed on the hot path")


def shard_clips(num_clips: int, world_size: int, rank: int) -> range:
    """Contiguous clip shard of one rank, remainder to the low ranks."""
    base, rem = divmod(num_clips, world_size)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))
