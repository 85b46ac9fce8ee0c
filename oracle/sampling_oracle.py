"""Plain-Python restatement of the reference's segment index math -- TEST INFRASTRUCTURE.

Follows reference core/dataset/dataset.py:156-172 and 194-239, core/dataset/epic_record.py:24-46,
written as scalar loops (no vectorisation) so that it is independent of the product's numpy code.
Pinned by ``tests/golden/sampling.npz`` which ``oracle/make_golden.py`` generates by running
the reference's own ``Video_Dataset._get_offsets`` on its ``annotations/epic_train_val.csv``.
"""
from __future__ import annotations

import numpy as np


def _frames(start_1b, stop_1b, m):
    s, e = start_1b - 1, stop_1b - 2
    if m == "Flow":
        s, e = s // 2, e // 2
    return s, e - s


def offsets(start_1b, stop_1b, m, n_seg, mode, flow_win=5, draw=None):
    """draw(seg_len, n) supplies the train-mode random offsets (np.random.randint semantics)."""
    frame_len = flow_win if m == "Flow" else 1
    start, nf = _frames(start_1b, stop_1b, m)
    seg_len = ((nf - frame_len + 1) // n_seg) if mode == "train" else (nf // n_seg)
    if seg_len <= 0:
        return [int(start)] * n_seg
    if mode == "train":
        off = [int(v) for v in draw(seg_len, n_seg)]
    else:
        o = seg_len // 2
        if m == "Flow":
            o = max(o - frame_len // 2, 0)
        off = [o] * n_seg
    return [int(start + i * seg_len + off[i]) for i in range(n_seg)]


def indices(start_1b, stop_1b, modality, n_seg, mode, sampling, flow_win=5, draw=None):
    out = {}
    for i, m in enumerate(modality):
        if i > 0 and sampling == "sync":
            base = out[modality[0]]
            # reference: (idx / 2).astype(np.int64) -- true division then truncation
            out[m] = [int(np.int64(np.float64(v) / 2)) for v in base] if m == "Flow" else list(base)
        else:
            out[m] = offsets(start_1b, stop_1b, m, n_seg, mode, flow_win, draw)
    return out


def flow_frames(idx, flow_win):
    return [v + j for v in idx for j in range(flow_win)]
