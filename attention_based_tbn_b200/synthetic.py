"""Seeded random-init weights and synthetic inputs for parity tests and the benchmark.

The reference's pretrained files (weights/download.sh) and the dataset are not available, so
parity and throughput are measured on random-init weights and synthetic tensors of the
BASELINE shapes.  Naive init makes the 1024-d feature input-independent
(hard part 1), which would make parity vacuous; the init here is
variance-preserving instead (He-normal convs, near-identity BatchNorm statistics, unit-variance
heads) so that features, attention weights and logits all depend on the input.

State-dict key names follow reference core/models/model.py:49-101 and bn_inception.py:38-107
(``Base_<M>.<cadene layer>``, ``pe.{0,1,2}``, ``attention_layer.attention_layer.*``,
``fusion.fusion_layer.0``, ``classifier.<key>``).  Everything is generated on the CPU with a
``torch.Generator`` so the same seed gives the same tensors on every box.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List

import torch

from .arch import FEATURE_SIZE, all_conv_specs

_IN_CH = {"RGB": 3, "Flow": 10, "Audio": 1}


def attn_win_size(audio_length: float) -> int:
    """round(audio_length * 25/4)  (reference model.py:60-61, dataset.py:540-541)."""
    return round(audio_length * (25 / 4))


def positional_encoding(dim_size: int, max_len: int) -> torch.Tensor:
    """Buffer ``pe.0.pe`` (1, dim, S): integer-frequency sin/cos (reference attention.py:25-33)."""
    pe = torch.zeros(max_len, dim_size)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1).expand(-1, dim_size // 2) * torch.arange(
        1, dim_size // 2 + 1
    )
    pe[:, 0::2] = torch.sin(position)
    pe[:, 1::2] = torch.cos(position)
    return pe.unsqueeze(0).transpose(1, 2).contiguous()


def make_state_dict(cfg, modality: List[str], seed: int = 0, calib: str | None = None) -> "OrderedDict[str, torch.Tensor]":
    """Random-init state dict.  ``calib``: path of an .npz with ``Base_*._bn.running_{mean,var}``
    entries (tests/golden/bn_calib.npz) that replace the random BatchNorm statistics."""
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def uniform(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for m in modality:
        p = f"Base_{m}."
        for c in all_conv_specs(_IN_CH[m]):
            fan_in = c.cin * c.k * c.k
            sd[p + c.name + ".weight"] = randn(c.cout, c.cin, c.k, c.k, std=math.sqrt(2.0 / fan_in))
            sd[p + c.name + ".bias"] = randn(c.cout, std=0.05)
            sd[p + c.name + "_bn.weight"] = uniform(c.cout, lo=0.8, hi=1.2)
            sd[p + c.name + "_bn.bias"] = randn(c.cout, std=0.1)
            sd[p + c.name + "_bn.running_mean"] = randn(c.cout, std=0.1)
            sd[p + c.name + "_bn.running_var"] = uniform(c.cout, lo=0.6, hi=1.4)
            sd[p + c.name + "_bn.num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    att = cfg.model.attention
    n_mod = len(modality)
    if n_mod > 1:
        if att.enable and not att.use_fixed:
            S = attn_win_size(cfg.data.audio.audio_length)
            if att.use_pe:
                sd["pe.0.pe"] = positional_encoding(10, S)
                sd["pe.1.weight"] = randn(1024, 1034, 1, std=1.0 / math.sqrt(1034))
                sd["pe.1.bias"] = randn(1024, std=0.05)
                sd["pe.2.weight"] = uniform(1024, lo=0.8, hi=1.2)
                sd["pe.2.bias"] = randn(1024, std=0.1)
            if att.type == "mha":
                a = "attention_layer.attention_layer."
                w_in = randn(3072, 1024, std=1.0 / math.sqrt(1024))
                w_in[:2048] *= 2.0  # sharper q.k scores so the attention weights are far from uniform
                sd[a + "in_proj_weight"] = w_in
                sd[a + "in_proj_bias"] = randn(3072, std=0.05)
                sd[a + "out_proj.weight"] = randn(1024, 1024, std=1.0 / math.sqrt(1024))
                sd[a + "out_proj.bias"] = randn(1024, std=0.05)
            elif att.type in ("unimodal", "proto"):
                a = "attention_layer.seq."
                n_out = S if att.type == "unimodal" else 3
                sd[a + "0.weight"] = randn(256, 1024, std=1.0 / math.sqrt(1024))
                sd[a + "0.bias"] = randn(256, std=0.05)
                sd[a + "2.weight"] = randn(n_out, 256, std=1.0 / math.sqrt(256))
                sd[a + "2.bias"] = randn(n_out, std=0.05)
                if att.type == "proto":  # buffer of 3 Gaussian prototypes (reference attention.py:122-133)
                    from .sampling import attention_prior

                    w1 = torch.from_numpy(attention_prior("gaussian", cfg.data.audio.audio_length)[:, 0].astype("float64"))
                    shift = S // 2 - 2
                    sd["attention_layer.prototype_wts"] = torch.stack(
                        (w1, torch.roll(w1, -shift), torch.roll(w1, shift))).float()
        sd["fusion.fusion_layer.0.weight"] = randn(512, FEATURE_SIZE * n_mod, std=1.0 / math.sqrt(FEATURE_SIZE * n_mod))
        sd["fusion.fusion_layer.0.bias"] = randn(512, std=0.05)
        cls_in = 512
    else:
        cls_in = FEATURE_SIZE
    for key, n_cls in cfg.model.num_classes.items():
        sd[f"classifier.{key}.weight"] = randn(n_cls, cls_in, std=1.0 / math.sqrt(cls_in))
        sd[f"classifier.{key}.bias"] = randn(n_cls, std=0.05)
    if calib is not None:
        import numpy as np

        with np.load(calib) as z:
            for k in z.files:
                if k in sd:
                    sd[k] = torch.from_numpy(z[k]).to(sd[k].dtype)
    return sd


def audio_frames(audio_length: float) -> int:
    """Spectrogram width: 1 + floor(len*24000/120) hop-120 frames (reference dataset.py:483-495)."""
    return 1 + int(audio_length * 24000) // 120


def make_inputs(cfg, modality: List[str], batch: int, num_segments: int, seed: int = 0,
                hw: int = 224) -> Dict[str, torch.Tensor]:
    """Synthetic input dict with the value ranges of the reference's transforms.

    RGB/Flow: ``x/255 - mean`` in about [-0.5, 0.6] (reference transform.py:495-497); Audio:
    ``log(power + 1e-6)`` in about [-13.8, 10] (dataset.py:483-495); ``weights`` (B,N,S,1) for
    the fixed-attention configs (dataset.py:540-548)."""
    g = torch.Generator().manual_seed(1000 + seed)
    out: Dict[str, torch.Tensor] = {}

    def field(c, h, w, coarse):
        """Per-frame smooth random field (bilinear-upsampled coarse noise) + white noise, so that
        different frames give different pooled features (white noise alone looks identical to
        a random conv net after global pooling)."""
        f = batch * num_segments
        lo = torch.randn(f, c, coarse[0], coarse[1], generator=g)
        lo = torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False)
        gain = torch.rand(f, 1, 1, 1, generator=g) * 1.5 + 0.25
        off = torch.randn(f, c, 1, 1, generator=g) * 0.5
        x = lo * gain + off + 0.25 * torch.randn(f, c, h, w, generator=g)
        return x.view(batch, num_segments, c, h, w)

    for m in modality:
        if m == "RGB":
            out[m] = (field(3, hw, hw, (6, 6)) * 0.25 + 0.05).clamp_(-0.5, 0.6)
        elif m == "Flow":
            out[m] = (field(10, hw, hw, (5, 5)) * 0.2).clamp_(-0.5, 0.5)
        else:
            t = audio_frames(cfg.data.audio.audio_length)
            out[m] = (field(1, 256, t, (8, 12)) * 3.0 - 5.0).clamp_(-13.8, 10.0)
    att = cfg.model.attention
    if att.enable and att.use_fixed and "Audio" in modality:
        S = attn_win_size(cfg.data.audio.audio_length)
        if att.prior_type == "uniform":
            w = torch.full((S, 1), 1.0 / S)
        else:  # gaussian, sigma=1 (cv2.getGaussianKernel(S, 1) restated)
            xs = torch.arange(S, dtype=torch.float64) - (S - 1) / 2.0
            k = torch.exp(-(xs ** 2) / 2.0)
            w = (k / k.sum()).float().unsqueeze(1)
        out["weights"] = w.expand(batch, num_segments, S, 1).contiguous()
    return out
