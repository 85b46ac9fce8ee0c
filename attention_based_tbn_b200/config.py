"""Attribute-access config mirroring the reference's Hydra/OmegaConf groups.

The reference model reads its configuration by attribute access only
(reference core/models/model.py:35-96, 215-229, 300-331), so any attr-dict satisfies it; the
key names and defaults below are those of reference config/model/tbn_model.yaml:2-36,
config/data/tbn_data.yaml:2-33 and config/config.yaml:8-9.  Hydra itself (CLI, config groups)
is out of scope.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, List


class Cfg(dict):
    """dict with attribute access, recursively (``cfg.model.attention.enable``)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, Cfg):
            v = Cfg(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover - mirrors omegaconf's AttributeError
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return Cfg({k: copy.deepcopy(v, memo) for k, v in self.items()})


_DEFAULTS: Dict[str, Any] = {
    "num_workers": 8,
    "gpu_ids": [],
    "model": {
        "arch": "bninception",
        "attention": {
            "enable": True,
            "use_pe": True,
            "type": "mha",
            "use_gumbel": True,
            "use_fixed": False,
            "prior_type": "gaussian",
            "attn_heads": 4,
            "attn_dropout": 0.5,
            "use_prior": False,
            "wt_loss": "kl",
            "wt_decay": 0.25,
            "loss_reduction": "batchmean",
            "use_contrast": False,
            "contrast_thresh": 0.1,
            "contrast_decay": 0.25,
            "use_entropy": False,
            "entropy_decay": 0.25,
            "entropy_thresh": 0.2,
            "decay_step": 10,
        },
        "resnet": {"depth": 101},
        "vgg": {"type": "16"},
        "freeze_base": True,
        "freeze_mode": "partialbn",
        "num_classes": {"verb": 125, "noun": 352},
        "agg_type": "avg",
        "fusion_dropout": 0.5,
        "loss_fn": "crossentropy",
        "checkpoint_dir": "tbn_weights",
    },
    "data": {
        "dataset": "epic",
        "sampling": "sync",
        "rgb": {"enable": True, "mean": [0.408, 0.459, 0.502], "std": [1.0, 1.0, 1.0]},
        "flow": {"enable": True, "read_flow_pickle": False, "win_length": 5, "mean": [0.502], "std": [1.0]},
        "audio": {
            "enable": True,
            "read_audio_pickle": False,
            "sampling_rate": 24000,
            "audio_length": 2.1,
            "spec_type": "stft",
            "dropout": 0,
        },
        "vid_fps": 60,
        "train_scale_size": 256,
        "train_crop_size": 224,
        "test_scale_size": 256,
        "test_crop_size": 224,
        "manual_seed": 0,
    },
    "train": {"num_segments": 3},
    "val": {"num_segments": 25, "batch_size": 2, "topk": [1, 5]},
    "test": {"num_segments": 25, "batch_size": 2, "topk": [1, 5], "save_results": False,
             "results_file": ["seen.json", "unseen.json"]},
}


def default_cfg(**overrides) -> Cfg:
    """Reference defaults; ``overrides`` uses dotted keys: ``default_cfg(**{"model.attention.type": "mha"})``."""
    cfg = Cfg(copy.deepcopy(_DEFAULTS))
    for dotted, v in overrides.items():
        node = cfg
        parts = dotted.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return cfg


def get_modality(cfg) -> List[str]:
    """Fixed order RGB, Flow, Audio (reference core/utils/misc.py:7-26)."""
    m = []
    if cfg.data.rgb.enable:
        m.append("RGB")
    if cfg.data.flow.enable:
        m.append("Flow")
    if cfg.data.audio.enable:
        m.append("Audio")
    return m


# The five BASELINE.json configurations.
def baseline_config(idx: int) -> Cfg:
    if idx == 1:  # audio-only, no attention, 1.279 s
        return default_cfg(**{"data.rgb.enable": False, "data.flow.enable": False,
                              "model.attention.enable": False, "data.audio.audio_length": 1.279})
    if idx == 2:  # RGB+Audio, trainable MHA + entropy, sync
        return default_cfg(**{"data.flow.enable": False, "model.attention.type": "mha",
                              "model.attention.use_entropy": True, "data.sampling": "sync"})
    if idx == 3:  # RGB+Audio, fixed uniform prior attention
        return default_cfg(**{"data.flow.enable": False, "model.attention.use_fixed": True,
                              "model.attention.prior_type": "uniform", "data.sampling": "sync"})
    if idx == 4:  # RGB+Flow+Audio, no attention, async sampling
        return default_cfg(**{"model.attention.enable": False, "data.sampling": "async"})
    if idx == 5:  # RGB+Flow+Audio MHA, 25 segments
        return default_cfg(**{"model.attention.type": "mha", "model.attention.use_entropy": True})
    raise ValueError(f"BASELINE config index must be 1..5, got {idx}")
