"""``build_model`` -- drop-in for reference core/models/model_builder.py:25-81.

Same signature and return value ``(model, criterion: OrderedDict, num_gpus)`` and the same
assertions.  The reference wraps the model in a single-process ``nn.DataParallel`` when several
GPUs are configured (model_builder.py:73-75: per-step parameter broadcast + scatter + gather);
the B200 design is one process per GPU instead (``parallel.py``), so under ``torchrun`` every
rank builds its own replica here and ``num_gpus`` reports the world size.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .model import TBNModel

_MODEL_TYPES = {"bninception": TBNModel}  # vgg / resnet bases: out of scope

_LOSS_TYPES = {
    "crossentropy": torch.nn.CrossEntropyLoss,
    "nll": torch.nn.NLLLoss,
    "kl": torch.nn.KLDivLoss,
    "mse": torch.nn.MSELoss,
    "smoothl1": torch.nn.SmoothL1Loss,
}


class ContrastLoss(torch.nn.Module):
    """Optional training-only loss on the attention weights (reference core/models/contrast_loss.py):
    weights below the threshold are pushed down, the others up.  Host glue, off by default."""

    def __init__(self, threshold=0.5, reduction=None):
        super().__init__()
        if reduction not in ("mean", "batchmean", "sum"):
            raise Exception(f"{reduction} type reduction not supported for Contrast Loss")
        self.threshold, self.reduction = threshold, reduction

    def forward(self, wts):
        sign = torch.where(wts.detach() >= self.threshold, -1.0, 1.0)
        loss = (wts * sign).sum(dim=1)  # per row: sum(below) - sum(at/above)
        return loss.mean() if self.reduction in ("mean", "batchmean") else loss


def _criteria(cfg, device) -> "OrderedDict[str, torch.nn.Module]":
    """Loss modules keyed like the reference's ``criterion`` dict: the classification loss under its config name, plus
    ``"prior"`` / ``"contrast"`` for the optional attention losses (model_builder.py:56-70)."""
    att = cfg.model.attention
    wanted = [(cfg.model.loss_fn, lambda: _LOSS_TYPES[cfg.model.loss_fn]())]
    if att.enable and att.use_prior:
        wanted.append(("prior", lambda: _LOSS_TYPES[att.wt_loss](reduction=att.loss_reduction)))
    if att.enable and att.use_contrast:
        wanted.append(("contrast", lambda: ContrastLoss(threshold=att.contrast_thresh, reduction=att.loss_reduction)))
    return OrderedDict((name, make().to(device)) for name, make in wanted)


def _gpu_count(cfg, device: torch.device) -> int:
    """World size under torchrun (one process per GPU); otherwise what the reference reports: len(cfg.gpu_ids) or every
    visible device (model_builder.py:46-53), with the same availability assertion."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size()
    n = len(cfg.gpu_ids) if len(cfg.gpu_ids) > 0 else torch.cuda.device_count()
    if device.type == "cuda":
        assert n <= torch.cuda.device_count(), "Cannot use more GPU devices than available"
    return n


def build_model(cfg, modality, device):
    arch, loss_fn = cfg.model.arch, cfg.model.loss_fn
    assert arch in _MODEL_TYPES, "Model type '{}' not supported".format(arch)
    assert loss_fn in _LOSS_TYPES, "Loss type '{}' not supported".format(loss_fn)
    device = torch.device(device)
    num_gpus = _gpu_count(cfg, device)
    model = _MODEL_TYPES[arch](cfg, modality, device).to(device)
    return model, _criteria(cfg, device), num_gpus
