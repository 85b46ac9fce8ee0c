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


def build_model(cfg, modality, device):
    assert cfg.model.arch in _MODEL_TYPES.keys(), "Model type '{}' not supported".format(cfg.model.arch)
    assert cfg.model.loss_fn in _LOSS_TYPES.keys(), "Loss type '{}' not supported".format(cfg.model.loss_fn)
    device = torch.device(device)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        num_gpus = torch.distributed.get_world_size()
    elif len(cfg.gpu_ids) > 0:
        num_gpus = len(cfg.gpu_ids)
    else:
        num_gpus = torch.cuda.device_count()
    if device.type == "cuda" and not torch.distributed.is_initialized():
        assert num_gpus <= torch.cuda.device_count(), "Cannot use more GPU devices than available"

    model = _MODEL_TYPES[cfg.model.arch](cfg, modality, device)

    criterion = OrderedDict()
    criterion[cfg.model.loss_fn] = _LOSS_TYPES[cfg.model.loss_fn]()
    if cfg.model.attention.enable:
        if cfg.model.attention.use_prior:
            criterion["prior"] = _LOSS_TYPES[cfg.model.attention.wt_loss](reduction=cfg.model.attention.loss_reduction)
        if cfg.model.attention.use_contrast:
            criterion["contrast"] = ContrastLoss(threshold=cfg.model.attention.contrast_thresh,
                                                 reduction=cfg.model.attention.loss_reduction)
    model = model.to(device)
    for key in criterion.keys():
        criterion[key] = criterion[key].to(device)
    return model, criterion, num_gpus
