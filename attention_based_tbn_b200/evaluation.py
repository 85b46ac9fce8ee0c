"""Score consumers of ``TBNModel.forward`` -- the eval-loop side of the hot path.

Same names, arguments and results as the reference:

* ``Metric`` (core/utils/metric.py:4-157): top-k accuracy, per-class confusion matrices, running losses.
  The reference finds the top-k with ``Tensor.topk`` and then walks the batch in a Python loop that indexes
  device tensors element by element (metric.py:154-155: two device->host reads per clip and class key, plus one
  ``.item()`` per loss and accuracy).  Here one ``tbn_topk_confusion`` launch per class key produces the
  correctness flags and updates the confusion matrix on the device; counts and losses stay on the device until
  ``get_metrics()`` and are then folded on the host with the reference's exact arithmetic (float32 count times
  float32 ``100 / batch_size``, python-float accumulation, ``round``), so the reported numbers are identical.
* ``save_scores`` (core/utils/misc.py:115-165): the EPIC-Kitchens action-recognition JSON wire format.
* ``test`` (core/tools/test.py:22-95): the evaluation loop over a data loader.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import Dict, List

import numpy as np
import torch

from . import _lib


def topk_correct(out: torch.Tensor, target: torch.Tensor, maxk: int, conf_mat: torch.Tensor = None, rank: torch.Tensor = None):
    """``Metric._get_correct_score`` (metric.py:132-157) on the device: returns ``correct`` (maxk, B) bool and the
    predictions (maxk, B) int64; ``conf_mat`` (C, C) fp32 is updated in place when given; ``rank`` (B) int32 receives
    the position of the target among the top ``maxk`` classes (``maxk`` = missed)."""
    if not out.is_cuda:
        raise _lib.TbnKernelError("topk_correct needs CUDA tensors (there is no CPU fallback)")
    out = out.detach()
    if out.dtype != torch.float32 or not out.is_contiguous():
        out = out.float().contiguous()
    tgt = target.detach().reshape(-1).to(device=out.device, dtype=torch.int64).contiguous()
    b, c = out.shape
    if tgt.numel() != b:
        raise ValueError(f"target has {tgt.numel()} entries for {b} rows of scores")
    if conf_mat is not None and (conf_mat.dtype != torch.float32 or tuple(conf_mat.shape) != (c, c) or not conf_mat.is_contiguous()):
        raise ValueError("conf_mat must be a contiguous fp32 (C, C) tensor")
    correct = torch.empty((maxk, b), dtype=torch.uint8, device=out.device)
    preds = torch.empty((maxk, b), dtype=torch.int64, device=out.device)
    _lib.call("tbn_topk_confusion", out.data_ptr(), tgt.data_ptr(), b, c, maxk, correct.data_ptr(), preds.data_ptr(),
              _lib.ptr(rank), _lib.ptr(conf_mat), torch.cuda.current_stream(out.device).cuda_stream)
    return correct.bool(), preds


class Metric(object):
    """Drop-in for the reference ``Metric(cfg, no_batches, device)``; see the module docstring."""

    def __init__(self, cfg, no_batches, device=torch.device("cuda")):
        self.cfg = cfg
        self.topk = list(self.cfg.val.topk)
        self.device = device
        self.no_batches = no_batches
        self.multi_class = True if len(self.cfg.model.num_classes.keys()) > 1 else False

        self.loss = {}
        self.accuracy = {}
        self.conf_mat = {}
        for key, no_cls in self.cfg.model.num_classes.items():
            self.accuracy[key] = [0] * (len(self.topk))
            self.conf_mat[key] = torch.zeros((no_cls, no_cls), device=device)
            self.loss[key] = 0
        if self.multi_class:
            self.loss["all_class"] = 0
            self.accuracy["all_class"] = [0] * (len(self.topk))
        att = self.cfg.model.attention
        if att.enable and not att.use_fixed:
            if att.use_prior:
                self.loss["prior"] = 0
            if att.use_contrast:
                self.loss["contrast"] = 0
            if att.use_entropy:
                self.loss["entropy"] = 0
        self.loss["total"] = 0
        # device-side results of the batches seen since the last fold: (counts (n_acc_keys, n_topk) int64,
        # stacked losses in self.loss order, batch_size)
        self._pending: List = []
        self._ks = None

    def set_metrics(self, out, target, batch_size, batch_loss):
        maxk = max(self.topk)
        cls_keys = [k for k in self.accuracy.keys() if k != "all_class"]
        dev = out[cls_keys[0]].device
        nb = out[cls_keys[0]].shape[0]
        # ranks[r, i] = position of clip i's target among the maxk best classes of key r ("correct within the top k"
        # is rank < k); the last row serves "all_class": every key right within the top k (metric.py:97-103)
        ranks = torch.empty((len(self.accuracy), nb), dtype=torch.int32, device=dev)
        for r, key in enumerate(cls_keys):
            topk_correct(out[key], target["class"][key], maxk, self.conf_mat[key], ranks[r])
        if self.multi_class:
            torch.amax(ranks[:-1], dim=0, out=ranks[-1])
        if self._ks is None:
            self._ks = torch.tensor(self.topk, dtype=torch.int32, device=dev)
        counts = (ranks[:, None, :] < self._ks[None, :, None]).sum(2)
        losses = torch.stack([batch_loss[k].detach().reshape(()).float() for k in self.loss.keys()])
        self._pending.append((counts, losses, batch_size))

    def _fold(self):
        """One device->host read for everything pending, then the reference's host arithmetic batch by batch."""
        if not self._pending:
            return
        counts = torch.stack([p[0] for p in self._pending]).cpu().numpy()
        losses = torch.stack([p[1] for p in self._pending]).cpu().numpy()
        for bi, (_, _, batch_size) in enumerate(self._pending):
            scale = np.float32(100.0 / batch_size)
            for li, key in enumerate(self.loss.keys()):
                self.loss[key] += float(losses[bi, li])
            for ai, key in enumerate(self.accuracy.keys()):
                for i in range(len(self.topk)):
                    self.accuracy[key][i] += float(np.float32(counts[bi, ai, i]) * scale)
        self._pending = []

    def get_metrics(self):
        self._fold()
        for key in self.accuracy.keys():
            self.accuracy[key] = [round(x / self.no_batches, 2) for x in self.accuracy[key]]
        for key in self.loss.keys():
            self.loss[key] = round(self.loss[key] / self.no_batches, 5)
        return self.loss, self.accuracy, self.conf_mat

    _get_correct_score = staticmethod(
        lambda out, target, topk, device: _correct_and_confusion(out, target, topk, device))


def _correct_and_confusion(out, target, topk, device):
    conf = torch.zeros((out.size(1), out.size(1)), device=device)
    correct, _ = topk_correct(out, target, max(topk), conf)
    return correct, conf


def save_scores(scores: Dict[str, list], file_name: str, action_names) -> None:
    """The reference's ``save_scores`` (misc.py:115-165): concatenates the per-batch lists in ``scores`` in place and
    writes ``{"version": "0.1", "challenge": "action_recognition", "results": {action_id: {key: {class: score}}}}``
    (``indent=4``).  ``"action"`` keeps the 100 best classes by name.  One device->host copy per key instead of one
    ``.item()`` per score."""
    out_result = {}
    out_result["version"] = "0.1"
    out_result["challenge"] = "action_recognition"
    for key in scores.keys():
        scores[key] = torch.cat(scores[key], dim=0)
    host = {}
    for key in scores.keys():
        if key == "action_id":
            host[key] = scores[key].cpu().tolist()
        elif key == "action":
            top_s, top_i = scores[key].topk(100, dim=1, largest=True, sorted=True)
            host[key] = (top_i.cpu().tolist(), top_s.cpu().tolist())
        else:
            host[key] = scores[key].cpu().tolist()
    results = {}
    for idx in range(len(host["action_id"])):
        a_id = str(host["action_id"][idx])
        results[a_id] = {}
        for key in scores.keys():
            if key == "action_id":
                continue
            elif key == "action":
                ids, vals = host[key][0][idx], host[key][1][idx]
                results[a_id][key] = {action_names[i]: s for i, s in zip(ids, vals)}
            else:
                results[a_id][key] = {str(i): s for i, s in enumerate(host[key][idx])}
    out_result["results"] = results
    os.makedirs(os.path.split(file_name)[0], exist_ok=True)
    with open(file_name, "w") as f:
        json.dump(out_result, f, indent=4)


def _to_device(d, device):
    for k, v in d.items():
        if isinstance(v, dict):
            _to_device(v, device)
        elif torch.is_tensor(v):
            d[k] = v.to(device, non_blocking=True)
    return d


def test(cfg, model, data_loader, criterion, modality, logger=None, device=torch.device("cuda")):
    """The reference's evaluation loop (core/tools/test.py:22-95): returns ``(loss, accuracy, conf_mat[, output])``."""
    no_batches = round(len(data_loader.dataset) / data_loader.batch_size)
    metric = Metric(cfg, no_batches, device)
    model.eval()
    save = bool(cfg.test.save_results)
    if save:
        output = {"action_id": []}
        for key in cfg.model.num_classes.keys():
            output[key] = []
    with torch.no_grad():
        for data, target, action_id in data_loader:
            data, target = _to_device(data, device), _to_device(target, device)
            out = model(data)
            if isinstance(target["class"], dict):
                loss, batch_size = model.get_loss(criterion, target, out)
                metric.set_metrics(out, target, batch_size, loss)
            if save:
                output["action_id"].extend([action_id])
                for key in cfg.model.num_classes.keys():
                    output[key].extend([out[key]])
    test_loss, test_acc, conf_mat = metric.get_metrics()
    if save:
        return (test_loss, test_acc, conf_mat, output)
    return (test_loss, test_acc, conf_mat)
