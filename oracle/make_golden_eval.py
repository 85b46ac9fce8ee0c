"""Generate tests/golden/eval_metric.npz and tests/golden/save_scores.json by running the UNMODIFIED reference
``Metric`` (core/utils/metric.py) and ``save_scores`` (core/utils/misc.py) on seeded inputs -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):  ``python -m oracle.make_golden_eval``
The inputs are regenerated from the seed by ``tests/cases.py::eval_case`` wherever the vectors are checked.

One shim for torch-version drift: metric.py:104 calls ``.view(-1)`` on ``preds.t().eq(...)[:k]``; torch 1.4 (the
reference's pin) returned a contiguous tensor from ``eq`` there, torch 2.x keeps the transposed strides and ``view``
raises.  For the duration of the run ``Tensor.view`` falls back to ``reshape`` (same values) -- the reference source is
not modified.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import eval_case, save_scores_case  # noqa: E402
from attention_based_tbn_b200.config import default_cfg  # noqa: E402

REF_ROOT = os.environ.get("TBN_REFERENCE_ROOT", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _view_or_reshape(self, *shape, **kw):
    try:
        return _ORIG_VIEW(self, *shape, **kw)
    except RuntimeError:
        return self.reshape(*shape)


_ORIG_VIEW = torch.Tensor.view


def main():
    torch.Tensor.view = _view_or_reshape
    metric_mod = _load("_tbn_ref_metric", "core/utils/metric.py")
    misc_mod = _load("_tbn_ref_misc", "core/utils/misc.py")
    store = {}
    for name in ("two_class", "one_class"):
        cfg, batches = eval_case(name)
        m = metric_mod.Metric(cfg, len(batches), torch.device("cpu"))
        for out, target, bs, loss in batches:
            m.set_metrics(out, target, bs, loss)
        loss, acc, conf = m.get_metrics()
        for k, v in loss.items():
            store[f"{name}/loss/{k}"] = np.float64(v)
        for k, v in acc.items():
            store[f"{name}/acc/{k}"] = np.asarray(v, np.float64)
        for k, v in conf.items():
            store[f"{name}/conf/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "eval_metric.npz"), **store)
    scores, names = save_scores_case()
    misc_mod.save_scores(scores, os.path.join(GOLDEN, "save_scores.json"), names)
    print("wrote eval_metric.npz, save_scores.json")


if __name__ == "__main__":
    main()
