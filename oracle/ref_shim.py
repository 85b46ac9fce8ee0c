"""Import the UNMODIFIED reference model in the build container -- TEST INFRASTRUCTURE.

Works only where ``/root/reference`` exists (the build container); nothing on the GPU box
imports this module.  Used by ``oracle/make_golden.py`` to produce ``tests/golden/*.npz``.

Two shims:
  1. ``pretrainedmodels`` is not installed and not vendored.  A stand-in
     ``pretrainedmodels.models.bninception.BNInception`` is derived AT RUN TIME from the
     reference's own vendored layer table (core/models/bn_inception_audio.py): the class is
     subclassed, its experimental split stem (lines 11-20 / 408-416) is swapped for the
     Cadene 7x7 stem that the same file keeps as comments (35-39 / 424-426), and the 3a..5b
     dataflow (437-1003) is executed verbatim from the reference source.  No reference source
     is copied into this repository.
  2. ``bninception()`` does ``torch.load`` of weight files that are absent
     (bn_inception.py:65-70): ``torch.load`` is patched, for the duration of the model
     construction only, to return the state-dict of a freshly constructed stand-in backbone.
     The test weights are then applied with the model's own ``load_state_dict``.
"""
from __future__ import annotations

import importlib.util
import inspect
import os
import re
import sys
import textwrap
import types
from contextlib import contextmanager

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("TBN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "core", "models", "model.py"))


def _build_cadene_standin():
    path = os.path.join(REF_ROOT, "core", "models", "bn_inception_audio.py")
    spec = importlib.util.spec_from_file_location("_tbn_ref_bn_inception_audio", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Audio = mod.BNInception_Audio

    # features(): keep the reference dataflow from conv2_3x3_reduce on; replace the split stem.
    src = textwrap.dedent(inspect.getsource(Audio.features))
    marker = "conv2_3x3_reduce_out = self.conv2_3x3_reduce(pool1_3x3_s2_out)"
    head, tail = src.split(marker, 1)
    assert head.lstrip().startswith("def features(self, input):")
    stem = (
        "def features(self, input):\n"
        "    conv1_7x7_s2_out = self.conv1_7x7_s2(input)\n"
        "    conv1_7x7_s2_bn_out = self.conv1_7x7_s2_bn(conv1_7x7_s2_out)\n"
        "    conv1_relu_7x7_out = self.conv1_relu_7x7(conv1_7x7_s2_bn_out)\n"
        "    pool1_3x3_s2_out = self.pool1_3x3_s2(conv1_relu_7x7_out)\n"
        "    "
    )
    ns = {"torch": torch, "nn": nn, "F": torch.nn.functional}
    exec(compile(stem + marker + tail, path + ":features[cadene-stem]", "exec"), ns)
    features_fn = ns["features"]

    class BNInception(Audio):  # stand-in for pretrainedmodels.models.bninception.BNInception
        def __init__(self, num_classes=1000):
            super().__init__(num_classes=num_classes)
            for nm in ("conv1_1x3_s2", "conv1_1x3_s2_bn", "conv1_relu_1x3",
                       "conv1_3x1_s2", "conv1_3x1_s2_bn", "conv1_relu_3x1"):
                delattr(self, nm)
            self.conv1_7x7_s2 = nn.Conv2d(3, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3))
            self.conv1_7x7_s2_bn = nn.BatchNorm2d(64, affine=True)
            self.conv1_relu_7x7 = nn.ReLU(True)
            # Cadene defines the stem first: keep its child order (reference model.py:170-176 freezes BatchNorm layers by
            # their position among the children -- "except first": the stem's BatchNorm, child 1, stays trainable)
            first = ("conv1_7x7_s2", "conv1_7x7_s2_bn", "conv1_relu_7x7")
            mods = dict(self._modules)
            self._modules.clear()
            for nm in first:
                self._modules[nm] = mods[nm]
            for nm, m in mods.items():
                if nm not in first:
                    self._modules[nm] = m

        features = features_fn

        def logits(self, features):  # overridden by the reference subclass
            x = self.global_pool(features)
            return self.last_linear(x.view(x.size(0), -1))

        def forward(self, input):
            return self.logits(self.features(input))

    return BNInception


def install_shims():
    if "pretrainedmodels.models.bninception" in sys.modules:
        return sys.modules["pretrainedmodels.models.bninception"].BNInception
    cls = _build_cadene_standin()
    ptm = types.ModuleType("pretrainedmodels")
    ptm_models = types.ModuleType("pretrainedmodels.models")
    ptm_bni = types.ModuleType("pretrainedmodels.models.bninception")
    ptm_bni.BNInception = cls
    ptm_models.bninception = ptm_bni
    ptm.models = ptm_models
    sys.modules["pretrainedmodels"] = ptm
    sys.modules["pretrainedmodels.models"] = ptm_models
    sys.modules["pretrainedmodels.models.bninception"] = ptm_bni
    if "librosa" not in sys.modules:  # core/dataset/dataset.py:10; never called by the index math
        try:
            import librosa  # noqa: F401
        except Exception:
            sys.modules["librosa"] = types.ModuleType("librosa")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return cls


@contextmanager
def _patched_torch_load(standin_cls):
    real = torch.load

    def fake(f, *a, **k):
        name = os.path.basename(str(f))
        if name == "kinetics_bninception_flow.pth":
            m = standin_cls(num_classes=400)
            m.conv1_7x7_s2 = nn.Conv2d(10, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3))
            return m.state_dict()
        if name == "imagenet_bninception_rgb.pth":
            return standin_cls(num_classes=1000).state_dict()
        return real(f, *a, **k)

    torch.load = fake
    try:
        yield
    finally:
        torch.load = real


def build_reference_model(cfg, modality, state_dict=None):
    """Reference ``TBNModel(cfg, modality, cpu).eval()`` with ``state_dict`` loaded (strict)."""
    cls = install_shims()
    from core.models.model import TBNModel  # the unmodified reference class

    with _patched_torch_load(cls):
        model = TBNModel(cfg, modality, torch.device("cpu"))
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model.eval()


def reference_video_dataset_cls():
    """Reference ``Video_Dataset`` class (for the segment-index oracle)."""
    install_shims()
    from core.dataset.dataset import Video_Dataset

    return Video_Dataset
