"""``TBNModel`` -- drop-in for the reference's model class, executed by the sm_100a kernels.

Same constructor ``TBNModel(cfg, modality, device)``, same config keys, same parameter / buffer
names (checkpoints interchange with the reference), same ``forward(input: dict) -> OrderedDict``
contract: ``{class key: (B, n_cls) fp32, ..., "weights": (B*N, 1, S)}``
(reference core/models/model.py:20-101, 205-262).  The module objects below are parameter
CONTAINERS only: the forward never calls a torch op on them, it runs the execution plan
(``plan.py``) through the C ABI.  There is no CPU / PyTorch fallback -- a forward on a non-CUDA
device, or without libtbn_sm100.so, raises.

Extra (non-reference) knobs, all optional:
  * ``precision``: "bf16" (default; tcgen05 tensor-core backbone, fp32 heads), "fp32"
    (CUDA-core exact path, the 1e-4 parity mode) or "bf16_simt" (bf16 storage, CUDA-core math:
    the on-device checker of the tensor-core path).  Also read from ``cfg.model.precision``.
  * ``chunk``: frames per backbone pass (L2 residency), ``cfg.model.frame_chunk``.
  * ``use_cuda_graph``: replay the whole forward as one CUDA graph when the input tensors keep
    their addresses.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .arch import FEATURE_SIZE, all_conv_specs, block_specs, stem_specs
from .backbone_train import TrainBackbone
from .head_train import HeadBackward, HeadGradFn, head_parameter_names
from .plan import PRECISIONS, BackbonePlan, HeadPlan, JointPhaseB
from .synthetic import attn_win_size, positional_encoding

_IN_CH = {"RGB": 3, "Flow": 10, "Audio": 1}


class BNInceptionParams(nn.Module):
    """Parameter container with the Cadene BNInception names and child order
    (reference bn_inception.py:38-107; layer table bn_inception_audio.py:21-39, 58-401)."""

    def __init__(self, in_channels: int, is_audio: bool = False, attend: bool = False):
        super().__init__()
        self.is_audio, self.attend, self.feature_size = is_audio, attend, FEATURE_SIZE
        self.in_channels = in_channels

        def triple(spec, relu_name):
            self.add_module(spec.name, nn.Conv2d(spec.cin, spec.cout, spec.k, spec.stride, spec.pad))
            self.add_module(spec.name + "_bn", nn.BatchNorm2d(spec.cout, affine=True))
            self.add_module(relu_name, nn.ReLU(True))

        c1, c2r, c2 = stem_specs(in_channels)
        triple(c1, "conv1_relu_7x7")
        self.pool1_3x3_s2 = nn.MaxPool2d(3, 2, ceil_mode=True)
        triple(c2r, "conv2_relu_3x3_reduce")
        triple(c2, "conv2_relu_3x3")
        self.pool2_3x3_s2 = nn.MaxPool2d(3, 2, ceil_mode=True)
        for b in block_specs():
            p = f"inception_{b.name}_"
            for spec in ([b.b1x1] if b.b1x1 else []) + list(b.b3x3) + list(b.bdbl):
                triple(spec, p + "relu_" + spec.name[len(p):])
            if b.pool == "avg":
                self.add_module(p + "pool", nn.AvgPool2d(3, 1, 1, ceil_mode=True, count_include_pad=True))
            elif b.pool == "max_s1":
                self.add_module(p + "pool", nn.MaxPool2d(3, 1, 1, ceil_mode=True))
            else:
                self.add_module(p + "pool", nn.MaxPool2d(3, 2, ceil_mode=True))
            if b.bpool:
                triple(b.bpool, p + "relu_pool_proj")
        for m in self.modules():  # variance-preserving init; real use loads a checkpoint
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
                nn.init.zeros_(m.bias)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("BNInceptionParams is a parameter container; TBNModel.forward runs the CUDA plan")


class _PEBuffer(nn.Module):
    """``pe.0``: holds the (1, 10, S) positional-encoding buffer (reference attention.py:8-45)."""

    def __init__(self, dim_size: int, max_len: int):
        super().__init__()
        self.dim_size, self.max_len = dim_size, max_len
        self.register_buffer("pe", positional_encoding(dim_size, max_len))


class _MHAParams(nn.Module):
    """``attention_layer.attention_layer.*`` of the reference wrapper (attention.py:48-57)."""

    def __init__(self, embed_dim: int, heads: int, dropout: float):
        super().__init__()
        self.attention_layer = nn.MultiheadAttention(embed_dim, heads, dropout=dropout, bias=True)


class _MLPAttnParams(nn.Module):
    """``attention_layer.seq.*`` of UniModalAttention / PrototypeAttention (attention.py:60-145)."""

    def __init__(self, in_size: int, n_out: int, hidden: int = 256, prototypes: Optional[torch.Tensor] = None):
        super().__init__()
        self.seq = nn.Sequential(nn.Linear(in_size, hidden), nn.ReLU(), nn.Linear(hidden, n_out))
        if prototypes is not None:
            self.register_buffer("prototype_wts", prototypes)


class Fusion(nn.Module):
    """Linear(in, 512)+ReLU(+Dropout in training) parameters (reference model.py:337-362)."""

    def __init__(self, in_size: int, out_size: int, dropout: float = 0):
        super().__init__()
        self.in_size, self.out_size, self.dropout = in_size, out_size, dropout
        self.fusion_layer = nn.Sequential(nn.Linear(in_size, out_size), nn.ReLU())
        nn.init.normal_(self.fusion_layer[0].weight, 0, 1e-3)
        nn.init.constant_(self.fusion_layer[0].bias, 0)
        if dropout > 0:
            self.dropout_layer = nn.Dropout(p=dropout)


class Classifier(nn.Module):
    """One Linear per class key (reference model.py:365-386)."""

    def __init__(self, num_classes, in_features: int):
        super().__init__()
        self.num_classes = num_classes
        for cls in num_classes.keys():
            self.add_module(cls, nn.Linear(in_features, num_classes[cls]))
            nn.init.normal_(getattr(self, cls).weight, 0, 1e-3)
            nn.init.constant_(getattr(self, cls).bias, 0)


class TBNModel(nn.Module):
    """Temporal Binding Network (see module docstring)."""

    def __init__(self, cfg, modality: List[str], device, precision: Optional[str] = None,
                 chunk: Optional[int] = None, use_cuda_graph: bool = True, chunk_lo: Optional[int] = None,
                 multi_stream: bool = True, fuse_1x1: bool = True, autotune: Optional[bool] = None,
                 group_launches: Optional[bool] = None, train_precision: Optional[str] = None):
        super().__init__()
        self.cfg = cfg
        self.modality = list(modality)
        self.base_model_name = cfg.model.arch
        self.num_classes = cfg.model.num_classes
        self.use_attention = cfg.model.attention.enable
        self.attention_type = cfg.model.attention.type
        self.device = torch.device(device)
        if self.base_model_name != "bninception":
            raise NotImplementedError(
                f"arch '{self.base_model_name}': only 'bninception' is on the B200 hot path (vgg/resnet are out of scope)")
        if cfg.model.agg_type.lower() == "avg":
            self.agg_type = "avg"
        else:
            print("Incorrect aggregation type")
            self.agg_type = None

        mcfg = cfg.model
        self.precision = precision or (mcfg["precision"] if "precision" in mcfg else os.environ.get("TBN_PRECISION", "bf16"))
        # arithmetic of the TRAINING step's convs: "fp32" (default: the reference trains in fp32 -- the parity mode) or "bf16"
        # (mixed precision: conv products on the tensor cores from bf16 operand copies, everything stored in fp32)
        self.train_precision = train_precision or (mcfg["train_precision"] if "train_precision" in mcfg
                                                   else os.environ.get("TBN_TRAIN_PRECISION", "fp32"))
        if self.train_precision not in ("fp32", "bf16"):
            raise ValueError(f"train_precision must be 'fp32' or 'bf16', got {self.train_precision!r}")
        if self.precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}, got {self.precision}")
        self.chunk = int(chunk or (mcfg["frame_chunk"] if "frame_chunk" in mcfg else 32))
        self.use_cuda_graph = use_cuda_graph
        self.chunk_lo, self.multi_stream, self.fuse_1x1 = chunk_lo, multi_stream, fuse_1x1
        # autotune: time the conv kernel's tiling variants per layer on the first forward of a new shape (bf16 mode)
        self.autotune = bool(int(os.environ.get("TBN_AUTOTUNE", "0"))) if autotune is None else bool(autotune)
        self._tune_cache: Dict[tuple, dict] = {}
        # experiment switch: run the modality backbones one after the other (each still on three stream lanes)
        self.sequential_modalities = bool(int(os.environ.get("TBN_SEQ_MODALITIES", "0")))
        # grouped launches (bf16 tensor-core mode): the same layer of all modalities + the independent 3x3 convs of a block
        # run as ONE tbn_conv_tc_group launch (plan.JointPhaseB).  None / TBN_GROUP=auto (default): with CUDA graphs both
        # plans are captured for a new shape and the faster replay is kept (grouped launches spend less kernel time, the
        # per-modality plan overlaps more kernels -- which wins depends on the batch); eager runs use the grouped plan.
        env = os.environ.get("TBN_GROUP", "auto")
        self.group_launches = (None if env == "auto" else bool(int(env))) if group_launches is None else bool(group_launches)
        self.plan_choice: Dict[tuple, str] = {}  # shape signature -> "grouped" / "per-modality" (auto mode)

        in_features = 0
        for m in self.modality:
            base = BNInceptionParams(_IN_CH[m], is_audio=(m == "Audio"), attend=bool(self.use_attention))
            self._maybe_load_pretrained(base, m)
            self.add_module(f"Base_{m}", base)
            in_features += base.feature_size
            if cfg.model.freeze_base:
                self._freeze_base_model(m, freeze_mode=cfg.model.freeze_mode)

        if len(self.modality) > 1:
            att = cfg.model.attention
            if self.use_attention and not att.use_fixed:
                win = attn_win_size(cfg.data.audio.audio_length)
                if att.use_pe:
                    self.pe = nn.Sequential(_PEBuffer(10, win), nn.Conv1d(1034, 1024, kernel_size=1), nn.GroupNorm(64, 1024))
                if self.attention_type == "mha":
                    self.attention_layer = _MHAParams(1024, att.attn_heads, att.attn_dropout)
                elif self.attention_type == "unimodal":
                    self.attention_layer = _MLPAttnParams(1024, win)
                elif self.attention_type == "proto":
                    from .sampling import attention_prior

                    w1 = torch.from_numpy(attention_prior("gaussian", cfg.data.audio.audio_length)[:, 0].astype(np.float64))
                    shift = win // 2 - 2
                    protos = torch.stack((w1, torch.roll(w1, -shift), torch.roll(w1, shift))).float()
                    self.attention_layer = _MLPAttnParams(1024, 3, prototypes=protos)
            self.add_module("fusion", Fusion(in_features, 512, dropout=cfg.model.fusion_dropout))
            self.add_module("classifier", Classifier(self.num_classes, 512))
        else:
            self.add_module("classifier", Classifier(self.num_classes, in_features))

        self._plans = None
        self._train_plans = None
        self._train_bbs = None
        self._train_token, self._head_bwd, self._train_step = 0, None, None
        self._joint = None
        self._graphs: Dict[tuple, tuple] = {}
        self._static_graphs: Dict[tuple, tuple] = {}
        self.last_launches = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())
        self.to(self.device)

    # -- construction helpers ------------------------------------------------------------------
    def _maybe_load_pretrained(self, base: BNInceptionParams, modality: str) -> None:
        """Reference bn_inception.py:58-101: ImageNet (RGB; Audio with conv1 averaged over the RGB channels and its keys
        reconciled with the model's) or Kinetics (Flow, 10-channel conv1) weights from ``cfg.model.weights_dir``.
        The reference hard-wires ``<repo>/weights`` and cannot build a model without the files; here the directory is a
        config key and leaving it unset keeps the random initialisation (the files are not redistributable).  With the
        key set the behaviour is the reference's: a missing file raises (``torch.load``), RGB / Flow load strictly
        (the classification layer ``last_linear``, which the reference deletes after loading, is dropped first)."""
        wdir = self.cfg.model["weights_dir"] if "weights_dir" in self.cfg.model else None
        if not wdir:
            return
        fname = "kinetics_bninception_flow.pth" if modality == "Flow" else "imagenet_bninception_rgb.pth"
        path = os.path.join(wdir, fname)
        if not os.path.isfile(path):
            raise FileNotFoundError(f"pretrained weights {path} not found (reference bn_inception.py:65-70)")
        data = dict(torch.load(path, map_location="cpu"))
        data = {k: v for k, v in data.items() if not k.startswith("last_linear.")}
        own = base.state_dict()
        if modality == "Audio":
            data["conv1_7x7_s2.weight"] = data["conv1_7x7_s2.weight"].mean(dim=1).unsqueeze(dim=1)
            for k, v in own.items():      # missing keys keep the model's values,
                data.setdefault(k, v)
            data = {k: v for k, v in data.items() if k in own}  # extra keys are dropped (bn_inception.py:81-88)
        base.load_state_dict(data, strict=True)
        print(f"Model initialized with {'kinetics' if modality == 'Flow' else 'imagenet'} weights")

    def _freeze_base_model(self, modality: str, freeze_mode: str) -> None:
        """Reference model.py:148-176 (only affects ``requires_grad``)."""
        base = getattr(self, f"Base_{modality}")
        if freeze_mode == "all":
            for p in base.parameters():
                p.requires_grad = False
        elif freeze_mode == "partialbn":
            for mod_no, mod in enumerate(base.children()):
                if isinstance(mod, nn.BatchNorm2d):
                    if (modality == "Audio" and mod_no > 6) or mod_no > 1:
                        mod.weight.requires_grad = False
                        mod.bias.requires_grad = False

    # -- plan management -----------------------------------------------------------------------
    def _invalidate(self) -> None:
        self._plans = None
        self._train_plans = None
        self._train_bbs = None
        self._head_bwd, self._train_step = None, None
        self._graphs = {}
        self._static_graphs = {}

    def refresh_weights(self) -> None:
        """Call after mutating parameters in place (load_state_dict / .to() do it themselves)."""
        self._invalidate()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if "_plans" in self.__dict__:
            self._invalidate()
        return out

    def train(self, mode: bool = True):
        if "_plans" in self.__dict__:
            self._invalidate()
        return super().train(mode)

    def _ensure_plans(self, device):
        if self._plans is None:
            sd = {k: v for k, v in self.state_dict().items()}
            backs = {m: BackbonePlan(sd, f"Base_{m}.", _IN_CH[m], self.precision, device, self.chunk, self.chunk_lo,
                                     self.fuse_1x1, self.multi_stream) for m in self.modality}
            self._mod_streams = {m: torch.cuda.Stream(device=device) for m in self.modality} if self.multi_stream else {}
            self._mod_events = {m: torch.cuda.Event() for m in self.modality}
            self._in_event = torch.cuda.Event()
            head = HeadPlan(sd, self.cfg, self.modality, device, tensor_cores=(self.precision == "bf16"))
            self._joint = JointPhaseB(backs) if (self.group_launches is not False and self.precision == "bf16" and self.fuse_1x1
                                                 and not self.autotune) else None
            self._use_joint = self._joint is not None
            self._joint_event = torch.cuda.Event()
            self._plans = (backs, head)
        return self._plans

    # -- forward -------------------------------------------------------------------------------
    def _forward_impl(self, inputs: Dict[str, torch.Tensor], cur: "torch.cuda.Stream"):
        """Enqueue the whole forward: every modality's backbone on its own stream (they are independent
        until the attention / fusion), then the heads on ``cur``."""
        backs, head = self._plans
        E = FEATURE_SIZE
        b = n = None
        launches = 0
        S = 0
        n_audio = None
        ld = E * len(self.modality)
        f_first = inputs[self.modality[0]].shape[0] * inputs[self.modality[0]].shape[1]
        tune = self._tune_cache if self.autotune else None
        if self._joint is not None and self._use_joint:
            return self._forward_grouped(inputs, cur)
        if self.multi_stream:
            self._in_event.record(cur)
        for m_no, m in enumerate(self.modality):
            x = inputs[m]
            b, n, c, h, w = x.shape
            F = b * n
            xf = x.reshape(F, c, h, w)
            X = head.fusion_row(f_first)
            if m == "Audio":
                n_audio = n
            main = cur
            if self.multi_stream and not self.sequential_modalities:
                main = self._mod_streams[m]
                main.wait_event(self._in_event)
            if m == "Audio" and head.use_att:
                S = backs[m].out_spatial(h, w)[1]
                launches += backs[m].run(xf, head.audio_positions(F, S), 0, E, True, main, tune,
                                         split=head.split_target(m_no, F, True, S))
            else:
                launches += backs[m].run(xf, X, m_no * E, ld, False, main, tune, split=head.split_target(m_no, F, False))
            if m_no == 0 and F == f_first and len(self.modality) > 1:
                launches += head.run_qk(F, main.cuda_stream)  # q.k projection of the visual query: off the serial tail
            if self.multi_stream and not self.sequential_modalities:
                self._mod_events[m].record(main)
        if self.multi_stream and not self.sequential_modalities:
            for m in self.modality:
                cur.wait_event(self._mod_events[m])
        fixed = inputs.get("weights") if head.fixed else None
        n_first = inputs[self.modality[0]].shape[1]
        scores, wts, ent, l2 = head.run(b, n_first, S, fixed, cur.cuda_stream, n_audio=n_audio)
        self.last_launches = launches + l2
        return scores, wts, ent

    def _forward_grouped(self, inputs: Dict[str, torch.Tensor], cur: "torch.cuda.Stream"):
        """Grouped plan: phase A (stem .. pool2) of every modality on its own stream, then ONE joint phase B whose conv
        launches span all modalities (plan.JointPhaseB), then the heads.  Frames are processed in windows of ``chunk_lo``."""
        backs, head = self._plans
        E = FEATURE_SIZE
        ld = E * len(self.modality)
        fork = self.multi_stream and not self.sequential_modalities
        geo, S, n_audio, launches = {}, 0, None, 0
        f_first = inputs[self.modality[0]].shape[0] * inputs[self.modality[0]].shape[1]
        X = head.fusion_row(f_first)
        for m_no, m in enumerate(self.modality):
            b, n, c, h, w = inputs[m].shape
            F = b * n
            att_audio = m == "Audio" and head.use_att
            if m == "Audio":
                n_audio = n
            if att_audio:
                S = backs[m].out_spatial(h, w)[1]
                feat, col, fld, rows = head.audio_positions(F, S), 0, E, S
            else:
                feat, col, fld, rows = X, m_no * E, ld, 1
            geo[m] = (inputs[m].reshape(F, c, h, w), F, h, w, feat, col, fld, rows, att_audio,
                      head.split_target(m_no, F, att_audio, S if att_audio else 0))
        win = min(bk.chunk_lo for bk in backs.values())
        fmax = max(g[1] for g in geo.values())
        if fork:
            self._in_event.record(cur)
        for g0 in range(0, fmax, win):
            spec, ptrs, sptrs = [], {}, {}
            for m in self.modality:
                xf, F, h, w, feat, col, fld, rows, att_audio, split = geo[m]
                n_lo = min(win, F - g0)
                if n_lo <= 0:
                    continue
                main = cur
                if fork:
                    main = self._mod_streams[m]
                    main.wait_event(self._in_event if g0 == 0 else self._joint_event)
                launches += backs[m].run_a(xf, g0, n_lo, main)
                if fork:
                    self._mod_events[m].record(main)
                spec.append((m, n_lo, h, w, fld, int(att_audio), split[2] if split is not None else 0))
                ptrs[m] = feat.data_ptr() + (g0 * rows * fld + col) * 4
                if split is not None:
                    sptrs[m] = split[0].data_ptr() + (g0 * rows * split[2] + split[1]) * 2
            if fork:
                for m, *_ in spec:
                    cur.wait_event(self._mod_events[m])
            launches += self._joint.run(tuple(spec), ptrs, cur, sptrs)
            if fork and g0 + win < fmax:
                self._joint_event.record(cur)
        if fork and head.att_type == "mha" and head.tc:
            # q.k projection on a side stream while the positional-encoding GEMM runs on `cur`
            side = self._mod_streams[self.modality[0]]
            self._joint_event.record(cur)
            side.wait_event(self._joint_event)
            launches += head.run_qk(f_first, side.cuda_stream)
            self._mod_events[self.modality[0]].record(side)
            join = lambda: cur.wait_event(self._mod_events[self.modality[0]])  # noqa: E731  (right before the attention core)
        else:
            join = None
        b, n_first = inputs[self.modality[0]].shape[:2]
        fixed = inputs.get("weights") if head.fixed else None
        scores, wts, ent, l2 = head.run(b, n_first, S, fixed, cur.cuda_stream, n_audio=n_audio, before_attention=join)
        self.last_launches = launches + l2
        return scores, wts, ent

    # -- training forward ------------------------------------------------------------------------
    def _forward_train(self, tensors: Dict[str, torch.Tensor], dev):
        """Forward in ``model.train()`` mode (reference core/tools/train.py:68-76 -> model.py:205-262): every BatchNorm of
        the backbones normalises with BATCH statistics and updates its running statistics (partial-BN only freezes the
        gradients of the affine parameters, model.py:164-176), the audio feature is dropped at random (model.py:215-222,
        numpy's global RNG, same draw order as the reference), the attention probabilities and the fusion output pass
        through dropout and the MLP attentions sample a hard Gumbel-softmax (torch's RNG stream).  It runs on the fp32
        CUDA-core path by default (``train_precision="fp32"``: the reference trains in fp32) or, with
        ``train_precision="bf16"``, in mixed precision with the conv products on the tensor cores -- eagerly (no CUDA graph: the step changes module state).  ``loss.backward()`` reaches every
        trainable parameter: heads (head_train.py) and backbone convs / BatchNorm affines (backbone_train.py).

        Backbones that the caller put in eval() -- ``model.train(); model.Base_RGB.eval(); ...``, the usual way to
        fine-tune on frozen backbones with FIXED BatchNorm statistics -- run on the inference plans instead (folded
        BatchNorm, ``precision`` as configured: the bf16 tensor-core path), and only the heads run in training mode."""
        head_ver = tuple(p._version for p in self._head_params())
        frozen_eval = all(not getattr(self, f"Base_{m}").training for m in self.modality)
        if self._train_plans is None:
            sd = {k: v for k, v in self.state_dict().items()}  # live tensors: running statistics are updated in place
            self._train_plans = ({}, HeadPlan(sd, self.cfg, self.modality, dev, tensor_cores=False))
            self._train_sd, self._head_ver = sd, head_ver
        elif head_ver != self._head_ver:
            # an optimizer step changed the head parameters: redo the plan's folded copies (workspaces are kept)
            old = self._train_plans[1]
            new = HeadPlan(self._train_sd, self.cfg, self.modality, dev, tensor_cores=False)
            new._ws, new._retired = old._ws, old._retired
            self._train_plans = (self._train_plans[0], new)
            self._head_ver = head_ver
        backs, head = self._train_plans
        # trainable backbones: the tape-recording implementation (backbone_train.py), whose backward loss.backward() reaches
        bb_live = (not frozen_eval and torch.is_grad_enabled()
                   and any(p.requires_grad for n, p in self.named_parameters() if n.startswith("Base_")))
        # tb: the tape-recording implementation runs the forward -- always when it has to record for a backward pass, and for
        # frozen backbones too when train_precision is "bf16" (its convs then run on the tensor cores; the tape is dropped)
        tb = not frozen_eval and (bb_live or self.train_precision == "bf16")
        if frozen_eval:
            self._ensure_plans(dev)
            backs = self._plans[0]
        elif tb:
            if self._train_bbs is None:
                # train_precision "bf16": the convs of the training step multiply on the tensor cores (mixed precision)
                self._train_bbs = {m: TrainBackbone(self._train_sd, f"Base_{m}.", _IN_CH[m], dev,
                                                    tensor_cores=(self.train_precision == "bf16")) for m in self.modality}
            backs = self._train_bbs
        elif not backs:
            backs.update({m: BackbonePlan(self._train_sd, f"Base_{m}.", _IN_CH[m], "fp32", dev, self.chunk, self.chunk_lo, False, False,
                                          train=True) for m in self.modality})
        att = self.cfg.model.attention
        E = FEATURE_SIZE
        ld = E * len(self.modality)
        cur = torch.cuda.current_stream(dev)
        f_first = tensors[self.modality[0]].shape[0] * tensors[self.modality[0]].shape[1]
        X = head.fusion_row(f_first)
        S, n_audio, launches, dropped = 0, None, 0, False
        for m_no, m in enumerate(self.modality):
            x = tensors[m]
            b, n, c, h, w = x.shape
            F = b * n
            if not frozen_eval:
                if not tb:
                    backs[m].refresh_weights()
                base = getattr(self, f"Base_{m}")
                for mod in base.modules():
                    if isinstance(mod, nn.BatchNorm2d) and mod.num_batches_tracked is not None:
                        mod.num_batches_tracked += 1
            xf = x.reshape(F, c, h, w)
            if m == "Audio":
                n_audio = n
            if m == "Audio" and head.use_att:
                S = backs[m].out_spatial(h, w)[1]
                if tb:
                    launches += backs[m].forward(xf, head.audio_positions(F, S), 0, E, True, cur.cuda_stream)
                else:
                    launches += backs[m].run(xf, head.audio_positions(F, S), 0, E, True, cur)
            elif tb:
                launches += backs[m].forward(xf, X, m_no * E, ld, False, cur.cuda_stream)
            else:
                launches += backs[m].run(xf, X, m_no * E, ld, False, cur)
            if tb and not bb_live:
                backs[m].backward_skip()  # frozen backbone: nothing will ask for its gradients, free the tape now
            if (m == "Audio" and len(self.modality) > 1 and self.cfg.data.audio.dropout > 0
                    and np.random.uniform() > self.cfg.data.audio.dropout):
                dropped = True  # (the backbone ran first, as in the reference: its BatchNorm statistics still move)
        opts = {"audio_dropped": dropped, "fusion_dropout": float(self.cfg.model.fusion_dropout) if len(self.modality) > 1 else 0.0}
        learned = self.use_attention and not att.use_fixed and "Audio" in self.modality and len(self.modality) > 1
        if dropped and learned:
            # reference quirk: with the audio feature dropped no attention runs, and model.py:259-260 then reads att_wts
            raise UnboundLocalError("local variable 'att_wts' referenced before assignment")
        if learned and self.attention_type == "mha" and att.attn_dropout > 0:
            H = int(att.attn_heads)
            # the mask torch's multi_head_attention_forward draws: dropout on the (F*H, 1, S) attention probabilities
            opts["attn_mask"] = torch.nn.functional.dropout(torch.ones(f_first * H, 1, S, device=dev), float(att.attn_dropout),
                                                            True).view(f_first, H, S).contiguous()
        if learned and self.attention_type in ("unimodal", "proto") and att.use_gumbel:
            opts["gumbel"] = True
        fixed = tensors.get("weights") if head.fixed else None
        opts["fixed_weights"] = fixed
        b, n_first = tensors[self.modality[0]].shape[:2]
        scores, wts, ent, l2 = head.run(b, n_first, S, fixed, cur.cuda_stream, n_audio=n_audio, train=opts)
        self.last_launches = launches + l2
        self._train_step = (b, n_first, S, opts, n_audio, bb_live)
        return scores.clone(), (wts.clone() if wts is not None else None), (ent.clone() if ent is not None else None)

    def _head_params(self) -> list:
        return [self.get_parameter(n) for n in head_parameter_names(self.cfg, self.modality)]

    def _grad_param_names(self, bb_live: bool) -> list:
        names = list(head_parameter_names(self.cfg, self.modality))
        if bb_live:
            for m in self.modality:
                names += self._train_bbs[m].parameter_names()
        return names

    def _grad_params(self, bb_live: bool) -> list:
        return [self.get_parameter(n) for n in self._grad_param_names(bb_live)]

    def _head_backward(self, token: int, d_scores: torch.Tensor, d_wts) -> list:
        """Gradients of ``_head_params()`` for the LAST training forward (its intermediates live in the head plan's
        workspaces, so a later forward invalidates the step).  The backbones must be frozen: their backward pass is not
        implemented, and returning head gradients alone for a model whose backbones expect gradients would silently train
        something else than the reference does."""
        if token != self._train_token:
            raise RuntimeError("backward of a stale training forward: run forward -> loss.backward() before the next forward "
                               "(the head workspaces hold one step)")
        b, n, S, opts, n_audio, bb_live = self._train_step
        live = [n_ for n_, p in self.named_parameters() if n_.startswith("Base_") and p.requires_grad]
        if live and not bb_live:
            raise NotImplementedError(
                f"{len(live)} backbone parameters require gradients (first: {live[0]}) but the backbones ran on the inference "
                "plans (their modules are in eval()), which keep nothing for a backward pass.  Either freeze them -- "
                "model.freeze_base with freeze_mode 'all' (reference model.py:160-163) or p.requires_grad_(False) -- or leave "
                "them in train() mode.")
        if n_audio is not None and n_audio != n:
            raise NotImplementedError("backward of the 10-crop forward (evaluation-only path, reference model.py:243-248)")
        head = self._train_plans[1]
        dev = d_scores.device
        E = FEATURE_SIZE
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if self._head_bwd is None or self._head_bwd.head is not head:
                self._head_bwd = HeadBackward(head, self._train_sd)
            hb = self._head_bwd
            grads = hb.run(b, n, S, d_scores, d_wts, opts, stream, input_grads=bb_live)
            launches = hb.launches
            if bb_live:
                ld = E * len(self.modality)
                for m_no, m in enumerate(self.modality):
                    bb = self._train_bbs[m]
                    if m == "Audio" and len(self.modality) > 1 and opts.get("audio_dropped"):
                        bb.backward_skip()  # the feature was replaced by zeros (model.py:215-222): no gradient reaches it
                        continue
                    if m == "Audio" and head.use_att:
                        grads.update(bb.backward(hb.d_a.data_ptr(), E, stream))
                    else:
                        grads.update(bb.backward(hb.d_x.data_ptr() + m_no * E * 4, ld, stream))
                    launches += bb.launches
        self.last_backward_launches = launches
        return [grads.get(nm) for nm in self._grad_param_names(bb_live)]

    def _check_inputs(self, input) -> Dict[str, torch.Tensor]:
        tensors = {}
        for m in self.modality:
            x = input[m]
            if not x.is_cuda:
                raise RuntimeError(f"input['{m}'] is on {x.device}: the TBN kernels run on CUDA only (no CPU fallback)")
            if x.dim() != 5 or x.shape[2] != _IN_CH[m]:
                raise ValueError(f"input['{m}'] must be (B, N, {_IN_CH[m]}, H, W), got {tuple(x.shape)}")
            # bf16 inputs pass through untouched when the fused bf16 stem consumes them (it rounds fp32 to bf16 on load
            # anyway: identical results, half the host-link bytes); everything else is the reference's fp32 contract
            keep_bf16 = x.dtype == torch.bfloat16 and self.precision == "bf16"
            tensors[m] = x.contiguous() if keep_bf16 else x.float().contiguous()
        b, n = tensors[self.modality[0]].shape[:2]
        att = self.cfg.model.attention
        for m in self.modality[1:]:
            bm, nm = tensors[m].shape[:2]
            if (bm, nm) == (b, n):
                continue
            # 10-crop path (reference model.py:243-248): only the AUDIO feature is tiled up to the first modality's rows;
            # any other mismatch fails in the reference's torch.cat (model.py:250), a learned attention in its batch check
            if m != "Audio" or bm != b or nm == 0 or n % nm != 0 or nm > n:
                raise RuntimeError(f"Sizes of tensors must match except in dimension 1: input['{m}'] has (B, N) = "
                                   f"{(bm, nm)}, input['{self.modality[0]}'] {(b, n)} (only audio rows are repeated)")
            if self.use_attention and not att.use_fixed:
                raise RuntimeError("learned attention needs the same number of visual and audio segments "
                                   f"(query rows {b * n}, key rows {bm * nm}): the reference fails here as well")
        if self.use_attention and att.use_fixed and "Audio" in self.modality and len(self.modality) > 1:
            fa = tensors["Audio"].shape[0] * tensors["Audio"].shape[1]
            tensors["weights"] = input["weights"].to(tensors[self.modality[0]].device).float().reshape(fa, -1).contiguous()
        return tensors

    def forward(self, input):
        tensors = self._check_inputs(input)
        dev = tensors[self.modality[0]].device
        _lib.load()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            if self.training:
                scores, wts, ent = self._forward_train({k: (v.float() if v.dtype != torch.float32 else v) for k, v in tensors.items()},
                                                       dev)
                params = self._grad_params(self._train_step[5])
                if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                    # autograd node of the head backward (head_train.py): loss.backward() reaches the head parameters
                    self._train_token += 1
                    scores, w2 = HeadGradFn.apply(self, self._train_token, scores,
                                                  wts if wts is not None else scores.new_empty(0), *params)
                    if wts is not None:
                        wts = w2
                    ent = None  # get_loss derives the entropy term from the weights so that its gradient flows
                return self._pack_outputs(scores, wts, ent, self._train_plans[1])
            self._ensure_plans(dev)
            if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
                scores, wts, ent = self._forward_graph(tensors, dev)
                # the graph writes into fixed result buffers: hand out copies (a few KB) so that results a caller keeps
                # -- test() collects the scores of every batch -- are not overwritten by the next replay
                scores = scores.clone()
                wts = wts.clone() if wts is not None else None
                ent = ent.clone() if ent is not None else None
            else:
                scores, wts, ent = self._forward_impl(tensors, stream)
        return self._pack_outputs(scores, wts, ent, self._plans[1])

    def _pack_outputs(self, scores, wts, ent, head):
        out = OrderedDict()
        col = 0
        for key, size in zip(head.class_keys, head.class_sizes):
            out[key] = scores[:, col:col + size]
            col += size
        if self.use_attention and not self.cfg.model.attention.use_fixed and wts is not None:
            if self.attention_type == "mha":
                out["weights"] = wts.view(wts.shape[0], 1, wts.shape[1])
            else:
                out["weights"] = wts
        self.last_entropy_rows = ent
        return out

    _MAX_POINTER_GRAPHS = 4

    def _capture(self, tensors: Dict[str, torch.Tensor], dev):
        if self._joint is None or self.group_launches is not None:
            return self._capture_one(tensors, dev)
        # auto mode: capture both plans, keep the faster replay (decided once per shape signature)
        skey = tuple((k, tuple(v.shape), v.dtype) for k, v in tensors.items())
        choice = self.plan_choice.get(skey)
        if choice is not None:
            self._use_joint = choice == "grouped"
            try:
                return self._capture_one(tensors, dev)
            finally:
                self._use_joint = True
        entries, times = {}, {}
        for name, joint in (("grouped", True), ("per-modality", False)):
            self._use_joint = joint
            entries[name] = self._capture_one(tensors, dev)
            g = entries[name][0]
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                g.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            times[name] = e0.elapsed_time(e1) / 3
        self._use_joint = True
        best = min(times, key=times.get)
        self.plan_choice[skey] = best
        self.plan_times = times
        return entries[best]

    def _capture_one(self, tensors: Dict[str, torch.Tensor], dev):
        # warm-up run on a side stream (allocates workspaces, builds op lists), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._forward_impl(tensors, side)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            res = self._forward_impl(tensors, torch.cuda.current_stream(dev))
        return (g, res, tensors, self.last_launches)

    def _forward_graph(self, tensors: Dict[str, torch.Tensor], dev):
        """One captured graph per (shapes, input addresses) for callers that reuse their input buffers (no copy at all);
        a caller that keeps bringing new tensors gets, after ``_MAX_POINTER_GRAPHS`` captures, one graph per shape
        signature on model-owned input buffers plus a device-to-device copy per call (never a re-capture per batch)."""
        key = tuple((k, tuple(v.shape), v.data_ptr()) for k, v in tensors.items())
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) < self._MAX_POINTER_GRAPHS:
                entry = self._graphs[key] = self._capture(tensors, dev)
            else:
                skey = tuple((k, tuple(v.shape), v.dtype) for k, v in tensors.items())
                entry = self._static_graphs.get(skey)
                if entry is None:
                    own = {k: torch.empty_like(v) for k, v in tensors.items()}
                    for k, v in tensors.items():
                        own[k].copy_(v)
                    entry = self._static_graphs[skey] = self._capture(own, dev)
                else:
                    for k, v in tensors.items():
                        entry[2][k].copy_(v)
        g, res, _keepalive, launches = entry
        g.replay()
        self.last_launches = launches
        return res

    # -- losses (host glue, same arithmetic as the reference) -------------------------------------
    def get_loss(self, criterion, target, preds, epoch=0):
        """Reference model.py:264-334.  Training-only multipliers are kept; the entropy term equals
        ``Categorical(probs=w+1e-6).entropy().mean()`` and is also produced per row by the fused
        attention kernel (``last_entropy_rows``)."""
        assert isinstance(target, dict) and isinstance(preds, dict) and isinstance(criterion, dict)
        loss = {"total": 0, "all_class": 0}
        batch_size = 0
        for key in target["class"].keys():
            labels = target["class"][key]
            batch_size = labels.shape[0]
            loss[key] = criterion["crossentropy"](preds[key], labels)
            loss["all_class"] += loss[key]
        loss["total"] += loss["all_class"]
        att = self.cfg.model.attention
        if self.use_attention and not att.use_fixed:
            if self.training and epoch + 1 < att.decay_step:
                prior_mult = contrast_mult = entropy_mult = 0
            else:
                prior_mult, contrast_mult, entropy_mult = att.wt_decay, att.contrast_decay, att.entropy_decay
            wts = preds["weights"].squeeze(1)
            if att.use_prior:
                b, n, _, _ = target["weights"].shape
                assert wts.shape[0] == b * n
                prior = target["weights"].reshape(b * n, -1)
                if att.wt_loss == "kl":
                    wts = torch.log(wts + 1e-7)
                loss["prior"] = criterion["prior"](wts, prior)
                loss["total"] += prior_mult * loss["prior"]
            if att.use_contrast:
                loss["contrast"] = criterion["contrast"](wts)
                loss["total"] += contrast_mult * loss["contrast"]
            if att.use_entropy:
                rows = getattr(self, "last_entropy_rows", None)
                if rows is not None and rows.shape[0] == wts.shape[0] and not att.use_prior:
                    loss["entropy"] = rows.mean()
                else:
                    p = wts + 1e-6
                    p = p / p.sum(-1, keepdim=True)
                    loss["entropy"] = (-(p * p.log()).sum(-1)).mean()
                if self.training and entropy_mult > 0 and loss["entropy"] < att.entropy_thresh:
                    entropy_mult = 0
                loss["total"] += entropy_mult * loss["entropy"]
        return loss, batch_size
