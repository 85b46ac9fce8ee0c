"""CPU oracle for the TBN forward hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product (``attention_based_tbn_b200``) never does.

It is a functional (state-dict driven) fp32 restatement of the reference's forward:

  * backbone      : third-party ``pretrainedmodels`` BNInception (UNPINNED in the reference's
                    install/requirements.txt:14; latest upstream 0.7.4), absent from
                    /root/reference.  Restated from the only in-tree statement of it,
                    reference core/models/bn_inception_audio.py:21-39 (stem, incl. the Cadene
                    7x7 stem kept there as comments), 58-401 (layer table), 437-1003 (dataflow).
  * feature pool  : reference core/models/bn_inception.py:16-35.
  * PE / MHA / fixed / unimodal / proto attention : reference core/models/attention.py:8-145,
                    torch ``F.multi_head_attention_forward`` (need_weights path).
  * fusion / classifier / consensus / forward switch : reference core/models/model.py:178-262,
                    337-386.
  * entropy term  : reference core/models/model.py:310, 323-324.

Pinning: the reference ships no tests or golden vectors for this path ("parity unpinned" by
the reference's own fixtures).  The oracle is pinned instead against the reference ITSELF run
in the build container: ``oracle/make_golden.py`` imports the unmodified reference
``TBNModel`` (with a ``pretrainedmodels`` shim text-spliced from the reference's own layer
table) and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those vectors.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

# (name, 1x1, 3x3_reduce, 3x3, dbl_reduce, dbl_1, dbl_2, pool, pool_proj) -- ref :58-401
_BLOCKS = [
    ("3a", 64, 64, 64, 64, 96, 96, "avg", 32),
    ("3b", 64, 64, 96, 64, 96, 96, "avg", 64),
    ("3c", 0, 128, 160, 64, 96, 96, "max2", 0),
    ("4a", 224, 64, 96, 96, 128, 128, "avg", 128),
    ("4b", 192, 96, 128, 96, 128, 128, "avg", 128),
    ("4c", 160, 128, 160, 128, 160, 160, "avg", 128),
    ("4d", 96, 128, 192, 160, 192, 192, "avg", 128),
    ("4e", 0, 128, 192, 192, 256, 256, "max2", 0),
    ("5a", 352, 192, 320, 160, 224, 224, "avg", 128),
    ("5b", 352, 192, 320, 192, 224, 224, "max1", 128),
]


_CALIBRATE = False  # set by oracle/calibrate.py only: BN running stats <- batch stats (momentum 1)
_TRAIN = False      # set by tbn_forward(train=True): BatchNorm in training mode (batch statistics, running stats updated
                    # IN PLACE in the state dict, momentum 0.1 -- ref core/tools/train.py:68-76 calls model.train())


def _cbr(sd, p, name, x, stride=1, pad=0):
    """conv(+bias) -> BatchNorm2d(eval, eps=1e-5) -> ReLU  (ref bn_inception_audio.py:24-31)."""
    y = F.conv2d(x, sd[p + name + ".weight"], sd[p + name + ".bias"], stride=stride, padding=pad)
    if _CALIBRATE:
        sd[p + name + "_bn.running_mean"] = y.mean(dim=(0, 2, 3))
        sd[p + name + "_bn.running_var"] = y.var(dim=(0, 2, 3), unbiased=True)
    y = F.batch_norm(
        y,
        sd[p + name + "_bn.running_mean"],
        sd[p + name + "_bn.running_var"],
        sd[p + name + "_bn.weight"],
        sd[p + name + "_bn.bias"],
        _TRAIN,
        0.1 if _TRAIN else 0.0,
        1e-5,
    )
    if _TRAIN and (p + name + "_bn.num_batches_tracked") in sd:
        sd[p + name + "_bn.num_batches_tracked"] += 1
    return F.relu(y)


def bninception_features(sd, prefix: str, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """Cadene BNInception.features(): (F,Cin,H,W) -> (F,1024,H/32,W/32)."""
    p = prefix
    x = _cbr(sd, p, "conv1_7x7_s2", x, 2, 3)
    if taps is not None:
        taps["conv1"] = x
    x = F.max_pool2d(x, 3, 2, ceil_mode=True)
    x = _cbr(sd, p, "conv2_3x3_reduce", x)
    x = _cbr(sd, p, "conv2_3x3", x, 1, 1)
    if taps is not None:
        taps["conv2"] = x
    x = F.max_pool2d(x, 3, 2, ceil_mode=True)
    if taps is not None:
        taps["pool2"] = x
    for (nm, n1, _, _, _, _, _, pool, npp) in _BLOCKS:
        b = f"inception_{nm}_"
        s = 2 if pool == "max2" else 1
        outs = []
        if n1:
            outs.append(_cbr(sd, p, b + "1x1", x))
        y = _cbr(sd, p, b + "3x3_reduce", x)
        outs.append(_cbr(sd, p, b + "3x3", y, s, 1))
        y = _cbr(sd, p, b + "double_3x3_reduce", x)
        y = _cbr(sd, p, b + "double_3x3_1", y, 1, 1)
        outs.append(_cbr(sd, p, b + "double_3x3_2", y, s, 1))
        if pool == "avg":
            y = F.avg_pool2d(x, 3, 1, 1, ceil_mode=True, count_include_pad=True)
        elif pool == "max1":
            y = F.max_pool2d(x, 3, 1, 1, ceil_mode=True)
        else:
            y = F.max_pool2d(x, 3, 2, ceil_mode=True)
        outs.append(_cbr(sd, p, b + "pool_proj", y) if npp else y)
        x = torch.cat(outs, 1)
        if taps is not None:
            taps[nm] = x
    return x


def backbone_forward(sd, prefix, x, is_audio: bool, attend: bool, taps=None) -> torch.Tensor:
    """features() + the overridden logits() (ref bn_inception.py:16-35)."""
    f = bninception_features(sd, prefix, x, taps)
    h, w = f.shape[2:]
    if is_audio and attend:
        return F.avg_pool2d(f, kernel_size=(h, 1), stride=(h, 1))  # (F,1024,1,S)
    return F.avg_pool2d(f, kernel_size=(h, w)).view(f.size(0), -1)  # (F,1024)


def pe_conv_groupnorm(sd, a: torch.Tensor) -> torch.Tensor:
    """PositionalEncoding(concat) -> Conv1d(1034,1024,1) -> GroupNorm(64,1024) (ref model.py:62-67)."""
    a = a.squeeze(2)  # (F,1024,S)
    pe = sd["pe.0.pe"]  # (1,10,S)
    z = torch.cat((a, pe.expand(a.shape[0], pe.shape[1], pe.shape[2])), dim=1)
    z = F.conv1d(z, sd["pe.1.weight"], sd["pe.1.bias"])
    return F.group_norm(z, 64, sd["pe.2.weight"], sd["pe.2.bias"], 1e-5)


def mha_single_query(sd, heads: int, q_in: torch.Tensor, z: torch.Tensor, dropout: float = 0.0, mask: Optional[torch.Tensor] = None):
    """nn.MultiheadAttention(1024, heads) in eval, L=1 query, S keys (ref attention.py:48-57).

    q_in (F,1024) visual query, z (F,1024,S) keys=values.  Returns (out (F,1024), wts (F,1,S))."""
    a = "attention_layer.attention_layer."
    W, bias = sd[a + "in_proj_weight"], sd[a + "in_proj_bias"]
    E = W.shape[1]
    d = E // heads
    Fn, _, S = z.shape
    zt = z.transpose(1, 2)  # (F,S,E)
    q = F.linear(q_in, W[:E], bias[:E]).view(Fn, heads, 1, d)
    k = F.linear(zt, W[E : 2 * E], bias[E : 2 * E]).view(Fn, S, heads, d).transpose(1, 2)  # (F,H,S,d)
    v = F.linear(zt, W[2 * E :], bias[2 * E :]).view(Fn, S, heads, d).transpose(1, 2)
    scores = torch.matmul(q * (1.0 / math.sqrt(d)), k.transpose(-1, -2))  # (F,H,1,S)
    p = torch.softmax(scores, dim=-1)
    if dropout > 0:  # training: torch's multi_head_attention_forward drops the (F*H, 1, S) probabilities
        p = F.dropout(p.reshape(Fn * heads, 1, S), dropout, True).reshape(Fn, heads, 1, S)
    if mask is not None:  # an explicit dropout mask (Fn, heads, S) holding 0 or 1 / (1 - p): see head_train_forward
        p = p * mask.view(Fn, heads, 1, S)
    o = torch.matmul(p, v).transpose(1, 2).reshape(Fn, E)
    out = F.linear(o, sd[a + "out_proj.weight"], sd[a + "out_proj.bias"])
    return out, p.mean(dim=1)  # head-averaged weights (F,1,S)


def gaussian_kernel(n: int, sigma: float = 1.0) -> torch.Tensor:
    """cv2.getGaussianKernel(n, sigma) for sigma>0: exp(-(i-(n-1)/2)^2/(2 sigma^2)), normalised."""
    xs = torch.arange(n, dtype=torch.float64) - (n - 1) / 2.0
    k = torch.exp(-(xs ** 2) / (2.0 * sigma * sigma))
    return k / k.sum()


def proto_weights(win: int) -> torch.Tensor:
    """PrototypeAttention._create_prototype_weights (ref attention.py:122-133): (3, win) fp32."""
    w1 = gaussian_kernel(win, 1.0)
    shift = win // 2 - 2
    return torch.stack((w1, torch.roll(w1, -shift), torch.roll(w1, shift))).float()


def entropy_term(w: torch.Tensor) -> torch.Tensor:
    """Categorical(probs=w+1e-6).entropy().mean() on (F,1,S) weights (ref model.py:310, 323-324)."""
    p = w.squeeze(1) + 1e-6
    p = p / p.sum(-1, keepdim=True)
    return (-(p * p.log()).sum(-1)).mean()


def tbn_forward(sd, cfg, modality: List[str], inputs: Dict[str, torch.Tensor], taps: Optional[dict] = None, train: bool = False):
    """TBNModel.forward (ref model.py:205-262) -> OrderedDict of (B,n_cls) [+ weights].  ``train``: the model.train()
    forward -- BatchNorm with batch statistics (the running statistics in ``sd`` are updated in place), audio feature
    dropout from numpy's global RNG (model.py:215-222), attention / fusion dropout and Gumbel-softmax from torch's."""
    global _TRAIN
    import numpy as np

    _TRAIN = bool(train)
    try:
        return _tbn_forward(sd, cfg, modality, inputs, taps, bool(train), np)
    finally:
        _TRAIN = False


def _tbn_forward(sd, cfg, modality, inputs, taps, train, np):
    att = cfg.model.attention
    use_att = bool(att.enable)
    feats = []
    att_wts = None
    b = n = None
    with torch.no_grad():
        for m in modality:
            b, n, c, h, w = inputs[m].shape
            t = None
            if taps is not None:
                t = taps.setdefault(m, {})
            f = backbone_forward(sd, f"Base_{m}.", inputs[m].reshape(b * n, c, h, w).float(), m == "Audio", use_att, t)
            if t is not None:
                t["backbone"] = f  # the backbone output: (F,1024), or (F,1024,1,S) positions for an attended Audio
            if (m == "Audio" and train and len(modality) > 1 and cfg.data.audio.dropout > 0
                    and np.random.uniform() > cfg.data.audio.dropout):
                f = torch.zeros_like(feats[0])  # drop the audio feature (ref model.py:215-222)
            elif m == "Audio" and use_att:
                if att.use_fixed:
                    f = (f.squeeze(2) * inputs["weights"].reshape(b * n, -1).unsqueeze(1)).sum(2)
                elif att.type == "mha":
                    z = pe_conv_groupnorm(sd, f)
                    if t is not None:
                        t["pe"] = z
                    f, att_wts = mha_single_query(sd, att.attn_heads, feats[0], z, float(att.attn_dropout) if train else 0.0)
                elif att.type in ("unimodal", "proto"):
                    hdn = F.relu(F.linear(feats[0], sd["attention_layer.seq.0.weight"], sd["attention_layer.seq.0.bias"]))
                    lg = F.linear(hdn, sd["attention_layer.seq.2.weight"], sd["attention_layer.seq.2.bias"])
                    mul = F.gumbel_softmax(lg, tau=1, hard=True) if (train and att.use_gumbel) else torch.softmax(lg, dim=1)
                    att_wts = mul if att.type == "unimodal" else mul @ proto_weights(f.shape[-1])
                    f = (f.squeeze(2) * att_wts.unsqueeze(1)).sum(2)
            if m == "Audio" and feats and feats[0].shape[0] > f.shape[0]:
                # 10-crop path (ref model.py:243-248, inside the Audio branch): the visual modalities carry k crops per
                # segment, audio does not -- tile the audio rows k times and count k*n segments in the consensus
                k = feats[0].shape[0] // f.shape[0]
                f = f.repeat(k, 1)
                n *= k
            if t is not None:
                t["feature"] = f
            feats.append(f)
        x = torch.cat(feats, dim=1)
        if len(modality) > 1:
            x = F.relu(F.linear(x, sd["fusion.fusion_layer.0.weight"], sd["fusion.fusion_layer.0.bias"]))
            if train and cfg.model.fusion_dropout > 0:
                x = F.dropout(x, float(cfg.model.fusion_dropout), True)
        out = OrderedDict()
        for key in cfg.model.num_classes.keys():
            y = F.linear(x, sd[f"classifier.{key}.weight"], sd[f"classifier.{key}.bias"])
            out[key] = y.view(b, n, -1).mean(dim=1)  # segment consensus (ref model.py:194-199)
        if use_att and not att.use_fixed and att_wts is not None:
            out["weights"] = att_wts
    return out


def head_train_forward(sd, cfg, modality: List[str], feats: Dict[str, torch.Tensor], b: int, n: int,
                       attn_mask: Optional[torch.Tensor] = None, fusion_mask: Optional[torch.Tensor] = None,
                       gumbel: Optional[torch.Tensor] = None):
    """The layers above the backbones (ref model.py:214-262) in training mode WITH autograd, on given backbone features:
    ``feats[m]`` = (F, 1024), or (F, 1024, 1, S) for an attended Audio.  The random draws are inputs, so that a CUDA run
    can be replayed exactly: ``attn_mask`` (F, H, S) and ``fusion_mask`` (F, 512) are inverted-dropout masks (0 or
    1 / (1 - p)), ``gumbel`` (F, n_out) the Gumbel noise of the hard Gumbel-softmax (tau = 1).  Gradients flow into every
    tensor of ``sd`` that requires them -- the checker of the head backward kernels (test infrastructure only)."""
    att = cfg.model.attention
    use_att = bool(att.enable)
    out_feats = []
    att_wts = None
    for m in modality:
        f = feats[m]
        if m == "Audio" and use_att and len(modality) > 1:
            if att.use_fixed:
                raise ValueError("fixed attention weights have no trainable layer: pass the weighted feature")
            if att.type == "mha":
                z = pe_conv_groupnorm(sd, f)
                f, att_wts = mha_single_query(sd, att.attn_heads, out_feats[0], z, 0.0, attn_mask)
            else:
                hdn = F.relu(F.linear(out_feats[0], sd["attention_layer.seq.0.weight"], sd["attention_layer.seq.0.bias"]))
                lg = F.linear(hdn, sd["attention_layer.seq.2.weight"], sd["attention_layer.seq.2.bias"])
                if gumbel is not None:  # torch.nn.functional.gumbel_softmax(hard=True) with the noise given
                    y_soft = (lg + gumbel).softmax(1)
                    y_hard = torch.zeros_like(lg).scatter_(1, y_soft.max(1, keepdim=True)[1], 1.0)
                    mul = y_hard - y_soft.detach() + y_soft
                else:
                    mul = torch.softmax(lg, dim=1)
                att_wts = mul if att.type == "unimodal" else mul @ sd["attention_layer.prototype_wts"]
                f = (f.squeeze(2) * att_wts.unsqueeze(1)).sum(2)
        out_feats.append(f)
    x = torch.cat(out_feats, dim=1)
    if len(modality) > 1:
        x = F.relu(F.linear(x, sd["fusion.fusion_layer.0.weight"], sd["fusion.fusion_layer.0.bias"]))
        if fusion_mask is not None:
            x = x * fusion_mask
    out = OrderedDict()
    for key in cfg.model.num_classes.keys():
        y = F.linear(x, sd[f"classifier.{key}.weight"], sd[f"classifier.{key}.bias"])
        out[key] = y.view(b, n, -1).mean(dim=1)
    if att_wts is not None:
        out["weights"] = att_wts
    return out
