"""Backward pass of the TBN heads: what ``loss.backward()`` does in the reference's training step
(core/tools/train.py:76-81) for the layers of model.py:214-262 -- attention / positional encoding / Fusion / classifiers.
With frozen backbones (``freeze_mode: "all"``, model.py:160-163) this is the whole backward; with trainable ones it also
produces the gradients of the backbone outputs that ``backbone_train.TrainBackbone.backward`` starts from.

The training forward (``TBNModel._forward_train``) leaves every intermediate the backward needs in the head plan's
workspaces (fusion row, attended audio positions, GroupNorm output, query, head outputs, Fusion output); what it
overwrote in place (the pre-GroupNorm conv output) or folded away (keys / values) is recomputed here with the same
kernels.  All arithmetic is ``csrc/head_bwd.cu`` (+ ``tbn_linear`` for the recomputation); torch only owns the memory.
The gradients come back keyed by state-dict names and are wired into autograd by ``HeadGradFn`` so that the reference's
loop -- ``loss.backward(); clip_grad_norm_; optimizer.step()`` -- runs unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib
from .plan import FEATURE_SIZE, HeadPlan

_A = "attention_layer.attention_layer."
_M = "attention_layer.seq."


def head_parameter_names(cfg, modality: List[str]) -> List[str]:
    """State-dict names of the head parameters, in the order ``HeadGradFn`` passes them."""
    names = []
    att = cfg.model.attention
    learned = bool(att.enable) and not att.use_fixed and "Audio" in modality and len(modality) > 1
    if learned and att.type == "mha":
        names += ["pe.1.weight", "pe.1.bias", "pe.2.weight", "pe.2.bias", _A + "in_proj_weight", _A + "in_proj_bias",
                  _A + "out_proj.weight", _A + "out_proj.bias"]
    elif learned and att.type in ("unimodal", "proto"):
        names += [_M + "0.weight", _M + "0.bias", _M + "2.weight", _M + "2.bias"]
    if len(modality) > 1:
        names += ["fusion.fusion_layer.0.weight", "fusion.fusion_layer.0.bias"]
    for key in cfg.model.num_classes.keys():
        names += [f"classifier.{key}.weight", f"classifier.{key}.bias"]
    return names


class HeadBackward:
    """Gradients of the head parameters from d(scores) (B, n_cls) and d(weights) (F, S)."""

    def __init__(self, head: HeadPlan, sd: Dict[str, torch.Tensor]):
        assert not head.tc, "the training step runs on the fp32 heads"
        self.head, self.sd = head, sd
        self._pe_rows: Optional[torch.Tensor] = None

    def _buf(self, name: str, *shape) -> torch.Tensor:
        return self.head.buf("bwd_" + name, *shape)

    def run(self, B: int, N: int, S: int, d_scores: torch.Tensor, d_wts: Optional[torch.Tensor], opts: dict,
            stream: int, input_grads: bool = False) -> Dict[str, torch.Tensor]:
        """``input_grads``: also leave the gradients of the backbone outputs in ``self.d_x`` ((F, ld) fusion row; the
        audio block is meaningless when the audio feature is attended) and ``self.d_a`` ((F*S, E) attended audio
        positions, or None) -- what the backbone backward starts from."""
        h = self.head
        self.d_x = self.d_a = None
        c = _lib.call
        F, E = B * N, FEATURE_SIZE
        ld = E * h.n_mod
        ncls = h.n_cls
        X = h.fusion_row(F)
        grads: Dict[str, torch.Tensor] = {}
        self.launches = 0

        def gemm(ta, tb, m, n, k, a, lda, b, ldb, out, ldo, beta=0.0):
            c("tbn_sgemm", ta, tb, m, n, k, 1.0, int(a), lda, int(b), ldb, beta, int(out), ldo, stream)
            self.launches += 1

        def colsum(x, ldx, m, n, out):
            c("tbn_colsum", int(x), ldx, m, n, 0.0, int(out), stream)
            self.launches += 1

        def new(*shape):
            return torch.empty(*shape, device=h.device, dtype=torch.float32)

        # ---- consensus + classifiers (model.py:194-199, 365-386) --------------------------------------------------
        d_scores = d_scores.float().contiguous()
        dl = self._buf("dlogits", F, ncls)
        c("tbn_consensus_bwd", d_scores.data_ptr(), ncls, dl.data_ptr(), B, N, ncls, stream)
        feat, k_in = (h.buf("hfus", F, 512), 512) if h.n_mod > 1 else (X, ld)
        g_wc, g_bc = new(ncls, k_in), new(ncls)
        gemm(1, 0, ncls, k_in, F, dl.data_ptr(), ncls, feat.data_ptr(), k_in, g_wc.data_ptr(), k_in)
        colsum(dl.data_ptr(), ncls, F, ncls, g_bc.data_ptr())
        row = 0
        for key, size in zip(h.class_keys, h.class_sizes):
            grads[f"classifier.{key}.weight"] = g_wc[row:row + size]
            grads[f"classifier.{key}.bias"] = g_bc[row:row + size]
            row += size
        self.launches += 1
        if h.n_mod == 1:
            if input_grads:
                self.d_x = self._buf("d_x", F, ld)
                gemm(0, 0, F, ld, ncls, dl.data_ptr(), ncls, h.w_cls.data_ptr(), ld, self.d_x.data_ptr(), ld)
            return grads

        # ---- Fusion: Linear -> ReLU -> Dropout (model.py:337-362) -------------------------------------------------
        dh = self._buf("dh", F, 512)
        gemm(0, 0, F, 512, ncls, dl.data_ptr(), ncls, h.w_cls.data_ptr(), 512, dh.data_ptr(), 512)
        p_fus = float(opts.get("fusion_dropout", 0.0))
        c("tbn_relu_dropout_bwd", dh.data_ptr(), feat.data_ptr(), 1.0 / (1.0 - p_fus), F * 512, stream)
        g_wf, g_bf = new(512, ld), new(512)
        gemm(1, 0, 512, ld, F, dh.data_ptr(), 512, X.data_ptr(), ld, g_wf.data_ptr(), ld)
        colsum(dh.data_ptr(), 512, F, 512, g_bf.data_ptr())
        grads["fusion.fusion_layer.0.weight"], grads["fusion.fusion_layer.0.bias"] = g_wf, g_bf
        self.launches += 1
        if input_grads:
            self.d_x = self._buf("d_x", F, ld)
            gemm(0, 0, F, ld, 512, dh.data_ptr(), 512, h.w_fus.data_ptr(), ld, self.d_x.data_ptr(), ld)
        if opts.get("audio_dropped") or not h.use_att or (h.att_type is None and not input_grads):
            return grads  # no trainable layer below the fusion row

        acol = h.modality.index("Audio") * E
        dxa = self._buf("dxa", F, E)  # gradient of the attended audio feature
        gemm(0, 0, F, E, 512, dh.data_ptr(), 512, h.w_fus.data_ptr() + acol * 4, ld, dxa.data_ptr(), E)
        a = h.audio_positions(F, S)
        self._input_grads = input_grads
        if h.att_type is None:  # fixed attention weights: feature = sum_s a[f][s] w[f][s]
            self.d_a = self._buf("d_a", F * S, E)
            c("tbn_weighted_sum_bwd_a", opts["fixed_weights"].data_ptr(), dxa.data_ptr(), E, self.d_a.data_ptr(), F, S, E, stream)
            return grads
        if h.att_type == "mha":
            self._mha(grads, F, S, X, ld, a, dxa, d_wts, opts, stream, gemm, colsum, new)
        else:
            self._mlp(grads, F, S, X, ld, a, dxa, d_wts, opts, stream, gemm, colsum, new)
        return grads

    # ---- nn.MultiheadAttention + pe (attention.py:48-57, model.py:62-67) ------------------------------------------
    def _mha(self, grads, F, S, X, ld, a, dxa, d_wts, opts, stream, gemm, colsum, new):
        h, c, E = self.head, _lib.call, FEATURE_SIZE
        H = h.heads
        FS = F * S
        W_in, b_in = self.sd[_A + "in_proj_weight"], self.sd[_A + "in_proj_bias"]
        assert W_in.is_contiguous() and W_in.dtype == torch.float32 and W_in.is_cuda
        wk, wv = W_in.data_ptr() + E * E * 4, W_in.data_ptr() + 2 * E * E * 4
        bk, bv = b_in.data_ptr() + E * 4, b_in.data_ptr() + 2 * E * 4
        z, q, o = h.buf("z", FS, E), h.buf("q", F, E), h.buf("o", F, E)
        # out projection
        g_wo, g_bo = new(E, E), new(E)
        gemm(1, 0, E, E, F, dxa.data_ptr(), E, o.data_ptr(), E, g_wo.data_ptr(), E)
        colsum(dxa.data_ptr(), E, F, E, g_bo.data_ptr())
        d_o = self._buf("d_o", F, E)
        gemm(0, 0, F, E, E, dxa.data_ptr(), E, h.w_o.data_ptr(), E, d_o.data_ptr(), E)
        # keys / values (the forward folds them into the query / output side: recompute)
        k, v = self._buf("k", FS, E), self._buf("v", FS, E)
        c("tbn_linear", z.data_ptr(), wk, bk, k.data_ptr(), FS, E, E, E, E, E, 1, 0, stream)
        c("tbn_linear", z.data_ptr(), wv, bv, v.data_ptr(), FS, E, E, E, E, E, 1, 0, stream)
        dq, dk, dv = self._buf("dq", F, E), self._buf("dk", FS, E), self._buf("dv", FS, E)
        mask = opts.get("attn_mask")
        dw = d_wts.float().contiguous() if d_wts is not None else None
        c("tbn_mha_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), mask.data_ptr() if mask is not None else 0, d_o.data_ptr(),
          dw.data_ptr() if dw is not None else 0, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), F, S, E, H, stream)
        g_in, g_inb = new(3 * E, E), new(3 * E)
        gemm(1, 0, E, E, F, dq.data_ptr(), E, X.data_ptr(), ld, g_in.data_ptr(), E)  # query = first modality's feature
        gemm(1, 0, E, E, FS, dk.data_ptr(), E, z.data_ptr(), E, g_in.data_ptr() + E * E * 4, E)
        gemm(1, 0, E, E, FS, dv.data_ptr(), E, z.data_ptr(), E, g_in.data_ptr() + 2 * E * E * 4, E)
        colsum(dq.data_ptr(), E, F, E, g_inb.data_ptr())
        colsum(dk.data_ptr(), E, FS, E, g_inb.data_ptr() + E * 4)
        colsum(dv.data_ptr(), E, FS, E, g_inb.data_ptr() + 2 * E * 4)
        dz = self._buf("dz", FS, E)
        gemm(0, 0, FS, E, E, dk.data_ptr(), E, wk, E, dz.data_ptr(), E)
        gemm(0, 0, FS, E, E, dv.data_ptr(), E, wv, E, dz.data_ptr(), E, beta=1.0)
        # GroupNorm: the forward normalised in place, recompute its input (Conv1d k=1 over [feature | PE])
        pre = self._buf("pre", FS, E)
        c("tbn_linear", a.data_ptr(), h.w_pe.data_ptr(), h.b_pe.data_ptr(), pre.data_ptr(), FS, E, E, E, E, E, S, 0, stream)
        dc, dgf, dbf = self._buf("dc", FS, E), self._buf("dgf", F, E), self._buf("dbf", F, E)
        c("tbn_groupnorm_bwd", pre.data_ptr(), dz.data_ptr(), h.gn_w.data_ptr(), dc.data_ptr(), dgf.data_ptr(), dbf.data_ptr(), F, S, E,
          64, 1e-5, stream)
        g_gw, g_gb = new(E), new(E)
        colsum(dgf.data_ptr(), E, F, E, g_gw.data_ptr())
        colsum(dbf.data_ptr(), E, F, E, g_gb.data_ptr())
        # Conv1d(E + n_pe, E, 1): columns [0, E) see the audio positions, the rest the positional-encoding rows
        pe = self.sd["pe.0.pe"]  # (1, n_pe, S)
        n_pe = pe.shape[1]
        if self._pe_rows is None or self._pe_rows.shape[0] != FS:
            self._pe_rows = pe[0].t().float().repeat(F, 1).contiguous().to(h.device)  # (F*S, n_pe): row (f, s) = pe[:, s]
        if self._input_grads:
            # query side: d(first modality's feature) += dq W_q;  key / value side: d(audio positions) = dc W_pe[:, :E]
            gemm(0, 0, F, E, E, dq.data_ptr(), E, h.w_q.data_ptr(), E, self.d_x.data_ptr(), ld, beta=1.0)
            self.d_a = self._buf("d_a", FS, E)
            gemm(0, 0, FS, E, E, dc.data_ptr(), E, h.w_pe.data_ptr(), E, self.d_a.data_ptr(), E)
        g_pw, g_pb = new(E, E + n_pe, 1), new(E)
        gemm(1, 0, E, E, FS, dc.data_ptr(), E, a.data_ptr(), E, g_pw.data_ptr(), E + n_pe)
        gemm(1, 0, E, n_pe, FS, dc.data_ptr(), E, self._pe_rows.data_ptr(), n_pe, g_pw.data_ptr() + E * 4, E + n_pe)
        colsum(dc.data_ptr(), E, FS, E, g_pb.data_ptr())
        self.launches += 5
        grads.update({"pe.1.weight": g_pw, "pe.1.bias": g_pb, "pe.2.weight": g_gw, "pe.2.bias": g_gb, _A + "in_proj_weight": g_in,
                      _A + "in_proj_bias": g_inb, _A + "out_proj.weight": g_wo, _A + "out_proj.bias": g_bo})

    # ---- UniModalAttention / PrototypeAttention (attention.py:60-145) ---------------------------------------------
    def _mlp(self, grads, F, S, X, ld, a, dxa, d_wts, opts, stream, gemm, colsum, new):
        h, c, E = self.head, _lib.call, FEATURE_SIZE
        n_out, hid = h.n_out, 256
        dw_in = d_wts.float().contiguous() if d_wts is not None else None
        dwt = self._buf("dwt", F, S)
        c("tbn_weighted_sum_bwd", a.data_ptr(), dxa.data_ptr(), E, dw_in.data_ptr() if dw_in is not None else 0, dwt.data_ptr(), F, S, E,
          stream)
        if h.att_type == "proto":  # weights = mul @ prototype_wts (a buffer, not trained)
            proto = self.sd["attention_layer.prototype_wts"].float().contiguous()  # (n_out, S)
            dmul = self._buf("dmul", F, n_out)
            gemm(0, 1, F, n_out, S, dwt.data_ptr(), S, proto.data_ptr(), S, dmul.data_ptr(), n_out)
        else:
            dmul = dwt
        # softmax / straight-through Gumbel-softmax (tau = 1): the gradient flows through the soft sample
        y = h.buf("ysoft", F, n_out)
        dlg = self._buf("dlg", F, n_out)
        c("tbn_softmax_bwd", y.data_ptr(), dmul.data_ptr(), dlg.data_ptr(), F, n_out, n_out, stream)
        hdn = h.buf("hdn", F, hid)
        w2 = self.sd[_M + "2.weight"]
        g_w2, g_b2 = new(n_out, hid), new(n_out)
        gemm(1, 0, n_out, hid, F, dlg.data_ptr(), n_out, hdn.data_ptr(), hid, g_w2.data_ptr(), hid)
        colsum(dlg.data_ptr(), n_out, F, n_out, g_b2.data_ptr())
        dhd = self._buf("dhdn", F, hid)
        gemm(0, 0, F, hid, n_out, dlg.data_ptr(), n_out, w2.data_ptr(), hid, dhd.data_ptr(), hid)
        c("tbn_relu_dropout_bwd", dhd.data_ptr(), hdn.data_ptr(), 1.0, F * hid, stream)
        if self._input_grads:
            w1 = self.sd[_M + "0.weight"]
            gemm(0, 0, F, E, hid, dhd.data_ptr(), hid, w1.data_ptr(), E, self.d_x.data_ptr(), ld, beta=1.0)
            self.d_a = self._buf("d_a", F * S, E)
            c("tbn_weighted_sum_bwd_a", h.last_wts.data_ptr(), dxa.data_ptr(), E, self.d_a.data_ptr(), F, S, E, stream)
        g_w1, g_b1 = new(hid, E), new(hid)
        gemm(1, 0, hid, E, F, dhd.data_ptr(), hid, X.data_ptr(), ld, g_w1.data_ptr(), E)
        colsum(dhd.data_ptr(), hid, F, hid, g_b1.data_ptr())
        self.launches += 3
        grads.update({_M + "0.weight": g_w1, _M + "0.bias": g_b1, _M + "2.weight": g_w2, _M + "2.bias": g_b2})


class HeadGradFn(torch.autograd.Function):
    """Puts the head backward behind the model outputs: ``forward`` hands the scores / weights of the training forward
    through, ``backward`` receives their gradients from the loss and returns the gradients of the head parameters."""

    @staticmethod
    def forward(ctx, model, token, scores, wts, *params):
        ctx.model, ctx.token, ctx.has_wts = model, token, wts.numel() > 0
        return scores.clone(), wts.clone()

    @staticmethod
    def backward(ctx, d_scores, d_wts):
        grads = ctx.model._head_backward(ctx.token, d_scores, d_wts if ctx.has_wts else None)
        return (None, None, None, None) + tuple(grads)
