"""Execution plans: weight preparation and the kernel launch sequence of the TBN forward.

A plan turns the module's parameters into kernel-ready operands once (BatchNorm folded into the
conv, weights repacked K-major, positional encoding folded into a per-position bias, the
attention key projection folded into the query) and records the forward as a flat list of C-ABI
calls on pre-allocated NHWC workspaces.  Running the plan is a loop over that list on the current
CUDA stream, so the whole forward is CUDA-graph capturable.

Data layout in HBM
  * activations: NHWC, bf16 (performance mode) or fp32 (parity mode); every inception block owns
    ONE (F, H, W, C_total) buffer and its branches write at their channel offsets, so the ten
    ``torch.cat`` of the reference (bn_inception_audio.py:485, 544, ... 994) never run.
  * mid-fusion row: one fp32 (F, 1024*M) matrix; each modality's pooled feature / attention
    output is written at column 1024*m (reference model.py:250).
  * frames are processed in chunks so that the inter-layer activations of a chunk stay L2
    resident (126 MB) instead of round-tripping HBM.

Reference for the arithmetic: SURVEY.md appendix B; reference core/models/model.py:205-262,
bn_inception.py:16-35, attention.py:8-57.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import POOL_AVG, POOL_MAX, TBN_BF16, TBN_F32
from .arch import FEATURE_SIZE, BlockSpec, ConvSpec, block_specs, pool_out_ceil, stem_specs, conv_out

PRECISIONS = ("bf16", "fp32", "bf16_simt")
BN_EPS = 1e-5
BN_MOMENTUM = 0.1  # nn.BatchNorm2d default (the reference's backbone layers are constructed with it)

# Optional instrumentation: ``LAUNCH_HOOK(name, args, flops, launch)`` replaces the plain call (bench.py
# brackets launches with CUDA events through it).  ``flops`` = algorithmic 2*MAC of a conv launch, else 0.
LAUNCH_HOOK = None


def _launch(name: str, *args, flops: float = 0.0) -> None:
    if LAUNCH_HOOK is None:
        _lib.call(name, *args)
    else:
        LAUNCH_HOOK(name, args, flops, lambda: _lib.call(name, *args))


def _round16(v: int) -> int:
    return (v + 15) // 16 * 16


def _round64(v: int) -> int:
    return (v + 63) // 64 * 64


class _Dyn:
    """Placeholder for a pointer that is only known when the plan runs."""

    def __init__(self, name: str, offset: int = 0):
        self.name, self.offset = name, offset


def pack_stem_rows(wf: torch.Tensor, bias: torch.Tensor, dtype=torch.float16) -> torch.Tensor:
    """(64, cin, 7, 7) + folded bias (64) -> fp16 (64, K64): conv1 operand layout of the fused stem kernel
    (tbn_stem_fused_tc; include/tbn_sm100.h).  Column (c*7 + r)*8 + s + 1 holds w[c, r, s]; column
    (c*7 + r)*8 + 0 pairs with input column 2*ox - 4 and stays zero; columns cin*56 and cin*56 + 1 hold the bias as
    bf16 hi and lo parts (the kernel's operand tile has ones there, so the bias add rides on the MMA)."""
    cout, cin, k, _ = wf.shape
    out = torch.zeros(cout, cin * 7, 8, dtype=torch.float64)
    out[:, :, 1:] = wf.double().reshape(cout, cin * 7, 7)
    kp = _round64(cin * 56 + 2)
    full = torch.zeros(cout, kp, dtype=torch.float64)
    full[:, :cin * 56] = out.reshape(cout, cin * 56)
    packed = full.to(dtype)
    hi = bias.double().to(dtype)
    packed[:, cin * 56] = hi
    packed[:, cin * 56 + 1] = (bias.double() - hi.double()).to(dtype)
    return packed


class PreparedConv:
    """Kernel operands of one conv+BN+ReLU triple (BN folded on the host in float64)."""

    def __init__(self, spec: ConvSpec, sd: Dict[str, torch.Tensor], prefix: str, precision: str, device, train: bool = False):
        self.train = train
        if train:
            # training forward: BatchNorm is NOT folded (it normalises with batch statistics); the conv keeps its own
            # bias, and the BatchNorm tensors are the module's LIVE parameters / buffers (running statistics are updated
            # in place by tbn_bn_train_relu)
            self.spec, self.strided = spec, spec.k not in (1, 3)
            self._w = sd[prefix + spec.name + ".weight"]
            self.bias = sd[prefix + spec.name + ".bias"].detach().float().contiguous().to(device)
            self.bn = tuple(sd[prefix + spec.name + "_bn." + nm] for nm in ("weight", "bias", "running_mean", "running_var"))
            for t in self.bn:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "BatchNorm tensors must be fp32 on the device"
            self.wk = None
            self.refresh()
            return
        w = sd[prefix + spec.name + ".weight"].detach().cpu().double()
        b = sd[prefix + spec.name + ".bias"].detach().cpu().double()
        g = sd[prefix + spec.name + "_bn.weight"].detach().cpu().double()
        beta = sd[prefix + spec.name + "_bn.bias"].detach().cpu().double()
        mu = sd[prefix + spec.name + "_bn.running_mean"].detach().cpu().double()
        var = sd[prefix + spec.name + "_bn.running_var"].detach().cpu().double()
        s = g / torch.sqrt(var + BN_EPS)
        wf = w * s.view(-1, 1, 1, 1)  # (cout, cin, k, k)
        self.spec = spec
        self.bias = (beta + s * (b - mu)).float().to(device)
        k, cin, cout = spec.k, spec.cin, spec.cout
        self.strided = k not in (1, 3)  # the 7x7/2 stem conv
        self.wk = self.wp = self.wp_rows_f16 = self.wp_f16 = None
        if precision == "bf16":
            if self.strided:  # conv1 lives in the fused stem, which multiplies fp16 operands (csrc/stem_fused.cu)
                self.wp_rows_f16 = pack_stem_rows(wf, (beta + s * (b - mu)), torch.float16).contiguous().to(device)
                return
            tapmajor = wf.permute(0, 2, 3, 1).contiguous()  # (cout, k, k, cin)
            wp = torch.zeros(cout, k * k, _round64(cin), dtype=torch.float64)
            wp[:, :, :cin] = tapmajor.reshape(cout, k * k, cin)
            wp = wp.reshape(cout, -1)
            if spec.name == "conv2_3x3_reduce":  # second GEMM of the fused stem: fp16 operands as well
                self.wp_f16 = wp.to(torch.float16).contiguous().to(device)
            self.wp = wp.to(torch.bfloat16).contiguous().to(device)
        else:
            if precision == "bf16_simt":  # same weight rounding as the tensor-core path
                wf = wf.to(torch.bfloat16).double()
            self.wk = wf.permute(2, 3, 1, 0).reshape(k * k * cin, cout).float().contiguous().to(device)


def _prepared_conv_refresh(self) -> None:
    """(training) re-pack the live conv weight (cout, cin, k, k) -> fp32 [tap * cin][cout] for tbn_conv_simt."""
    sp = self.spec
    wk = self._w.detach().float().permute(2, 3, 1, 0).reshape(sp.k * sp.k * sp.cin, sp.cout).contiguous()
    if self.wk is None:
        self.wk = wk
    else:
        self.wk.copy_(wk)  # same address: recorded launches stay valid


PreparedConv.refresh = _prepared_conv_refresh


class _Buf:
    """A channel slice of an NHWC workspace."""

    def __init__(self, t: torch.Tensor, h: int, w: int, cs: int, co: int = 0, c: Optional[int] = None):
        self.t, self.h, self.w, self.cs, self.co = t, h, w, cs, co
        self.c = cs if c is None else c

    def slice(self, co: int, c: int) -> "_Buf":
        return _Buf(self.t, self.h, self.w, self.cs, self.co + co, c)

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()


class FusedReduce:
    """The same-input 1x1 convs of an inception block (1x1, 3x3_reduce, double_3x3_reduce) as ONE
    GEMM: weights / biases concatenated along Cout (tensor-core mode only)."""

    def __init__(self, convs: List[PreparedConv]):
        self.convs = convs
        self.cin = convs[0].spec.cin
        self.wp = torch.cat([c.wp for c in convs], 0).contiguous()
        self.bias = torch.cat([c.bias for c in convs], 0).contiguous()
        self.cout = sum(c.spec.cout for c in convs)


class BackbonePlan:
    """One BN-Inception backbone (reference bn_inception.py:38-107 + Cadene features()).

    Execution shape
      * phase A (stem .. inception_3c, the high-resolution layers) runs on ``chunk`` frames at a time so
        that its activations stay L2 resident; phase B (4a .. 5b, small maps) runs on up to ``chunk_lo``
        frames at once so that its grids fill the 148 SMs.
      * inside a block the three independent branch chains run on three stream lanes (fork / join with
        events, which become parallel branches of the captured CUDA graph).
      * tensor-core mode fuses the same-input 1x1 convs of a block into one multi-destination GEMM.
    """

    LANES = 3

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, in_channels: int, precision: str, device,
                 chunk: int = 32, chunk_lo: Optional[int] = None, fuse_1x1: bool = True, multi_stream: bool = True,
                 train: bool = False):
        assert precision in PRECISIONS, precision
        assert not train or precision == "fp32", "the training forward runs on the fp32 CUDA-core path"
        self.train = train
        self.precision, self.device, self.in_channels = precision, device, in_channels
        self.chunk = int(chunk)
        self.chunk_lo = max(int(chunk_lo or 8 * chunk), self.chunk)
        self.chunk_lo = (self.chunk_lo // self.chunk) * self.chunk
        self.dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        self.dcode = TBN_F32 if precision == "fp32" else TBN_BF16
        self.esize = 4 if precision == "fp32" else 2
        self.fuse_1x1 = fuse_1x1 and precision == "bf16"
        self.multi_stream = multi_stream
        self.stem = [PreparedConv(c, sd, prefix, precision, device, train) for c in stem_specs(in_channels)]
        self.blocks: List[Tuple[BlockSpec, Dict[str, PreparedConv], Optional[FusedReduce]]] = []
        for b in block_specs():
            convs = {}
            for c in ([b.b1x1] if b.b1x1 else []) + list(b.b3x3) + list(b.bdbl) + ([b.bpool] if b.bpool else []):
                convs[c.name] = PreparedConv(c, sd, prefix, precision, device, train)
            fused = None
            if self.fuse_1x1:
                members = ([convs[b.b1x1.name]] if b.b1x1 else []) + [convs[b.b3x3[0].name], convs[b.bdbl[0].name]]
                if b.pool == "avg":  # avgpool and the 1x1 pool_proj commute: project first, pool Cout channels
                    members.append(convs[b.bpool.name])
                fused = FusedReduce(members)
            self.blocks.append((b, convs, fused))
        self._ws: Dict[Tuple[int, int], Dict[str, torch.Tensor]] = {}
        self._ops: Dict[tuple, list] = {}
        self._retired: List[torch.Tensor] = []
        self._keep: list = []  # host-side segment arrays referenced by recorded ops
        self._events: Dict[int, torch.cuda.Event] = {}
        self._side: Optional[List[torch.cuda.Stream]] = None
        self._tuned: Dict[int, dict] = {}  # id(op) -> tuning knobs chosen by autotune()

    def refresh_weights(self) -> None:
        """(training) re-pack every conv weight from the live parameters (after an optimizer step)."""
        for pc in self.stem:
            pc.refresh()
        for _, convs, _ in self.blocks:
            for pc in convs.values():
                pc.refresh()

    # -- workspace -----------------------------------------------------------------------------
    def _alloc(self, ws: Dict[str, torch.Tensor], name: str, numel: int, dtype=None) -> torch.Tensor:
        dtype = dtype or self.dtype
        t = ws.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            if t is not None:  # recorded ops may still point at the old block: never free it
                self._retired.append(t)
            t = torch.empty(numel, dtype=dtype, device=self.device)
            ws[name] = t
        return t

    # -- op recording --------------------------------------------------------------------------
    # op = ("k", lane, entry point, args, flops) | ("rec", lane, event id) | ("wait", lane, event id)
    def _conv(self, ops, ws, lane: int, pc: PreparedConv, n: int, x: _Buf, y: _Buf, x_dyn=None, y_dyn=None):
        sp = pc.spec
        assert x.c == sp.cin and y.c == sp.cout, (sp.name, x.c, sp.cin, y.c, sp.cout)
        xptr = x_dyn if x_dyn is not None else x.ptr
        yptr = y_dyn if y_dyn is not None else y.ptr
        oh, ow = conv_out(x.h, sp.k, sp.stride, sp.pad), conv_out(x.w, sp.k, sp.stride, sp.pad)
        flops = 2.0 * n * oh * ow * sp.cout * sp.cin * sp.k * sp.k
        if self.precision != "bf16":
            ops.append(("k", lane, "tbn_conv_simt", (xptr, pc.wk.data_ptr(), pc.bias.data_ptr(), yptr, n, x.h, x.w, sp.cin,
                                                     x.cs, x.co, sp.cout, y.cs, y.co, sp.k, sp.stride, sp.pad,
                                                     0 if self.train else 1, self.dcode), flops))
            if self.train:  # BatchNorm with batch statistics + ReLU, in place on the output slice (running stats updated)
                g, b, rm, rv = pc.bn
                scratch = self._alloc(ws, f"{self._tag}bnws{lane}", 2 * 1056, torch.float64)
                ops.append(("k", lane, "tbn_bn_train_relu", (yptr, g.data_ptr(), b.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                                             scratch.data_ptr(), BN_MOMENTUM, BN_EPS, n * oh * ow, sp.cout, y.cs, y.co, 1),
                            0.0))
            return
        if not pc.strided and sp.stride == 2:  # 3x3/2 of 3c, 4e: TMA boxes with element stride 2
            ops.append(("k", lane, "tbn_conv_tc_s2", (xptr, pc.wp.data_ptr(), pc.bias.data_ptr(), yptr, n, x.h, x.w, sp.cin,
                                                      x.cs, x.co, sp.cout, y.cs, y.co, 1, TBN_BF16), flops))
            return
        assert not pc.strided, sp.name  # (the 7x7/2 conv of the tensor-core path lives in the fused stem)
        ops.append(("k", lane, "tbn_conv_tc", (xptr, pc.wp.data_ptr(), pc.bias.data_ptr(), yptr, n, x.h, x.w, sp.cin,
                                               x.cs, x.co, sp.cout, y.cs, y.co, sp.k, 1, TBN_BF16, 1), flops))

    def _block(self, ops, ws, A, bi: int, n: int, cur: _Buf, out: _Buf, out_dyn=None):
        """Record one inception block: lane 0 = 1x1 + 3x3 chain, lane 1 = double-3x3 chain, lane 2 = pool chain."""
        b, cv, fused = self.blocks[bi]
        h, w = cur.h, cur.w
        off = b.out_offsets
        L1, L2 = (1, 2) if self.multi_stream else (0, 0)
        ev = 10 * bi
        r3, c3 = b.b3x3
        rd, d1, d2 = b.bdbl
        t3 = A("t_3r", h, w, r3.cout)
        td = A("t_dr", h, w, rd.cout)
        td1 = A("t_d1", h, w, d1.cout)
        oslice = lambda key, c: out.slice(off[key], c)  # noqa: E731
        if self.multi_stream:
            ops.append(("rec", 0, ev))  # block input is ready on lane 0
            ops.append(("wait", L2, ev))
        commuted = fused is not None and b.pool == "avg"
        tpp = A("t_pp", h, w, b.bpool.cout) if commuted else None
        # lane 2: pooling branch (depends on the block input only, unless the projection is commuted)
        if b.reduce:
            ops.append(("k", L2, "tbn_maxpool3s2", (cur.ptr, out_dyn if out_dyn is not None else out.ptr, n, h, w, b.cin,
                                                    cur.cs, cur.co, out.cs, off["pool"], self.dcode), 0.0))
        elif not commuted:
            tp = A("t_pool", h, w, b.cin)
            mode = POOL_AVG if b.pool == "avg" else POOL_MAX
            ops.append(("k", L2, "tbn_pool3s1", (cur.ptr, tp.ptr, n, h, w, b.cin, cur.cs, cur.co, tp.cs, 0, mode, self.dcode), 0.0))
            self._conv(ops, ws, L2, cv[b.bpool.name], n, tp, oslice("pool", b.bpool.cout), y_dyn=out_dyn)
        if self.multi_stream and not commuted:
            ops.append(("rec", L2, ev + 1))
        # lane 0: the same-input 1x1 convs
        if fused is not None:
            both = _lib.SEG_RELU | _lib.SEG_BIAS
            segs = []
            if b.b1x1:
                segs.append((out_dyn if out_dyn is not None else out.ptr, out.cs, off["1x1"], b.b1x1.cout, both))
            segs += [(t3.ptr, t3.cs, 0, r3.cout, both), (td.ptr, td.cs, 0, rd.cout, both)]
            if commuted:
                segs.append((tpp.ptr, tpp.cs, 0, b.bpool.cout, 0))  # raw projection: bias + ReLU come after the pool
            flops = 2.0 * n * h * w * fused.cout * fused.cin
            ops.append(("k", 0, "tbn_conv_tc_multi", (cur.ptr, fused.wp.data_ptr(), fused.bias.data_ptr(), _Segs(segs),
                                                      len(segs), n, h, w, fused.cin, cur.cs, cur.co, 1, TBN_BF16), flops))
        else:
            if b.b1x1:
                self._conv(ops, ws, 0, cv[b.b1x1.name], n, cur, oslice("1x1", b.b1x1.cout), y_dyn=out_dyn)
            self._conv(ops, ws, 0, cv[r3.name], n, cur, t3)
            self._conv(ops, ws, 0, cv[rd.name], n, cur, td)
        if self.multi_stream:
            ops.append(("rec", 0, ev + 2))
            ops.append(("wait", L1, ev + 2))
            if commuted:
                ops.append(("wait", L2, ev + 2))
        if commuted:  # lane 2: pool the projected (Cout-channel) pre-activation, then bias + ReLU into the concat slot
            pp = cv[b.bpool.name]
            ops.append(("k", L2, "tbn_avgpool3s1_bias_relu", (tpp.ptr, pp.bias.data_ptr(), out_dyn if out_dyn is not None else out.ptr,
                                                              n, h, w, b.bpool.cout, tpp.cs, 0, out.cs, off["pool"], self.dcode), 0.0))
            if self.multi_stream:
                ops.append(("rec", L2, ev + 1))
        # lane 1: double 3x3 chain; lane 0: 3x3
        self._conv(ops, ws, L1, cv[d1.name], n, td, td1)
        self._conv(ops, ws, L1, cv[d2.name], n, td1, oslice("dbl", d2.cout), y_dyn=out_dyn)
        if self.multi_stream:
            ops.append(("rec", L1, ev + 3))
        self._conv(ops, ws, 0, cv[c3.name], n, t3, oslice("3x3", c3.cout), y_dyn=out_dyn)
        if self.multi_stream:
            ops.append(("wait", 0, ev + 1))
            ops.append(("wait", 0, ev + 3))

    # -- grouped phase B: conv problems are DESCRIBED, not recorded, so that JointPhaseB can hand the same layer of several
    #    modalities, and the independent 3x3 convs of a block, to one tbn_conv_tc_group launch
    def _desc(self, wp: torch.Tensor, bias: torch.Tensor, cin: int, n: int, x: _Buf, segs: list, k: int, stride: int) -> dict:
        oh, ow = conv_out(x.h, k, stride, k // 2), conv_out(x.w, k, stride, k // 2)
        cout = sum(sg[3] for sg in segs)
        return dict(x=x.ptr, wp=wp.data_ptr(), bias=bias.data_ptr(), segs=segs, n=n, h=x.h, w=x.w, cin=cin, x_cs=x.cs, x_co=x.co,
                    k=k, stride=stride, out_dtype=TBN_BF16, flops=2.0 * n * oh * ow * cout * cin * k * k)

    def _conv_desc(self, pc: PreparedConv, n: int, x: _Buf, y: _Buf) -> dict:
        sp = pc.spec
        assert x.c == sp.cin and y.c == sp.cout and not pc.strided, (sp.name, x.c, sp.cin, y.c, sp.cout)
        return self._desc(pc.wp, pc.bias, sp.cin, n, x, [(y.ptr, y.cs, y.co, sp.cout, _lib.SEG_RELU | _lib.SEG_BIAS)], sp.k, sp.stride)

    def block_items(self, A, bi: int, n: int, cur: _Buf, out: _Buf) -> dict:
        """One inception block as data: ``pre`` (kernels that only need the block input: the pass-through max-pool of a
        reduce block, the 3x3/1 max-pool of 5b), ``fused`` (the same-input 1x1 convs as one multi-destination GEMM),
        ``post`` (the commuted avg-pool + bias + ReLU, after ``fused``), ``g1`` (the two independent 3x3 convs that follow
        the reduce convs) and ``g2`` (the last double-3x3 conv, plus 5b's pool_proj which waits for ``pre``)."""
        b, cv, fused = self.blocks[bi]
        assert fused is not None
        h, w = cur.h, cur.w
        off = b.out_offsets
        r3, c3 = b.b3x3
        rd, d1, d2 = b.bdbl
        t3, td, td1 = A("t_3r", h, w, r3.cout), A("t_dr", h, w, rd.cout), A("t_d1", h, w, d1.cout)
        oslice = lambda key, c: out.slice(off[key], c)  # noqa: E731
        commuted = b.pool == "avg"
        both = _lib.SEG_RELU | _lib.SEG_BIAS
        pre, post, g2 = [], [], []
        if b.reduce:
            pre.append(("tbn_maxpool3s2", (cur.ptr, out.ptr, n, h, w, b.cin, cur.cs, cur.co, out.cs, off["pool"], self.dcode)))
        elif not commuted:  # 5b: MaxPool(3, 1, 1) -> pool_proj
            tp = A("t_pool", h, w, b.cin)
            pre.append(("tbn_pool3s1", (cur.ptr, tp.ptr, n, h, w, b.cin, cur.cs, cur.co, tp.cs, 0, POOL_MAX, self.dcode)))
            g2.append(self._conv_desc(cv[b.bpool.name], n, tp, oslice("pool", b.bpool.cout)))
        segs = []
        if b.b1x1:
            segs.append((out.ptr, out.cs, off["1x1"], b.b1x1.cout, both))
        segs += [(t3.ptr, t3.cs, 0, r3.cout, both), (td.ptr, td.cs, 0, rd.cout, both)]
        if commuted:
            tpp = A("t_pp", h, w, b.bpool.cout)
            segs.append((tpp.ptr, tpp.cs, 0, b.bpool.cout, 0))  # raw projection: bias + ReLU come after the pool
            pp = cv[b.bpool.name]
            post.append(("tbn_avgpool3s1_bias_relu", (tpp.ptr, pp.bias.data_ptr(), out.ptr, n, h, w, b.bpool.cout, tpp.cs, 0, out.cs,
                                                      off["pool"], self.dcode)))
        fused_d = self._desc(fused.wp, fused.bias, fused.cin, n, cur, segs, 1, 1)
        g1 = [self._conv_desc(cv[c3.name], n, t3, oslice("3x3", c3.cout)), self._conv_desc(cv[d1.name], n, td, td1)]
        g2.insert(0, self._conv_desc(cv[d2.name], n, td1, oslice("dbl", d2.cout)))
        return dict(pre=pre, fused=fused_d, post=post, g1=g1, g2=g2)

    def _feat_op(self, cur: _Buf, n: int, h: int, w: int, feat_ld: int, freq_only: bool, split_ld: int):
        """Feature pooling (ref bn_inception.py:16-35) into the caller's fp32 matrix and, with tensor-core heads, also as
        the split operand of the first head GEMM (pointers ``feat`` / ``split`` are dynamic)."""
        if split_ld:
            return ("tbn_avgpool_feat_split", (cur.ptr, _Dyn("feat"), _Dyn("split"), split_ld, n, h, w, FEATURE_SIZE, feat_ld,
                                               1 if freq_only else 0, self.dcode))
        return ("tbn_avgpool_feat", (cur.ptr, _Dyn("feat"), n, h, w, FEATURE_SIZE, feat_ld, 1 if freq_only else 0, self.dcode))

    def phase_b_items(self, n: int, H: int, W: int, feat_ld: int, freq_only: bool, split_ld: int = 0) -> list:
        """Phase B of this backbone as a list of block items + the feature-pooling kernel (pointer ``feat`` is dynamic)."""
        ws = self._ws.setdefault((H, W), {})
        self._tag, self._cap = "b_", self.chunk_lo
        A = lambda name, h, w, c: _Buf(self._alloc(ws, "b_" + name, self.chunk_lo * h * w * c), h, w, c)  # noqa: E731
        h, w = self.mid_spatial(H, W)
        cur = _Buf(self._mid_buffer(H, W), h, w, self.mid_channels)
        for role, numel in self._temps(ws, h, w, self.N_HI_BLOCKS, len(self.blocks)).items():
            self._alloc(ws, "b_" + role, self.chunk_lo * numel)
        items = []
        for bi in range(self.N_HI_BLOCKS, len(self.blocks)):
            b = self.blocks[bi][0]
            oh, ow = (conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)) if b.reduce else (h, w)
            out = A(f"out_{b.name}", oh, ow, b.cout)
            items.append(self.block_items(A, bi, n, cur, out))
            cur, h, w = out, oh, ow
        return items, self._feat_op(cur, n, h, w, feat_ld, freq_only, split_ld)

    def _temps(self, ws, h: int, w: int, first: int, last: int):
        """Branch temporaries shared by blocks first..last-1, sized for the largest so addresses never move."""
        tmax = {"t_3r": 0, "t_dr": 0, "t_d1": 0, "t_pool": 0, "t_pp": 0}
        for b, _, _ in self.blocks[first:last]:
            if b.bpool:
                tmax["t_pp"] = max(tmax["t_pp"], h * w * b.bpool.cout)
            tmax["t_3r"] = max(tmax["t_3r"], h * w * b.b3x3[0].cout)
            tmax["t_dr"] = max(tmax["t_dr"], h * w * b.bdbl[0].cout)
            tmax["t_d1"] = max(tmax["t_d1"], h * w * b.bdbl[1].cout)
            if not b.reduce:
                tmax["t_pool"] = max(tmax["t_pool"], h * w * b.cin)
            else:
                h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        return tmax

    # Inception blocks that run in the chunked phase A.  0: phase A = stem .. pool2 only (the layers whose
    # activations are too large for L2 at full batch); 3a .. 5b then run on all frames at once, so that every
    # launch has enough tiles to amortise its fixed cost (pipeline fill + epilogue tail, ~8 us measured).
    N_HI_BLOCKS = int(__import__('os').environ.get('TBN_N_HI', '0'))  # (experiment switch)

    def _build_a(self, n: int, H: int, W: int, in_bf16: bool = False) -> list:
        """Phase A: stem .. 3c on n frames; the 3c concat buffer (phase B input) is a dynamic pointer."""
        ws = self._ws.setdefault((H, W), {})
        ops: list = []
        cin = self.in_channels
        self._tag, self._cap = "a_", self.chunk
        A = lambda name, h, w, c: _Buf(self._alloc(ws, "a_" + name, self.chunk * h * w * c), h, w, c)  # noqa: E731
        c1, c2r, c2 = self.stem
        h1, w1 = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
        if self.precision == "bf16":
            assert cin in (1, 3, 10), cin
            # whole stem in one kernel: conv1 + pool1 + conv2_3x3_reduce (neither intermediate touches memory)
            h2, w2 = pool_out_ceil(h1), pool_out_ceil(w1)
            y2r = A("conv2r", h2, w2, 64)
            ops.append(("k", 0, "tbn_stem_fused_tc", (_Dyn("input"), _lib.TBN_BF16 if in_bf16 else _lib.TBN_F32,
                                                      c1.wp_rows_f16.data_ptr(), c2r.wp_f16.data_ptr(), c2r.bias.data_ptr(),
                                                      y2r.ptr, n, cin, H, W, 64, 0),
                        2.0 * n * (h1 * w1 * 64 * 49 * cin + h2 * w2 * 64 * 64)))
            return self._build_a_tail(ops, ws, A, n, y2r, c2, h2, w2)
        if in_bf16:
            raise ValueError("bf16 inputs need the tensor-core path (precision='bf16')")
        y1 = A("conv1", h1, w1, 64)
        xin = _Buf(self._alloc(ws, "a_x_nhwc", self.chunk * H * W * cin), H, W, cin)
        ops.append(("k", 0, "tbn_nchw_to_nhwc", (_Dyn("input"), xin.ptr, n, cin, H, W, cin, self.dcode), 0.0))
        self._conv(ops, ws, 0, c1, n, xin, y1)
        h2, w2 = pool_out_ceil(h1), pool_out_ceil(w1)
        p1 = A("pool1", h2, w2, 64)
        ops.append(("k", 0, "tbn_maxpool3s2", (y1.ptr, p1.ptr, n, h1, w1, 64, 64, 0, 64, 0, self.dcode), 0.0))
        y2r = A("conv2r", h2, w2, 64)
        self._conv(ops, ws, 0, c2r, n, p1, y2r)
        return self._build_a_tail(ops, ws, A, n, y2r, c2, h2, w2)

    def _build_a_tail(self, ops, ws, A, n: int, y2r: "_Buf", c2: PreparedConv, h2: int, w2: int) -> list:
        """conv2_3x3 -> pool2 (-> inception blocks of phase A, if any)."""
        y2 = A("conv2", h2, w2, 192)
        self._conv(ops, ws, 0, c2, n, y2r, y2)
        h, w = pool_out_ceil(h2), pool_out_ceil(w2)
        if self.N_HI_BLOCKS == 0:  # phase A ends at pool2: its output is phase B's input (dynamic pointer)
            ops.append(("k", 0, "tbn_maxpool3s2", (y2.ptr, _Dyn("mid"), n, h2, w2, 192, 192, 0, 192, 0, self.dcode), 0.0))
            return ops
        cur = A("pool2", h, w, 192)
        ops.append(("k", 0, "tbn_maxpool3s2", (y2.ptr, cur.ptr, n, h2, w2, 192, 192, 0, 192, 0, self.dcode), 0.0))
        for role, numel in self._temps(ws, h, w, 0, self.N_HI_BLOCKS).items():
            self._alloc(ws, "a_" + role, self.chunk * numel)
        for bi in range(self.N_HI_BLOCKS):
            b = self.blocks[bi][0]
            oh, ow = (conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)) if b.reduce else (h, w)
            if bi == self.N_HI_BLOCKS - 1:
                out = _Buf(None, oh, ow, b.cout)  # lives in phase B's workspace: dynamic pointer
                self._block(ops, ws, A, bi, n, cur, out, out_dyn=_Dyn("mid"))
            else:
                out = A(f"out_{b.name}", oh, ow, b.cout)
                self._block(ops, ws, A, bi, n, cur, out)
            cur, h, w = out, oh, ow
        return ops

    def mid_spatial(self, H: int, W: int) -> Tuple[int, int]:
        h, w = pool_out_ceil(conv_out(H, 7, 2, 3)), pool_out_ceil(conv_out(W, 7, 2, 3))
        h, w = pool_out_ceil(h), pool_out_ceil(w)
        for b, _, _ in self.blocks[:self.N_HI_BLOCKS]:
            if b.reduce:
                h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        return h, w

    @property
    def mid_channels(self) -> int:
        return self.blocks[self.N_HI_BLOCKS - 1][0].cout if self.N_HI_BLOCKS > 0 else 192

    def _mid_buffer(self, H: int, W: int) -> torch.Tensor:
        ws = self._ws.setdefault((H, W), {})
        h, w = self.mid_spatial(H, W)
        return self._alloc(ws, "b_mid", self.chunk_lo * h * w * self.mid_channels)

    def _build_b(self, n: int, H: int, W: int, feat_ld: int, freq_only: bool, split_ld: int = 0) -> list:
        """Phase B: 4a .. 5b + feature pooling on n frames read from the 3c concat buffer."""
        ws = self._ws.setdefault((H, W), {})
        ops: list = []
        self._tag, self._cap = "b_", self.chunk_lo
        A = lambda name, h, w, c: _Buf(self._alloc(ws, "b_" + name, self.chunk_lo * h * w * c), h, w, c)  # noqa: E731
        h, w = self.mid_spatial(H, W)
        cur = _Buf(self._mid_buffer(H, W), h, w, self.mid_channels)
        for role, numel in self._temps(ws, h, w, self.N_HI_BLOCKS, len(self.blocks)).items():
            self._alloc(ws, "b_" + role, self.chunk_lo * numel)
        for bi in range(self.N_HI_BLOCKS, len(self.blocks)):
            b = self.blocks[bi][0]
            oh, ow = (conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)) if b.reduce else (h, w)
            out = A(f"out_{b.name}", oh, ow, b.cout)
            self._block(ops, ws, A, bi, n, cur, out)
            cur, h, w = out, oh, ow
        # feature pooling (ref bn_inception.py:16-35) straight into the caller's matrix
        name, args = self._feat_op(cur, n, h, w, feat_ld, freq_only, split_ld)
        ops.append(("k", 0, name, args, 0.0))
        return ops

    def out_spatial(self, H: int, W: int) -> Tuple[int, int]:
        h, w = self.mid_spatial(H, W)
        for b, _, _ in self.blocks[self.N_HI_BLOCKS:]:
            if b.reduce:
                h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        return h, w

    # -- execution -----------------------------------------------------------------------------
    def _lanes(self, main: "torch.cuda.Stream") -> List["torch.cuda.Stream"]:
        if not self.multi_stream:
            return [main]
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(self.LANES - 1)]
        return [main] + self._side

    def _event(self, i: int) -> "torch.cuda.Event":
        e = self._events.get(i)
        if e is None:
            e = self._events[i] = torch.cuda.Event()
        return e

    def _resolve(self, op, dyn: Dict[str, int]) -> list:
        args = []
        for a in op[3]:
            if isinstance(a, _Dyn):
                a = dyn[a.name]
            elif isinstance(a, _Segs):
                a = a.resolve(dyn, self._keep)
            args.append(a)
        return args

    def autotune(self, ops: list, dyn: Dict[str, int], stream: "torch.cuda.Stream", cache: Dict[tuple, dict]) -> None:
        """Per-launch choice of the conv kernel's tiling knobs (halo vs per-tap A loads, one or two CTAs per SM, widest
        N tile): every candidate is timed as a CUDA-graph replay of the real launch on the real buffers and the fastest
        is recorded for this op.  Identical problems (same entry point and dimensions) are tuned once."""
        for op in ops:
            if op[0] != "k" or op[2] not in TUNED_ENTRY_POINTS or id(op) in self._tuned:
                continue
            args = self._resolve(op, dyn)
            key = (op[2],) + tuple(a for a in args if isinstance(a, int) and abs(a) < (1 << 20))
            if isinstance(op[3][3], _Segs):
                key += tuple(sg[3] for sg in op[3][3].segs)
            best = cache.get(key)
            if best is None:
                k3s1 = op[2] == "tbn_conv_tc" and args[13] == 3
                cands = [dict(tc_halo=h, tc_occ2=o, tc_max_bn=b) for h in ((1, 0) if k3s1 else (1,)) for o in (1, 0)
                         for b in (256, 128, 64)]
                best, best_t = None, None
                for cand in cands:
                    _apply_tuning(cand)
                    t = _time_launch(op[2], args, stream)
                    if best_t is None or t < best_t * 0.98:  # prefer earlier (default-like) candidates on ties
                        best, best_t = cand, t
                _apply_tuning(TUNE_DEFAULTS)
                cache[key] = best
            self._tuned[id(op)] = best

    def _execute(self, ops: list, dyn: Dict[str, int], lanes: List["torch.cuda.Stream"]) -> int:
        launches = 0
        ptrs = [s.cuda_stream for s in lanes]
        for op in ops:
            kind, lane = op[0], op[1]
            if kind == "g":  # grouped conv launch: ("g", lane, address of the tbn_conv_desc array, count, flops, keepalive)
                _launch("tbn_conv_tc_group", op[2], op[3], ptrs[lane], flops=op[4])
                launches += 1
            elif kind == "k":
                args = self._resolve(op, dyn)
                tune = self._tuned.get(id(op))
                if tune:
                    _apply_tuning(tune)
                _launch(op[2], *args, ptrs[lane], flops=op[4])
                if tune:
                    _apply_tuning(TUNE_DEFAULTS)
                launches += 1
            elif kind == "rec":
                self._event(op[2]).record(lanes[lane])
            else:
                lanes[lane].wait_event(self._event(op[2]))
        return launches

    def run_a(self, x: torch.Tensor, g0: int, n_lo: int, main: "torch.cuda.Stream",
              tune_cache: Optional[Dict[tuple, dict]] = None) -> int:
        """Phase A (stem .. pool2) of frames [g0, g0 + n_lo) in chunks, into the phase-B input buffer."""
        F, C, H, W = x.shape
        in_bf16 = x.dtype == torch.bfloat16
        lanes = self._lanes(main)
        mid = self._mid_buffer(H, W)
        mh, mw = self.mid_spatial(H, W)
        mid_frame_bytes = mh * mw * self.mid_channels * self.esize
        launches = 0
        for f0 in range(g0, g0 + n_lo, self.chunk):
            n = min(self.chunk, g0 + n_lo - f0)
            key = ("a", n, H, W, in_bf16)
            ops = self._ops.get(key)
            if ops is None:
                ops = self._ops[key] = self._build_a(n, H, W, in_bf16)
            dyn = {"input": x.data_ptr() + f0 * C * H * W * x.element_size(), "mid": mid.data_ptr() + (f0 - g0) * mid_frame_bytes}
            if tune_cache is not None and self.precision == "bf16" and not torch.cuda.is_current_stream_capturing():
                self.autotune(ops, dyn, main, tune_cache)
            launches += self._execute(ops, dyn, lanes)
        return launches

    def run(self, x: torch.Tensor, feat: torch.Tensor, feat_col: int, feat_ld: int, freq_only: bool,
            main: "torch.cuda.Stream", tune_cache: Optional[Dict[tuple, dict]] = None,
            split: Optional[Tuple[torch.Tensor, int, int]] = None) -> int:
        """x: (F, C, H, W) fp32 (or bf16 with the fused bf16 stem) contiguous on the device.  feat: fp32 matrix receiving the pooled
        feature at column ``feat_col`` with leading dimension ``feat_ld`` -- rows = frames, or
        frames*S positions when ``freq_only``.  Returns the number of kernel launches."""
        F, C, H, W = x.shape
        assert C == self.in_channels and x.dtype in (torch.float32, torch.bfloat16) and x.is_contiguous()
        if self.train and (self.chunk != F or self.chunk_lo != F):
            # BatchNorm statistics span the WHOLE batch (reference: one bn call over all B*N frames): no frame chunks.
            # Recorded ops hold workspace addresses sized for the old chunk: rebuild them.
            self.chunk = self.chunk_lo = F
            self._ops.clear()
        in_bf16 = x.dtype == torch.bfloat16
        rows_per_frame = self.out_spatial(H, W)[1] if freq_only else 1
        lanes = self._lanes(main)
        launches = 0
        for g0 in range(0, F, self.chunk_lo):
            n_lo = min(self.chunk_lo, F - g0)
            launches += self.run_a(x, g0, n_lo, main, tune_cache)
            split_ld = split[2] if split is not None else 0
            key = ("b", n_lo, H, W, feat_ld, int(freq_only), split_ld)
            ops = self._ops.get(key)
            if ops is None:
                ops = self._ops[key] = self._build_b(n_lo, H, W, feat_ld, freq_only, split_ld)
            dyn = {"feat": feat.data_ptr() + (g0 * rows_per_frame * feat_ld + feat_col) * 4}
            if split is not None:  # (tensor, first column, leading dim) of the bf16 split operand
                dyn["split"] = split[0].data_ptr() + (g0 * rows_per_frame * split_ld + split[1]) * 2
            if tune_cache is not None and self.precision == "bf16" and not torch.cuda.is_current_stream_capturing():
                self.autotune(ops, dyn, main, tune_cache)
            launches += self._execute(ops, dyn, lanes)
        return launches


TUNE_DEFAULTS = dict(tc_halo=1, tc_occ2=1, tc_max_bn=256)
TUNED_ENTRY_POINTS = ("tbn_conv_tc", "tbn_conv_tc_multi", "tbn_conv_tc_s2")


def _apply_tuning(tune: Dict[str, int]) -> None:
    for k, v in tune.items():
        _lib.call("tbn_set_tuning", k.encode(), int(v))


def _time_launch(name: str, args: list, stream: "torch.cuda.Stream", reps: int = 4) -> float:
    """Milliseconds per launch, measured as a CUDA-graph replay (no host overhead in the number)."""
    with torch.cuda.stream(stream):
        _lib.call(name, *args, stream.cuda_stream)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(reps):
                _lib.call(name, *args, stream.cuda_stream)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        g.replay()
        g.replay()
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / (2 * reps)


class _Segs:
    """Output-segment list of a fused GEMM whose pointers may be dynamic."""

    def __init__(self, segs):
        self.segs = segs
        self._static = None

    def resolve(self, dyn: Dict[str, int], keep: list) -> int:
        if self._static is not None:
            return self._static
        has_dyn = any(isinstance(s[0], _Dyn) for s in self.segs)
        arr, addr = _lib.make_segs([(dyn[s[0].name] if isinstance(s[0], _Dyn) else s[0],) + tuple(s[1:]) for s in self.segs])
        self._arr = arr  # the library copies the segments during the call; keep the array alive until then
        if not has_dyn:
            self._static = addr
        return addr


class JointPhaseB:
    """Phase B (inception 3a .. 5b + feature pooling) of ALL modality backbones as ONE launch sequence (bf16 tensor-core
    mode).  The reference runs the modalities one after the other (model.py:208-213) and a block's branches one after
    the other; here every stage of a block is a single grouped launch over all modalities (tbn_conv_tc_group):
        fused 1x1 GEMMs  ->  {3x3, double_3x3_1}  ->  {double_3x3_2 (+ 5b's pool_proj)}
    i.e. 3 conv launches per block for any number of modalities instead of 4-5 per modality, with the pooling kernels on
    one side lane.  A persistent conv grid owns every SM, so separate launches serialise anyway; grouped, a launch has
    enough tiles for whole waves (the 7x7 / 8x13 maps have 1-2.6 waves per modality) and pays fill / drain once."""

    def __init__(self, backs: Dict[str, BackbonePlan]):
        self.backs = backs
        self.first = next(iter(backs.values()))
        self._ops: Dict[tuple, list] = {}
        self._keep: list = []

    def _build(self, spec: tuple) -> list:
        """spec = ((modality, n, H, W, feat_ld, freq_only, split_ld), ...)"""
        per = []
        for m, n, H, W, feat_ld, freq_only, split_ld in spec:
            items, feat = self.backs[m].phase_b_items(n, H, W, feat_ld, bool(freq_only), split_ld)
            per.append((m, items, feat))
        ops: list = []
        nblocks = len(per[0][1])
        import os
        PL = int(os.environ.get("TBN_GROUP_POOL_LANE", "2"))  # experiment: 2 = pooling kernels on a side lane, 0 = in line
        TWO = int(os.environ.get("TBN_GROUP_LANES", "1")) == 2 and self.first.multi_stream

        def group(lane, descs):
            arr, addr, keep = _lib.make_descs(descs)
            self._keep.append(keep)
            ops.append(("g", lane, addr, len(descs), sum(d["flops"] for d in descs)))

        for bi in range(nblocks):
            its = [items[bi] for _, items, _ in per]
            ev = 10 * bi
            ops.append(("rec", 0, ev))            # block inputs ready
            ops.append(("wait", PL, ev))
            for it in its:                          # side lane: pooling that only needs the block input
                for name, args in it["pre"]:
                    ops.append(("k", PL, name, args, 0.0))
            group(0, [it["fused"] for it in its])
            ops.append(("rec", 0, ev + 2))
            ops.append(("wait", PL, ev + 2))
            for it in its:                          # side lane: commuted avg-pool + bias + ReLU
                for name, args in it["post"]:
                    ops.append(("k", PL, name, args, 0.0))
            ops.append(("rec", PL, ev + 1))
            if TWO:  # two conv chains: lane 0 = the 3x3 convs, lane 1 = the double-3x3 chain (each launch spans the modalities)
                ops.append(("wait", 1, ev + 2))
                group(0, [it["g1"][0] for it in its])
                group(1, [it["g1"][1] for it in its])
                if any(len(it["g2"]) > 1 for it in its):
                    ops.append(("wait", 1, ev + 1))
                group(1, [d for it in its for d in it["g2"]])
                ops.append(("rec", 1, ev + 3))
                ops.append(("wait", 0, ev + 3))
            else:
                group(0, [d for it in its for d in it["g1"]])
                if any(len(it["g2"]) > 1 for it in its):  # 5b: pool_proj reads the side lane's max-pool
                    ops.append(("wait", 0, ev + 1))
                group(0, [d for it in its for d in it["g2"]])
            ops.append(("wait", 0, ev + 1))
        for m, _, feat in per:
            name, args = feat
            ops.append(("k", 0, name, tuple(_Dyn(a.name + "_" + m) if isinstance(a, _Dyn) else a for a in args), 0.0))
        return ops

    def run(self, spec: tuple, feat_ptrs: Dict[str, int], main: "torch.cuda.Stream", split_ptrs: Optional[Dict[str, int]] = None) -> int:
        ops = self._ops.get(spec)
        if ops is None:
            ops = self._ops[spec] = self._build(spec)
        lanes = self.first._lanes(main) if self.first.multi_stream else [main, main, main]
        dyn = {"feat_" + m: p for m, p in feat_ptrs.items()}
        dyn.update({"split_" + m: p for m, p in (split_ptrs or {}).items()})
        return self.first._execute(ops, dyn, lanes)


class HeadPlan:
    """Attention + mid-fusion + classifiers + consensus (reference model.py:214-262), fp32."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg, modality: List[str], device, tensor_cores: bool = False):
        self.cfg, self.modality, self.device = cfg, modality, device
        # tensor_cores: the GEMMs run as bf16 tcgen05 GEMMs over [hi | hi | lo] x [hi | lo | hi] split operands
        # (tbn_split_bf16x3 + tbn_conv_tc): fp32-class accuracy, ~10x faster than the CUDA-core linear kernel.
        self.tc = bool(tensor_cores)
        att = cfg.model.attention
        self.n_mod = len(modality)
        self.use_att = bool(att.enable) and "Audio" in modality and self.n_mod > 1
        self.fixed = self.use_att and bool(att.use_fixed)
        self.att_type = att.type if (self.use_att and not self.fixed) else None
        dev = lambda t: t.detach().float().contiguous().to(device)  # noqa: E731
        self.class_keys = list(cfg.model.num_classes.keys())
        self.class_sizes = [int(cfg.model.num_classes[k]) for k in self.class_keys]
        self.n_cls = sum(self.class_sizes)
        self.w_cls = dev(torch.cat([sd[f"classifier.{k}.weight"] for k in self.class_keys], 0))
        self.b_cls = dev(torch.cat([sd[f"classifier.{k}.bias"] for k in self.class_keys], 0))
        if self.n_mod > 1:
            self.w_fus = dev(sd["fusion.fusion_layer.0.weight"])
            self.b_fus = dev(sd["fusion.fusion_layer.0.bias"])
        if self.att_type == "mha":
            if "pe.1.weight" not in sd:
                # reference quirk: type == "mha" calls self.pe unconditionally (model.py:230)
                raise AttributeError("'TBNModel' object has no attribute 'pe' (attention.type='mha' needs use_pe=True)")
            E = FEATURE_SIZE
            self.heads = int(att.attn_heads)
            d = E // self.heads
            w_pe = sd["pe.1.weight"].detach().double()[:, :, 0]  # (1024, 1034)
            pe = sd["pe.0.pe"].detach().double()[0].to(w_pe.device)  # (10, S)
            self.S = pe.shape[1]
            self.w_pe = dev(w_pe[:, :E])
            # the PE rows of the concat are constants: fold them into a per-position bias table
            self.b_pe = dev((w_pe[:, E:] @ pe).t() + sd["pe.1.bias"].detach().double())  # (S, 1024)
            self.gn_w, self.gn_b = dev(sd["pe.2.weight"]), dev(sd["pe.2.bias"])
            a = "attention_layer.attention_layer."
            W = sd[a + "in_proj_weight"].detach().double()
            bvec = sd[a + "in_proj_bias"].detach().double()
            self.w_q, self.b_q = dev(W[:E]), dev(bvec[:E])
            # scores_h,p = (W_k,h^T q_h / sqrt(d)) . z_p  (+ a per-(f,h) constant that softmax ignores)
            wk = W[E:2 * E].view(self.heads, d, E)
            self.w_kt = dev((wk.transpose(1, 2) / math.sqrt(d)).reshape(self.heads * E, d))  # (H*E, d)
            self.w_v, self.b_v = dev(W[2 * E:]), dev(bvec[2 * E:])
            self.w_o, self.b_o = dev(sd[a + "out_proj.weight"]), dev(sd[a + "out_proj.bias"])
        elif self.att_type in ("unimodal", "proto"):
            a = "attention_layer.seq."
            self.w_h, self.b_h = dev(sd[a + "0.weight"]), dev(sd[a + "0.bias"])
            w2, b2 = sd[a + "2.weight"].detach().float().cpu(), sd[a + "2.bias"].detach().float().cpu()
            n_out = w2.shape[0]
            self.n_logit = (n_out + 3) // 4 * 4  # pad to a multiple of 4; padded logits get -inf-like bias
            w2p = torch.zeros(self.n_logit, w2.shape[1]); w2p[:n_out] = w2
            b2p = torch.full((self.n_logit,), -1e30); b2p[:n_out] = b2
            self.w_2, self.b_2 = dev(w2p), dev(b2p)
            self.n_out = n_out
            if self.att_type == "proto":
                pw = sd["attention_layer.prototype_wts"].detach().float().cpu()  # (3, S)
                pt = torch.zeros(pw.shape[1], self.n_logit); pt[:, :n_out] = pw.t()
                self.proto_t = dev(pt)  # (S, n_logit)
        self._ws: Dict[str, torch.Tensor] = {}
        self._retired: List[torch.Tensor] = []
        self._x3_blocks: set = set()  # modality blocks of the fusion row's split operand already written this forward
        self._pe_split = False         # the attended audio positions were written as the PE GEMM's split operand
        if self.tc:
            self._prepare_tc(sd)

    # -- tensor-core head GEMMs ----------------------------------------------------------------------
    def _pack3(self, w: torch.Tensor, rows_pad: int = 0) -> torch.Tensor:
        """(N, K) float64 -> bf16 (N [padded], 3K) = [hi | lo | hi] matching activations split as [hi | hi | lo]."""
        w = w.double()
        hi = w.to(torch.bfloat16)
        lo = (w - hi.double()).to(torch.bfloat16)
        out = torch.cat([hi, lo, hi], 1)
        if rows_pad > out.shape[0]:
            out = torch.cat([out, torch.zeros(rows_pad - out.shape[0], out.shape[1], dtype=torch.bfloat16)], 0)
        return out.contiguous().to(self.device)

    def _prepare_tc(self, sd) -> None:
        dev32 = lambda t: t.float().contiguous().to(self.device)  # noqa: E731
        E = FEATURE_SIZE
        self.n_cls_pad = _round16(self.n_cls)
        w_cls = torch.cat([sd[f"classifier.{k}.weight"].detach().cpu().double() for k in self.class_keys], 0)
        b_cls = torch.cat([sd[f"classifier.{k}.bias"].detach().cpu().double() for k in self.class_keys], 0)
        self.t_cls = self._pack3(w_cls, self.n_cls_pad)
        self.t_cls_b = dev32(torch.cat([b_cls, torch.zeros(self.n_cls_pad - self.n_cls, dtype=torch.float64)]))
        if self.n_mod > 1:
            # the fusion row's split operand holds one [hi | hi | lo] block per modality (its producers write it): pack
            # the weights block-wise to match
            wf = sd["fusion.fusion_layer.0.weight"].detach().cpu().double()
            self.t_fus = torch.cat([self._pack3(wf[:, g * E:(g + 1) * E]) for g in range(self.n_mod)], 1).contiguous()
        if self.att_type == "mha":
            H, d = self.heads, E // self.heads
            self.t_pe = self._pack3(sd["pe.1.weight"].detach().cpu().double()[:, :E, 0])
            a = "attention_layer.attention_layer."
            W = sd[a + "in_proj_weight"].detach().cpu().double()
            bvec = sd[a + "in_proj_bias"].detach().cpu().double()
            Wo = sd[a + "out_proj.weight"].detach().cpu().double()
            bo = sd[a + "out_proj.bias"].detach().cpu().double()
            # scores need  qk[f,h,:] = W_k,h^T (W_q,h x + b_q,h) / sqrt(d)  =  M_h x + c_h : ONE GEMM with N = H*E
            Ms, cs, Ps = [], [], []
            for h in range(H):
                Wq, bq = W[h * d:(h + 1) * d], bvec[h * d:(h + 1) * d]
                Wk = W[E + h * d:E + (h + 1) * d]
                Wv = W[2 * E + h * d:2 * E + (h + 1) * d]
                Ms.append(Wk.t() @ Wq / math.sqrt(d))
                cs.append(Wk.t() @ bq / math.sqrt(d))
                Ps.append(Wo[:, h * d:(h + 1) * d] @ Wv)  # out = sum_h (W_o,h W_v,h) zbar_h + (W_o b_v + b_o)
            self.t_qk = self._pack3(torch.cat(Ms, 0))                       # (H*E, 3E)
            self.t_qk_b = dev32(torch.cat(cs, 0))
            # zbar is (F, H*E) with head-major columns: split per head block so that K stays head-contiguous
            self.t_out = torch.cat([self._pack3(P) for P in Ps], 1).contiguous()  # (E, H*3E)
            self.t_out_b = dev32(Wo @ bvec[2 * E:] + bo)

    def _grow(self, name: str, numel: int, dtype) -> torch.Tensor:
        """Workspace ``name`` with at least ``numel`` elements.  A block that has to grow is RETIRED, never freed:
        captured CUDA graphs (TBNModel._graphs) of smaller batches keep replaying into the old addresses."""
        t = self._ws.get(name)
        if t is None or t.numel() < numel:
            if t is not None:
                self._retired.append(t)
            t = self._ws[name] = torch.empty(numel, dtype=dtype, device=self.device)
        return t

    def bufh(self, name: str, *shape) -> torch.Tensor:
        numel = 1
        for s in shape:
            numel *= s
        return self._grow(name, numel, torch.bfloat16)[:numel].view(*shape)

    def _split(self, x_ptr: int, m: int, k: int, ldx: int, out_ptr: int, out_ld: int, stream: int, group: int = 0) -> int:
        """fp32 (m, k; leading dim ldx) -> bf16 [hi | hi | lo] blocks in rows of out_ld elements at out_ptr."""
        _launch("tbn_split_bf16x3_ld", x_ptr, out_ptr, out_ld, m, k, ldx, group or k, stream)
        return 1

    def _gemm3(self, a3_ptr: int, a3_ld: int, m: int, k3: int, w3: torch.Tensor, bias: torch.Tensor, y: Optional[torch.Tensor],
               n: int, ldy: int, relu: int, stream: int, bias_period: int = 1, tag: str = "", nseg: int = 1,
               split_out: Optional[Tuple[int, int]] = None, y_ptr: Optional[int] = None) -> int:
        """y (m / nseg, n; leading dim ldy) = mean_nseg act(A (m, k3 bf16 split operand; leading dim a3_ld) @ W3^T + bias) on
        tcgen05; ``split_out`` = (pointer, leading dim): the result is (also) written as the split operand of the next
        GEMM.  Tall-K / few-row products run split-K over 8x more CTAs and the reduction kernel does the rest; otherwise
        the conv kernel writes fp32 and the consensus / split run as their own (small) launches."""
        yp = y_ptr if y_ptr is not None else (y.data_ptr() if y is not None else 0)
        kblocks = k3 // 64
        if m <= 512 and kblocks >= 48:
            ksplit = next(d for d in (8, 6, 4, 3, 2, 1) if kblocks % d == 0)
            part = self.buf("part_" + tag, ksplit, m, n)
            _launch("tbn_gemm_tc_splitk_ld", a3_ptr, a3_ld, w3.data_ptr(), part.data_ptr(), m, k3, n, ksplit, stream)
            sp, sld = split_out if split_out is not None else (0, 0)
            _launch("tbn_splitk_reduce_ex", part.data_ptr(), bias.data_ptr(), yp, m, n, ldy, ksplit, bias_period, relu, nseg, sp, sld,
                    stream)
            return 2
        launches = 1
        need_tmp = nseg > 1 or yp == 0
        full = self.buf("full_" + tag, m, n) if need_tmp else None
        _launch("tbn_conv_tc", a3_ptr, w3.data_ptr(), bias.data_ptr(), full.data_ptr() if need_tmp else yp, 1, 1, m, k3, a3_ld, 0, n,
                n if need_tmp else ldy, 0, 1, relu, TBN_F32, bias_period, stream)  # (no flops: bench.py accounts the backbone convs)
        src, src_ld, rows = (full.data_ptr(), n, m) if need_tmp else (yp, ldy, m)
        if nseg > 1:
            assert yp == 0 or ldy == n, "the segment mean writes a contiguous matrix"
            cons_ptr = yp if yp else self.buf("cons_" + tag, m // nseg, n).data_ptr()
            _launch("tbn_consensus", src, cons_ptr, m // nseg, nseg, n, stream)
            src, src_ld, rows = cons_ptr, n, m // nseg
            launches += 1
        if split_out is not None:
            launches += self._split(src, rows, n, src_ld, split_out[0], split_out[1], stream)
        return launches

    def buf(self, name: str, *shape) -> torch.Tensor:
        numel = 1
        for s in shape:
            numel *= s
        return self._grow(name, numel, torch.float32)[:numel].view(*shape)

    def fusion_row(self, F: int) -> torch.Tensor:
        return self.buf("fusion_in", F, FEATURE_SIZE * self.n_mod)

    def audio_positions(self, F: int, S: int) -> torch.Tensor:
        return self.buf("audio_pos", F * S, FEATURE_SIZE)

    # -- split operands written by the producers (tensor-core heads) -----------------------------------------
    def x3(self, F: int) -> torch.Tensor:
        """Split operand of the fusion row: (F, n_mod * 3E) bf16, one [hi | hi | lo] block of E columns per modality."""
        return self.bufh("x3", F, 3 * FEATURE_SIZE * self.n_mod)

    def split_target(self, m_no: int, F: int, positions: bool, S: int = 0) -> Optional[Tuple[torch.Tensor, int, int]]:
        """Where a backbone's feature pooling writes its rows as a split GEMM operand: (tensor, first column, leading dim)
        -- the positional-encoding GEMM's operand for the attended audio positions, else the modality's block of the
        fusion row's operand.  None without tensor-core heads."""
        if not self.tc:
            return None
        E = FEATURE_SIZE
        if positions:
            self._pe_split = self.att_type == "mha"
            return (self.bufh("a3_pe", F * S, 3 * E), 0, 3 * E) if self._pe_split else None
        self._x3_blocks.add(m_no)
        return self.x3(F), m_no * 3 * E, 3 * E * self.n_mod

    def run_qk(self, F: int, stream: int) -> int:
        """The folded q.k projection of the MHA (scores need M_h x + c_h of the VISUAL query only): callers enqueue it as soon
        as the first modality's features are in the fusion row -- on that modality's stream, or on a side stream -- so that it
        overlaps the audio backbone / the positional-encoding GEMM instead of sitting on the serial tail.  ``run`` skips it."""
        if not (self.att_type == "mha" and self.tc):
            return 0
        E, H = FEATURE_SIZE, self.heads
        qk = self.buf("qk", F, H * E)
        self._qk_ready = F
        n = 0
        if 0 not in self._x3_blocks:  # (the query modality's block was not written by its backbone)
            n = self._split(self.fusion_row(F).data_ptr(), F, E, E * self.n_mod, self.x3(F).data_ptr(), 3 * E * self.n_mod, stream)
            self._x3_blocks.add(0)
        return n + self._gemm3(self.x3(F).data_ptr(), 3 * E * self.n_mod, F, 3 * E, self.t_qk, self.t_qk_b, qk, H * E, H * E, 0, stream,
                               tag="qk")

    def run(self, B: int, N: int, S: int, fixed_weights: Optional[torch.Tensor], stream: int, n_audio: Optional[int] = None,
            before_attention=None, train: Optional[dict] = None):
        """Consumes fusion_row (non-audio columns filled) and audio_positions; returns
        (scores (B, n_cls), weights (F, S) or None, entropy (F,) or None, launches).  ``n_audio`` < N: the 10-crop path
        (reference model.py:243-248) -- the audio feature has B*n_audio rows and is tiled N / n_audio times."""
        F = B * N
        Fa = B * (n_audio if n_audio else N)  # audio rows
        crops = F // Fa
        E = FEATURE_SIZE
        ld = E * self.n_mod
        X = self.fusion_row(F)
        acol = self.modality.index("Audio") * E if "Audio" in self.modality else 0
        c = _launch
        launches = 0
        wts = ent = None
        train = train or {}
        assert not train or not self.tc, "the training forward uses the fp32 CUDA-core heads"
        if train.get("audio_dropped"):
            # reference model.py:215-222: the audio feature is replaced by zeros (the attention is skipped)
            X[:, acol:acol + E].zero_()
        elif self.use_att:
            a = self.audio_positions(Fa, S)
            ydst = X.data_ptr() + acol * 4
            if self.fixed:
                assert fixed_weights is not None and fixed_weights.numel() == Fa * S
                c("tbn_attn_weighted_sum", a.data_ptr(), fixed_weights.data_ptr(), ydst, Fa, S, E, ld, stream)
                launches += 1
            elif self.att_type == "mha" and self.tc:
                assert S == self.S, f"audio window {S} != positional-encoding length {self.S}"
                H = self.heads
                z = self.buf("z", F * S, E)
                a3pe = self.bufh("a3_pe", F * S, 3 * E)
                if not self._pe_split:  # (the audio backbone did not write the operand: split the fp32 positions)
                    launches += self._split(a.data_ptr(), F * S, E, E, a3pe.data_ptr(), 3 * E, stream)
                launches += self._gemm3(a3pe.data_ptr(), 3 * E, F * S, 3 * E, self.t_pe, self.b_pe, z, E, E, 0, stream, bias_period=S,
                                        tag="pe")
                qk = self.buf("qk", F, H * E)
                if getattr(self, "_qk_ready", None) == F:  # already enqueued by run_qk on another stream
                    self._qk_ready = None
                    if before_attention is not None:
                        before_attention()  # join that stream
                else:
                    launches += self.run_qk(F, stream)
                    self._qk_ready = None
                wts = self.buf("wts", F, S)
                ent = self.buf("entropy", F)
                # GroupNorm (pe[2]) runs inside the attention core on the staged z; the weighted value sums come out as the
                # split operand of the out projection (no tbn_groupnorm / tbn_split_bf16x3 launches)
                zbar3 = self.bufh("a3_out", F, 3 * H * E)
                c("tbn_mha_core_gn", qk.data_ptr(), z.data_ptr(), self.gn_w.data_ptr(), self.gn_b.data_ptr(), 64, 1e-5, 0,
                  zbar3.data_ptr(), wts.data_ptr(), ent.data_ptr(), F, S, E, H, stream)
                # out projection: fp32 into the fusion row AND as its split block (the Fusion GEMM's operand)
                launches += self._gemm3(zbar3.data_ptr(), 3 * H * E, F, 3 * H * E, self.t_out, self.t_out_b, None, E, ld, 0, stream,
                                        tag="out", y_ptr=ydst,
                                        split_out=(self.x3(F).data_ptr() + acol * 3 * 2, 3 * E * self.n_mod))
                self._x3_blocks.add(acol // E)
                launches += 1
            elif self.att_type == "mha":
                assert S == self.S, f"audio window {S} != positional-encoding length {self.S}"
                H = self.heads
                d = E // H
                z = self.buf("z", F * S, E)
                c("tbn_linear", a.data_ptr(), self.w_pe.data_ptr(), self.b_pe.data_ptr(), z.data_ptr(), F * S, E, E, E, E,
                  E, S, 0, stream)
                c("tbn_groupnorm", z.data_ptr(), self.gn_w.data_ptr(), self.gn_b.data_ptr(), F, S, E, 64, 1e-5, stream)
                q = self.buf("q", F, E)
                c("tbn_linear", X.data_ptr(), self.w_q.data_ptr(), self.b_q.data_ptr(), q.data_ptr(), F, E, E, ld, E, E,
                  1, 0, stream)
                qk = self.buf("qk", F, H * E)
                c("tbn_linear_grouped", q.data_ptr(), self.w_kt.data_ptr(), 0, qk.data_ptr(), F, E, d, E, d, H * E, H,
                  d, E * d, 0, E, 0, stream)
                zbar = self.buf("zbar", F, H * E)
                wts = self.buf("wts", F, S)
                ent = self.buf("entropy", F)
                mask = train.get("attn_mask")  # training: dropout on the attention probabilities (F, H, S)
                if mask is not None:
                    assert mask.numel() == F * H * S and mask.dtype == torch.float32 and mask.is_contiguous()
                    psum = self.buf("psum", F, H)
                    c("tbn_mha_core_train", qk.data_ptr(), z.data_ptr(), mask.data_ptr(), zbar.data_ptr(), wts.data_ptr(),
                      ent.data_ptr(), psum.data_ptr(), F, S, E, H, stream)
                else:
                    c("tbn_mha_core", qk.data_ptr(), z.data_ptr(), zbar.data_ptr(), wts.data_ptr(), ent.data_ptr(), F, S, E,
                      H, stream)
                o = self.buf("o", F, E)
                c("tbn_linear_grouped", zbar.data_ptr(), self.w_v.data_ptr(), self.b_v.data_ptr(), o.data_ptr(), F, d, E,
                  H * E, E, E, H, E, d * E, d, d, 0, stream)
                if mask is not None:  # the value bias is weighted by the kept probability mass of its (frame, head)
                    c("tbn_value_bias_fix", o.data_ptr(), self.b_v.data_ptr(), psum.data_ptr(), F, E, H, stream)
                    launches += 1
                c("tbn_linear", o.data_ptr(), self.w_o.data_ptr(), self.b_o.data_ptr(), ydst, F, E, E, E, E, ld, 1, 0,
                  stream)
                launches += 8
            else:  # unimodal / proto (reference attention.py:60-145, eval: plain softmax)
                hdn = self.buf("hdn", F, 256)
                c("tbn_linear", X.data_ptr(), self.w_h.data_ptr(), self.b_h.data_ptr(), hdn.data_ptr(), F, 256, E, ld, E,
                  256, 1, 1, stream)
                lg = self.buf("lg", F, self.n_logit)
                c("tbn_linear", hdn.data_ptr(), self.w_2.data_ptr(), self.b_2.data_ptr(), lg.data_ptr(), F, self.n_logit,
                  256, 256, 256, self.n_logit, 1, 0, stream)
                if train.get("gumbel"):
                    # training with use_gumbel (attention.py:82-84, 136-137): a hard Gumbel-softmax sample instead of the
                    # softmax -- torch's own sampler on the logits, so that the RNG stream is the reference's
                    # (spelled out as torch.nn.functional.gumbel_softmax does it, to keep the soft sample for the backward)
                    lgt = lg[:, :self.n_out]
                    gum = -torch.empty_like(lgt, memory_format=torch.legacy_contiguous_format).exponential_().log()
                    y_soft = ((lgt + gum) / 1.0).softmax(1)
                    y_hard = torch.zeros_like(lgt, memory_format=torch.legacy_contiguous_format).scatter_(
                        1, y_soft.max(1, keepdim=True)[1], 1.0)
                    lg[:, :self.n_out] = y_hard - y_soft + y_soft
                    lg[:, self.n_out:] = 0
                    self.buf("ysoft", F, self.n_out).copy_(y_soft)
                    train["gumbel_noise"] = gum
                else:
                    c("tbn_softmax_rows", lg.data_ptr(), lg.data_ptr(), F, self.n_logit, stream)
                    if train:  # (kept for the backward)
                        self.buf("ysoft", F, self.n_out).copy_(lg[:, :self.n_out])
                if self.att_type == "proto":
                    wts = self.buf("wts", F, S)
                    c("tbn_linear", lg.data_ptr(), self.proto_t.data_ptr(), 0, wts.data_ptr(), F, S, self.n_logit,
                      self.n_logit, self.n_logit, S, 1, 0, stream)
                    launches += 1
                else:
                    assert self.n_logit == S or self.n_out == S
                    wts = lg if self.n_logit == S else None
                    if wts is None:  # compact the padded softmax to (F, S)
                        wts = self.buf("wts", F, S)
                        c("tbn_linear", lg.data_ptr(), self._eye(S).data_ptr(), 0, wts.data_ptr(), F, S, self.n_logit,
                          self.n_logit, self.n_logit, S, 1, 0, stream)
                        launches += 1
                c("tbn_attn_weighted_sum", a.data_ptr(), wts.data_ptr(), ydst, F, S, E, ld, stream)
                launches += 4
        if crops > 1:  # audio rows [0, Fa) -> [j*Fa, (j+1)*Fa): feature.repeat(crops, 1)
            assert "Audio" in self.modality and (not self.use_att or self.fixed) and Fa * crops == F
            c("tbn_repeat_rows", X.data_ptr() + acol * 4, Fa, E, ld, crops, stream)
            self._x3_blocks.discard(acol // E)  # (its split block holds Fa rows only)
            launches += 1
        if self.tc:
            x3 = self.x3(F)
            x3_ld = 3 * E * self.n_mod
            for g in range(self.n_mod):  # blocks no producer wrote (fixed / MLP attention output, 10-crop repeat, fp32 features)
                if g not in self._x3_blocks:
                    launches += self._split(X.data_ptr() + g * E * 4, F, E, ld, x3.data_ptr() + g * 3 * E * 2, x3_ld, stream)
            self._x3_blocks = set()
            self._pe_split = False
            scores = self.buf("scores", B, self.n_cls_pad)
            if self.n_mod > 1:
                # Fusion (Linear + ReLU) -> segment consensus -> split operand of the classifier, in the reduction kernel:
                # the mean over a clip's segments commutes with the affine classifier (reference model.py:194-199, 365-386)
                a3c = self.bufh("a3_cls", B, 3 * 512)
                launches += self._gemm3(x3.data_ptr(), x3_ld, F, x3_ld, self.t_fus, self.b_fus, None, 512, 512, 1, stream, tag="fus",
                                        nseg=N, split_out=(a3c.data_ptr(), 3 * 512))
                launches += self._gemm3(a3c.data_ptr(), 3 * 512, B, 3 * 512, self.t_cls, self.t_cls_b, scores, self.n_cls_pad,
                                        self.n_cls_pad, 0, stream, tag="cls")
            else:  # single modality: classifier on the feature, then the consensus
                launches += self._gemm3(x3.data_ptr(), x3_ld, F, x3_ld, self.t_cls, self.t_cls_b, scores, self.n_cls_pad,
                                        self.n_cls_pad, 0, stream, tag="cls", nseg=N)
            return scores, wts, ent, launches
        feat, k_in = X, ld
        if self.n_mod > 1:
            hfus = self.buf("hfus", F, 512)
            c("tbn_linear", X.data_ptr(), self.w_fus.data_ptr(), self.b_fus.data_ptr(), hfus.data_ptr(), F, 512, ld, ld,
              ld, 512, 1, 1, stream)
            if train.get("fusion_dropout", 0) > 0:  # nn.Dropout after the ReLU (model.py:354-360), torch's own RNG stream
                # (drawn as a mask of the same shape: the same Philox offsets as dropout on the tensor itself)
                train["fusion_mask"] = torch.nn.functional.dropout(torch.ones_like(hfus), train["fusion_dropout"], True)
                hfus.mul_(train["fusion_mask"])
            feat, k_in = hfus, 512
            launches += 1
        self.last_wts = wts  # (the backward of the weighted-sum attentions reads them)
        logits = self.buf("logits", F, self.n_cls)
        c("tbn_linear", feat.data_ptr(), self.w_cls.data_ptr(), self.b_cls.data_ptr(), logits.data_ptr(), F, self.n_cls,
          k_in, k_in, k_in, self.n_cls, 1, 0, stream)
        scores = self.buf("scores", B, self.n_cls)
        c("tbn_consensus", logits.data_ptr(), scores.data_ptr(), B, N, self.n_cls, stream)
        launches += 2
        return scores, wts, ent, launches

    def _eye(self, S: int) -> torch.Tensor:
        t = self._ws.get("eye")
        if t is None:
            e = torch.zeros(S, self.n_logit)
            e[:, :S] = torch.eye(S)
            t = self._ws["eye"] = e.to(self.device)
        return t
