"""One BN-Inception backbone in training mode WITH a backward pass (reference bn_inception.py:38-107 + the Cadene
``features()`` dataflow of bn_inception_audio.py:437-1003 under ``model.train()``; gradients as torch autograd derives
them for core/tools/train.py:81).

Differences from the inference plan (plan.BackbonePlan): nothing is fused, folded or chunked -- BatchNorm normalises with
the statistics of the WHOLE batch, every layer keeps its input, its pre-norm conv output and the batch mean / rstd for
the backward, and the forward records a tape that ``backward`` replays in reverse:

  conv + BN + ReLU : tbn_conv_simt (bias, no ReLU) -> tbn_bn_train_fwd          | tbn_bn_train_bwd -> wgrad (tbn_im2col_f32 +
                                                                                 | tbn_sgemm_splitk) -> dgrad (tbn_conv_simt on the
                                                                                 | gradient with flipped weights; stride 2 via
                                                                                 | tbn_upsample_zero2)
  max pools        : tbn_maxpool3s2 / tbn_pool3s1                               | tbn_maxpool3_bwd
  avg pool 3x3/1   : tbn_pool3s1                                                | tbn_pool3s1 on the gradient (self-adjoint)
  feature pooling  : tbn_avgpool_feat                                           | tbn_avgpool_feat_bwd

fp32 on CUDA cores throughout (the reference trains in fp32).  torch owns memory and the weight re-layouts.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import _lib
from .arch import FEATURE_SIZE, ConvSpec, block_specs, conv_out, pool_out_ceil, stem_specs
from .plan import BN_EPS, BN_MOMENTUM, POOL_AVG, POOL_MAX, TBN_F32

_SM_TARGET = 148 * 4  # CTAs a weight-gradient GEMM should spread over
PROFILE = {} if __import__("os").environ.get("TBN_TRAIN_PROFILE") else None  # {(entry point, fwd | bwd): milliseconds}


class _Act:
    """A channel slice of an NHWC fp32 activation stored as a (rows, cs) matrix."""

    def __init__(self, t: torch.Tensor, n: int, h: int, w: int, cs: int, co: int = 0, c: Optional[int] = None):
        self.t, self.n, self.h, self.w, self.cs, self.co = t, n, h, w, cs, co
        self.c = cs if c is None else c

    def slice(self, co: int, c: int) -> "_Act":
        return _Act(self.t, self.n, self.h, self.w, self.cs, self.co + co, c)

    @property
    def rows(self) -> int:
        return self.n * self.h * self.w

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()


class TrainBackbone:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, in_channels: int, device, tensor_cores: bool = False):
        """``tensor_cores``: mixed precision -- the forward convs and the input-gradient convs (k = 1, 3) multiply bf16
        copies of their fp32 operands on the tcgen05 path (tbn_conv_tc, fp32 accumulation, fp32 results); BatchNorm, the
        weight gradients, all stored activations / gradients stay fp32."""
        self.sd, self.prefix, self.cin, self.device = sd, prefix, in_channels, device
        self.tc = bool(tensor_cores)
        self._zero_bias = torch.zeros(1056, device=device, dtype=torch.float32)
        self.stem = stem_specs(in_channels)
        self.blocks = block_specs()
        self._bn_ws = torch.zeros(2 * 1056, dtype=torch.float64, device=device)
        self._cols: Optional[torch.Tensor] = None
        self._colst: Optional[torch.Tensor] = None
        self._tape: List[Callable[[], None]] = []
        self._grads: Dict[int, torch.Tensor] = {}
        self.param_grads: Dict[str, torch.Tensor] = {}
        self.launches = 0

    # -- helpers -----------------------------------------------------------------------------------------------------
    def _new(self, n: int, h: int, w: int, c: int, zero: bool = False) -> _Act:
        mk = torch.zeros if zero else torch.empty
        return _Act(mk(n * h * w, c, device=self.device, dtype=torch.float32), n, h, w, c)

    def _g(self, a: _Act) -> _Act:
        """Gradient view of an activation view (the gradient buffer of its tensor is created, zeroed, on first use)."""
        g = self._grads.get(id(a.t))
        if g is None:
            g = self._grads[id(a.t)] = torch.zeros_like(a.t)
        return _Act(g, a.n, a.h, a.w, a.cs, a.co, a.c)

    def _call(self, name: str, *args) -> None:
        if PROFILE is not None:  # TBN_TRAIN_PROFILE=1: serialised per-entry-point timing (scripts/train_step_profile.py)
            import time
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _lib.call(name, *args, self._stream)
            torch.cuda.synchronize()
            key = (name, "bwd" if self._in_backward else "fwd")
            PROFILE[key] = PROFILE.get(key, 0.0) + (time.perf_counter() - t0) * 1e3
        else:
            _lib.call(name, *args, self._stream)
        self.launches += 1

    def _bf16(self, a: _Act) -> torch.Tensor:
        """bf16 contiguous copy (rows, c) of an fp32 view; cached per forward (a block input feeds several convs)."""
        key = (id(a.t), a.co, a.c)
        t = self._bf16_cache.get(key)
        if t is None:
            t = torch.empty(a.rows, a.c, device=self.device, dtype=torch.bfloat16)
            self._call("tbn_cast_bf16_view", a.ptr, a.cs, a.co, t.data_ptr(), a.rows, a.c)
            self._bf16_cache[key] = t
        return t

    @staticmethod
    def _pack_tc(w4: torch.Tensor) -> torch.Tensor:
        """(cout, cin, k, k) fp32 -> bf16 [cout][k*k][cin padded to 64] (the tbn_conv_tc weight layout)."""
        cout, cin, k, _ = w4.shape
        cp = (cin + 63) // 64 * 64
        wp = torch.zeros(cout, k * k, cp, device=w4.device, dtype=torch.bfloat16)
        wp[:, :, :cin] = w4.permute(0, 2, 3, 1).reshape(cout, k * k, cin)
        return wp

    def _conv_tc(self, xb: torch.Tensor, w4: torch.Tensor, bias_ptr: int, out: torch.Tensor, n: int, h: int, w: int, stride: int) -> None:
        """out (fp32, contiguous (rows, cout)) = conv(xb bf16 (n, h, w, cin), w4) on the tensor cores; k in (1, 3), pad k // 2."""
        cout, cin, k, _ = w4.shape
        wp = self._pack_tc(w4)
        self._keep.append(wp)
        if stride == 1:
            self._call("tbn_conv_tc", xb.data_ptr(), wp.data_ptr(), bias_ptr, out.data_ptr(), n, h, w, cin, cin, 0, cout, cout, 0, k, 0,
                       TBN_F32, 1)
        else:
            self._call("tbn_conv_tc_s2", xb.data_ptr(), wp.data_ptr(), bias_ptr, out.data_ptr(), n, h, w, cin, cin, 0, cout, cout, 0, 0,
                       TBN_F32)

    def _p(self, name: str) -> torch.Tensor:
        t = self.sd[self.prefix + name]
        assert t.is_cuda and t.dtype == torch.float32
        return t

    # -- layers ------------------------------------------------------------------------------------------------------
    def _conv(self, sp: ConvSpec, x: _Act, y: Optional[_Act] = None, need_dx: bool = True) -> _Act:
        assert x.c == sp.cin, (sp.name, x.c, sp.cin)
        n, k = x.n, sp.k
        oh, ow = conv_out(x.h, k, sp.stride, sp.pad), conv_out(x.w, k, sp.stride, sp.pad)
        m = n * oh * ow
        w, bias = self._p(sp.name + ".weight"), self._p(sp.name + ".bias")
        gamma, beta = self._p(sp.name + "_bn.weight"), self._p(sp.name + "_bn.bias")
        rm, rv = self._p(sp.name + "_bn.running_mean"), self._p(sp.name + "_bn.running_var")
        u = torch.empty(m, sp.cout, device=self.device, dtype=torch.float32)
        use_tc = self.tc and k in (1, 3) and sp.cin % 16 == 0 and sp.pad == k // 2
        if use_tc:
            self._conv_tc(self._bf16(x), w.detach(), bias.data_ptr(), u, n, x.h, x.w, sp.stride)
        else:
            wk = w.detach().permute(2, 3, 1, 0).reshape(k * k * sp.cin, sp.cout).contiguous()
            self._keep.append(wk)
            self._call("tbn_conv_simt", x.ptr, wk.data_ptr(), bias.data_ptr(), u.data_ptr(), n, x.h, x.w, sp.cin, x.cs, x.co, sp.cout,
                       sp.cout, 0, k, sp.stride, sp.pad, 0, TBN_F32)
        if y is None:
            y = self._new(n, oh, ow, sp.cout)
        assert y.c == sp.cout and (y.h, y.w) == (oh, ow), sp.name
        mean = torch.empty(sp.cout, device=self.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        self._call("tbn_bn_train_fwd", u.data_ptr(), y.ptr, y.cs, y.co, gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                   mean.data_ptr(), rstd.data_ptr(), self._bn_ws.data_ptr(), BN_MOMENTUM, BN_EPS, m, sp.cout)
        self._tape.append(lambda: self._conv_bwd(sp, x, u, y, mean, rstd, w, gamma, need_dx))
        return y

    def _conv_bwd(self, sp: ConvSpec, x: _Act, u: torch.Tensor, y: _Act, mean, rstd, w, gamma, need_dx: bool) -> None:
        n, k, cin, cout = x.n, sp.k, sp.cin, sp.cout
        oh, ow = y.h, y.w
        m = n * oh * ow
        dy = self._g(y)
        dev, f32 = self.device, torch.float32
        du = torch.empty(m, cout, device=dev, dtype=f32)
        dgamma, dbeta = torch.empty(cout, device=dev, dtype=f32), torch.empty(cout, device=dev, dtype=f32)
        self._call("tbn_bn_train_bwd", u.data_ptr(), dy.ptr, dy.cs, dy.co, y.ptr, y.cs, y.co, gamma.data_ptr(), mean.data_ptr(),
                   rstd.data_ptr(), du.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), self._bn_ws.data_ptr(), m, cout)
        # the conv bias sits in front of a batch-statistics BatchNorm: it cancels in (u - mean), its gradient sum(du) is
        # identically zero (the reference's autograd returns rounding noise of ~1e-8 here)
        db = torch.zeros(cout, device=dev, dtype=f32)
        # weight gradient: dWk (k*k*cin, cout) = A^T du, A = the conv's im2col rows
        kk = k * k * cin
        pg = self.param_grads
        pg[self.prefix + sp.name + ".bias"] = db
        pg[self.prefix + sp.name + "_bn.weight"] = dgamma
        pg[self.prefix + sp.name + "_bn.bias"] = dbeta
        if self.tc and cout % 16 == 0 and cin >= 16:  # (the 1 / 3 / 10-channel stem conv keeps the fp32 GEMM: its transposed
            # im2col would move 32-channel tiles that are almost empty)
            pg[self.prefix + sp.name + ".weight"] = self._wgrad_tc(sp, x, du, n, oh, ow).view(k, k, cin, cout).permute(3, 2, 0, 1).contiguous()
            if need_dx:
                self._input_grad(sp, x, du, n, oh, ow, w)
            return
        if k == 1 and sp.stride == 1:
            a_ptr, lda = x.ptr + x.co * 4, x.cs
        else:
            need = m * kk
            if self._cols is None or self._cols.numel() < need:
                self._cols = None
                self._cols = torch.empty(need, device=dev, dtype=f32)
            self._call("tbn_im2col_f32", x.ptr, self._cols.data_ptr(), n, x.h, x.w, cin, x.cs, x.co, k, sp.stride, sp.pad)
            a_ptr, lda = self._cols.data_ptr(), kk
        tiles = ((kk + 127) // 128) * ((cout + 63) // 64)
        ksplit = max(1, min(m // 256, (_SM_TARGET + tiles - 1) // tiles))
        dwk = torch.empty(kk, cout, device=dev, dtype=f32)
        self._call("tbn_sgemm_splitk", 1, 0, kk, cout, m, 1.0, a_ptr, lda, du.data_ptr(), cout, dwk.data_ptr(), cout, ksplit)
        self.launches += 1  # (the zeroing kernel inside)
        pg[self.prefix + sp.name + ".weight"] = dwk.view(k, k, cin, cout).permute(3, 2, 0, 1).contiguous()
        if need_dx:
            self._input_grad(sp, x, du, n, oh, ow, w)

    def _input_grad(self, sp: ConvSpec, x: _Act, du: torch.Tensor, n: int, oh: int, ow: int, w: torch.Tensor) -> None:
        tmp = self._dgrad(sp, du, n, oh, ow, x.h, x.w, w)
        dx = self._g(x)
        self._call("tbn_add_view", dx.ptr, dx.cs, dx.co, tmp.data_ptr(), sp.cin, 0, x.rows, sp.cin)

    def _wgrad_tc(self, sp: ConvSpec, x: _Act, du: torch.Tensor, n: int, oh: int, ow: int) -> torch.Tensor:
        """Weight gradient (k*k*cin, cout) on the tensor cores: a split-K GEMM whose reduction axis is the n*oh*ow output
        positions -- both operands are laid out position-major in bf16 (tbn_im2col_t_bf16), the fp32 partial sums of the
        splits are added by tbn_splitk_reduce."""
        k, cin, cout = sp.k, sp.cin, sp.cout
        kk, m = k * k * cin, n * oh * ow
        tiles = ((kk + 127) // 128) * ((cout + 127) // 128)
        ksplit = max(1, min(64, (2 * 148 + tiles - 1) // tiles, (m + 255) // 256))
        step = 64 * ksplit
        mpad = (m + step - 1) // step * step
        dev = self.device
        need = (kk + cout) * mpad
        if self._colst is None or self._colst.numel() < need:
            self._colst = None
            self._colst = torch.empty(need, device=dev, dtype=torch.bfloat16)
        colst, dut = self._colst[:kk * mpad], self._colst[kk * mpad:kk * mpad + cout * mpad]
        self._call("tbn_im2col_t_bf16", x.ptr, x.cs, x.co, colst.data_ptr(), mpad, n, x.h, x.w, cin, k, sp.stride, sp.pad)
        self._call("tbn_im2col_t_bf16", du.data_ptr(), cout, 0, dut.data_ptr(), mpad, 1, 1, m, cout, 1, 1, 0)
        partial = torch.empty(ksplit, kk, cout, device=dev, dtype=torch.float32)
        self._call("tbn_gemm_tc_splitk", colst.data_ptr(), dut.data_ptr(), partial.data_ptr(), kk, mpad, cout, ksplit)
        dwk = torch.empty(kk, cout, device=dev, dtype=torch.float32)
        self._call("tbn_splitk_reduce", partial.data_ptr(), self._zero_bias.data_ptr(), dwk.data_ptr(), kk, cout, cout, ksplit, 1, 0)
        return dwk

    def _dgrad(self, sp: ConvSpec, du: torch.Tensor, n: int, oh: int, ow: int, xh: int, xw: int, w: torch.Tensor) -> torch.Tensor:
        """Input gradient (n * xh * xw, cin) of a conv from its output gradient du (n * oh * ow, cout): a stride-1 conv of
        the (for stride 2: zero-spread) output gradient with the spatially flipped, channel-transposed filter."""
        k, cin, cout = sp.k, sp.cin, sp.cout
        dev, f32 = self.device, torch.float32
        if sp.stride == 1:
            src, sh, sw = du, oh, ow
        else:
            assert sp.stride == 2
            src, sh, sw = torch.empty(n * xh * xw, cout, device=dev, dtype=f32), xh, xw
            self._call("tbn_upsample_zero2", du.data_ptr(), src.data_ptr(), n, oh, ow, xh, xw, cout)
        tmp = torch.empty(n * xh * xw, cin, device=dev, dtype=f32)
        if self.tc and k in (1, 3) and cout % 16 == 0 and cin % 16 == 0 and sp.pad == k // 2:
            srcb = torch.empty(src.shape, device=dev, dtype=torch.bfloat16)
            self._call("tbn_cast_bf16_view", src.data_ptr(), cout, 0, srcb.data_ptr(), src.shape[0], cout)
            self._conv_tc(srcb, w.detach().flip(2, 3).transpose(0, 1), self._zero_bias.data_ptr(), tmp, n, sh, sw, 1)
        else:
            wd = w.detach().flip(2, 3).permute(2, 3, 0, 1).reshape(k * k * cout, cin).contiguous()
            self._call("tbn_conv_simt", src.data_ptr(), wd.data_ptr(), 0, tmp.data_ptr(), n, sh, sw, cout, cout, 0, cin, cin, 0, k, 1,
                       k - 1 - sp.pad, 0, TBN_F32)
        return tmp

    def _maxpool_s2(self, x: _Act, y: Optional[_Act] = None) -> _Act:
        oh, ow = pool_out_ceil(x.h), pool_out_ceil(x.w)
        if y is None:
            y = self._new(x.n, oh, ow, x.c)
        assert (y.h, y.w, y.c) == (oh, ow, x.c)
        self._call("tbn_maxpool3s2", x.ptr, y.ptr, x.n, x.h, x.w, x.c, x.cs, x.co, y.cs, y.co, TBN_F32)
        self._tape.append(lambda: self._maxpool_bwd(x, y, 1))
        return y

    def _maxpool_bwd(self, x: _Act, y: _Act, stride2: int) -> None:
        dy, dx = self._g(y), self._g(x)
        self._call("tbn_maxpool3_bwd", x.ptr, x.cs, x.co, dy.ptr, dy.cs, dy.co, dx.ptr, dx.cs, dx.co, x.n, x.h, x.w, x.c, stride2)

    def _pool_s1(self, x: _Act, mode: int) -> _Act:
        y = self._new(x.n, x.h, x.w, x.c)
        self._call("tbn_pool3s1", x.ptr, y.ptr, x.n, x.h, x.w, x.c, x.cs, x.co, y.cs, 0, mode, TBN_F32)
        if mode == POOL_MAX:
            self._tape.append(lambda: self._maxpool_bwd(x, y, 0))
        else:
            self._tape.append(lambda: self._avgpool_bwd(x, y))
        return y

    def _avgpool_bwd(self, x: _Act, y: _Act) -> None:
        dy, dx = self._g(y), self._g(x)
        tmp = torch.empty(x.rows, x.c, device=self.device, dtype=torch.float32)
        self._call("tbn_pool3s1", dy.ptr, tmp.data_ptr(), x.n, x.h, x.w, x.c, dy.cs, dy.co, x.c, 0, POOL_AVG, TBN_F32)
        self._call("tbn_add_view", dx.ptr, dx.cs, dx.co, tmp.data_ptr(), x.c, 0, x.rows, x.c)

    # -- the network -------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, feat: torch.Tensor, feat_col: int, feat_ld: int, freq_only: bool, stream: int) -> int:
        """x (F, C, H, W) fp32 -> pooled features into ``feat`` (column block / positions, as BackbonePlan.run)."""
        F, C, H, W = x.shape
        assert C == self.cin and x.dtype == torch.float32 and x.is_contiguous()
        self._stream, self.launches, self._in_backward = stream, 0, False
        self._tape, self._grads, self._keep, self.param_grads, self._bf16_cache = [], {}, [], {}, {}
        xin = self._new(F, H, W, C)
        self._call("tbn_nchw_to_nhwc", x.data_ptr(), xin.ptr, F, C, H, W, C, TBN_F32)
        c1, c2r, c2 = self.stem
        cur = self._maxpool_s2(self._conv(c1, xin, need_dx=False))
        cur = self._maxpool_s2(self._conv(c2, self._conv(c2r, cur)))
        for b in self.blocks:
            off = b.out_offsets
            oh, ow = (conv_out(cur.h, 3, 2, 1), conv_out(cur.w, 3, 2, 1)) if b.reduce else (cur.h, cur.w)
            out = self._new(F, oh, ow, b.cout)
            if b.b1x1:
                self._conv(b.b1x1, cur, out.slice(off["1x1"], b.b1x1.cout))
            r3, c3 = b.b3x3
            self._conv(c3, self._conv(r3, cur), out.slice(off["3x3"], c3.cout))
            rd, d1, d2 = b.bdbl
            self._conv(d2, self._conv(d1, self._conv(rd, cur)), out.slice(off["dbl"], d2.cout))
            if b.reduce:
                self._maxpool_s2(cur, out.slice(off["pool"], b.cin))
            else:
                tp = self._pool_s1(cur, POOL_AVG if b.pool == "avg" else POOL_MAX)
                self._conv(b.bpool, tp, out.slice(off["pool"], b.bpool.cout))
            cur = out
        assert cur.c == FEATURE_SIZE
        self._call("tbn_avgpool_feat", cur.ptr, feat.data_ptr() + feat_col * 4, F, cur.h, cur.w, FEATURE_SIZE, feat_ld,
                   1 if freq_only else 0, TBN_F32)
        self._last = (cur, freq_only)
        self._bf16_cache = {}
        return self.launches

    def backward(self, dfeat_ptr: int, ld: int, stream: int) -> Dict[str, torch.Tensor]:
        """dfeat: gradient of the pooled features ((F, ld) rows, or (F * S, ld) positions).  Returns {state-dict name:
        gradient} for every conv / BatchNorm parameter of this backbone and frees the tape."""
        self._stream, self.launches, self._in_backward = stream, 0, True
        cur, freq_only = self._last
        g = self._g(cur)
        self._call("tbn_avgpool_feat_bwd", dfeat_ptr, ld, g.ptr, cur.n, cur.h, cur.w, FEATURE_SIZE, 1 if freq_only else 0)
        for step in reversed(self._tape):
            step()
        out = self.param_grads
        self._tape, self._grads, self._keep, self.param_grads, self._last = [], {}, [], {}, None
        return out

    def out_spatial(self, H: int, W: int):
        h, w = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
        for _ in range(2):
            h, w = pool_out_ceil(h), pool_out_ceil(w)
        for b in self.blocks:
            if b.reduce:
                h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        return h, w

    def backward_skip(self) -> None:
        """Drop the tape of a forward whose output received no gradient."""
        self._tape, self._grads, self._keep, self.param_grads, self._last = [], {}, [], {}, None
        self.launches = 0

    def parameter_names(self) -> List[str]:
        names = []
        convs = list(self.stem)
        for b in self.blocks:
            convs += ([b.b1x1] if b.b1x1 else []) + list(b.b3x3) + list(b.bdbl) + ([b.bpool] if b.bpool else [])
        for sp in convs:
            names += [self.prefix + sp.name + s for s in (".weight", ".bias", "_bn.weight", "_bn.bias")]
        return names
