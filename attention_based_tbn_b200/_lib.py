"""ctypes binding of libtbn_sm100.so (the C ABI declared in include/tbn_sm100.h).

There is no CPU or PyTorch fallback: if the shared library is missing or does not export a
symbol, importing/using the kernels raises.  Pointers come from ``tensor.data_ptr()``; the stream
is ``torch.cuda.current_stream().cuda_stream``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtbn_sm100.so")

TBN_F32, TBN_BF16 = 0, 1
POOL_AVG, POOL_MAX = 0, 1

_I, _P, _LL, _F = c_int, c_void_p, c_longlong, c_float

# name -> argtypes (all return int except tbn_last_error); keep in sync with include/tbn_sm100.h
SIGNATURES = {
    "tbn_version": [],
    "tbn_device_info": [_P, _P, _P],
    "tbn_set_tuning": [c_char_p, _I],
    "tbn_debug_trace": [_P, _P, _P, _P],
    "tbn_debug_stamp": [_P, _P],
    "tbn_debug_trace_stem": [_P],
    "tbn_nchw_to_nhwc": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "tbn_conv_simt": [_P, _P, _P, _P] + [_I] * 14 + [_P],
    "tbn_conv_tc": [_P, _P, _P, _P] + [_I] * 13 + [_P],
    "tbn_conv_tc_s2": [_P, _P, _P, _P] + [_I] * 11 + [_P],
    "tbn_conv_tc_multi": [_P, _P, _P, _P] + [_I] * 9 + [_P],
    "tbn_conv_tc_group": [_P, _I, _P],
    "tbn_stem_fused_tc": [_P, _I, _P, _P, _P, _P] + [_I] * 6 + [_P],
    "tbn_avgpool3s1_bias_relu": [_P, _P, _P] + [_I] * 9 + [_P],
    "tbn_im2col": [_P, _P] + [_I] * 12 + [_P],
    "tbn_maxpool3s2": [_P, _P] + [_I] * 9 + [_P],
    "tbn_pool3s1": [_P, _P] + [_I] * 10 + [_P],
    "tbn_avgpool_feat": [_P, _P] + [_I] * 7 + [_P],
    "tbn_avgpool_feat_split": [_P, _P, _P, _LL] + [_I] * 7 + [_P],
    "tbn_linear": [_P, _P, _P, _P] + [_I] * 8 + [_P],
    "tbn_split_bf16x3": [_P, _P, _I, _I, _I, _I, _P],
    "tbn_split_bf16x3_ld": [_P, _P, _LL, _I, _I, _I, _I, _P],
    "tbn_gemm_tc_splitk": [_P, _P, _P, _I, _I, _I, _I, _P],
    "tbn_splitk_reduce": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "tbn_splitk_reduce_ex": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _LL, _P],
    "tbn_gemm_tc_splitk_ld": [_P, _I, _P, _P, _I, _I, _I, _I, _P],
    "tbn_linear_grouped": [_P, _P, _P, _P] + [_I] * 7 + [_LL] * 4 + [_I, _P],
    "tbn_groupnorm": [_P, _P, _P, _I, _I, _I, _I, _F, _P],
    "tbn_bn_train_relu": [_P, _P, _P, _P, _P, _P, _F, _F, _LL, _I, _I, _I, _I, _P],
    "tbn_mha_core": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "tbn_mha_core_train": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "tbn_sgemm": [_I, _I, _I, _I, _I, _F, _P, _LL, _P, _LL, _F, _P, _LL, _P],
    "tbn_colsum": [_P, _LL, _I, _I, _F, _P, _P],
    "tbn_consensus_bwd": [_P, _LL, _P, _I, _I, _I, _P],
    "tbn_relu_dropout_bwd": [_P, _P, _F, _LL, _P],
    "tbn_value_bias_fix": [_P, _P, _P, _I, _I, _I, _P],
    "tbn_groupnorm_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "tbn_mha_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "tbn_weighted_sum_bwd": [_P, _P, _LL, _P, _P, _I, _I, _I, _P],
    "tbn_softmax_bwd": [_P, _P, _P, _I, _I, _LL, _P],
    "tbn_weighted_sum_bwd_a": [_P, _P, _LL, _P, _I, _I, _I, _P],
    "tbn_bn_train_fwd": [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _F, _F, _LL, _I, _P],
    "tbn_bn_train_bwd": [_P, _P, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _P],
    "tbn_im2col_f32": [_P, _P] + [_I] * 9 + [_P],
    "tbn_sgemm_splitk": [_I, _I, _I, _I, _LL, _F, _P, _LL, _P, _LL, _P, _LL, _I, _P],
    "tbn_upsample_zero2": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "tbn_maxpool3_bwd": [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "tbn_add_view": [_P, _I, _I, _P, _I, _I, _LL, _I, _P],
    "tbn_avgpool_feat_bwd": [_P, _LL, _P, _I, _I, _I, _I, _I, _P],
    "tbn_cast_bf16_view": [_P, _I, _I, _P, _LL, _I, _P],
    "tbn_im2col_t_bf16": [_P, _I, _I, _P, _LL, _I, _I, _I, _I, _I, _I, _I, _P],
    "tbn_mha_core_gn": [_P, _P, _P, _P, _I, _F, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "tbn_attn_weighted_sum": [_P, _P, _P, _I, _I, _I, _I, _P],
    "tbn_repeat_rows": [_P, _I, _I, _I, _I, _P],
    "tbn_softmax_rows": [_P, _P, _I, _I, _P],
    "tbn_consensus": [_P, _P, _I, _I, _I, _P],
    "tbn_topk_confusion": [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "tbn_stft_frames": [_P, _LL, _I, _I, _I, _I, _I, _P, _P],
    "tbn_stft_logpower": [_P, _P, _I, _I, _I, _F, _P],
}



class OutSeg(ctypes.Structure):
    """``tbn_out_seg`` of include/tbn_sm100.h."""

    _fields_ = [("y", c_void_p), ("y_cs", c_int), ("y_co", c_int), ("cols", c_int), ("flags", c_int)]


SEG_RELU, SEG_BIAS = 1, 2


class ConvDesc(ctypes.Structure):
    """``tbn_conv_desc`` of include/tbn_sm100.h (one problem of a grouped launch)."""

    _fields_ = [("x", c_void_p), ("wp", c_void_p), ("bias", c_void_p), ("segs", c_void_p), ("nseg", c_int),
                ("n", c_int), ("h", c_int), ("w", c_int), ("cin", c_int), ("x_cs", c_int), ("x_co", c_int),
                ("k", c_int), ("stride", c_int), ("out_dtype", c_int)]


def make_descs(items):
    """[dict(x, wp, bias, segs=[(ptr, y_cs, y_co, cols[, flags]), ...], n, h, w, cin, x_cs, x_co, k, stride, out_dtype), ...]
    -> (ctypes array, address, keepalive list).  Keep the returned objects alive while the array is in use."""
    keep, descs = [], []
    for it in items:
        arr, addr = make_segs(it["segs"])
        keep.append(arr)
        descs.append(ConvDesc(int(it["x"]), int(it["wp"]), int(it["bias"]), addr, len(it["segs"]), int(it["n"]), int(it["h"]),
                              int(it["w"]), int(it["cin"]), int(it["x_cs"]), int(it["x_co"]), int(it["k"]), int(it["stride"]),
                              int(it["out_dtype"])))
    darr = (ConvDesc * len(descs))(*descs)
    keep.append(darr)
    return darr, ctypes.addressof(darr), keep


def make_segs(segs):
    """[(ptr, y_cs, y_co, cols[, flags]), ...] -> (ctypes array, address); keep the array alive while in use.
    flags default to bias + ReLU."""
    items = []
    for sg in segs:
        p, cs, co, c = sg[:4]
        fl = sg[4] if len(sg) > 4 else (SEG_RELU | SEG_BIAS)
        items.append(OutSeg(int(p), int(cs), int(co), int(c), int(fl)))
    arr = (OutSeg * len(items))(*items)
    return arr, ctypes.addressof(arr)


_lib = None


class TbnKernelError(RuntimeError):
    pass


def load():
    """Load the library once; raises if it is absent (build with ``python -m attention_based_tbn_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TbnKernelError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m attention_based_tbn_b200.build`."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.tbn_last_error.argtypes = []
    lib.tbn_last_error.restype = c_char_p
    _lib = lib
    # experiment switches: TBN_TUNE="key=value,key=value" -> tbn_set_tuning (see csrc/conv_tc.cu for the keys)
    for kv in filter(None, os.environ.get("TBN_TUNE", "").split(",")):
        k, v = kv.split("=")
        if lib.tbn_set_tuning(k.strip().encode(), int(v)) != 0:
            raise TbnKernelError(f"TBN_TUNE: {lib.tbn_last_error().decode()}")
    return lib


def last_error() -> str:
    return load().tbn_last_error().decode("utf-8", "replace")


def call(name: str, *args) -> None:
    """Invoke an entry point; a non-zero status raises with the library's message."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise TbnKernelError(f"{name} failed (status {rc}): {last_error()}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
