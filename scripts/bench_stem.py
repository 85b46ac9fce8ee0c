"""Micro-benchmark of tbn_stem_fused_tc on the chunked phase-A shapes (32 frames per launch): graph-timed us per launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import _lib  # noqa: E402
from attention_based_tbn_b200.plan import pack_stem_rows  # noqa: E402

_lib.load()


def graph_time(fn, reps=10, iters=5):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(st.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn(st.cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(iters):
            g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


for name, n, cin, h, w in (("rgb 32f", 32, 3, 224, 224), ("aud 32f", 32, 1, 256, 421), ("flow 32f", 32, 10, 224, 224)):
    x = torch.randn(n, cin, h, w, device="cuda")
    w1 = torch.randn(64, cin, 7, 7) / 7
    w1p = pack_stem_rows(w1, torch.randn(64)).cuda()
    w2p = (torch.randn(64, 64) / 8).to(torch.float16).cuda()  # [cout][cin] K-major, fp16 like w1p
    b2 = torch.randn(64, device="cuda")
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    ph, pw = -(-(oh - 3) // 2) + 1, -(-(ow - 3) // 2) + 1
    y = torch.empty(n, ph, pw, 64, device="cuda", dtype=torch.bfloat16)
    us = graph_time(lambda s: _lib.call("tbn_stem_fused_tc", x.data_ptr(), _lib.TBN_F32, w1p.data_ptr(), w2p.data_ptr(), b2.data_ptr(),
                                        y.data_ptr(), n, cin, h, w, 64, 0, s))
    xb = x.to(torch.bfloat16)
    usb = graph_time(lambda s: _lib.call("tbn_stem_fused_tc", xb.data_ptr(), _lib.TBN_BF16, w1p.data_ptr(), w2p.data_ptr(),
                                         b2.data_ptr(), y.data_ptr(), n, cin, h, w, 64, 0, s))
    gf = 2.0 * n * oh * ow * 64 * 49 * cin / 1e9
    print(f"stem_fused {name}: {us:7.1f} us/launch   conv1 {gf / us * 1e3:6.1f} TF/s   input {x.numel() * 4 / us / 1e3:6.1f} GB/s   (bf16 input: {usb:6.1f} us)")
