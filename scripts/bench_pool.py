"""Micro-benchmark of the pooling kernels on the bench workload's shapes (graph-timed, L2-warm like in the plan)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import _lib  # noqa: E402

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


tot = 0.0
print("maxpool 3x3/2 ceil (bf16 NHWC)")
for name, n, h, w, c, count in (("rgb pool2 (32f)", 32, 56, 56, 192, 6), ("aud pool2 (32f)", 32, 64, 105, 192, 6)):
    x = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
    oh, ow = -(-(h - 3) // 2) + 1, -(-(w - 3) // 2) + 1
    y = torch.empty(n, oh, ow, c, device="cuda", dtype=torch.bfloat16)
    us = graph_time(lambda s: _lib.call("tbn_maxpool3s2", x.data_ptr(), y.data_ptr(), n, h, w, c, c, 0, c, 0, _lib.TBN_BF16, s))
    gb = (x.numel() + y.numel()) * 2 / 1e9
    tot += us * count
    print(f"  {name:22s} {us:7.1f} us  {gb / us * 1e6:7.0f} GB/s  (x{count} per step)")
print("avgpool 3x3/1 + bias + relu (bf16 NHWC, channel slice of the concat buffer)")
for name, n, h, w, c, cs, count in (("rgb 3a proj 32", 192, 28, 28, 32, 256, 1), ("rgb 3b proj 64", 192, 28, 28, 64, 320, 1),
                                    ("rgb 4x proj 128", 192, 14, 14, 128, 576, 4), ("rgb 5a proj 128", 192, 7, 7, 128, 1024, 1),
                                    ("aud 3a proj 32", 192, 32, 52, 32, 256, 1), ("aud 3b proj 64", 192, 32, 52, 64, 320, 1),
                                    ("aud 4x proj 128", 192, 16, 26, 128, 576, 4), ("aud 5a proj 128", 192, 8, 13, 128, 1024, 1)):
    x = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
    b = torch.randn(c, device="cuda")
    y = torch.empty(n, h, w, cs, device="cuda", dtype=torch.bfloat16)
    us = graph_time(lambda s: _lib.call("tbn_avgpool3s1_bias_relu", x.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, w, c, c, 0,
                                        cs, cs - c, _lib.TBN_BF16, s))
    gb = 2 * x.numel() * 2 / 1e9
    tot += us * count
    print(f"  {name:22s} {us:7.1f} us  {gb / us * 1e6:7.0f} GB/s  (x{count} per step)")
print(f"per step: {tot:.0f} us")
