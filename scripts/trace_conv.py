"""Role-level timeline of the persistent tcgen05 conv kernel (clock64 stamps written by the kernel itself)
plus CUDA-graph timing at several problem sizes (slope = throughput, intercept = fixed launch cost)."""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import _lib  # noqa: E402

lib = _lib.load()
DEV = "cuda"
SHAPES = [
    ("rgb conv2_3x3", 32, 56, 56, 64, 192, 3),
]


def mk(n, h, w, cin, cout, k):
    cp = (cin + 63) // 64 * 64
    x = torch.randn(n, h, w, cin, device=DEV).to(torch.bfloat16)
    wp = (torch.randn(cout, k * k * cp, device=DEV) / math.sqrt(cin * k * k)).to(torch.bfloat16)
    b = torch.zeros(cout, device=DEV)
    y = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    return x, wp, b, y


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


ctas, stride, tn = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
_lib.call("tbn_debug_trace", 0, ctypes.addressof(ctas), ctypes.addressof(stride), ctypes.addressof(tn))
ctas, stride, tn = ctas.value, stride.value, tn.value
buf = torch.zeros(ctas * stride, dtype=torch.int64, device=DEV)

for name, n, h, w, cin, cout, k in SHAPES:
    print(f"==== {name}: n={n} {h}x{w} {cin}->{cout} k{k}")
    # --- graph timing vs problem size
    for pers in (0, 1):
        _lib.call("tbn_set_tuning", b"tc_bres", pers)
        row = []
        for mult in (1, 2, 4):
            x, wp, b, y = mk(n * mult, h, w, cin, cout, k)
            us = graph_time(lambda s: _lib.call("tbn_conv_tc", x.data_ptr(), wp.data_ptr(), b.data_ptr(), y.data_ptr(), n * mult, h, w,
                                                cin, cin, 0, cout, cout, 0, k, 1, _lib.TBN_BF16, 1, s))
            fl = 2.0 * n * mult * h * w * cin * cout * k * k
            row.append(f"x{mult}: {us:7.1f} us {fl / us / 1e6:6.0f} TF/s")
        print(f"  graph-timed bres={pers}: " + " | ".join(row))
    # --- trace (persistent)
    _lib.call("tbn_set_tuning", b"tc_halo", 1)
    x, wp, b, y = mk(n, h, w, cin, cout, k)
    s = torch.cuda.current_stream().cuda_stream
    for dbg, c2 in ((0, 0), (0, 1), (2, 1)):
        _lib.call("tbn_set_tuning", b"tc_debug", dbg)
        _lib.call("tbn_set_tuning", b"tc_bres", c2)
        for _ in range(2):
            _lib.call("tbn_conv_tc", x.data_ptr(), wp.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, w, cin, cin, 0, cout, cout, 0, k, 1, _lib.TBN_BF16, 1, s)
        torch.cuda.synchronize()
        buf.zero_()
        _lib.call("tbn_debug_trace", buf.data_ptr(), 0, 0, 0)
        _lib.call("tbn_conv_tc", x.data_ptr(), wp.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, w, cin, cin, 0, cout, cout, 0, k, 1, _lib.TBN_BF16, 1, s)
        torch.cuda.synchronize()
        _lib.call("tbn_debug_trace", 0, 0, 0, 0)
        t = buf.cpu().view(ctas, stride)
        for c in range(2):
            r = t[c]
            t0, tend = int(r[0]), int(r[1])
            arr = [r[8 + i * tn: 8 + (i + 1) * tn].tolist() for i in range(4)]
            nkb = sum(1 for v in arr[1] if v)
            print(f"  [dbg={dbg} bres={c2}] CTA {c}: total {tend - t0} cycles, k-blocks {nkb}")
            rel = lambda v: (v - t0) if v else -1  # noqa: E731
            print("    kb: producer-issue  mma-full  mma-commit (cycles since CTA start)")
            for kb in list(range(0, min(nkb, 22))) + list(range(max(22, nkb - 4), nkb)):
                print(f"    {kb:3d} {rel(arr[0][kb]):9d} {rel(arr[1][kb]):9d} {rel(arr[2][kb]):9d}")
            ep = [(rel(arr[3][2 * i]), rel(arr[3][2 * i + 1])) for i in range(tn // 2) if arr[3][2 * i]]
            print("    epilogue (start,end) per tile:", ep[:10])
    _lib.call("tbn_set_tuning", b"tc_debug", 0)
