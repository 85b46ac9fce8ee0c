"""Micro-benchmark of tbn_conv_tc on the layer shapes of the bench workload.  Every measurement replays a CUDA
graph of `reps` back-to-back launches (no host overhead in the number; operands L2-warm as in the real plan).
With --ablate, each shape is also timed with parts of the kernel switched off (tbn_set_tuning("tc_debug"))."""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ablate", type=int, default=0)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()

_lib.load()
DEV = "cuda"
# name, n, h, w, cin, cout, k
SHAPES = [
    # chunked phase A (32 frames per launch)
    ("rgb conv2r 1x1 64->64 (32f)", 32, 56, 56, 64, 64, 1),
    ("rgb conv2_3x3 (32f)", 32, 56, 56, 64, 192, 3),
    ("aud conv2_3x3 (32f)", 32, 64, 105, 64, 192, 3),
    # phase B: all 192 frames of the step
    ("rgb 3a fused1x1 192->224", 192, 28, 28, 192, 224, 1),
    ("rgb 3a 3x3 64->64", 192, 28, 28, 64, 64, 3),
    ("rgb 3b dbl 96->96", 192, 28, 28, 96, 96, 3),
    ("aud 3a fused1x1 192->224", 192, 32, 52, 192, 224, 1),
    ("aud 3b dbl 96->96", 192, 32, 52, 96, 96, 3),
    ("rgb 4a fused1x1 576->512", 192, 14, 14, 576, 512, 1),
    ("rgb 4c 3x3 128->160", 192, 14, 14, 128, 160, 3),
    ("rgb 4e 3x3 192->256", 192, 14, 14, 192, 256, 3),
    ("aud 4a fused1x1 576->512", 192, 16, 26, 576, 512, 1),
    ("aud 4c 3x3 128->160", 192, 16, 26, 128, 160, 3),
    ("rgb 5a fused1x1 1056->832", 192, 7, 7, 1056, 832, 1),
    ("rgb 5a 3x3 192->320", 192, 7, 7, 192, 320, 3),
    ("aud 5a fused1x1 1056->832", 192, 8, 13, 1056, 832, 1),
    ("aud 5b 3x3 224->224", 192, 8, 13, 224, 224, 3),
]
# name, halo, debug bits, cta2 policy
MODES = [("default", 1, 0, 3), ("streamed weights", 1, 512, 3), ("2 epi teams", 1, 128, 3), ("no pairs", 1, 0, 0), ("tap", 0, 0, 3)]
if args.ablate:
    MODES = [("default", 1, 0, 0), ("nostore", 1, 1, 0), ("noepi", 1, 2, 0), ("nomma", 1, 4, 0), ("tma-only", 1, 6, 0),
             ("mma+epi", 1, 24, 0), ("mma-only", 1, 26, 0)]
if args.ablate == 2:  # CTA pairs (cta_group::2 for N tiles wider than 128 columns) against independent CTAs
    MODES = [("default", 1, 0, 0), ("cta2", 1, 0, 1), ("noepi", 1, 2, 0), ("cta2 noepi", 1, 2, 1), ("tma-only", 1, 6, 0),
             ("cta2 tma-only", 1, 6, 1), ("mma-only", 1, 26, 0), ("cta2 mma-only", 1, 26, 1)]


def graph_time(fn, reps, iters=5):
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


tot = {m[0]: 0.0 for m in MODES}
tot_f = 0.0
print(f"{'shape':30s}" + "".join(f"{m[0]:>14s}" for m in MODES) + "   (us per launch)   TF/s of the first two")
for name, n, h, w, cin, cout, k in SHAPES:
    cp = (cin + 63) // 64 * 64
    x = torch.randn(n, h, w, cin, device=DEV).to(torch.bfloat16)
    wp = (torch.randn(cout, k * k * cp, device=DEV) / math.sqrt(cin * k * k)).to(torch.bfloat16)
    b = torch.zeros(cout, device=DEV)
    y = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    fl = 2.0 * n * h * w * cin * cout * k * k
    tot_f += fl
    line = f"{name:30s}"
    res = []
    for mname, halo, dbg, cta2 in MODES:
        _lib.call("tbn_set_tuning", b"tc_halo", halo)
        _lib.call("tbn_set_tuning", b"tc_cta2", cta2)
        _lib.call("tbn_set_tuning", b"tc_debug", dbg & 63)
        _lib.call("tbn_set_tuning", b"tc_occ2", 0 if dbg & 64 else 1)
        _lib.call("tbn_set_tuning", b"tc_teams3", 0 if dbg & 128 else 1)
        _lib.call("tbn_set_tuning", b"tc_bres", 0 if dbg & 512 else 2)
        us = graph_time(lambda s: _lib.call("tbn_conv_tc", x.data_ptr(), wp.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, w, cin,
                                            cin, 0, cout, cout, 0, k, 1, _lib.TBN_BF16, 1, s), args.reps)
        tot[mname] += us
        res.append(us)
        line += f"{us:14.1f}"
    _lib.call("tbn_set_tuning", b"tc_debug", 0)
    _lib.call("tbn_set_tuning", b"tc_occ2", 1)
    _lib.call("tbn_set_tuning", b"tc_teams3", 1)
    _lib.call("tbn_set_tuning", b"tc_bres", 2)
    _lib.call("tbn_set_tuning", b"tc_halo", 1)
    _lib.call("tbn_set_tuning", b"tc_cta2", 3)
    print(line + f"   {fl / 1e6:9.0f} MF  {fl / res[0] / 1e6:6.0f} {fl / res[1] / 1e6:6.0f}")
print(f"{'TOTAL':30s}" + "".join(f"{tot[m[0]]:14.1f}" for m in MODES))
print(f"{'TOTAL TF/s':30s}" + "".join(f"{tot_f / tot[m[0]] / 1e6:14.1f}" for m in MODES))
