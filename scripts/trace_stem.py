"""Phase breakdown of the fused stem: clock64 stamps of CTA 0 over its first tiles (tbn_debug_trace_stem)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import _lib  # noqa: E402
from attention_based_tbn_b200.plan import pack_stem_rows  # noqa: E402

_lib.load()
names = ["tile start", "staged", "row0 built", "row1 built", "mma0 done", "drain0 done", "row2 built", "mma1 done", "drain1 done",
         "(no build)", "mma2 done", "drain2 done", "pooled", "gemm2 done", "epilogue+store"]
for label, n, cin, h, w in (("rgb 32f", 32, 3, 224, 224), ("aud 32f", 32, 1, 256, 421)):
    x = torch.randn(n, cin, h, w, device="cuda")
    w1p = pack_stem_rows(torch.randn(64, cin, 7, 7) / 7, torch.randn(64)).cuda()
    w2p = (torch.randn(64, 64) / 8).to(torch.float16).cuda()
    b2 = torch.randn(64, device="cuda")
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    ph, pw = -(-(oh - 3) // 2) + 1, -(-(ow - 3) // 2) + 1
    y = torch.empty(n, ph, pw, 64, device="cuda", dtype=torch.bfloat16)
    buf = torch.zeros(8 * 16 + 2048, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    call = lambda: _lib.call("tbn_stem_fused_tc", x.data_ptr(), _lib.TBN_F32, w1p.data_ptr(), w2p.data_ptr(), b2.data_ptr(), y.data_ptr(),
                             n, cin, h, w, 64, 0, s)
    call()
    _lib.call("tbn_debug_trace_stem", buf.data_ptr())
    call()
    torch.cuda.synchronize()
    _lib.call("tbn_debug_trace_stem", 0)
    full = buf.cpu()
    t = full[:128].view(8, 16)
    se = full[128:].view(1024, 2)
    se = se[se[:, 0] > 0]
    t0 = int(se[:, 0].min())
    starts = sorted((se[:, 0] - t0).tolist())
    ends = sorted((se[:, 1] - t0).tolist())
    dur = sorted((se[:, 1] - se[:, 0]).tolist())
    print(f"== {label}: cycles since the tile start, CTA 0, tiles 1..4 (tile 0 includes the weight load)")
    for i in range(1, 15):
        print(f"  {names[i]:16s} " + " ".join(f"{int(t[k, i] - t[k, 0]):7d}" for k in range(1, 5) if t[k, i] > 0))
    print("  next tile start  " + " ".join(f"{int(t[k + 1, 0] - t[k, 0]):7d}" for k in range(1, 5)))
    print("  tile 0 (with the weight load): staged %d, row0 built %d, pooled %d, end %d cycles; tiles 0..6 start at %s" % (
        int(t[0, 1] - t[0, 0]), int(t[0, 2] - t[0, 0]), int(t[0, 12] - t[0, 0]), int(t[1, 0] - t[0, 0]),
        [int(t[k, 0] - t[0, 0]) for k in range(8) if t[k, 0] > 0]))
    q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))] / 1e3  # noqa: E731
    print(f"  {len(dur)} CTAs: start min/median/max {q(starts, 0):.1f}/{q(starts, .5):.1f}/{q(starts, 1):.1f} us, end min/median/max "
          f"{q(ends, 0):.1f}/{q(ends, .5):.1f}/{q(ends, 1):.1f} us, CTA lifetime min/median/max {q(dur, 0):.1f}/{q(dur, .5):.1f}/{q(dur, 1):.1f} us")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    print(f"  launch: {e0.elapsed_time(e1) * 100:.1f} us = {e0.elapsed_time(e1) * 100 * 1.92e3:.0f} cycles at 1.92 GHz")
