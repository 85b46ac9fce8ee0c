"""Device-side timeline of one CUDA-graph replay of the bench workload: every kernel launch of the plan is bracketed
by one-thread %globaltimer stamps on its own stream (captured into the graph with it).  Prints per-stream busy
time, the busiest kernels and an ASCII occupancy strip, so that overlap / serialisation between the stream lanes is
visible without nsys."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import TBNModel, _lib  # noqa: E402
from attention_based_tbn_b200 import plan as plan_mod  # noqa: E402
from attention_based_tbn_b200.config import baseline_config, get_modality  # noqa: E402
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--rows", type=int, default=0, help="print the first N launches")
args = ap.parse_args()

_lib.load()
cfg = baseline_config(args.cfg)
mod = get_modality(cfg)
sd = make_state_dict(cfg, mod, seed=0, calib=os.path.join(ROOT, "tests", "golden", "bn_calib.npz"))
model = TBNModel(cfg, mod, "cuda", precision="bf16", use_cuda_graph=True)
model.load_state_dict(sd)
model.eval()
inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, args.clips, 3, seed=1).items()}

slots = torch.zeros(4096, dtype=torch.int64, device="cuda")
recs = []  # (name, stream, slot index)
state = {"on": False}


def hook(name, a, flops, launch):
    if not state["on"]:
        return launch()
    stream = a[-1]
    i = len(recs)
    _lib.call("tbn_debug_stamp", slots.data_ptr() + 16 * i, stream)
    launch()
    _lib.call("tbn_debug_stamp", slots.data_ptr() + 16 * i + 8, stream)
    recs.append((name, stream, a))


plan_mod.LAUNCH_HOOK = hook
with torch.no_grad():
    # the model warms up once (uncaptured) and then captures: record only during the capture pass
    orig = model._forward_impl

    def wrapped(t, s):
        state["on"] = torch.cuda.is_current_stream_capturing()
        if state["on"]:
            recs.clear()
        try:
            return orig(t, s)
        finally:
            state["on"] = False

    model._forward_impl = wrapped
    for _ in range(3):
        model(inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model(inp)
    e1.record()
    torch.cuda.synchronize()
t = slots.cpu().view(-1, 2)[:len(recs)]
t0 = int(t[:, 0].min())
span = (int(t[:, 1].max()) - t0) / 1e3
print(f"launches {len(recs)}  graph replay (with stamps) {e0.elapsed_time(e1) * 1e3:.0f} us  stamped span {span:.0f} us")
streams = collections.OrderedDict()
for (name, stream, a), (s, e) in zip(recs, t.tolist()):
    streams.setdefault(stream, []).append((name, (s - t0) / 1e3, (e - t0) / 1e3))
W = 120
for si, (stream, items) in enumerate(streams.items()):
    busy = sum(e - s for _, s, e in items)
    strip = [" "] * W
    for name, s, e in items:
        for x in range(int(s / span * W), min(W, int(e / span * W) + 1)):
            strip[x] = "s" if "stem" in name else ("p" if "pool" in name else ("c" if "conv" in name else "h"))
    print(f"stream {si}: {len(items):3d} launches busy {busy:7.0f} us ({100 * busy / span:4.1f}%) |{''.join(strip)}|")
agg = collections.defaultdict(float)
for name, stream, a in recs:
    pass
for (name, stream, a), (s, e) in zip(recs, t.tolist()):
    agg[name] += (e - s) / 1e3
print("in-graph time by entry point (overlapping lanes double count):")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {v:8.0f} us")
if args.rows:
    for (name, stream, a), (s, e) in list(zip(recs, t.tolist()))[:args.rows]:
        print(f"  {list(streams).index(stream)} {name:26s} {(s - t0) / 1e3:8.1f} -> {(e - t0) / 1e3:8.1f}  ({(e - s) / 1e3:6.1f})")
