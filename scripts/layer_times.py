"""Per-launch times of one bench-workload forward: every kernel launch of the eager single-stream plan is repeated
`reps` times back to back between two CUDA events (launches are idempotent: same inputs, outputs overwritten), so the
number is the kernel's L2-warm duration without launch gaps.  Prints one line per launch and totals per entry point.

    python scripts/layer_times.py [--cfg 2] [--clips 64] [--reps 5] [--tune key=value ...]
"""
import argparse
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import TBNModel, _lib  # noqa: E402
from attention_based_tbn_b200 import plan as plan_mod  # noqa: E402
from attention_based_tbn_b200.config import baseline_config, get_modality  # noqa: E402
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--segments", type=int, default=3)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tune", nargs="*", default=[])
ap.add_argument("--quiet", type=int, default=0)
args = ap.parse_args()

_lib.load()
for kv in args.tune:
    k, v = kv.split("=")
    _lib.call("tbn_set_tuning", k.encode(), int(v))
cfg = baseline_config(args.cfg)
mod = get_modality(cfg)
sd = make_state_dict(cfg, mod, seed=0, calib=os.path.join(ROOT, "tests", "golden", "bn_calib.npz"))
model = TBNModel(cfg, mod, "cuda", precision="bf16", use_cuda_graph=False, multi_stream=False)
model.load_state_dict(sd)
model.eval()
inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, args.clips, args.segments, seed=1).items()}
rows = []


def hook(name, a, flops, launch):
    launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        launch()
    e1.record()
    rows.append((name, a, flops, e0, e1))


with torch.no_grad():
    model(inp)
    torch.cuda.synchronize()
    plan_mod.LAUNCH_HOOK = hook
    model(inp)
    plan_mod.LAUNCH_HOOK = None
torch.cuda.synchronize()


def describe(name, a):
    ints = [x for x in a if isinstance(x, int) and abs(x) < (1 << 24)]
    if name == "tbn_conv_tc":       # n h w cin x_cs x_co cout y_cs y_co k
        return f"n{ints[0]} {ints[1]}x{ints[2]} {ints[3]}->{ints[6]} k{ints[9]}"
    if name == "tbn_conv_tc_s2":
        return f"n{ints[0]} {ints[1]}x{ints[2]} {ints[3]}->{ints[6]} k3s2"
    if name == "tbn_conv_tc_multi":  # nseg n h w cin
        return f"n{ints[1]} {ints[2]}x{ints[3]} cin{ints[4]} segs{ints[0]}"
    return " ".join(str(x) for x in ints[:6])


tot = defaultdict(float)
totf = defaultdict(float)
for name, a, flops, e0, e1 in rows:
    us = e0.elapsed_time(e1) * 1e3 / args.reps
    tot[name] += us
    totf[name] += flops
    if not args.quiet:
        print(f"{name:28s} {describe(name, a):34s} {us:8.1f} us" + (f" {flops / us / 1e6:7.0f} TF/s" if flops else ""))
print("---- totals per entry point (us, TF/s)")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:28s} {tot[k]:9.1f}" + (f" {totf[k] / tot[k] / 1e6:7.0f}" if totf[k] else ""))
conv = [k for k in tot if k.startswith("tbn_conv_tc")]
cu, cf = sum(tot[k] for k in conv), sum(totf[k] for k in conv)
print(f"conv_tc_kernel all: {cu:.1f} us, {cf / cu / 1e6:.0f} TF/s;  everything: {sum(tot.values()):.1f} us")
