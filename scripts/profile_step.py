"""One eager (no CUDA graph, single stream) forward of the bench workload, for ncu launch lists.

    python scripts/profile_step.py [--clips 64] [--reps 2] [--multi-stream 0]
Prints the number of kernel launches per forward so that ncu can skip the warm-up (-s N -c N)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import TBNModel  # noqa: E402
from attention_based_tbn_b200.config import baseline_config, get_modality  # noqa: E402
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--chunk", type=int, default=32)
ap.add_argument("--multi-stream", type=int, default=0)
ap.add_argument("--fuse", type=int, default=1)
ap.add_argument("--group", type=int, default=1, help="1: grouped launch plan (JointPhaseB), 0: per-modality plan")
ap.add_argument("--segments", type=int, default=3)
args = ap.parse_args()

cfg = baseline_config(args.cfg)
mod = get_modality(cfg)
sd = make_state_dict(cfg, mod, seed=0, calib=os.path.join(ROOT, "tests", "golden", "bn_calib.npz"))
model = TBNModel(cfg, mod, "cuda", precision="bf16", use_cuda_graph=False, chunk=args.chunk,
                 multi_stream=bool(args.multi_stream), fuse_1x1=bool(args.fuse), group_launches=bool(args.group))
model.load_state_dict(sd)
model.eval()
inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, args.clips, args.segments, seed=1).items()}
with torch.no_grad():
    for _ in range(args.reps):
        out = model(inp)
        torch.cuda.synchronize()
print("launches_per_forward", model.last_launches)
