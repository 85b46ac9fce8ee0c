"""Where a full training step spends its time: per entry point, forward / backward, serialised (one sync per launch).
    TBN_TRAIN_PROFILE=1 python scripts/train_step_profile.py [--clips 64]"""
import argparse
import os
import sys

os.environ["TBN_TRAIN_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import TBNModel, backbone_train  # noqa: E402
from attention_based_tbn_b200.config import baseline_config, get_modality  # noqa: E402
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--precision", default="fp32")
args = ap.parse_args()
cfg = baseline_config(2)
cfg.model.attention.decay_step = 1
mod = get_modality(cfg)
model = TBNModel(cfg, mod, "cuda", train_precision=args.precision)
model.load_state_dict(make_state_dict(cfg, mod, seed=0, calib=os.path.join(ROOT, "tests", "golden", "bn_calib.npz")))
model.train()
B = args.clips
inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, B, 3, seed=1).items()}
g = torch.Generator().manual_seed(2)
target = {"class": {k: torch.randint(0, int(v), (B,), generator=g).cuda() for k, v in cfg.model.num_classes.items()}}
crit = {"crossentropy": torch.nn.CrossEntropyLoss()}
for it in range(2):
    backbone_train.PROFILE.clear()
    out = model(inp)
    loss, _ = model.get_loss(crit, target, out, epoch=20)
    loss["total"].backward()
    torch.cuda.synchronize()
tot = sum(backbone_train.PROFILE.values())
print(f"backbone kernels of one training step, {B} clips x 3 segments, precision {args.precision} (serialised): {tot:.1f} ms")
for (name, phase), ms in sorted(backbone_train.PROFILE.items(), key=lambda kv: -kv[1]):
    print(f"  {phase}  {name:28s} {ms:9.2f} ms  {100 * ms / tot:5.1f} %")
