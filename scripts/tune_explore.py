"""Exhaustive per-layer sweep of the conv kernel's tiling knobs on the bench workload (design aid).  Every distinct conv
launch of one eager forward (per-modality plan) is timed, as a CUDA-graph replay on its real buffers, for all
combinations of tc_max_bn x tc_occ2 x tc_cta2 x tc_halo; prints default vs best and the summed potential."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import TBNModel, _lib  # noqa: E402
from attention_based_tbn_b200 import plan as plan_mod  # noqa: E402
from attention_based_tbn_b200.config import baseline_config, get_modality  # noqa: E402
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

_lib.load()
cfg = baseline_config(2)
mod = get_modality(cfg)
sd = make_state_dict(cfg, mod, seed=0, calib=os.path.join(ROOT, "tests", "golden", "bn_calib.npz"))
model = TBNModel(cfg, mod, "cuda", precision="bf16", use_cuda_graph=False, multi_stream=False, group_launches=False)
model.load_state_dict(sd)
model.eval()
inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, 64, 3, seed=1).items()}
ops = []


def hook(name, a, flops, launch):
    launch()
    if name in ("tbn_conv_tc", "tbn_conv_tc_s2", "tbn_conv_tc_multi") and flops > 0:
        ops.append((name, a, flops))


with torch.no_grad():
    model(inp)
    torch.cuda.synchronize()
    plan_mod.LAUNCH_HOOK = hook
    model(inp)
    plan_mod.LAUNCH_HOOK = None
torch.cuda.synchronize()
st = torch.cuda.Stream()
DEF = dict(tc_max_bn=256, tc_occ2=1, tc_cta2=3, tc_halo=1)
seen = {}
tot_def = tot_best = 0.0
for name, a, flops in ops:
    ints = tuple(x for x in a[:-1] if isinstance(x, int) and abs(x) < (1 << 20))
    key = (name,) + ints
    if key in seen:
        d, b = seen[key]
        tot_def += d
        tot_best += b
        continue
    args = list(a[:-1])
    res = {}
    k3s1 = name == "tbn_conv_tc" and ints[9] == 3
    for bn, occ2, cta2, halo in itertools.product((256, 192, 160, 128, 112, 96, 80, 64), (1, 0), (3, 0, 1, 2), ((1, 0) if k3s1 else (1,))):
        cand = dict(tc_max_bn=bn, tc_occ2=occ2, tc_cta2=cta2, tc_halo=halo)
        plan_mod._apply_tuning(cand)
        try:
            res[(bn, occ2, cta2, halo)] = plan_mod._time_launch(name, args, st) * 1e3
        except Exception:
            pass
    plan_mod._apply_tuning(DEF)
    d = res[(256, 1, 3, 1)]
    bk = min(res, key=res.get)
    b = res[bk]
    seen[key] = (d, b)
    tot_def += d
    tot_best += b
    desc = f"{name[4:]:14s} " + " ".join(str(x) for x in ints[:10])
    print(f"{desc:64s} default {d:6.1f} us  best {b:6.1f} us ({100 * (1 - b / d):4.1f} %)  bn/occ2/cta2/halo = {bk}", flush=True)
print(f"TOTAL default {tot_def:.1f} us  best-per-layer {tot_best:.1f} us  ({100 * (1 - tot_best / tot_def):.1f} %)")
