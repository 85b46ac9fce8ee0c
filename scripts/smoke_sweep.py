import os, sys, torch
sys.path.insert(0, "/root/repo")
from attention_based_tbn_b200 import TBNModel, _lib
from attention_based_tbn_b200.config import baseline_config, get_modality
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict
from oracle import tbn_oracle
cfg = baseline_config(2); mod = get_modality(cfg)
sd = make_state_dict(cfg, mod, seed=0, calib="/root/repo/tests/golden/bn_calib.npz")
model = TBNModel(cfg, mod, "cuda:0", precision="bf16", use_cuda_graph=False); model.load_state_dict(sd); model.eval()
m32 = TBNModel(cfg, mod, "cuda:0", precision="fp32", use_cuda_graph=False); m32.load_state_dict(sd); m32.eval()
for seed in (9, 10, 11, 12):
    inp = make_inputs(cfg, mod, 1, 3, seed=seed)
    ref = tbn_oracle.tbn_forward(sd, cfg, mod, inp)
    d = {k: v.cuda() for k, v in inp.items()}
    with torch.no_grad():
        out = model(d); o32 = m32(d)
    e = {k: ((out[k].float().cpu() - r).abs().max() / r.abs().max()).item() for k, r in ref.items()}
    e32 = {k: ((o32[k].float().cpu() - r).abs().max() / r.abs().max()).item() for k, r in ref.items()}
    print(seed, {k: f"{v:.3e}" for k, v in e.items()}, "fp32:", {k: f"{v:.1e}" for k, v in e32.items()}, "wmax", ref["weights"].max().item())
