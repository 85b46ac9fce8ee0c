"""Shared case table of the golden fixtures (mirrors oracle/make_golden.py::CASES)."""
import os

import numpy as np
import torch

from attention_based_tbn_b200.config import baseline_config, default_cfg, get_modality
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
CALIB = os.path.join(GOLDEN, "bn_calib.npz")

CASES = {
    "cfg1_audio_b1n3": (lambda: baseline_config(1), 1, 3),
    "cfg2_mha_b2n3": (lambda: baseline_config(2), 2, 3),
    "cfg3_uniform_b1n3": (lambda: baseline_config(3), 1, 3),
    "cfg3_gaussian_b1n3": (lambda: default_cfg(**{"data.flow.enable": False, "model.attention.use_fixed": True,
                                                  "model.attention.prior_type": "gaussian"}), 1, 3),
    "cfg4_3mod_b1n3": (lambda: baseline_config(4), 1, 3),
    "cfg5_3mod_mha_b1n25": (lambda: baseline_config(5), 1, 25),
    "unimodal_b1n3": (lambda: default_cfg(**{"data.flow.enable": False, "model.attention.type": "unimodal"}), 1, 3),
    "proto_b1n3": (lambda: default_cfg(**{"data.flow.enable": False, "model.attention.type": "proto"}), 1, 3),
}


def load_case(name):
    mk, B, N = CASES[name]
    cfg = mk()
    mod = get_modality(cfg)
    sd = make_state_dict(cfg, mod, seed=0, calib=CALIB)
    inp = make_inputs(cfg, mod, B, N, seed=0)
    with np.load(os.path.join(GOLDEN, f"model_{name}.npz")) as z:
        gold = {k: torch.from_numpy(z[k]) for k in z.files}
    return cfg, mod, sd, inp, gold


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max|a - ref| / max|ref| (the parity metric of BASELINE.md section 4)."""
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
