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
    # 10-crop path (ref model.py:243-248): the visual modality has 2 crops per segment, audio has none -> (B, N, n_audio)
    "crop2_noattn_b2n4": (lambda: default_cfg(**{"data.flow.enable": False, "model.attention.enable": False}), 2, 4, 2),
    "crop2_fixed_b1n6": (lambda: baseline_config(3), 1, 6, 3),
}


def case_inputs(cfg, mod, B, N, n_audio=None, seed=0):
    """Inputs of a case; with ``n_audio`` the audio tensor (and the fixed attention weights) have fewer segments."""
    inp = make_inputs(cfg, mod, B, N, seed=seed)
    if n_audio is not None:
        aud = make_inputs(cfg, mod, B, n_audio, seed=seed + 500)
        inp["Audio"] = aud["Audio"]
        if "weights" in aud:
            inp["weights"] = aud["weights"]
    return inp


def load_case(name):
    case = CASES[name]
    mk, B, N = case[:3]
    cfg = mk()
    mod = get_modality(cfg)
    sd = make_state_dict(cfg, mod, seed=0, calib=CALIB)
    inp = case_inputs(cfg, mod, B, N, case[3] if len(case) > 3 else None)
    with np.load(os.path.join(GOLDEN, f"model_{name}.npz")) as z:
        gold = {k: torch.from_numpy(z[k]) for k in z.files}
    return cfg, mod, sd, inp, gold


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max|a - ref| / max|ref| (the parity metric of BASELINE.md section 4)."""
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


# ---- evaluation-loop fixtures (oracle/make_golden_eval.py ran the reference Metric / save_scores on exactly these) ----
def eval_case(name, device="cpu"):
    """-> (cfg, [(out, target, batch_size, batch_loss), ...]) regenerated from the seed."""
    from collections import OrderedDict

    g = torch.Generator().manual_seed(1234 if name == "two_class" else 4321)
    if name == "two_class":
        cfg = baseline_config(2)  # verb 125 + noun 352, MHA with the entropy term
        sizes = [16, 16, 16, 16, 7]
    else:
        cfg = default_cfg(**{"data.flow.enable": False, "model.attention.enable": False, "val.topk": [1, 3]})
        cfg.model.num_classes = OrderedDict([("verb", 125)])
        sizes = [8, 8, 5]
    loss_keys = list(cfg.model.num_classes.keys())
    if len(loss_keys) > 1:
        loss_keys.append("all_class")
    att = cfg.model.attention
    if att.enable and not att.use_fixed:
        loss_keys += [k for k, on in (("prior", att.use_prior), ("contrast", att.use_contrast), ("entropy", att.use_entropy)) if on]
    loss_keys.append("total")
    batches = []
    for bs in sizes:
        out, tgt = OrderedDict(), {}
        for key, ncls in cfg.model.num_classes.items():
            t = torch.randint(0, ncls, (bs,), generator=g)
            s = torch.randn(bs, ncls, generator=g)
            boost = (torch.rand(bs, generator=g) < 0.6).float() * 2.5  # the target often lands in the top k
            s[torch.arange(bs), t] += boost
            out[key], tgt[key] = s.to(device), t.to(device)
        if att.enable and not att.use_fixed:
            out["weights"] = torch.rand(bs * 3, 1, 13, generator=g).to(device)
        loss = {k: (torch.rand((), generator=g) * 3).to(device) for k in loss_keys}
        batches.append((out, {"class": tgt}, bs, loss))
    return cfg, batches


def save_scores_case(device="cpu"):
    g = torch.Generator().manual_seed(99)
    names = [f"verb{i % 17} noun{i}" for i in range(300)]
    scores = {"action_id": [], "verb": [], "noun": [], "action": []}
    next_id = 7000
    for bs in (3, 2):
        scores["action_id"].append(torch.arange(next_id, next_id + bs))
        next_id += bs + 5
        scores["verb"].append(torch.randn(bs, 125, generator=g).to(device))
        scores["noun"].append(torch.randn(bs, 352, generator=g).to(device))
        scores["action"].append(torch.randn(bs, 300, generator=g).to(device))
    return scores, names


def sampling_golden():
    """tests/golden/sampling.npz -> dict of int64 index arrays (the file stores int32 offsets from each segment's first
    frame of the modality; all 28,472 rows of the reference's annotations/epic_train_val.csv)."""
    z = np.load(os.path.join(GOLDEN, "sampling.npz"))
    out = {k: z[k] for k in ("rows", "start_frame", "stop_frame")}
    s0 = out["start_frame"].astype(np.int64) - 1
    for k in z.files:
        if k in out:
            continue
        base = s0 // 2 if k.endswith("_Flow") else s0
        out[k] = z[k].astype(np.int64) + base[:, None]
    return out


# ---- training-mode fixtures (oracle/make_golden.py::make_train ran the unmodified reference in model.train()) ----
TRAIN_STATS = ("conv1_7x7_s2_bn", "inception_3a_1x1_bn", "inception_4e_double_3x3_2_bn", "inception_5b_pool_proj_bn")


def train_cases():
    c2 = baseline_config(2)
    c2.model.fusion_dropout = 0
    c2.model.attention.attn_dropout = 0
    c4 = baseline_config(4)
    c4.model.fusion_dropout = 0
    c4.data.audio.dropout = 0.5
    return {"train_cfg2_b2n3": (c2, 2, 3, 1), "train_cfg4_audiodrop_b1n3": (c4, 1, 3, 4)}


GRAD_SAMPLES = 2048  # gradient entries kept per parameter (evenly strided) next to its norm


def grad_cases():
    """Training-STEP cases (forward + get_loss + backward with frozen backbones, reference model.py:160-163): dropouts off
    so that the CPU reference and a CUDA run see the same function."""
    base = {"data.flow.enable": False, "model.fusion_dropout": 0, "model.freeze_mode": "all", "model.attention.decay_step": 1,
            "model.attention.use_entropy": True, "model.attention.attn_dropout": 0, "model.attention.use_gumbel": False}
    c2 = default_cfg(**dict(base, **{"model.attention.type": "mha"}))
    cu = default_cfg(**dict(base, **{"model.attention.type": "unimodal"}))
    cp = default_cfg(**dict(base, **{"model.attention.type": "proto"}))
    c4 = baseline_config(4)
    c4.model.fusion_dropout = 0
    c4.model.freeze_mode = "all"
    return {"grad_mha_b2n3": (c2, 2, 3), "grad_unimodal_b2n3": (cu, 2, 3), "grad_proto_b2n3": (cp, 2, 3),
            "grad_cfg4_b1n3": (c4, 1, 3)}


def grad_targets(cfg, B, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"class": {k: torch.randint(0, int(v), (B,), generator=g) for k, v in cfg.model.num_classes.items()}}


def grad_sample(g: torch.Tensor) -> torch.Tensor:
    flat = g.detach().reshape(-1)
    step = max(flat.numel() // GRAD_SAMPLES, 1)
    return flat[::step][:GRAD_SAMPLES].clone()


FULL_GRAD_SAMPLES = 256


def full_grad_cases():
    """Training-step cases with TRAINABLE backbones (the reference's default freeze_mode "partialbn": every conv trains,
    BatchNorm affine parameters only in the first layer): gradients of every parameter of the model."""
    base = {"data.flow.enable": False, "model.fusion_dropout": 0, "model.attention.decay_step": 1, "model.attention.use_entropy": True,
            "model.attention.attn_dropout": 0, "model.attention.use_gumbel": False}
    c2 = default_cfg(**dict(base, **{"model.attention.type": "mha"}))
    cu = default_cfg(**dict(base, **{"model.attention.type": "unimodal"}))
    cf = default_cfg(**dict(base, **{"model.attention.use_fixed": True, "model.attention.prior_type": "gaussian"}))
    c4 = baseline_config(4)
    c4.model.fusion_dropout = 0
    c1 = baseline_config(1)
    return {"fullgrad_mha_b1n3": (c2, 1, 3), "fullgrad_unimodal_b1n2": (cu, 1, 2), "fullgrad_fixed_b1n2": (cf, 1, 2),
            "fullgrad_cfg4_b1n2": (c4, 1, 2), "fullgrad_cfg1_b2n2": (c1, 2, 2)}


def full_grad_sample(g: torch.Tensor) -> torch.Tensor:
    flat = g.detach().reshape(-1)
    step = max(flat.numel() // FULL_GRAD_SAMPLES, 1)
    return flat[::step][:FULL_GRAD_SAMPLES].clone()


def load_full_grad_case(name):
    cfg, B, N = full_grad_cases()[name]
    mod = get_modality(cfg)
    sd = make_state_dict(cfg, mod, seed=0, calib=CALIB)
    inputs = make_inputs(cfg, mod, B, N, seed=40)
    with np.load(os.path.join(GOLDEN, f"model_{name}.npz")) as z:
        gold = {k: torch.from_numpy(z[k]) for k in z.files}
    return cfg, mod, sd, inputs, grad_targets(cfg, B), gold


def load_grad_case(name):
    cfg, B, N = grad_cases()[name]
    mod = get_modality(cfg)
    sd = make_state_dict(cfg, mod, seed=0, calib=CALIB)
    inputs = make_inputs(cfg, mod, B, N, seed=30)
    with np.load(os.path.join(GOLDEN, f"model_{name}.npz")) as z:
        gold = {k: torch.from_numpy(z[k]) for k in z.files}
    return cfg, mod, sd, inputs, grad_targets(cfg, B), gold


def load_train_case(name):
    cfg, B, N, forwards = train_cases()[name]
    mod = get_modality(cfg)
    sd = make_state_dict(cfg, mod, seed=0, calib=CALIB)
    inputs = [make_inputs(cfg, mod, B, N, seed=20 + i) for i in range(forwards)]
    with np.load(os.path.join(GOLDEN, f"model_{name}.npz")) as z:
        gold = {k: torch.from_numpy(z[k]) for k in z.files}
    return cfg, mod, sd, inputs, gold
