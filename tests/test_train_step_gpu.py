"""GPU: the training STEP with frozen backbones -- model.train() forward -> get_loss -> loss.backward() -> clip -> SGD
(reference core/tools/train.py:68-94 with model.freeze_base / freeze_mode "all", model.py:160-163).

  (a) gradients of every head parameter against the ones the unmodified reference produced (tests/golden/model_grad_*.npz:
      norm + strided sample per parameter; dropouts off so that CPU and CUDA see the same function);
  (b) with dropout / Gumbel sampling ON, element-wise against the autograd restatement of the heads
      (oracle/tbn_oracle.py::head_train_forward, itself pinned on the reference's gradients when the goldens are made),
      replaying the masks the CUDA run drew;
  (c) the backward kernels one by one against torch autograd;
  (d) a few optimizer steps: the loss falls and the refreshed plans follow the updated parameters."""
import math

import pytest
import torch
import torch.nn.functional as F

from cases import (CALIB, full_grad_cases, full_grad_sample, grad_cases, grad_sample, grad_targets, load_full_grad_case,
                   load_grad_case, rel_err)
from attention_based_tbn_b200 import TBNModel
from attention_based_tbn_b200.config import default_cfg, get_modality
from attention_based_tbn_b200.head_train import head_parameter_names
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict
from oracle import tbn_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
FULL_GRAD_MAX = 0.25  # worst sampled entry of any parameter gradient vs float64 (the reference's own fp32 run reaches 0.17)
GRAD_TOL = 2e-3  # max|d| / max|ref| per parameter gradient (fp32 kernels vs fp32 CPU autograd through ~70 conv layers)


def _s():
    return torch.cuda.current_stream().cuda_stream


def _model(cfg, mod, sd):
    model = TBNModel(cfg, mod, "cuda", precision="fp32")
    model.load_state_dict(sd, strict=True)
    return model.train()


def _crit():
    return {"crossentropy": torch.nn.CrossEntropyLoss()}


def _cuda_target(t):
    return {"class": {k: v.cuda() for k, v in t["class"].items()}}


@pytest.mark.parametrize("name", list(grad_cases()))
def test_training_step_gradients_match_reference(lib, name):
    cfg, mod, sd, inp, target, gold = load_grad_case(name)
    model = _model(cfg, mod, sd)
    out = model({k: v.cuda() for k, v in inp.items()})
    for k in out:
        assert rel_err(out[k].detach().cpu(), gold["out_" + k]) < 1e-4, k
    loss, bs = model.get_loss(_crit(), _cuda_target(target), out, epoch=20)
    assert abs(float(loss["total"]) - float(gold["loss_total"])) < 1e-4 * abs(float(gold["loss_total"]))
    loss["total"].backward()
    names = [k[len("gsample_"):] for k in gold if k.startswith("gsample_")]
    assert sorted(names) == sorted(head_parameter_names(cfg, mod))
    params = dict(model.named_parameters())
    for n in names:
        g = params[n].grad
        assert g is not None and g.shape == params[n].shape, n
        e = rel_err(grad_sample(g.cpu()), gold["gsample_" + n])
        gn = abs(g.norm().item() - float(gold["gnorm_" + n])) / max(float(gold["gnorm_" + n]), 1e-30)
        print(f"[grad] {name} {n}: sample rel err {e:.2e}, norm rel err {gn:.2e}")
        assert e < GRAD_TOL and gn < GRAD_TOL, (n, e, gn)
    assert all(p.grad is None for n, p in params.items() if n.startswith("Base_"))
    assert model.last_backward_launches > 0


@pytest.mark.parametrize("name", list(full_grad_cases()))
def test_full_training_step_gradients_match_reference(lib, name):
    """TRAINABLE backbones (the reference's default freeze_mode "partialbn"): one training step of the unmodified reference
    against the CUDA step -- loss, outputs, and the gradient of every parameter that receives one (69 convs per backbone
    + the unfrozen BatchNorm affines + the heads), norm and strided sample each."""
    cfg, mod, sd, inp, target, gold = load_full_grad_case(name)
    model = _model(cfg, mod, sd)
    out = model({k: v.cuda() for k, v in inp.items()})
    for k in out:
        assert rel_err(out[k].detach().cpu(), gold["out_" + k]) < 1e-4, k
    loss, _ = model.get_loss(_crit(), _cuda_target(target), out, epoch=20)
    assert abs(float(loss["total"]) - float(gold["loss_total"])) < 1e-4 * abs(float(gold["loss_total"]))
    loss["total"].backward()
    torch.cuda.synchronize()
    names = sorted(k[len("gsample_"):] for k in gold if k.startswith("gsample_"))
    params = dict(model.named_parameters())
    have = sorted(n for n, p in params.items() if p.grad is not None)
    assert have == names, (set(have) ^ set(names))
    # Parity bar: as close to the float64 gradients as the reference's own fp32 gradients are.  (Through ~70 conv +
    # batch-statistics BatchNorm layers on a handful of frames fp32 gradients are reproducible to ~1e-2 in max-norm only:
    # the goldens hold the reference step in fp32 AND in float64.)
    mine, ref, mine_n, ref_n = [], [], [], []
    for n in names:
        g = params[n].grad
        if n.startswith("Base_") and n.endswith(".bias") and not n.endswith("_bn.bias"):
            # a conv bias in front of a batch-statistics BatchNorm has NO effect on the output: its gradient is zero up to
            # rounding in the reference as well (both sides: noise far below the weight gradient of the same layer)
            wn = float(gold["gnorm_" + n[:-len("bias")] + "weight"])
            assert g.norm().item() < 1e-4 * wn and float(gold["gnorm_" + n]) < 1e-4 * wn, (n, g.norm().item(), wn)
            continue
        g64, n64 = gold["gsample64_" + n], float(gold["gnorm64_" + n])
        mine.append(rel_err(full_grad_sample(g.cpu()), g64))
        ref.append(rel_err(gold["gsample_" + n], g64))
        mine_n.append(abs(g.norm().item() - n64) / n64)
        ref_n.append(abs(float(gold["gnorm_" + n]) - n64) / n64)
    q = lambda v, f: sorted(v)[min(int(len(v) * f), len(v) - 1)]  # noqa: E731
    print(f"[full grad] {name}: {len(mine)} parameters, {model.last_backward_launches} backward launches; max-norm error of the "
          f"sampled entries vs float64, median / p90 / max: CUDA {q(mine, .5):.1e} / {q(mine, .9):.1e} / {max(mine):.1e}, reference "
          f"fp32 {q(ref, .5):.1e} / {q(ref, .9):.1e} / {max(ref):.1e}; norm error: CUDA {q(mine_n, .5):.1e} / {q(mine_n, .9):.1e} / "
          f"{max(mine_n):.1e}, reference fp32 {q(ref_n, .5):.1e} / {q(ref_n, .9):.1e} / {max(ref_n):.1e}")
    assert q(mine, .5) <= 2 * q(ref, .5) + 1e-4 and q(mine, .9) <= 3 * q(ref, .9) + 1e-4 and max(mine) <= FULL_GRAD_MAX
    assert q(mine_n, .5) <= 3 * q(ref_n, .5) + 1e-4 and q(mine_n, .9) <= 3 * q(ref_n, .9) + 1e-3 and max(mine_n) <= 2e-2


@pytest.mark.parametrize("name", ["fullgrad_mha_b1n3", "fullgrad_cfg4_b1n2"])
def test_mixed_precision_training_step(lib, name):
    """train_precision="bf16": the forward, input-gradient and weight-gradient convs of the training step run on the tensor cores (bf16 products,
    fp32 accumulation); BatchNorm, weight gradients and everything stored stay fp32.  Outputs within the bf16 bar of the
    inference path (2e-2), gradients against float64 within a small multiple of what fp32 itself achieves."""
    cfg, mod, sd, inp, target, gold = load_full_grad_case(name)
    model = TBNModel(cfg, mod, "cuda", train_precision="bf16")
    model.load_state_dict(sd, strict=True)
    model.train()
    out = model({k: v.cuda() for k, v in inp.items()})
    for k in out:
        e = rel_err(out[k].detach().cpu(), gold["out_" + k])
        print(f"[amp] {name} {k}: forward rel err {e:.2e}")
        assert e < (2e-1 if k == "weights" else 4e-2), (k, e)  # (batch statistics over 2-3 frames amplify the bf16 error)
    loss, _ = model.get_loss(_crit(), _cuda_target(target), out, epoch=20)
    loss["total"].backward()
    torch.cuda.synchronize()
    params = dict(model.named_parameters())
    # The sampled ENTRIES of these gradients are not a usable yardstick here: on this random-weight network a relative
    # perturbation of 1e-7 (fp32 rounding) already moves single entries by 1e-2 (ReLU / max-pool decisions flip), and the bf16
    # forward perturbs by 1e-2.  Direction and length of each parameter's gradient are: cosine against float64 and norm.
    cos, nerr, head_cos = [], [], []
    for n in sorted(k[len("gsample_"):] for k in gold if k.startswith("gsample_")):
        if n.startswith("Base_") and n.endswith(".bias") and not n.endswith("_bn.bias"):
            continue
        a, b_ = full_grad_sample(params[n].grad.cpu()).double(), gold["gsample64_" + n].double()
        c_ = float((a * b_).sum() / (a.norm() * b_.norm()))
        (cos if n.startswith("Base_") else head_cos).append(c_)
        nerr.append(abs(params[n].grad.norm().item() - float(gold["gnorm64_" + n])) / float(gold["gnorm64_" + n]))
    q = lambda v, f: sorted(v)[min(int(len(v) * f), len(v) - 1)]  # noqa: E731
    print(f"[amp] {name}: gradient cosine vs float64 -- heads min {min(head_cos):.4f}; backbone convs p10 / median / max "
          f"{q(cos, .1):.3f} / {q(cos, .5):.3f} / {max(cos):.3f}; norm error median / p90 {q(nerr, .5):.1e} / {q(nerr, .9):.1e}")
    assert min(head_cos) > 0.97 and q(cos, .5) > 0.6 and max(cos) > 0.9 and q(nerr, .5) < 6e-2 and q(nerr, .9) < 1.5e-1


def _oracle_step(model, cfg, mod, B, N, target, opts, epoch=20):
    """The autograd checker (on the CPU: no TF32 anywhere) on the features the CUDA forward produced, with the masks it
    drew; the loss follows get_loss (reference model.py:264-334) including its entropy threshold."""
    head = model._train_plans[1]
    F_ = B * N
    E = 1024
    X = head.fusion_row(F_).cpu()
    feats = {}
    for i, m in enumerate(mod):
        feats[m] = X[:, i * E:(i + 1) * E]
    S = model._train_step[2]
    if head.att_type is not None:
        a = head.audio_positions(F_, S).cpu()
        feats["Audio"] = a.view(F_, S, E).permute(0, 2, 1).unsqueeze(2).contiguous()  # (F, E, 1, S)
    names = head_parameter_names(cfg, mod)
    par = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    cpu = lambda t: t.cpu() if t is not None else None  # noqa: E731
    for n in names:
        par[n].requires_grad_(True)
    out = tbn_oracle.head_train_forward(par, cfg, mod, feats, B, N, attn_mask=cpu(opts.get("attn_mask")),
                                        fusion_mask=cpu(opts.get("fusion_mask")), gumbel=cpu(opts.get("gumbel_noise")))
    loss = sum(F.cross_entropy(out[k], target["class"][k].cpu()) for k in target["class"])
    if "weights" in out and cfg.model.attention.use_entropy and epoch + 1 >= cfg.model.attention.decay_step:
        ent = tbn_oracle.entropy_term(out["weights"].reshape(F_, 1, -1))
        if not ent < cfg.model.attention.entropy_thresh:  # (training: the term is switched off below the threshold)
            loss = loss + cfg.model.attention.entropy_decay * ent
    loss.backward()
    return out, loss, {n: par[n].grad for n in names}


@pytest.mark.parametrize("att_type", ["mha", "unimodal", "proto"])
def test_training_step_with_dropout_matches_autograd_checker(lib, att_type):
    over = {"data.flow.enable": False, "model.attention.type": att_type, "model.freeze_mode": "all", "model.attention.use_entropy": True,
            "model.attention.decay_step": 1}  # fusion_dropout 0.5, attn_dropout 0.5, use_gumbel True: the reference defaults
    cfg = default_cfg(**over)
    mod = get_modality(cfg)
    B, N = 3, 3
    sd = make_state_dict(cfg, mod, seed=0, calib=CALIB)
    model = _model(cfg, mod, sd)
    inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, B, N, seed=31).items()}
    target = _cuda_target(grad_targets(cfg, B))
    torch.manual_seed(5)
    out = model(inp)
    loss, _ = model.get_loss(_crit(), target, out, epoch=20)
    loss["total"].backward()
    opts = model._train_step[3]
    assert "fusion_mask" in opts and (("attn_mask" in opts) if att_type == "mha" else ("gumbel_noise" in opts))
    o2, l2, g2 = _oracle_step(model, cfg, mod, B, N, target, opts)
    for k in out:  # the FORWARD under dropout (value bias weighted by the kept probability mass, hard Gumbel sample)
        e = rel_err(out[k].detach().float().cpu().reshape(-1), o2[k].detach().float().cpu().reshape(-1))
        print(f"[train fwd] {att_type} {k}: rel err {e:.2e}")
        assert e < 1e-4, (k, e)
    assert abs(float(loss["total"]) - float(l2)) < 1e-4 * abs(float(l2))
    params = dict(model.named_parameters())
    for n, g in g2.items():
        e = rel_err(params[n].grad.cpu(), g.cpu())
        print(f"[grad] {att_type}+dropout {n}: rel err {e:.2e}")
        assert e < 5e-4, (n, e)


def test_head_finetune_steps(lib):
    """The reference loop (train.py:68-94): zero_grad, forward, loss, backward, clip, step -- a few times on one batch.
    The loss falls, and after the parameters moved the plans' folded weight copies follow (second-step gradients agree
    with the checker evaluated at the LIVE parameters)."""
    cfg = default_cfg(**{"data.flow.enable": False, "model.attention.type": "mha", "model.freeze_mode": "all", "model.fusion_dropout": 0,
                         "model.attention.attn_dropout": 0, "model.attention.use_entropy": True, "model.attention.decay_step": 1})
    mod = get_modality(cfg)
    B, N = 2, 3
    model = _model(cfg, mod, make_state_dict(cfg, mod, seed=0, calib=CALIB))
    inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, B, N, seed=32).items()}
    target = _cuda_target(grad_targets(cfg, B))
    train_params = [p for p in model.parameters() if p.requires_grad]
    assert len(train_params) == len(head_parameter_names(cfg, mod))
    opt = torch.optim.SGD(train_params, lr=0.01, momentum=0.9)
    losses = []
    for step in range(4):
        opt.zero_grad()
        out = model(inp)
        loss, _ = model.get_loss(_crit(), target, out, epoch=20)
        loss["total"].backward()
        if step == 2:
            _, l2, g2 = _oracle_step(model, cfg, mod, B, N, target, model._train_step[3])
            params = dict(model.named_parameters())
            for n, g in g2.items():
                assert rel_err(params[n].grad.cpu(), g.cpu()) < 5e-4, n
            assert abs(float(loss["total"]) - float(l2)) < 1e-4 * abs(float(l2))
        torch.nn.utils.clip_grad_norm_(model.parameters(), 20.0)
        opt.step()
        losses.append(float(loss["total"]))
    print("[finetune] losses", [round(v, 4) for v in losses])
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_training_loop_with_trainable_backbones(lib, precision):
    """The reference loop (train.py:68-94) with its default freeze_mode "partialbn" -- every conv of both backbones and all
    head layers train: zero_grad, forward (dropouts on), get_loss, backward, clip_grad_norm_(20), SGD step.  The loss falls
    on a fixed batch, every trainable parameter received a gradient and moved, the frozen BatchNorm affines did not."""
    cfg = default_cfg(**{"data.flow.enable": False, "model.attention.type": "mha", "model.attention.use_entropy": True,
                         "model.attention.decay_step": 1})
    mod = get_modality(cfg)
    B, N = 2, 3
    model = TBNModel(cfg, mod, "cuda", train_precision=precision)  # "bf16": the convs of the step run on the tensor cores
    model.load_state_dict(make_state_dict(cfg, mod, seed=0, calib=CALIB))
    model.train()
    inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, B, N, seed=35).items()}
    target = _cuda_target(grad_targets(cfg, B))
    train_params = {n: p for n, p in model.named_parameters() if p.requires_grad}
    frozen = {n: p.detach().clone() for n, p in model.named_parameters() if not p.requires_grad}
    assert len(frozen) == 2 * 68 * 2  # BatchNorm weight + bias of all but the first layer, two backbones
    before = {n: p.detach().clone() for n, p in train_params.items()}
    opt = torch.optim.SGD(train_params.values(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    torch.manual_seed(3)
    losses = []
    for step in range(5):
        opt.zero_grad()
        out = model(inp)
        loss, _ = model.get_loss(_crit(), target, out, epoch=20)
        loss["total"].backward()
        if step == 0:
            unused = {"pe.0.pe"}
            missing = [n for n, p in train_params.items() if p.grad is None and n not in unused]
            assert not missing, missing[:5]
        torch.nn.utils.clip_grad_norm_(model.parameters(), 20.0)
        opt.step()
        losses.append(float(loss["total"]))
    print(f"[full training loop] {precision} losses", [round(v, 4) for v in losses])
    assert losses[-1] < 0.7 * losses[0]
    assert all(not torch.equal(p.detach(), before[n]) for n, p in train_params.items())
    assert all(torch.equal(p.detach(), frozen[n]) for n, p in model.named_parameters() if n in frozen)
    live = model.state_dict()
    assert int(live["Base_RGB.inception_4a_1x1_bn.num_batches_tracked"]) == 5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_finetune_on_eval_mode_backbones(lib, precision):
    """model.train() with the backbones put in eval(): they run on the inference plans (folded BatchNorm with the running
    statistics, bf16 tensor-core path when configured) and only the heads train.  With the dropouts off the outputs are the
    eval forward's; the gradients agree with the checker on the features that path produced; no BatchNorm state moves."""
    cfg = default_cfg(**{"data.flow.enable": False, "model.attention.type": "mha", "model.freeze_mode": "all", "model.fusion_dropout": 0,
                         "model.attention.attn_dropout": 0, "model.attention.use_entropy": True, "model.attention.decay_step": 1})
    mod = get_modality(cfg)
    B, N = 2, 3
    model = TBNModel(cfg, mod, "cuda", precision=precision)
    model.load_state_dict(make_state_dict(cfg, mod, seed=0, calib=CALIB))
    inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, B, N, seed=34).items()}
    target = _cuda_target(grad_targets(cfg, B))
    model.eval()
    with torch.no_grad():
        ref = {k: v.clone() for k, v in model(inp).items()}
    model.train()
    for m in mod:
        getattr(model, f"Base_{m}").eval()
    nbt = {n: b.clone() for n, b in model.named_buffers() if n.endswith("num_batches_tracked")}
    out = model(inp)
    for k in ref:
        e = rel_err(out[k].detach().cpu().reshape(-1), ref[k].cpu().reshape(-1))
        print(f"[eval-backbone train fwd] {precision} {k}: rel err vs eval forward {e:.2e}")
        assert e < 1e-4, (k, e)
    loss, _ = model.get_loss(_crit(), target, out, epoch=20)
    loss["total"].backward()
    _, l2, g2 = _oracle_step(model, cfg, mod, B, N, target, model._train_step[3])
    params = dict(model.named_parameters())
    for n, g in g2.items():
        assert rel_err(params[n].grad.cpu(), g.cpu()) < 5e-4, n
    assert all(torch.equal(b, nbt[n]) for n, b in model.named_buffers() if n in nbt)
    # the heads moved: the next eval() forward must see the new parameters
    with torch.no_grad():
        for n in head_parameter_names(cfg, mod):
            params[n].add_(0.05 * torch.randn_like(params[n]))
    model.eval()
    with torch.no_grad():
        moved = model(inp)
    assert rel_err(moved["verb"].cpu(), ref["verb"].cpu()) > 1e-3


def test_backward_contract_errors(lib):
    cfg = default_cfg(**{"data.flow.enable": False, "model.attention.enable": False, "model.fusion_dropout": 0})  # partialbn: live backbones
    mod = get_modality(cfg)
    model = _model(cfg, mod, make_state_dict(cfg, mod, seed=0, calib=CALIB))
    inp = {k: v.cuda() for k, v in make_inputs(cfg, mod, 1, 3, seed=33).items()}
    for m in mod:  # eval-mode backbones run on the inference plans, which keep nothing for a backward pass
        getattr(model, f"Base_{m}").eval()
    out = model(inp)
    with pytest.raises(NotImplementedError, match="backbone parameters"):
        out["verb"].sum().backward()
    for n, p in model.named_parameters():
        if n.startswith("Base_"):
            p.requires_grad_(False)
    first = model(inp)
    model(inp)
    with pytest.raises(RuntimeError, match="stale training forward"):
        first["verb"].sum().backward()
    with torch.no_grad():  # no autograd node without grad mode
        assert not model(inp)["verb"].requires_grad


# ---- kernels ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_sgemm_variants(lib, ta, tb):
    g = torch.Generator().manual_seed(1)
    for (m, n, k) in ((70, 130, 50), (477, 512, 6), (1024, 1034, 77), (6, 3, 13), (64, 64, 16)):
        a = torch.randn((k, m + 3) if ta else (m, k + 5), generator=g).to(DEV)
        b = torch.randn((n, k + 2) if tb else (k, n + 7), generator=g).to(DEV)
        c0 = torch.randn(m, n + 1, generator=g).to(DEV)
        opa = a[:, :m].t() if ta else a[:, :k]
        opb = b[:, :k].t() if tb else b[:, :n]
        for beta in (0.0, 1.0):
            c = c0.clone()
            lib.call("tbn_sgemm", ta, tb, m, n, k, 0.5, a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], beta, c.data_ptr(), n + 1, _s())
            ref = 0.5 * (opa.double() @ opb.double()) + beta * c0[:, :n].double()
            assert rel_err(c[:, :n].cpu(), ref.cpu()) < 1e-5, (m, n, k, beta)
            assert torch.equal(c[:, n], c0[:, n])  # padding column untouched


def test_small_backward_kernels(lib):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(37, 70, generator=g).to(DEV)
    out = torch.ones(66, device=DEV)
    lib.call("tbn_colsum", x.data_ptr(), 70, 37, 66, 0.0, out.data_ptr(), _s())
    assert rel_err(out.cpu(), x[:, :66].double().sum(0).cpu()) < 1e-6
    ds = torch.randn(3, 9, generator=g).to(DEV)
    dl = torch.empty(12, 7, device=DEV)
    lib.call("tbn_consensus_bwd", ds.data_ptr(), 9, dl.data_ptr(), 3, 4, 7, _s())
    assert torch.allclose(dl, (ds[:, :7] / 4).repeat_interleave(4, 0))
    h = torch.relu(torch.randn(5, 512, generator=g)).to(DEV)
    dh = torch.randn(5, 512, generator=g).to(DEV)
    ref = torch.where(h > 0, dh * 2.0, torch.zeros_like(dh))
    lib.call("tbn_relu_dropout_bwd", dh.data_ptr(), h.data_ptr(), 2.0, 5 * 512, _s())
    assert torch.equal(dh, ref)
    y = torch.softmax(torch.randn(11, 13, generator=g), 1).to(DEV)
    dy = torch.randn(11, 13, generator=g).to(DEV)
    dx = torch.empty(11, 13, device=DEV)
    lib.call("tbn_softmax_bwd", y.data_ptr(), dy.data_ptr(), dx.data_ptr(), 11, 13, 13, _s())
    assert rel_err(dx.cpu(), (y * (dy - (y * dy).sum(1, keepdim=True))).cpu()) < 1e-5
    a = torch.randn(6 * 5, 64, generator=g).to(DEV)
    dyf = torch.randn(6, 80, generator=g).to(DEV)
    dwi = torch.randn(6, 5, generator=g).to(DEV)
    dw = torch.empty(6, 5, device=DEV)
    lib.call("tbn_weighted_sum_bwd", a.data_ptr(), dyf.data_ptr(), 80, dwi.data_ptr(), dw.data_ptr(), 6, 5, 64, _s())
    ref = dwi + torch.einsum("fsc,fc->fs", a.view(6, 5, 64), dyf[:, :64])
    assert rel_err(dw.cpu(), ref.cpu()) < 1e-5


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("cin,cout,k,stride,h,w", [(64, 96, 3, 1, 14, 17), (128, 160, 3, 2, 28, 28), (192, 64, 1, 1, 9, 13),
                                                   (96, 96, 3, 2, 15, 21)])
def test_conv_dgrad_matches_autograd(lib, tc, cin, cout, k, stride, h, w):
    """The input-gradient conv of the backbone backward (flipped / transposed filter; stride 2 through the zero-spread
    gradient), CUDA cores and tensor cores, against autograd of F.conv2d."""
    from attention_based_tbn_b200.arch import ConvSpec
    from attention_based_tbn_b200.backbone_train import TrainBackbone

    g = torch.Generator().manual_seed(11)
    n = 3
    x = torch.randn(n, cin, h, w, generator=g).requires_grad_(True)  # (reference on the CPU: cuDNN would use TF32)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    y = F.conv2d(x, wt, None, stride, k // 2)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    oh, ow = y.shape[2:]
    dy, wt = dy.to(DEV), wt.to(DEV)
    bb = TrainBackbone({}, "", 3, torch.device("cuda"), tensor_cores=tc)
    bb._stream, bb._keep, bb.launches, bb._in_backward = _s(), [], 0, True
    du = dy.permute(0, 2, 3, 1).reshape(n * oh * ow, cout).contiguous()
    tmp = bb._dgrad(ConvSpec("t", cin, cout, k, stride, k // 2), du, n, oh, ow, h, w, wt)
    torch.cuda.synchronize()
    got = tmp.view(n, h, w, cin).permute(0, 3, 1, 2).cpu()
    e = rel_err(got, x.grad)
    a, b_ = got.flatten().double(), x.grad.flatten().double()
    cos = float((a * b_).sum() / (a.norm() * b_.norm()))
    print(f"[dgrad] tc={tc} k={k} s={stride}: rel err {e:.2e}, cosine {cos:.6f}")
    assert e < (2e-2 if tc else 1e-5) and cos > (0.9999 if tc else 0.9999999)


@pytest.mark.parametrize("cin,cout,k,stride,h,w", [(64, 96, 3, 1, 14, 17), (128, 160, 3, 2, 28, 28), (192, 64, 1, 1, 9, 13), (3, 64, 7, 2, 37, 41)])
def test_conv_wgrad_tensor_cores_matches_autograd(lib, cin, cout, k, stride, h, w):
    """Weight gradient as a tensor-core split-K GEMM over the output positions (transposed bf16 im2col operands)."""
    from attention_based_tbn_b200.arch import ConvSpec
    from attention_based_tbn_b200.backbone_train import TrainBackbone, _Act

    g = torch.Generator().manual_seed(12)
    n = 5
    x = torch.randn(n, cin, h, w, generator=g)
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).requires_grad_(True)
    y = F.conv2d(x, wt, None, stride, k // 2)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    oh, ow = y.shape[2:]
    bb = TrainBackbone({}, "", 3, torch.device("cuda"), tensor_cores=True)
    bb._stream, bb._keep, bb.launches, bb._in_backward = _s(), [], 0, True
    xa = _Act(x.permute(0, 2, 3, 1).reshape(n * h * w, cin).contiguous().to(DEV), n, h, w, cin)
    du = dy.permute(0, 2, 3, 1).reshape(n * oh * ow, cout).contiguous().to(DEV)
    dwk = bb._wgrad_tc(ConvSpec("t", cin, cout, k, stride, k // 2), xa, du, n, oh, ow)
    torch.cuda.synchronize()
    got = dwk.view(k, k, cin, cout).permute(3, 2, 0, 1).cpu()
    e = rel_err(got, wt.grad)
    print(f"[wgrad tc] k={k} s={stride} cin={cin}: rel err {e:.2e}")
    assert e < 1e-2


@pytest.mark.parametrize("f,s,e,groups", [(5, 8, 1024, 64), (3, 13, 256, 8)])
def test_groupnorm_bwd_matches_autograd(lib, f, s, e, groups):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(f, s, e, generator=g).to(DEV).requires_grad_(True)  # rows (f, s) x channels
    gamma = torch.randn(e, generator=g).to(DEV).requires_grad_(True)
    beta = torch.randn(e, generator=g).to(DEV).requires_grad_(True)
    dy = torch.randn(f, s, e, generator=g).to(DEV)
    y = F.group_norm(x.permute(0, 2, 1), groups, gamma, beta, 1e-5)  # (f, e, s)
    y.backward(dy.permute(0, 2, 1))
    dx, dgf, dbf = torch.empty(f * s, e, device=DEV), torch.empty(f, e, device=DEV), torch.empty(f, e, device=DEV)
    lib.call("tbn_groupnorm_bwd", x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), dx.data_ptr(), dgf.data_ptr(), dbf.data_ptr(), f, s, e,
             groups, 1e-5, _s())
    assert rel_err(dx.view(f, s, e).cpu(), x.grad.cpu()) < 1e-4
    assert rel_err(dgf.sum(0).cpu(), gamma.grad.cpu()) < 1e-4
    assert rel_err(dbf.sum(0).cpu(), beta.grad.cpu()) < 1e-4


@pytest.mark.parametrize("f,s,e,heads,drop", [(4, 8, 1024, 4, True), (3, 13, 1024, 8, False), (2, 5, 256, 16, True), (2, 32, 128, 1, False)])
def test_mha_bwd_matches_autograd(lib, f, s, e, heads, drop):
    g = torch.Generator().manual_seed(6)
    d = e // heads
    q = torch.randn(f, e, generator=g).to(DEV).requires_grad_(True)
    k = torch.randn(f * s, e, generator=g).to(DEV).requires_grad_(True)
    v = torch.randn(f * s, e, generator=g).to(DEV).requires_grad_(True)
    mask = ((torch.rand(f, heads, s, generator=g) > 0.5).float() * 2.0).to(DEV) if drop else None
    d_o = torch.randn(f, e, generator=g).to(DEV)
    d_w = torch.randn(f, s, generator=g).to(DEV)
    qh = q.view(f, heads, 1, d)
    kh = k.view(f, s, heads, d).transpose(1, 2)
    vh = v.view(f, s, heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d), -1)
    if drop:
        p = p * mask.view(f, heads, 1, s)
    o = (p @ vh).transpose(1, 2).reshape(f, e)
    w = p.mean(1).view(f, s)
    ((o * d_o).sum() + (w * d_w).sum()).backward()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    lib.call("tbn_mha_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), mask.data_ptr() if drop else 0, d_o.data_ptr(), d_w.data_ptr(),
             dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), f, s, e, heads, _s())
    assert rel_err(dq.cpu(), q.grad.cpu()) < 1e-4
    assert rel_err(dk.cpu(), k.grad.cpu()) < 1e-4
    assert rel_err(dv.cpu(), v.grad.cpu()) < 1e-4
