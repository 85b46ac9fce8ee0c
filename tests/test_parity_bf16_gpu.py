"""GPU: bf16 tensor-core path against the CPU oracle / the reference's golden vectors at the sizes the benchmark runs.

north_star: bf16-backbone logits within 2e-2 relative (max|d| / max|ref| per tensor) with identical top-1 class.
The top-1 rule is asserted as an explicit AGREEMENT COUNT over all decisions of the batch; a tolerated flip must sit
inside the error band (reference margin between its two best classes smaller than the permitted logit error) and is
printed with its margin -- nothing is filtered silently."""
import pytest
import torch

from cases import load_case, rel_err
from attention_based_tbn_b200 import TBNModel
from attention_based_tbn_b200.synthetic import make_inputs
from oracle import tbn_oracle

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
BF16_WTS_TOL = 1e-1  # attention weights: not bounded by north_star in bf16 mode (softmax amplifies the feature error)
# 25-segment, 3-modality config: an emulated IDEAL all-bf16 backbone (CPU, fp32 accumulation, bf16 weights/activations)
# already differs from the fp32 reference by 2.6e-2 / 2.9e-2 on the logits and 0.77 (max-norm) on the attention weights
# for the golden clip -- the q.k scores of the test weights are sharp -- so the weights are bounded by their MEAN absolute
# error and a loose max-norm here (measured on the B200: logits 1.2e-2 / 1.4e-2, weights 0.27 max-norm)
CFG5_WTS_MAX, CFG5_WTS_MEAN_ABS = 5e-1, 1.5e-2


def _model(cfg, mod, sd, precision="bf16", **kw):
    m = TBNModel(cfg, mod, "cuda", precision=precision, **kw)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def _fwd(model, inp):
    with torch.no_grad():
        out = model({k: v.cuda() for k, v in inp.items()})
    torch.cuda.synchronize()
    return {k: v.float().cpu().clone() for k, v in out.items()}


def top1_report(out, ref, keys, tol=BF16_TOL):
    """-> (decisions, agreements, flips) with flips = [(key, row, ref margin / max|ref|, ref class, got class)]."""
    n = agree = 0
    flips = []
    for k in keys:
        r, o = ref[k], out[k]
        scale = r.abs().max().item()
        top2 = r.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1]) / scale
        same = o.argmax(1) == r.argmax(1)
        n += r.shape[0]
        agree += int(same.sum())
        for row in (~same).nonzero().flatten().tolist():
            flips.append((k, row, margin[row].item(), int(r[row].argmax()), int(o[row].argmax())))
    return n, agree, flips


def check_top1(name, out, ref, keys, min_rate):
    n, agree, flips = top1_report(out, ref, keys)
    print(f"[top-1] {name}: {agree}/{n} decisions identical to the reference")
    for k, row, margin, rc, gc in flips:
        print(f"[top-1]   flip {k}[{row}]: reference margin {margin:.2e} of max|ref| (band {2 * BF16_TOL:.0e}); {rc} -> {gc}")
        # a flip is only explicable by the permitted logit error when the reference's two best classes are closer
        # than twice that error; anything else is a defect
        assert margin <= 2 * BF16_TOL, (name, k, row, margin)
    assert agree >= min_rate * n, (name, agree, n)
    return agree, n


def test_bf16_bench_size_batch_vs_live_oracle(lib):
    """BASELINE config 2 at the benchmarked size (64 clips x 3 segments, CUDA graph, bf16 backbone) against the oracle
    run live on the same weights and inputs (a few seconds of CPU)."""
    cfg, mod, sd, _, _ = load_case("cfg2_mha_b2n3")
    inp = make_inputs(cfg, mod, 64, 3, seed=64)
    ref = tbn_oracle.tbn_forward(sd, cfg, mod, inp)
    out = _fwd(_model(cfg, mod, sd, use_cuda_graph=True, chunk=32), inp)
    for k in ("verb", "noun"):
        err = rel_err(out[k], ref[k])
        print(f"[parity] cfg2 B=64 bf16 {k}: rel err {err:.3e} (tol {BF16_TOL:g})")
        assert err <= BF16_TOL, (k, err)
    werr = rel_err(out["weights"], ref["weights"])
    print(f"[parity] cfg2 B=64 bf16 weights: rel err {werr:.3e} (tol {BF16_WTS_TOL:g})")
    assert werr <= BF16_WTS_TOL
    check_top1("cfg2 B=64", out, ref, ("verb", "noun"), min_rate=0.95)


@pytest.mark.parametrize("seed", list(range(40, 48)))
def test_bf16_seed_sweep_vs_live_oracle(lib, seed):
    """Eight further input seeds at 2 clips: the 2e-2 logit bound must not depend on one lucky batch."""
    cfg, mod, sd, _, _ = load_case("cfg2_mha_b2n3")
    inp = make_inputs(cfg, mod, 2, 3, seed=seed)
    ref = tbn_oracle.tbn_forward(sd, cfg, mod, inp)
    out = _fwd(_model(cfg, mod, sd, use_cuda_graph=False), inp)
    for k in ("verb", "noun"):
        err = rel_err(out[k], ref[k])
        print(f"[parity] cfg2 seed {seed} bf16 {k}: rel err {err:.3e}")
        assert err <= BF16_TOL, (seed, k, err)
    assert rel_err(out["weights"], ref["weights"]) <= BF16_WTS_TOL
    check_top1(f"cfg2 seed {seed}", out, ref, ("verb", "noun"), min_rate=0.0)


def _check_cfg5_weights(name, w, r):
    mx, mean_abs = rel_err(w, r), (w - r).abs().mean().item()
    print(f"[parity] {name} bf16 weights: max-norm rel err {mx:.3e} (bound {CFG5_WTS_MAX:g}), mean abs err {mean_abs:.3e} "
          f"(bound {CFG5_WTS_MEAN_ABS:g})")
    assert mx <= CFG5_WTS_MAX and mean_abs <= CFG5_WTS_MEAN_ABS, (name, mx, mean_abs)
    assert torch.allclose(w.sum(-1), torch.ones_like(w.sum(-1)), atol=1e-5)


def test_bf16_cfg5_25_segments(lib):
    """BASELINE config 5 (RGB + Flow + Audio MHA, 25 segments): the reference's golden vector at 1 clip and the live
    oracle at 2 clips."""
    cfg, mod, sd, inp, gold = load_case("cfg5_3mod_mha_b1n25")
    out = _fwd(_model(cfg, mod, sd, use_cuda_graph=False), inp)
    ref = {k[4:]: v for k, v in gold.items() if k.startswith("out_")}
    for k in ("verb", "noun"):
        err = rel_err(out[k], ref[k])
        print(f"[parity] cfg5 golden bf16 {k}: rel err {err:.3e}")
        assert err <= BF16_TOL, (k, err)
    _check_cfg5_weights("cfg5 golden", out["weights"], ref["weights"])
    check_top1("cfg5 golden", out, ref, ("verb", "noun"), min_rate=0.0)
    inp2 = make_inputs(cfg, mod, 2, 25, seed=55)
    ref2 = tbn_oracle.tbn_forward(sd, cfg, mod, inp2)
    out2 = _fwd(_model(cfg, mod, sd, use_cuda_graph=True), inp2)
    for k in ("verb", "noun"):
        err = rel_err(out2[k], ref2[k])
        print(f"[parity] cfg5 B=2 N=25 bf16 {k}: rel err {err:.3e}")
        assert err <= BF16_TOL, (k, err)
    assert out2["weights"].shape == (50, 1, 13)
    _check_cfg5_weights("cfg5 B=2", out2["weights"], ref2["weights"])
    check_top1("cfg5 B=2", out2, ref2, ("verb", "noun"), min_rate=0.0)


@pytest.mark.parametrize("name", ["unimodal_b1n3", "proto_b1n3", "cfg3_gaussian_b1n3", "crop2_noattn_b2n4", "crop2_fixed_b1n6"])
def test_bf16_attention_variants_vs_reference_golden(lib, name):
    """UniModal / Prototype attention, the Gaussian fixed prior and the 10-crop audio repeat path (reference
    model.py:243-248) on the bf16 backbone."""
    cfg, mod, sd, inp, gold = load_case(name)
    out = _fwd(_model(cfg, mod, sd, use_cuda_graph=False), inp)
    ref = {k[4:]: v for k, v in gold.items() if k.startswith("out_")}
    assert sorted(ref) == sorted(out)
    for k in ref:
        err = rel_err(out[k], ref[k])
        print(f"[parity] {name} bf16 {k}: rel err {err:.3e}")
        assert err <= (BF16_WTS_TOL if k == "weights" else BF16_TOL), (name, k, err)
    check_top1(name, out, ref, [k for k in ref if k != "weights"], min_rate=0.0)
