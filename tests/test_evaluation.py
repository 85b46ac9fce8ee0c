"""Evaluation-loop consumers of the scores: Metric, save_scores.

CPU: the numpy oracle and the host-side JSON writer against fixtures produced by the unmodified reference
(oracle/make_golden_eval.py).  GPU: the product ``Metric`` (tbn_topk_confusion kernel) against the same fixtures."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from cases import GOLDEN, eval_case, save_scores_case
from attention_based_tbn_b200 import evaluation


def _gold():
    with np.load(os.path.join(GOLDEN, "eval_metric.npz")) as z:
        return {k: z[k] for k in z.files}


def _check(name, loss, acc, conf, gold):
    for k, v in loss.items():
        assert float(gold[f"{name}/loss/{k}"]) == v, (name, "loss", k)
    for k, v in acc.items():
        assert gold[f"{name}/acc/{k}"].tolist() == list(v), (name, "acc", k)
    for k, v in conf.items():
        got = v.cpu().numpy() if torch.is_tensor(v) else v
        assert np.array_equal(got, gold[f"{name}/conf/{k}"]), (name, "conf", k)


@pytest.mark.parametrize("name", ["two_class", "one_class"])
def test_metric_oracle_matches_reference_fixture(name):
    from oracle.eval_oracle import MetricOracle

    gold = _gold()
    cfg, batches = eval_case(name)
    m = MetricOracle(OrderedDict(cfg.model.num_classes.items()), cfg.val.topk, len(batches), list(batches[0][3].keys()))
    for out, target, bs, loss in batches:
        m.set_metrics({k: v.numpy() for k, v in out.items()}, {k: v.numpy() for k, v in target["class"].items()}, bs,
                      {k: v.item() for k, v in loss.items()})
    _check(name, *m.get_metrics(), gold)


def test_save_scores_json_is_byte_identical_to_reference(tmp_path):
    scores, names = save_scores_case()
    out = os.path.join(str(tmp_path), "sub", "seen.json")
    evaluation.save_scores(scores, out, names)
    with open(out) as f, open(os.path.join(GOLDEN, "save_scores.json")) as g:
        got, ref = f.read(), g.read()
    assert got == ref
    d = json.loads(got)
    assert list(d.keys()) == ["version", "challenge", "results"] and len(d["results"]) == 5
    assert len(d["results"]["7000"]["action"]) == 100 and list(d["results"]["7000"].keys()) == ["verb", "noun", "action"]
    assert torch.is_tensor(scores["verb"]) and scores["verb"].shape == (5, 125)  # concatenated in place like the reference


def test_metric_requires_the_cuda_library():
    with pytest.raises(Exception):
        evaluation.topk_correct(torch.zeros(2, 5), torch.zeros(2, dtype=torch.int64), 1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["two_class", "one_class"])
def test_metric_gpu_matches_reference_fixture(name):
    gold = _gold()
    cfg, batches = eval_case(name, device="cuda")
    m = evaluation.Metric(cfg, len(batches), torch.device("cuda"))
    for out, target, bs, loss in batches:
        m.set_metrics(out, target, bs, loss)
    _check(name, *m.get_metrics(), gold)


@pytest.mark.gpu
@pytest.mark.parametrize("b,c,maxk", [(1, 5, 5), (7, 125, 5), (64, 352, 5), (33, 2513, 10), (200, 16, 3)])
def test_topk_confusion_kernel_vs_oracle(b, c, maxk):
    from oracle.eval_oracle import correct_and_confusion

    g = torch.Generator().manual_seed(b * 1000 + c)
    s = torch.randn(b, c, generator=g)
    if c == 16:
        s = s.round()  # many ties: lower class index first
    if b == 7:
        s[3, 10] = float("-inf"); s[4] = float("-inf")
    t = torch.randint(0, c, (b,), generator=g)
    conf = torch.full((c, c), 2.0, device="cuda")
    correct, preds = evaluation.topk_correct(s.cuda(), t.cuda(), maxk, conf)
    rc, rp, rconf = correct_and_confusion(s.numpy(), t.numpy(), maxk)
    assert np.array_equal(preds.cpu().numpy(), rp)
    assert np.array_equal(correct.cpu().numpy(), rc)
    assert np.array_equal(conf.cpu().numpy(), rconf + 2.0)


@pytest.mark.gpu
def test_save_scores_from_device_tensors(tmp_path):
    scores, names = save_scores_case(device="cuda")
    out = os.path.join(str(tmp_path), "seen.json")
    evaluation.save_scores(scores, out, names)
    with open(out) as f, open(os.path.join(GOLDEN, "save_scores.json")) as g:
        assert f.read() == g.read()


@pytest.mark.gpu
def test_eval_loop_end_to_end(tmp_path):
    """evaluation.test(): the model forward, get_loss, Metric and save_scores chained like core/tools/test.py, checked
    against per-batch results recomputed with the numpy oracle from the model's own outputs."""
    from cases import CALIB
    from attention_based_tbn_b200 import baseline_config, build_model, get_modality
    from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict
    from oracle.eval_oracle import MetricOracle

    cfg = baseline_config(2)
    cfg.test.save_results = True
    cfg.test.batch_size = 2
    mod = get_modality(cfg)
    model, criterion, _ = build_model(cfg, mod, "cuda")
    model.load_state_dict(make_state_dict(cfg, mod, seed=0, calib=CALIB))
    model.eval()
    g = torch.Generator().manual_seed(5)
    batches = []
    for i in range(3):
        data = make_inputs(cfg, mod, 2, 3, seed=50 + i)
        target = {"class": {k: torch.randint(0, n, (2,), generator=g) for k, n in cfg.model.num_classes.items()}}
        batches.append((data, target, torch.tensor([100 + 2 * i, 101 + 2 * i])))

    class Loader(list):
        dataset = list(range(6))
        batch_size = 2

    loss, acc, conf, output = evaluation.test(cfg, model, Loader(batches), criterion, mod, device=torch.device("cuda"))
    assert set(loss) == {"verb", "noun", "all_class", "entropy", "total"} and set(acc) == {"verb", "noun", "all_class"}
    assert all(float(conf[k].sum()) == 6.0 for k in conf)
    # recompute from the collected outputs
    ref = MetricOracle(OrderedDict(cfg.model.num_classes.items()), cfg.val.topk, 3, list(loss.keys()))
    for i, (_, target, _) in enumerate(batches):
        out = {k: output[k][i].float().cpu().numpy() for k in cfg.model.num_classes}
        tgt = {k: v.cpu().numpy() for k, v in target["class"].items()}
        ref.set_metrics(out, tgt, 2, {k: 0.0 for k in loss})
    _, racc, rconf = ref.get_metrics()
    for k in acc:
        assert list(acc[k]) == list(racc[k]), k
    for k in conf:
        assert np.array_equal(conf[k].cpu().numpy(), rconf[k]), k
    assert all(np.isfinite(v) and v > 0 for v in loss.values())
    path = os.path.join(str(tmp_path), "res", "seen.json")
    evaluation.save_scores(output, path, None)
    d = json.load(open(path))
    assert sorted(d["results"]) == [str(100 + j) for j in range(6)] and len(d["results"]["100"]["noun"]) == 352


@pytest.mark.parametrize("name", ["two_class", "one_class"])
def test_metric_host_fold_matches_reference_fixture(name):
    """CPU: the host half of the product Metric (deferred fold of per-batch hit counts and losses with the reference's
    float32 / python-float arithmetic) against the reference fixture; the device half (tbn_topk_confusion) is replaced
    here by hit counts computed with the oracle."""
    from oracle.eval_oracle import correct_and_confusion

    gold = _gold()
    cfg, batches = eval_case(name)
    m = evaluation.Metric(cfg, len(batches), torch.device("cpu"))
    for out, target, bs, loss in batches:
        correct = OrderedDict()
        for key in cfg.model.num_classes:
            correct[key], _, cm = correct_and_confusion(out[key].numpy(), target["class"][key].numpy(), max(m.topk))
            m.conf_mat[key] += torch.from_numpy(cm)
        rows = []
        for key in m.accuracy:
            if key == "all_class":
                rows.append([int(np.prod([c[:k].sum(0) for c in correct.values()], axis=0).sum()) for k in m.topk])
            else:
                rows.append([int(correct[key][:k].sum()) for k in m.topk])
        m._pending.append((torch.tensor(rows, dtype=torch.int64), torch.stack([loss[k].float() for k in m.loss]), bs))
    _check(name, *m.get_metrics(), gold)
