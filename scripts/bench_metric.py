"""Time Metric.set_metrics per batch: the product (one tbn_topk_confusion launch per class key, no host sync) against
the reference's algorithm as restated by the oracle's structure on the same device tensors (topk + the per-clip Python
loop of metric.py:154-155).  Prints microseconds per batch of 64 clips (verb 125 + noun 352, topk [1, 5])."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from attention_based_tbn_b200 import baseline_config, evaluation  # noqa: E402


def reference_style(out, target, topk, conf):
    maxk = max(topk)
    _, preds = out.topk(maxk, 1, largest=True, sorted=True)
    preds = preds.t()
    correct = preds.eq(target.view(1, -1).expand_as(preds))
    for i1, i2 in zip(target.view(-1), preds[0, :].view(-1)):
        conf[i1, i2] += 1
    return correct


cfg = baseline_config(2)
B = 64
g = torch.Generator().manual_seed(0)
out = {k: torch.randn(B, n, generator=g).cuda() for k, n in cfg.model.num_classes.items()}
tgt = {"class": {k: torch.randint(0, n, (B,), generator=g).cuda() for k, n in cfg.model.num_classes.items()}}
loss = {k: torch.rand((), device="cuda") for k in ("verb", "noun", "all_class", "entropy", "total")}
iters = 50
m = evaluation.Metric(cfg, iters, torch.device("cuda"))
for _ in range(5):
    m.set_metrics(out, tgt, B, loss)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    m.set_metrics(out, tgt, B, loss)
m.get_metrics()
torch.cuda.synchronize()
t_new = (time.perf_counter() - t0) / iters * 1e6
conf = {k: torch.zeros(n, n, device="cuda") for k, n in cfg.model.num_classes.items()}
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    for k in out:
        c = reference_style(out[k], tgt["class"][k], [1, 5], conf[k])
        _ = float(c[:1].reshape(-1).float().sum().mul_(100.0 / B)); _ = float(c[:5].reshape(-1).float().sum().mul_(100.0 / B))
    for k in loss:
        loss[k].item()
torch.cuda.synchronize()
t_ref = (time.perf_counter() - t0) / 5 * 1e6
print(f"Metric.set_metrics, batch {B}: product {t_new:.0f} us/batch (host wall, incl. the final fold), "
      f"reference algorithm on the same GPU {t_ref:.0f} us/batch -> {t_ref / t_new:.1f}x")
