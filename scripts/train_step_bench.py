"""Fine-tuning step benchmark (supplementary to bench.py, which measures the inference forward of BASELINE.json).

One step = the reference's training iteration (core/tools/train.py:68-94) with frozen backbones (model.freeze_base,
freeze_mode "all"): zero_grad -> model.train() forward -> get_loss -> loss.backward() -> gradient all-reduce (N > 1) ->
clip_grad_norm_ -> SGD step.  Workload: cfg2 (RGB + Audio, MHA attention + entropy loss), 64 clips x 3 segments per GPU,
synthetic inputs resident in HBM.  Two set-ups:
  eval-bn : backbones in eval() (fixed BatchNorm statistics) -> inference plans, bf16 tensor-core path; heads fp32
  batch-bn: everything in train() as the reference's loop leaves it -> batch-statistics BatchNorm, fp32 CUDA-core path
  full    : the reference's default freeze_mode "partialbn" -- all conv layers of both backbones train as well (backbone
            backward pass, backbone_train.py), fp32 CUDA-core path

    python scripts/train_step_bench.py [--steps 10] [--warmup 3] [--clips 64]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_step_bench.py

Prints one JSON line per set-up (rank 0).  With N > 1 every rank trains its own shard and the replicas must stay
identical: the parameter checksums of all ranks are compared after the timed steps."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from attention_based_tbn_b200 import TBNModel, parallel  # noqa: E402
from attention_based_tbn_b200.config import baseline_config, get_modality  # noqa: E402
from attention_based_tbn_b200.synthetic import make_inputs, make_state_dict  # noqa: E402


def reference_arm(args):
    """The reference's own training step (train.py:68-94) on the CPU: unmodified TBNModel from baseline/_ref, fp32 PyTorch,
    all host threads, a bounded sample of the same workload (cfg2, --ref-clips clips x 3 segments)."""
    import contextlib
    import statistics
    import time

    os.environ["TBN_REFERENCE_ROOT"] = os.path.join(ROOT, "baseline", "_ref")
    from oracle import ref_shim

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    calib = os.path.join(ROOT, "tests", "golden", "bn_calib.npz")
    B, N = args.ref_clips, 3
    crit = {"crossentropy": torch.nn.CrossEntropyLoss()}
    for mode in args.modes.split(","):
        if mode == "eval-bn":
            continue  # (not a set-up of the reference's loop: it leaves every module in train())
        cfg = baseline_config(2)
        cfg.model.freeze_mode = "partialbn" if mode == "full" else "all"
        cfg.model.attention.decay_step = 1
        mod = get_modality(cfg)
        sd = make_state_dict(cfg, mod, seed=0, calib=calib)
        with contextlib.redirect_stdout(sys.stderr):
            model = ref_shim.build_reference_model(cfg, mod, sd).train()
        inp = make_inputs(cfg, mod, B, N, seed=100)
        g = torch.Generator().manual_seed(200)
        target = {"class": {k: torch.randint(0, int(v), (B,), generator=g) for k, v in cfg.model.num_classes.items()}}
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.SGD(params, lr=0.01, momentum=0.9)
        ts = []
        for it in range(3):
            t0 = time.perf_counter()
            opt.zero_grad()
            out = model(dict(inp))
            loss, _ = model.get_loss(crit, target, out, 20)
            loss["total"].backward()
            torch.nn.utils.clip_grad_norm_(params, 20.0)
            opt.step()
            if it:
                ts.append(time.perf_counter() - t0)
        med = statistics.median(ts)
        print(json.dumps({"impl": "reference", "metric": "clips/sec TBN RGB+Audio MHA training step" + ("" if mode == "full" else
                          " (frozen backbones)"), "mode": mode, "value": B / med, "unit": "clips/s", "ms_per_step": med * 1e3,
                          "cores": cores, "sample": f"{B} clips x {N} segments, fp32 PyTorch CPU, unmodified reference TBNModel "
                          "(baseline/_ref), 1 warm-up + 2 timed steps, median"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--modes", default="eval-bn,batch-bn,full")
    ap.add_argument("--train-precision", default="bf16", choices=["fp32", "bf16"],
                    help="arithmetic of the training convs: bf16 = mixed precision on the tensor cores (default here), fp32 = the parity mode")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"],
                    help="reference: the UNMODIFIED reference model (baseline/_ref) doing the same step on the host CPU cores")
    ap.add_argument("--ref-clips", type=int, default=8)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mod = None
    calib = os.path.join(ROOT, "tests", "golden", "bn_calib.npz")
    cfg = baseline_config(2)
    mod = get_modality(cfg)
    sd = make_state_dict(cfg, mod, seed=0, calib=calib)
    B, N = args.clips, 3
    inp = {k: v.to(dev) for k, v in make_inputs(cfg, mod, B, N, seed=100 + rank).items()}  # every rank its own shard
    g = torch.Generator().manual_seed(200 + rank)
    target = {"class": {k: torch.randint(0, int(v), (B,), generator=g).to(dev) for k, v in cfg.model.num_classes.items()}}
    crit = {"crossentropy": torch.nn.CrossEntropyLoss()}
    for mode in args.modes.split(","):
        cfg = baseline_config(2)
        cfg.model.freeze_mode = "partialbn" if mode == "full" else "all"
        cfg.model.attention.decay_step = 1
        mod = get_modality(cfg)
        model = TBNModel(cfg, mod, dev, precision="bf16", train_precision=args.train_precision)
        model.load_state_dict(sd)
        model.train()
        if mode == "eval-bn":
            for m in mod:
                getattr(model, f"Base_{m}").eval()
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.SGD(params, lr=0.01, momentum=0.9)
        torch.manual_seed(1 + rank)

        def step():
            opt.zero_grad(set_to_none=False)
            out = model(inp)
            loss, _ = model.get_loss(crit, target, out, epoch=20)
            loss["total"].backward()
            n_coll = parallel.all_reduce_gradients(params)
            torch.nn.utils.clip_grad_norm_(params, 20.0)
            opt.step()
            return loss["total"], n_coll

        first = None
        for _ in range(args.warmup):
            l, _ = step()
            first = float(l) if first is None else first
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            l, n_coll = step()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        check = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
        lo, hi = check.clone(), check.clone()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({
                "metric": "clips/sec TBN RGB+Audio MHA training step" + ("" if mode == "full" else " (frozen backbones)"), "mode": mode, "n_gpus": world,
                "value": B * world / (ms.item() / 1e3), "unit": "clips/s", "ms_per_step": ms.item(), "steps": args.steps,
                "warmup": args.warmup, "clips_per_gpu": B, "segments": N, "trainable_parameters": sum(p.numel() for p in params),
                "train_precision": args.train_precision, "gradient_collectives_per_step": n_coll, "loss_first": first, "loss_last": float(l),
                "forward_launches": model.last_launches, "backward_launches": model.last_backward_launches,
                "replicas_in_sync": bool((hi - lo).abs().item() <= 1e-9 * max(abs(hi.item()), 1.0)),
                "backbone": {"eval-bn": "bf16 tensor-core inference plans (BatchNorm folded, eval statistics)",
                             "batch-bn": "batch-statistics BatchNorm, training forward",
                             "full": "batch-statistics BatchNorm, forward + backward (all convs train)"}[mode],
                "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                "heads": "fp32 forward + backward kernels (csrc/heads.cu, csrc/head_bwd.cu)", "data": "synthetic"}), flush=True)
        del model, opt
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
