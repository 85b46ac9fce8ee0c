"""Aggregate pinned host -> device copy bandwidth with N ranks copying at the same time (torchrun, one rank per GPU).
Every rank copies the cfg-2 step input (198 MB, fp32) `iters` times from its own pinned buffer; device-timed per rank,
rank 0 prints per-rank and aggregate GB/s.  This is the ceiling of the fp32-input `e2e` leg of bench.py at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_probe_ranks.py
"""
import datetime
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
nbytes, iters = 198_377_472, 20
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = torch.tensor([nbytes * iters / e0.elapsed_time(e1) / 1e6], device="cuda", dtype=torch.float64)
allg = [torch.zeros_like(gbs) for _ in range(world)]
if world > 1:
    dist.all_gather(allg, gbs)
else:
    allg = [gbs]
if rank == 0:
    per = [round(float(g.item()), 1) for g in allg]
    # weak scaling: a step ends when the SLOWEST rank has its 198 MB, so the ceiling is set by the lowest per-rank rate
    step_ms = nbytes / (min(per) * 1e9) * 1e3
    print(f"h2d_probe ranks={world} per-rank GB/s={per} aggregate={sum(per):.1f} GB/s -> slowest rank needs {step_ms:.2f} ms "
          f"per 198 MB step: fp32 e2e ceiling {world * 64 / step_ms * 1e3:.0f} clips/s", flush=True)
if world > 1:
    dist.destroy_process_group()
