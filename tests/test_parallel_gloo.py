"""CPU, world_size 2, gloo: the N>1 path (clip sharding + score gather) without a GPU."""
import os
import socket
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeModel:
    """Stands in for the CUDA model: deterministic function of the clips it is given."""

    def __call__(self, inp):
        x = inp["RGB"]  # (b, n, 3, h, w)
        b, n = x.shape[:2]
        s = x.reshape(b, n, -1).mean(-1)  # (b, n)
        out = OrderedDict()
        out["verb"] = s.mean(1, keepdim=True).repeat(1, 5)
        out["noun"] = s.mean(1, keepdim=True).repeat(1, 7) * 2
        out["weights"] = s.reshape(b * n, 1, 1).repeat(1, 1, 13)
        return out


def _worker(rank, world, port, B, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from attention_based_tbn_b200 import parallel

    g = torch.Generator().manual_seed(0)
    full = {"RGB": torch.rand(B, N, 3, 4, 4, generator=g), "vid_id": [f"v{i}" for i in range(B)]}
    local = parallel.shard_input(full)
    out = parallel.data_parallel_forward(_FakeModel(), full)
    ref = _FakeModel()(full)
    ok = all(torch.allclose(out[k], ref[k]) for k in ref) and len(local["vid_id"]) == local["RGB"].shape[0]
    q.put((rank, ok, local["RGB"].shape[0], tuple(out["weights"].shape)))
    dist.barrier()
    dist.destroy_process_group()


def _run_world2(B, N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    return res


def test_shard_and_gather_world2():
    B, N = 5, 3  # uneven split: 3 + 2 clips
    res = _run_world2(B, N)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [3, 2]
    assert res[0][3] == (B * N, 1, 13)


def test_fewer_clips_than_ranks_world2():
    """Last partial batch of an evaluation: 1 clip on 2 ranks.  The idle rank runs a stand-in clip, contributes no
    rows, and both ranks still receive the full result (no rank may skip the collective)."""
    B, N = 1, 3
    res = _run_world2(B, N)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [1, 0]
    assert res[1][3] == (B * N, 1, 13)


def test_pack_unpack_round_trip():
    from attention_based_tbn_b200 import parallel

    g = torch.Generator().manual_seed(1)
    local = OrderedDict(verb=torch.randn(4, 125, generator=g), noun=torch.randn(4, 352, generator=g),
                        weights=torch.rand(12, 1, 13, generator=g))
    mat, layout = parallel.pack_rows(local, 3, 6)
    assert mat.shape == (6, 125 + 352 + 39) and not mat[3:].any()
    back = parallel.unpack_rows(mat[:3], layout)
    assert torch.equal(back["verb"], local["verb"][:3]) and torch.equal(back["noun"], local["noun"][:3])
    assert back["weights"].shape == (9, 1, 13) and torch.equal(back["weights"], local["weights"][:9])


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from attention_based_tbn_b200 import parallel

    g = torch.Generator().manual_seed(5)
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((300, 40), (7,), (64, 64), (5, 3, 3))]
    params[1].requires_grad = False
    for i, p in enumerate(params):
        if p.requires_grad and not (rank == 1 and i == 2):  # rank 1 has no gradient for one parameter: counts as zeros
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = parallel.all_reduce_gradients(params, bucket_bytes=10000)
    exp0 = (1 + 2) / 2 * 1
    exp2 = (1 * 3 + 0) / 2
    exp3 = (1 + 2) / 2 * 4
    ok = (n >= 2 and torch.allclose(params[0].grad, torch.full_like(params[0], exp0)) and params[1].grad is None
          and torch.allclose(params[2].grad, torch.full_like(params[2], exp2))
          and torch.allclose(params[3].grad, torch.full_like(params[3], exp3)))
    q.put((rank, ok, n))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_all_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True], res
