"""world_size-2 data-parallel plumbing on CPU (gloo): the gradient bucket all-reduce and batch sharding."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tdvc.dp import GradAverager, broadcast_parameters, shard_batch
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    if rank == 1:
        with torch.no_grad():
            lin.weight.add_(1.0)
    broadcast_parameters(lin)
    w0 = lin.weight.detach().clone()
    params = list(lin.parameters())
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    avg = GradAverager()
    avg("net", params)
    avg("net", params)   # idempotent on already-averaged grads, reuses the bucket
    full = {"x": torch.arange(8.).view(4, 2), "lab": torch.arange(4), "neg": [torch.arange(12).view(4, 3)], "k": 3}
    sh = shard_batch(full, rank, world)
    # plain lists: tensors would travel by fd-sharing and this process exits before the parent reads them
    sh_l = {k: (v.tolist() if torch.is_tensor(v) else [t.tolist() for t in v] if isinstance(v, list) else v)
            for k, v in sh.items()}
    q.put((rank, w0.tolist(), [p.grad.tolist() for p in params], sh_l))
    dist.destroy_process_group()


def test_grad_average_and_shard_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w_a, g_a, s_a), (_, w_b, g_b, s_b) = res
    assert w_a == w_b                                          # broadcast made the replicas identical
    for i, (ga, gb) in enumerate(zip(g_a, g_b)):
        ga_t = torch.tensor(ga)
        assert torch.allclose(ga_t, torch.full_like(ga_t, 1.5 * (i + 1)))   # mean of (1, 2) * (i+1)
        assert ga == gb
    full = torch.arange(8.).view(4, 2)
    assert s_a["x"] == full[:2].tolist() and s_b["x"] == full[2:].tolist()
    assert s_b["lab"] == [2, 3] and s_a["k"] == 3
    assert s_b["neg"][0] == torch.arange(12).view(4, 3)[2:].tolist()
