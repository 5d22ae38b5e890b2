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


def _overlap_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tdvc.dp import BucketedReducer
    from tdvc.optim import FusedAdamW
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                              torch.nn.Linear(40, 3))
    unused = torch.nn.Parameter(torch.ones(7))                  # never receives a gradient
    params = list(net.parameters()) + [unused]
    opt = FusedAdamW(params, 1e-3)
    red = BucketedReducer(opt, bucket_mb=40 * 40 * 4 / (1 << 20))   # ~one 40x40 weight per bucket: several buckets
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 6)
    fired_before_finish = []
    for it in range(2):                                          # twice: re-arming works
        for p in params:
            p.grad = None
        red.arm()
        net(x).square().sum().backward()
        fired_before_finish.append(list(red.order))
        red.finish()
    local = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in params]
    bank = opt.bank(0).clone()
    q.put((rank, [g.tolist() for g in local], bank.tolist(), len(red.buckets), fired_before_finish))
    dist.destroy_process_group()


def test_bucketed_reducer_overlaps_and_averages_two_ranks():
    """tdvc.dp.BucketedReducer: buckets are cut in reverse parameter order, each is all-reduced from the gradient hook of
    its last parameter (so before the backward has finished), parameters without gradient travel as zeros, and the bank
    ends up holding the mean over ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g_a, bank_a, nb_a, fired_a), (_, g_b, bank_b, nb_b, fired_b) = res
    assert nb_a == nb_b and nb_a >= 3
    flat_a = torch.cat([torch.tensor(g).reshape(-1) for g in g_a])
    flat_b = torch.cat([torch.tensor(g).reshape(-1) for g in g_b])
    want = 0.5 * (flat_a + flat_b)
    assert torch.allclose(torch.tensor(bank_a), want, rtol=1e-6, atol=1e-7)
    assert bank_a == bank_b
    assert torch.all(want[-7:] == 0)                             # the unused parameter
    for fired in (fired_a, fired_b):
        for order in fired:
            # every bucket whose parameters all received gradients was reduced from a hook, during the backward, and in
            # reverse layer order (bucket 0 holds the LAST parameters but also the unused one: it waits for finish())
            assert order == sorted(order) and len(order) >= nb_a - 1 and 0 not in order
