"""CPU, world_size 2 over gloo: the N>1 host path — batch sharding and the bucketed gradient
all-reduce that bench.py uses over NCCL (same code, different backend)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transfusion_b200.parallel import BucketedGradAllReduce, shard_range


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w1 = torch.nn.Parameter(torch.randn(5, 3))
    w2 = torch.nn.Parameter(torch.randn(3))
    frozen = torch.nn.Parameter(torch.randn(2), requires_grad=False)
    unused = torch.nn.Parameter(torch.randn(4))   # requires grad but takes no part in the graph (the reference's
    red = BucketedGradAllReduce([[w1, frozen, unused], [w2]])   # heatmap_token): must not switch the exchange off
    x_all = torch.arange(8 * 5, dtype=torch.float32).reshape(8, 5) / 10.0
    lo, hi = shard_range(8, rank, world)
    for it in range(2):  # two steps: buffers are reused
        w1.grad = None
        w2.grad = None
        red.reset()
        loss = ((x_all[lo:hi] @ w1 + w2) ** 2).sum()
        loss.backward()
        red.finish()
    # single-process reference: mean over ranks of per-rank sums
    w1r, w2r = w1.detach().clone().requires_grad_(True), w2.detach().clone().requires_grad_(True)
    tot = sum(((x_all[slice(*shard_range(8, r, world))] @ w1r + w2r) ** 2).sum() for r in range(world)) / world
    tot.backward()
    ok = torch.allclose(w1.grad, w1r.grad, rtol=1e-5, atol=1e-5) and torch.allclose(w2.grad, w2r.grad, rtol=1e-5, atol=1e-5)
    ok = ok and unused.grad is None
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)
    assert sorted(r[2] for r in res) == [(0, 4), (4, 8)]


def _worker_accumulate(rank, world, port, q):
    """accumulate_grad_batches = 2 (ego_nao_res50_ego4dv2.yml:125): the first micro-step only accumulates
    (reducer.active = False, DDP no_sync), the second all-reduces the accumulated gradients."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w1 = torch.nn.Parameter(torch.randn(5, 3))
    w2 = torch.nn.Parameter(torch.randn(3))
    red = BucketedGradAllReduce([[w1], [w2]])
    x_all = torch.arange(16 * 5, dtype=torch.float32).reshape(16, 5) / 20.0
    micro = [x_all[:8], x_all[8:]]
    lo, hi = shard_range(8, rank, world)
    w1.grad = None
    w2.grad = None
    for j, xb in enumerate(micro):
        red.active = j == len(micro) - 1
        if red.active:
            red.reset()
        ((xb[lo:hi] @ w1 + w2) ** 2).sum().backward()
        if red.active:
            red.finish()
    w1r, w2r = w1.detach().clone().requires_grad_(True), w2.detach().clone().requires_grad_(True)
    tot = sum(((xb[slice(*shard_range(8, r, world))] @ w1r + w2r) ** 2).sum() for xb in micro for r in range(world)) / world
    tot.backward()
    ok = torch.allclose(w1.grad, w1r.grad, rtol=1e-5, atol=1e-5) and torch.allclose(w2.grad, w2r.grad, rtol=1e-5, atol=1e-5)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradient_accumulation_reduces_once_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_accumulate, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)
