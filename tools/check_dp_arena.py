"""2-rank check of the in-place gradient-arena all-reduce (parallel.BucketedGradAllReduce) against the
flatten / scatter fallback on the same inputs: run with torchrun --nproc-per-node 2.  Dev tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from tests.fusion_testlib import build_module, param_dict, run_module
from transfusion_b200.parallel import BucketedGradAllReduce, level_buckets

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
m = build_module(256, [(16, 24), (8, 12)], [32, 64], [2, 1], [2, 2], 4, seed=3)   # same weights on both ranks
m.train()
g = torch.Generator().manual_seed(100 + rank)
feats = {"0": torch.relu(torch.randn(2, 32, 16, 24, generator=g)).cuda(), "1": torch.relu(torch.randn(2, 64, 8, 12, generator=g)).cuda()}
lang = (0.5 * torch.randn(2, 12, 256, generator=g)).cuda()
mask = torch.ones(2, 12, dtype=torch.int64).cuda()


def step(use_arena):
    red = BucketedGradAllReduce(level_buckets(m))
    if not use_arena:
        red._arena_of = lambda bucket: None
    m.zero_grad(set_to_none=True)
    red.reset()
    out, _ = run_module(m, {k: v.clone() for k, v in feats.items()}, lang.clone(), mask)
    sum(o.float().sum() for o in out.values()).backward()
    red.finish()
    torch.cuda.synchronize()
    red.remove()
    return {k: p.grad.detach().clone() for k, p in param_dict(m).items() if p.grad is not None}


a = step(True)
b = step(False)
worst = max(float((a[k] - b[k]).norm() / (b[k].norm() + 1e-12)) for k in b)
# identical on both ranks after the reduction
chk = torch.stack([v.double().sum() for v in a.values()]).sum().reshape(1).cuda()
allc = [torch.zeros_like(chk) for _ in range(2)]
dist.all_gather(allc, chk)
if rank == 0:
    print(f"arena vs fallback worst rel diff {worst:.3e}; rank checksums {[float(x) for x in allc]}")
    assert worst < 1e-5 and abs(float(allc[0]) - float(allc[1])) <= 1e-6 * abs(float(allc[0]))
    print("check_dp_arena ok")
dist.destroy_process_group()
