"""Data-parallel plumbing for the fusion path (one process per GPU, torch.distributed / NCCL).

The path shards by batch exactly like the reference's Lightning DDP (run_experiment.py:373-374,452):
rank r takes samples [r*B, (r+1)*B), there is no forward collective, and the only exchange is one
sum-all-reduce of the parameter gradients per step.  ``BucketedGradAllReduce`` launches that
all-reduce per bucket (one bucket per FPN level by default) as soon as every gradient of the bucket
has been accumulated, so it overlaps the backward of the remaining levels; NCCL runs it over
NVLink 5 / NVSwitch.  Works with any backend (gloo on CPU for the tests)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Per-rank slice of a global batch: the reference splits ``bs // n_devices`` per process
    (run_experiment.py:373-374); the remainder is dropped, as there."""
    per = global_batch // world
    return rank * per, (rank + 1) * per


class BucketedGradAllReduce:
    def __init__(self, buckets: Sequence[Iterable[torch.nn.Parameter]], group: Optional[dist.ProcessGroup] = None,
                 average: bool = True, compress: Optional[str] = None):
        """compress = "bf16": the level arenas are exchanged as bf16 (half the NVLink bytes and half the time NCCL's
        channel CTAs hold SMs the persistent compute kernels want); the sum runs in NCCL's bf16 reduction, the result is
        unpacked to fp32 with the 1 / world factor.  Default None = fp32 like the reference's DDP."""
        self.group = group
        self.average = average
        if compress not in (None, "bf16"):
            raise ValueError("compress must be None or 'bf16'")
        self.compress = compress
        self._packed = {}
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets: List[List[torch.nn.Parameter]] = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self.flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self.pending = [0] * len(self.buckets)
        self.works = []
        self._handles = []
        # False = accumulate only (DDP no_sync: the non-final micro-steps of accumulate_grad_batches, run_experiment.py:444)
        self.active = True
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
        self.reset()

    def _offsets(self, bi):
        off, out = 0, []
        for p in self.buckets[bi]:
            out.append((off, p.numel()))
            off += p.numel()
        return out, off

    def _make_hook(self, bi):
        def hook(param):
            if not self.active:
                return
            self.pending[bi] -= 1
            if self.pending[bi] == 0:
                self._launch(bi)
        return hook

    @staticmethod
    def _arena_of(bucket):
        """The level's gradient arena (cross_fusion/level_fn.py) if EVERY gradient of the bucket is a view into it
        (true when autograd adopted the freshly produced gradients, i.e. after zero_grad(set_to_none=True))."""
        arena = getattr(bucket[0], "_xf_grad_arena", None)
        if arena is None or not arena.is_cuda:
            return None
        lo, hi = arena.data_ptr(), arena.data_ptr() + arena.numel() * arena.element_size()
        for p in bucket:
            g = p.grad
            if g is None:
                continue   # took no part in this backward: nothing to reduce (every rank runs the same graph)
            if g.dtype != arena.dtype or not (lo <= g.data_ptr() and g.data_ptr() + g.numel() * g.element_size() <= hi):
                return None
        return arena

    def _launch(self, bi):
        bucket = self.buckets[bi]
        arena = self._arena_of(bucket)
        if arena is not None:
            self._arena_buckets.add(bi)
        if arena is not None and self.compress == "bf16" and arena.is_cuda and arena.numel() % 64 == 0:
            # bf16-compressed exchange: pack (one cast launch), all-reduce half the bytes, unpack + average in finish()
            from . import ops
            n = arena.numel()
            buf = self._packed.get(bi)
            if buf is None or buf.numel() != n or buf.device != arena.device:
                buf = self._packed[bi] = torch.empty(n, dtype=torch.bfloat16, device=arena.device)
            ops.cast_pad(arena.view(n // 64, 64), buf.view(n // 64, 64), n // 64, 64)
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if self.world > 1 else None
            self.works.append((bi, work, ("bf16", arena, buf)))
            return
        if arena is not None:   # one in-place all-reduce of the level's arena: no flatten / scatter copies
            if self.world > 1:
                # NCCL averages inside the collective; other backends sum and finish() divides
                avg = self.average and dist.get_backend(self.group) == "nccl"
                op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
                self.works.append((bi, dist.all_reduce(arena, op=op, group=self.group, async_op=True), None if avg else arena))
            else:
                self.works.append((bi, None, None))
            return
        offs, total = self._offsets(bi)
        dev = next((p.grad.device for p in bucket if p.grad is not None), bucket[0].device)
        if self.flat[bi] is None or self.flat[bi].device != dev:
            self.flat[bi] = torch.empty(total, dtype=torch.float32, device=dev)
        flat = self.flat[bi]
        for p, (o, n) in zip(bucket, offs):
            if p.grad is None:          # unused on this rank (e.g. the reference's heatmap_token): contributes zeros
                flat[o:o + n].zero_()
            else:
                flat[o:o + n].copy_(p.grad.reshape(-1))
        if self.world > 1:
            self.works.append((bi, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), None))
        else:
            self.works.append((bi, None, None))

    def reset(self):
        """Call before each backward."""
        self.pending = [len(b) for b in self.buckets]
        self.works = []
        self._arena_buckets = set()

    def finish(self):
        """Wait for the outstanding all-reduces and scatter the (averaged) result back into .grad.  Buckets whose
        hooks did not all fire (a parameter that took no part in this backward) are reduced here, so an unused
        parameter cannot silently switch the exchange off."""
        launched = {w[0] for w in self.works}
        for bi, bucket in enumerate(self.buckets):
            if bi not in launched and any(p.grad is not None for p in bucket):
                self._launch(bi)
        for bi, work, arena in self.works:
            if work is not None:
                work.wait()
            if isinstance(arena, tuple):   # bf16-compressed: unpack into the fp32 arena the .grad views live in
                from . import ops
                _, dst, buf = arena
                ops.bf16_to_f32(buf, dst, 1.0 / self.world if (self.average and self.world > 1) else 1.0)
                continue
            if arena is not None:
                if self.average and self.world > 1:
                    arena.div_(self.world)
                continue
            if bi in self._arena_buckets:
                continue   # reduced (and averaged) in place
            flat = self.flat[bi]
            if self.average and self.world > 1:
                flat.div_(self.world)
            offs, _ = self._offsets(bi)
            for p, (o, n) in zip(self.buckets[bi], offs):
                if p.grad is not None:
                    p.grad = flat[o:o + n].view_as(p)
        self.works = []

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def level_buckets(wrapper) -> List[List[torch.nn.Parameter]]:
    """One gradient bucket per FPN level (patch-embed + encoder + back-projection of that level)."""
    out = []
    for i in range(len(wrapper.cross_fusion_encoders)):
        ps = list(wrapper.patches_to_token[i].parameters()) + list(wrapper.cross_fusion_encoders[i].parameters()) + \
            list(wrapper.tokens_to_features[i].parameters())
        out.append(ps)
    if hasattr(wrapper, "lm_layer"):
        out.append(list(wrapper.lm_layer.parameters()))
    return out
