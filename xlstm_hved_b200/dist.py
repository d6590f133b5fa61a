"""Multi-GPU plumbing for the hot path: one process per GPU (torchrun), volumes sharded across ranks, and ONE
flat-bucket all-reduce of the parameter gradients per training step (replaces the reference's single-process
``nn.DataParallel`` replicate/scatter/gather/reduce_add_coalesced, train.py:148-151).  Backend NCCL on GPUs (NVLink 5 /
NVSwitch); the same code runs over gloo on CPU for tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items independent units (volumes) for this rank; sizes differ by at most 1."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradBucket:
    """Sum (or average) the gradients of ``params`` across ranks with a single collective on one contiguous buffer.
    Parameters without a gradient contribute zeros (the reference model leaves ~50 tensors without one)."""

    def __init__(self, params, average: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, device=p0.device, dtype=torch.float32)
        self.average = average
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def reduce(self, group=None):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, group=group)
            if self.average:
                self.flat.div_(dist.get_world_size(group))
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat
