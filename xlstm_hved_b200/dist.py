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

    The reference model leaves ~50 parameter tensors without a gradient, and its optimizer (Adam with weight decay,
    train.py:177) skips parameters whose ``grad`` is None.  To keep that behaviour the bucket carries one has-gradient flag
    per parameter behind the gradient data (same buffer, same collective): a parameter that received a gradient on no rank
    keeps ``grad = None``; one that received a gradient on some ranks gets the reduced value everywhere."""

    def __init__(self, params, average: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel + len(self.params), device=p0.device, dtype=torch.float32)
        self.grads = self.flat[:self.numel]
        self.flags = self.flat[self.numel:]
        self.average = average
        self._has = None
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.grads[off:off + p.numel()].view_as(p))
            off += p.numel()

    def reduce(self, group=None):
        """Gather p.grad into the bucket, all-reduce it, write the result back into p.grad.  Returns the gradient part.
        Two multi-tensor copies and one collective per call, whatever the number of parameters."""
        has = tuple(p.grad is not None for p in self.params)
        if has != self._has:                      # upload the flags only when the pattern changes (normally once)
            self.flags.copy_(torch.tensor(has, dtype=torch.float32))
            self._has = has
        with_grad = [(p.grad, v) for p, v in zip(self.params, self.views) if p.grad is not None]
        without = [v for p, v in zip(self.params, self.views) if p.grad is None]
        if with_grad:
            torch._foreach_copy_([v for _, v in with_grad], [g for g, _ in with_grad])
        if without:
            torch._foreach_zero_(without)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        anyone = has
        if world > 1:
            dist.all_reduce(self.flat, group=group)
            if self.average:
                self.grads.div_(world)
            if not all(has):                      # one small D2H only when some gradient is missing on this rank
                anyone = tuple(f > 0 for f in self.flags.tolist())
                self._has = None                  # the flags now hold counts: re-upload next time
        if with_grad:
            torch._foreach_copy_([g for g, _ in with_grad], [v for _, v in with_grad])
        for p, v, h, a in zip(self.params, self.views, has, anyone):
            if not h and a:
                p.grad = v.clone()
        return self.grads
