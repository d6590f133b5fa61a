"""Host-side multi-rank logic on CPU (gloo, world_size 2): batch sharding and the flat-bucket gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xlstm_hved_b200 import ViLBlock, SequenceTraversal
    from xlstm_hved_b200.dist import FlatGradBucket, shard_range
    torch.manual_seed(0)
    blk = ViLBlock(32, SequenceTraversal.ROWWISE_FROM_TOP_LEFT)          # parameter container only (no CUDA compute here)
    params = list(blk.parameters())
    lo, hi = shard_range(5, rank, world)
    # a fake per-volume gradient: volume v contributes (v+1) to every element.  Tensor 3 has no grad on any rank (it must
    # stay None: the reference's Adam skips such parameters, train.py:177); tensor 5 has a grad on rank 1 only (every
    # rank must end up with the sum)
    for i, p in enumerate(params):
        if i == 3 or (i == 5 and rank == 0):
            continue
        p.grad = torch.full_like(p, float(sum(v + 1 for v in range(lo, hi))))
    bucket = FlatGradBucket(params, average=False)
    flat = bucket.reduce()
    ok = bucket.numel == sum(p.numel() for p in params) == 8936
    ok &= params[3].grad is None
    ok &= torch.all(params[5].grad == 9.0).item()                         # volumes 3, 4 live on rank 1: 4 + 5
    ok &= all(torch.all(p.grad == 15.0).item() for i, p in enumerate(params) if i not in (3, 5))
    ok &= flat.numel() == 8936
    out[rank] = (bool(ok), (lo, hi))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert out[0][0] and out[1][0]
        assert out[0][1] == (0, 3) and out[1][1] == (3, 5)


def test_shard_range_covers_everything():
    from xlstm_hved_b200.dist import shard_range
    for n in (1, 7, 8, 15, 64):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
