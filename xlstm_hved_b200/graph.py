"""CUDA-graph replay of a step that calls the package's ops.

A forward+backward pass over the hot path is ~25 kernel launches of 15-100 us each: issued one by one from Python they cost
more host time (1.6 ms) than the kernels take on the device (1.2 ms).  The kernels keep no host-visible state, allocate
nothing themselves and run on the current stream, so the whole step -- autograd included -- can be stream-captured once and
replayed (``torch.cuda.graph``).  Everything the step reads must live in buffers that stay at the same address between
replays; refill them in place (``tensor.copy_``) before calling the graphed step again.
"""
from __future__ import annotations

import torch


class GraphedStep:
    """Capture ``fn()`` (no arguments: it reads static device buffers) once and replay it.

    ``fn`` is run ``warmup`` times eagerly on a side stream first (lazy initialisation, allocator warm-up), then recorded.
    ``__call__`` replays the graph on the current stream and returns the captured outputs (tensors that are overwritten by
    the next replay).  Collectives and host-side decisions belong outside ``fn``.
    """

    def __init__(self, fn, warmup: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("xlstm_hved_b200 has no CPU path")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

    def __call__(self):
        self.graph.replay()
        return self.outputs
