"""Model-level driver for the reference's missing-modality evaluation (BASELINE configs[2]).

The reference evaluates a volume under each of the 15 modality subsets with 15 separate forwards of batch 1
(test.py:78-102, evaluation.py:306-342: zero the missing modalities, ``model(x, [idx], valid=True)``).  The model already
supports per-SAMPLE missing modalities (``instance_missing=True`` with a (B, 4) ``drop`` mask, RA_HVED.py:510-513: the fusion
then runs through ``ProductOfExperts2``, buildingblocks.py:875-886), and every normalisation layer is per-sample in eval
mode, so the 15 masked copies of a volume can ride through the model as ONE batch: one forward, one S-MVAE launch per latent
level with the per-sample drop mask (``xhved_poe_fwd(drop=...)``), convolutions with 15x the batch.  The results are the
ones the 15 separate forwards give (``tests/test_gpu_model_e2e.py::test_all_subsets_in_one_batch_matches_15_forwards``).
"""
from __future__ import annotations

import contextlib
import io

import torch

from .ops import SUBSETS_MODALITIES


def subset_batch(x: torch.Tensor, subsets=None):
    """x: (B, 4, D, H, W).  Returns (x_all, drop): x_all (len(subsets) * B, 4, D, H, W) holds, subset-major, a copy of the
    batch per subset with the missing modalities zeroed (evaluation.py:306-307); drop (len(subsets) * B, 4) bool marks them."""
    subsets = list(range(len(SUBSETS_MODALITIES))) if subsets is None else list(subsets)
    B = x.shape[0]
    present = torch.zeros(len(subsets), 4, dtype=torch.bool, device=x.device)
    for r, idx in enumerate(subsets):
        present[r, list(SUBSETS_MODALITIES[idx])] = True
    x_all = x.unsqueeze(0) * present.view(len(subsets), 1, 4, 1, 1, 1).to(x.dtype)
    drop = (~present).unsqueeze(1).expand(len(subsets), B, 4)
    return x_all.reshape(len(subsets) * B, *x.shape[1:]), drop.reshape(len(subsets) * B, 4).contiguous()


@torch.no_grad()
def all_subsets_forward(model, x: torch.Tensor, subsets=None, max_batch: int = 15):
    """Segmentation of every volume of ``x`` under every requested modality subset: (len(subsets), B, classes, D, H, W).
    ``model``: a reference ``XLSTM_HVED`` in eval mode (patched or not); the masked copies run ``max_batch`` at a time."""
    subsets = list(range(len(SUBSETS_MODALITIES))) if subsets is None else list(subsets)
    x_all, drop = subset_batch(x, subsets)
    outs = []
    for lo in range(0, x_all.shape[0], max_batch):
        with contextlib.redirect_stdout(io.StringIO()):                 # the reference prints from inside its forward
            seg, _ = model(x_all[lo:lo + max_batch], [14], instance_missing=True, drop=drop[lo:lo + max_batch], valid=True)
        outs.append(seg)
    seg = torch.cat(outs, 0)
    return seg.reshape(len(subsets), x.shape[0], *seg.shape[1:])


class GraphedSubsetsForward:
    """The evaluation forward of a (patched) reference ``XLSTM_HVED`` replayed as ONE CUDA graph per input shape.

    Eager, a 128^3 forward of the patched model is ~900 launches issued from Python (14.6 ms for 9.5 ms of device work); the
    reference's subset-index call form also builds the drop mask on the host and copies it inside the forward (RA_HVED.py:515-520),
    which a stream capture refuses -- so the graph is recorded over the per-sample form (``instance_missing=True`` with a device
    ``drop`` mask, as :func:`all_subsets_forward`), with static input / mask buffers that ``__call__`` refills.

        fwd = xh.GraphedSubsetsForward(model)             # model.eval(), patched
        seg = fwd(x, subsets=[14])                        # (1, B, classes, D, H, W); the graph of this shape is recorded on first use
        seg_all = fwd(x)                                  # all 15 subsets as one batch of 15 B

    The returned tensor is a clone (the static output is overwritten by the next replay).  Parameters are read in place:
    ``load_state_dict`` / optimizer steps are seen by later replays; adding or replacing parameter tensors is not.
    """

    def __init__(self, model, warmup: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("xlstm_hved_b200 has no CPU path")
        if model.training:
            raise RuntimeError("GraphedSubsetsForward records the evaluation forward: call model.eval() first")
        self.model, self.warmup, self._graphs = model, warmup, {}

    def _record(self, x_all, drop):
        static_x, static_drop = x_all.clone(), drop.clone()

        def run():
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                return self.model(static_x, [14], instance_missing=True, drop=static_drop, valid=True)[0]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                run()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = run()
        return graph, static_x, static_drop, out

    def __call__(self, x: torch.Tensor, subsets=None):
        subsets = list(range(len(SUBSETS_MODALITIES))) if subsets is None else list(subsets)
        x_all, drop = subset_batch(x, subsets)
        key = (x_all.device, tuple(x_all.shape), x_all.dtype)
        if key not in self._graphs:
            self._graphs[key] = self._record(x_all, drop)
        graph, static_x, static_drop, out = self._graphs[key]
        static_x.copy_(x_all)
        static_drop.copy_(drop)
        graph.replay()
        return out.clone().reshape(len(subsets), x.shape[0], *out.shape[1:])
