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
