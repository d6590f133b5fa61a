#!/usr/bin/env python
"""S-MVAE fusion microbench (BASELINE config 3): PoE + sampling + KL over the latent levels of 128^3 volumes, one subset
vs all 15 missing-modality subsets in ONE launch; achieved GB/s against the algorithmic bytes:
read 5 x (mu, logvar) once (+ noise per subset), write mu, logvar, z per subset."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlstm_hved_b200 import ops  # noqa: E402

ALL15 = ops.SUBSETS_MODALITIES


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = "cuda"
    for B in (8, 32):
        for C, d in ((1, 64), (2, 32), (4, 16), (8, 8)):
            mu = torch.randn(5, B, C, d, d, d, device=dev)
            lv = torch.randn(5, B, C, d, d, d, device=dev)
            n = mu[0].numel()
            for name, subsets in (("1 subset", [(0, 1, 2, 3)]), ("15 subsets", ALL15)):
                ns = len(subsets)
                noise = torch.randn(ns, B, C, d, d, d, device=dev)
                gz = torch.randn(ns, B, C, d, d, d, device=dev)
                f_ms = timeit(lambda: ops.poe_fwd(mu, lv, subsets, noise=noise, want_kld=True))
                b_ms = timeit(lambda: ops.poe_bwd(mu, lv, subsets, noise=noise, g_z=gz, kld_scale=[0.1] * ns))
                fb, bb = n * (40 + ns * 16), n * (40 + ns * 8 + 40)
                print(json.dumps({"B": B, "level": [C, d], "subsets": ns, "elements": n,
                                  "fwd_us": round(f_ms * 1e3, 1), "fwd_gbs": round(fb / f_ms / 1e6), "fwd_frac_of_hbm": round(fb / f_ms / 1e6 / peak, 3),
                                  "bwd_us": round(b_ms * 1e3, 1), "bwd_gbs": round(bb / b_ms / 1e6), "bwd_frac_of_hbm": round(bb / b_ms / 1e6 / peak, 3)}))


def fused_levels():
    """Config 3 as SURVEY 8d words it: ONE launch fusing all 4 latent levels x 15 subsets (inference: no sampling), B in {1, 8, 64};
    bytes: 32 B in per latent element (constant prior declared, not read) + 8 B out per subset."""
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for B in (1, 8, 64):
        levels = []
        for C, d in ((1, 64), (2, 32), (4, 16), (8, 8)):
            mu = torch.randn(5, B, C, d, d, d, device="cuda")
            lv = torch.randn(5, B, C, d, d, d, device="cuda")
            mu[0].zero_(), lv[0].zero_()
            levels.append((mu, lv))
        n = sum(m[0].numel() for m, _ in levels)
        ms = timeit(lambda: ops.poe_fwd_levels(levels, ALL15, standard_prior=True))
        nbytes = n * (32 + 15 * 8)
        print(json.dumps({"fused_levels": 4, "B": B, "subsets": 15, "elements": n, "fwd_us": round(ms * 1e3, 1),
                          "fwd_gbs": round(nbytes / ms / 1e6), "fwd_frac_of_hbm": round(nbytes / ms / 1e6 / peak, 3)}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "levels":
        fused_levels()
    else:
        main()
